"""TEST INFRASTRUCTURE ONLY.  The reference's per-instance body on the host CPU, assembled from the oracle
pieces: SHOT (oracle/shot_oracle.cpp), heads (oracle/heads_torch.py, torch-CPU float32 like the reference's
nn.Linear stacks), decode + votes + pose (oracle/cppf_oracle.c via oracle/cpu.py).  bench.py times it as
the CPU baseline / `--impl reference` arm; tests use it as the end-to-end checker."""
from __future__ import annotations

import time

import numpy as np
import torch

from . import cpu as oracle
from .heads_torch import Ref


def instance_pose_cpu(pc, idx, cfg: dict, state_dicts: dict, desc=None, bins=None, seed=0, threads=0, timings=None,
                      num_rots=180, sym_y_only=False, opt=False):
    """eval.py:207-372 for one instance; `opt` adds the online refinement (eval.py:319-355, oracle/refine_torch.py).
    `state_dicts` = {"dino": sd, "shot": sd} (either optional).
    `bins[branch]` injects the multinomial draws; otherwise torch.multinomial on the CPU with `seed`.
    Returns {branch: oracle.instance_body dict} plus the winning branch under key "best"."""
    t0 = time.perf_counter()
    res = float(cfg["res"])
    shot_threads = threads if threads > 0 else oracle.num_threads()
    d, n = oracle.shot_compute(pc, res * 10, res * 10, threads=shot_threads)
    shot_feat = np.nan_to_num(d.reshape(-1, 352), nan=0.0)            # eval.py:215-216
    normal = np.nan_to_num(n.reshape(-1, 3), nan=0.0)
    t1 = time.perf_counter()
    out, best, best_loss, dino_scale = {}, None, np.inf, None
    tpc, tidx = torch.from_numpy(pc), torch.from_numpy(np.asarray(idx, dtype=np.int64))
    t_heads = t_vote = t_ref_total = 0.0
    for branch in ("dino", "shot"):
        sd = state_dicts.get(branch)
        if sd is None or (branch == "dino" and desc is None):
            continue
        th = time.perf_counter()
        with torch.no_grad():
            ref = Ref(branch, sd, num_more=int(cfg.get("num_more", 3)))
            if branch == "dino":
                cls, scales = ref.forward_dino(tpc, torch.from_numpy(desc), tidx)
            else:
                cls, scales = ref.forward_shot(tpc, tidx, torch.from_numpy(shot_feat), torch.from_numpy(normal))
            if bins is not None and branch in bins:
                b = np.asarray(bins[branch], dtype=np.uint8)
            else:
                torch.manual_seed(seed)
                prob = torch.softmax(cls, -1)                           # eval.py:227-229
                b = torch.multinomial(prob.reshape(-1, prob.shape[-1]), 1).reshape(-1, 6).numpy().astype(np.uint8)
        tv = time.perf_counter()
        body = oracle.instance_body(pc, np.asarray(idx, dtype=np.int64), b, scales.numpy(), cfg["up"], cfg["right"],
                                    cfg["front"], res, num_rots=num_rots, sym_y_only=sym_y_only,
                                    scale_override=dino_scale if branch == "shot" else None)   # eval.py:308-310
        if branch == "dino":
            dino_scale = body["pred_scale"]
        t_ref0 = time.perf_counter()
        if opt and body["pairs_mask"].any():                            # eval.py:319-355, then the loss with the refined pose
            from .refine_torch import final_loss, refine_pose
            mask = body["pairs_mask"]
            pair_idx = np.asarray(idx, dtype=np.int64)[mask][:, :2]
            scaled = (body["pred_pairs"] * body["pair_scale"][:, None, None]).astype(np.float32)
            body["T_voted"], body["R_voted"] = body["T_est"], body["R_est"]
            body["T_est"], body["R_est"] = refine_pose(pc, pair_idx, scaled[mask], body["T_est"], body["R_est"], sym_y_only)
            body["loss"] = final_loss(pc, pair_idx, body["pred_pairs"][mask], body["T_est"], body["R_est"],
                                      np.float32(np.linalg.norm(body["pred_scale"])), sym_y_only)
        t_refine = time.perf_counter() - t_ref0
        body["bins"] = b
        out[branch] = body
        if body["loss"] < best_loss:
            best, best_loss = branch, body["loss"]
        t_heads += tv - th
        t_vote += time.perf_counter() - tv - t_refine
        t_ref_total += t_refine
    out["best"] = best
    if timings is not None:
        timings["shot"] = timings.get("shot", 0.0) + (t1 - t0)
        timings["heads"] = timings.get("heads", 0.0) + t_heads
        timings["vote"] = timings.get("vote", 0.0) + t_vote
        if opt:
            timings["refine"] = timings.get("refine", 0.0) + t_ref_total
    return out
