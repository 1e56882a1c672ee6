#!/usr/bin/env python
"""BASELINE config 1: the reference's single-frame demo case -- example_data/{rgb,depth,mask}.png, SHOT branch, T = 50 000
tuples x 180 rotations, opt=False (notebook cell 13 / demo.py:126-300) -- end to end on both arms, with timings.

    python tools/example_data.py [--reps 5]

The cloud (depth/10000 m, K of notebook cell 11, 2 mm voxels: 4 251 points, centre grid 118 x 51 x 132) and the tuple
indices (np.random.seed(0), eval.py:207) are read from tests/golden/example_instance.npz, which oracle/make_golden.py minted
by running the reference's own functions in the build container (/root/reference does not exist on the GPU box).  The SHOT
head is the reference architecture with a seeded default initialisation (ckpts/shot ships no weights), same state_dict on
both arms.

  CPU arm : oracle/pipeline_cpu.py -- PCL-semantics SHOT (C++), torch-CPU float32 heads, torch.multinomial, C voting chain.
  GPU arm : PoseEstimator (SHOT-352 kernel, bf16 tcgen05 heads with fused decode, batched vote chain), (a) with its own
            draws, timed through the public call with host buffers; (b) with the CPU arm's draws injected, for the pose
            comparison: translation, kept count and sphere bins must be identical, R within 0.1 deg.
  golden  : the pose the REFERENCE's own functions produce from the golden's draws, against the GPU vote chain on the same
            draws (grid sha256, centre, kept set, R, scale, loss).
Prints one JSON line.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    from cppf2_b200.config import default_category_cfg
    from cppf2_b200.estimator import Instance, PoseEstimator
    from cppf2_b200.heads import BeyondCPPFSHOT
    from cppf2_b200.heads_spec import init_state_dict
    from cppf2_b200.pipeline import PoseVoter, VoteConfig
    g = np.load(os.path.join(ROOT, "tests", "golden", "example_instance.npz"))
    pc, idx, T = np.ascontiguousarray(g["pc"]), g["idx"].astype(np.int64), int(g["num_tuples"])
    cfg = default_category_cfg("custom")            # config/custom.yaml: res 0.002, axes up=y right=x front=z
    sd = init_state_dict("shot", 1234)
    out = {"config": "BASELINE configs[0]: example_data single frame, SHOT branch, T=50000 x R=180, opt=False",
           "points": int(pc.shape[0]), "tuples": T, "grid_shape": [int(v) for v in g["grid_shape"]]}

    # ---- GPU arm ---------------------------------------------------------------------------------------------------
    model = BeyondCPPFSHOT(cfg, precision=1)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    est = PoseEstimator({"custom": {"shot": model}}, {"custom": cfg}, num_pairs=T, max_points=pc.shape[0], seed=0)
    inst = Instance(pc=pc, category="custom", point_idxs=idx.astype(np.int32))
    for _ in range(3):
        pose = est.estimate([inst])[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.reps):
        pose = est.estimate([inst])[0]            # host cloud + indices in, pose record out
    gpu_s = (time.perf_counter() - t0) / args.reps
    r = pose.results["shot"]
    out["gpu"] = {"seconds_per_frame": gpu_s, "tuples_per_sec": T / gpu_s, "t": r.t.tolist(), "kept": r.kept, "status": r.status,
                  "launches": est.launches, "api": "PoseEstimator.estimate([Instance(host cloud, host indices)])"}

    # ---- the reference's own pose on the golden's draws vs the GPU vote chain on the same draws ------------------------
    voter = PoseVoter(T, pc.shape[0])
    rg = voter.vote(pc, idx, VoteConfig(res=cfg["res"]), pred_scales=g["pred_scales"].astype(np.float32), bins=g["bins"]).result()
    grid = np.ascontiguousarray(voter.intermediates()["grid"].astype(np.int64))
    sha = np.frombuffer(hashlib.sha256(grid.tobytes()).digest(), dtype=np.uint8)
    ang = float(np.degrees(np.arccos(np.clip((np.trace(rg.R.T @ g["R_est"]) - 1) / 2, -1, 1))))
    out["vs_reference_golden"] = {"grid_sha256_equal": bool(np.array_equal(sha, g["grid_sha256"])),
                                  "t_equal": bool(np.array_equal(rg.t, g["T_est"])),
                                  "kept_equal": bool(rg.kept == int(np.unpackbits(g["pairs_mask"])[:T].sum())),
                                  "R_deg": ang, "scale_equal": bool(np.array_equal(rg.scale, g["pred_scale"])),
                                  "loss_rel_diff": float(abs(rg.loss - float(g["loss_all"])) / float(g["loss_all"]))}

    # ---- CPU arm, then the GPU arm on the CPU arm's draws ---------------------------------------------------------------
    if not args.no_cpu:
        from oracle import cpu as oracle
        from oracle.pipeline_cpu import instance_pose_cpu
        torch.set_num_threads(os.cpu_count() or 1)
        oracle.set_num_threads(os.cpu_count() or 1)
        timings = {}
        t0 = time.perf_counter()
        c = instance_pose_cpu(pc, idx, cfg, {"shot": sd}, seed=0, timings=timings)
        cpu_s = time.perf_counter() - t0
        o = c["shot"]
        out["cpu"] = {"seconds_per_frame": cpu_s, "tuples_per_sec": T / cpu_s, "cores": os.cpu_count(), "kind": "port",
                      "stage_seconds": {k: round(v, 3) for k, v in timings.items()}, "t": o["T_est"].tolist(),
                      "kept": int(o["pairs_mask"].sum())}
        gi = est.estimate([inst], draws=[{"shot": o["bins"]}])[0].results["shot"]
        ang = float(np.degrees(np.arccos(np.clip((np.trace(gi.R.T @ o["R_est"]) - 1) / 2, -1, 1))))
        out["gpu_on_cpu_draws"] = {"t_equal": bool(np.array_equal(gi.t, o["T_est"])), "kept_equal": bool(gi.kept == int(o["pairs_mask"].sum())),
                                   "bins_equal": bool(gi.bin_up == o["bin_up"] and gi.bin_right == o["bin_right"]), "R_deg": ang,
                                   "scale_rel_diff": float(np.abs(gi.scale - o["pred_scale"]).max() / np.abs(o["pred_scale"]).max())}
        out["speedup_e2e"] = cpu_s / gpu_s
    print(json.dumps(out))


if __name__ == "__main__":
    main()
