#!/usr/bin/env python
"""BASELINE config 4: vote-aggregation sweep with the tuples of ONE (instance, branch) sharded over the ranks.

    python tools/vote_sweep.py                                                       # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/vote_sweep.py   # tuple-sharded

For T in 2^16 .. 2^24 (bounded by --min-log2 / --max-log2) and two clouds -- the N = 4096 half cylinder (grid ~40x50x20,
privatised in shared memory) and the example_data cloud (grid 118x51x133 = 0.8 M cells, voted through L2) -- every rank
runs the reference's per-branch body eval.py:219-313 on its contiguous block of T/g tuples: SHOT-branch heads (bf16
tcgen05, decode fused in), targets, centre votes, back-vote errors, mask, rotation votes, loss terms; five exchange steps
cross NVLink (cppf2_b200.sharded: grid all-reduce, 4 B/tuple error all-gather, importance + scale histogram, sphere bins,
loss).  One JSON line per (cloud, T) from rank 0: tuples/s over all ranks (CUDA events, max over ranks), the time inside
the collectives, and a parity record against the UNSHARDED chain run by rank 0 on the same tuples and the same draws
(bit-exact grid, kept set, bins).  `--no-heads` sweeps the vote chain alone on injected draws; `--cpu-max-log2 L` adds the
CPU oracle's vote chain (chunked at 2^16 tuples as BASELINE.md section 3 prescribes) up to T = 2^L.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

_G = 0x9E3779B97F4A7C15
_MASK = 0xFFFFFFFFFFFFFFFF


def load_cloud(name: str):
    from cppf2_b200 import synth
    if name == "halfcyl":
        return synth.half_cylinder_cloud(4096, seed=1)
    if name == "example":          # the reference's example_data cloud (4.2 k points, 2 mm voxels), as minted into the golden
        z = np.load(os.path.join(ROOT, "tests", "golden", "vote_center_example.npz"))
        return np.ascontiguousarray(z["pc"], dtype=np.float32)
    raise ValueError(name)


def device_tuples(n: int, first: int, count: int, seed: int, dev) -> torch.Tensor:
    """Rows [first, first+count) of the T x 5 index matrix cppf_sample_tuples(n, T, 5, seed) would draw (counter-based
    generator: a shift of the flat counter by 5*first is a shift of the seed)."""
    from cppf2_b200 import _lib
    idx = torch.empty((count, 5), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().cppf_sample_tuples(n, count, 5, (seed + _G * 5 * first) & _MASK, idx.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream), "cppf_sample_tuples")
    return idx


def injected_draws(pc, idx_host, seed=11):
    """Noisy draws around the true canonical coordinates, so that a vote peak exists (vote-only sweep)."""
    rng = np.random.default_rng(seed)
    ctr = 0.5 * (pc.max(0) + pc.min(0))
    diag = float(np.linalg.norm(pc.max(0) - pc.min(0)))
    canon = (pc[idx_host[:, :2]].astype(np.float32) - ctr.astype(np.float32)) / np.float32(diag)
    bins = np.clip(np.rint((canon + 0.5) * 31) + rng.integers(-1, 2, canon.shape), 0, 31).reshape(-1, 6).astype(np.uint8)
    scales = (np.array([0.57, 0.71, 0.41], np.float32) + 0.02 * rng.standard_normal((idx_host.shape[0], 3)).astype(np.float32))
    return bins, scales


def run_sharded(T: int, cloud: str, dev, rank: int, world: int, reps: int = 5, warm: int = 2, with_heads: bool = True,
                check: bool = True, seed: int = 11, res: float = 0.002):
    """One point of the sweep; returns the JSON record on rank 0, None elsewhere."""
    from cppf2_b200 import shot
    from cppf2_b200.heads import BeyondCPPFSHOT
    from cppf2_b200.pipeline import PoseVoter, VoteConfig
    from cppf2_b200.sharded import ShardedPoseVoter, shard_bounds
    pc = load_cloud(cloud)
    n = pc.shape[0]
    cfg = VoteConfig(res=res)
    cells = PoseVoter.grid_cells_on_host(pc, res)
    lo, hi = shard_bounds(T, world, rank)
    pc_d = torch.from_numpy(pc).to(dev)
    idx_l = device_tuples(n, lo, hi - lo, seed, dev)
    model = desc = normals = bins_l = scales_l = None
    if with_heads:
        model = BeyondCPPFSHOT.load_from_checkpoint("/nonexistent/a/b/last.ckpt", cfg=dict(num_more=3), precision=1, seed=4321)
        desc, normals = shot.compute_device(pc_d, res * 10, res * 10)              # eval.py:210, outside the branch loop
    else:
        b, s = injected_draws(pc, idx_l.cpu().numpy(), seed + rank)
        bins_l, scales_l = torch.from_numpy(b).to(dev), torch.from_numpy(s).to(dev)
    sv = ShardedPoseVoter(hi - lo, n, grid_capacity=max(1 << 22, 32 * cells), device=dev)

    def one(lazy=False):
        if with_heads:
            return sv.vote_with_heads(model, pc_d, idx_l, cfg, first_tuple=lo, seed=seed, shot_feat=desc, normal=normals,
                                      cells_hint=cells, lazy=lazy)
        return sv.vote(pc_d, idx_l, cfg, scales_l, bins_l, cells_hint=cells, lazy=lazy)

    res_s = None
    for _ in range(warm):
        res_s = one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    sv.timing = []
    sv.stage_marks = []
    ev[0].record()
    # streamed votes: each returns its un-synchronised handle, so the next vote is queued while this one runs; the poses are
    # read back after the timed region (the last one is the one checked)
    handles = [one(lazy=True) for _ in range(reps)]
    ev[1].record()
    torch.cuda.synchronize()
    res_s = [h.result() for h in handles][-1]
    ms = torch.tensor([ev[0].elapsed_time(ev[1]) / reps], device=dev, dtype=torch.float64)
    # per-stage breakdown of this rank (mean over the reps): interval from each mark to the next
    stages = {}
    marks = sv.stage_marks
    sv.stage_marks = None
    for (la, ea), (lb, eb) in zip(marks[:-1], marks[1:]):
        name = "heads" if la == "heads begin" else (lb if lb != "heads begin" else "finish+next")
        if lb == "heads begin" or name == "finish+next":
            continue
        stages[name] = stages.get(name, 0.0) + ea.elapsed_time(eb) / reps
    coll = {}
    for label, a, b in sv.timing:
        coll[label] = coll.get(label, 0.0) + a.elapsed_time(b) / reps
    sv.timing = None
    coll_ms = torch.tensor([sum(coll.values())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(coll_ms, op=dist.ReduceOp.MAX)
    mid = sv.stages.intermediates()
    n_coll = sv.n_collectives
    parity = None
    if check:
        mask_all = sv.gather_mask()                      # collective: every rank takes part
        if rank == 0:
            # the unsharded chain on the same tuples and the same draws (same seed => same uniforms per global tuple)
            idx_all = device_tuples(n, 0, T, seed, dev)
            v = PoseVoter(T, n, grid_capacity=max(1 << 22, 32 * cells), device=dev)
            if with_heads:
                bins_a, scales_a = model.forward_sampled(pc_d, idx_all, desc, normals, seed=seed)
            else:
                bs, ss = [], []
                for r in range(world):
                    l2, h2 = shard_bounds(T, world, r)
                    b, s = injected_draws(pc, idx_all[l2:h2].cpu().numpy(), seed + r)
                    bs.append(b)
                    ss.append(s)
                bins_a, scales_a = torch.from_numpy(np.concatenate(bs)).to(dev), torch.from_numpy(np.concatenate(ss)).to(dev)
            single = v.vote(pc_d, idx_all, cfg, pred_scales=scales_a, bins=bins_a, cells_hint=cells).result()
            smid = v.intermediates()
            parity = dict(grid_bit_exact=bool(np.array_equal(mid["grid"], smid["grid"])),
                          kept_set_equal=bool(np.array_equal(mask_all, smid["pairs_mask"])),
                          imp_equal=bool(np.array_equal(mid["imp"], smid["imp"][:n])),
                          centre_equal=bool(np.array_equal(res_s.t, single.t)),
                          bins_equal=bool(res_s.bin_up == single.bin_up and res_s.bin_right == single.bin_right),
                          sphere_counts_max_abs_diff=float(max(np.abs(mid["counts_up"] - smid["counts_up"]).max(),
                                                               np.abs(mid["counts_right"] - smid["counts_right"]).max())),
                          scale_equal=bool(np.array_equal(res_s.scale, single.scale)),
                          R_max_abs_diff=float(np.abs(res_s.R - single.R).max()),
                          loss_rel_diff=float(abs(res_s.loss - single.loss) / max(abs(single.loss), 1e-30)),
                          grid_checksum=int(mid["grid"].astype(np.uint64).sum()))
            del v
    out = None
    if rank == 0:
        t = float(ms.item())
        out = {"config": "vote sweep (BASELINE configs[3]): one (instance, branch), tuples sharded over the ranks",
               "cloud": cloud, "points": int(n), "grid_cells": int(cells), "log2_T": int(np.log2(T)), "tuples": int(T),
               "n_gpus": world, "heads": "SHOT branch, bf16 tcgen05, decode fused" if with_heads else None,
               "ms": t, "tuples_per_sec": T / (t * 1e-3), "collective_ms": float(coll_ms.item()),
               "collective_share": float(coll_ms.item()) / t, "collectives": {k: round(v, 4) for k, v in coll.items()},
               "stages_ms_rank0": {k: round(v, 4) for k, v in stages.items()},
               "n_collectives": n_coll, "kept": int(res_s.kept), "parity": parity}
    del sv
    torch.cuda.empty_cache()
    return out


def cpu_vote_chain(T: int, cloud: str, seed: int = 11, res: float = 0.002):
    """The CPU oracle's vote chain on injected draws, centre votes chunked at 2^16 tuples with the integer grids summed
    (exact: the corners depend on the cloud only) -- the reference materialises 6.5 KB per tuple."""
    from oracle import cpu as oracle
    pc = load_cloud(cloud)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, pc.shape[0], (T, 5)).astype(np.int64)
    bins, _ = injected_draws(pc, idx, seed)
    t0 = time.perf_counter()
    grid = None
    for c0 in range(0, T, 1 << 16):
        sl = slice(c0, min(T, c0 + (1 << 16)))
        _, scaled, _ = oracle.decode_pairs(pc, idx[sl], bins[sl], 32)
        tr, _ = oracle.generate_target_pairs(scaled, [0, 1, 0], [0, 0, 1], [1, 0, 0])
        g, _ = oracle.vote_center(pc, tr, res, idx[sl, :2], 180)
        grid = g if grid is None else grid + g
    dt = time.perf_counter() - t0
    return {"impl": "reference", "config": "vote sweep (BASELINE configs[3]), CPU oracle: decode + targets + centre votes, 2^16-tuple chunks",
            "cloud": cloud, "log2_T": int(np.log2(T)), "tuples": T, "seconds": dt, "tuples_per_sec": T / dt,
            "cpu_threads": os.cpu_count(), "grid_checksum": int(grid.sum())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=16)
    ap.add_argument("--max-log2", type=int, default=22)
    ap.add_argument("--step", type=int, default=2)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--clouds", default="halfcyl,example")
    ap.add_argument("--no-heads", action="store_true")
    ap.add_argument("--check-max-log2", type=int, default=22, help="largest T for which rank 0 also runs the unsharded chain")
    ap.add_argument("--cpu-max-log2", type=int, default=0, help="also time the CPU oracle's vote chain up to this T (rank 0)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    for cloud in args.clouds.split(","):
        for lg in range(args.min_log2, args.max_log2 + 1, args.step):
            rec = run_sharded(1 << lg, cloud, dev, rank, world, reps=args.reps, with_heads=not args.no_heads,
                              check=lg <= args.check_max_log2)
            if rank == 0:
                print(json.dumps(rec), flush=True)
            if rank == 0 and lg <= args.cpu_max_log2:
                from oracle import cpu as oracle
                oracle.set_num_threads(os.cpu_count() or 1)
                print(json.dumps(cpu_vote_chain(1 << lg, cloud)), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
