#!/usr/bin/env python
"""BASELINE config 4: vote-aggregation sweep, tuples sharded over the ranks with NCCL grid all-reduce.

    python tools/vote_sweep.py                                   # 1 GPU
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/vote_sweep.py   # tuple-sharded

For T in 2^16 .. 2^24 (bounded by --max-log2): a half-cylinder cloud (N = 4096, grid ~40x50x20) and noisy
draws around the true canonical coordinates, so that a vote peak exists.  Every rank decodes, votes and
back-votes its contiguous block of tuples (cppf2_b200.sharded); the centre grid, the kept-pair data and the
sphere bins are exchanged with NCCL.  Prints one JSON line per T (rank 0): tuples/s over all ranks (CUDA
events, max over ranks), and a parity record: the all-reduced grid's checksum and arg-max against the
unsharded single-GPU chain computed by rank 0 on the same inputs (bit-exact, integer votes).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_inputs(T: int, n: int = 4096, seed: int = 11):
    from cppf2_b200 import synth
    pc = synth.half_cylinder_cloud(n, seed=1)
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, pc.shape[0], (T, 5), dtype=np.int32)
    canon = (pc[idx[:, :2]].astype(np.float32) - np.array([0.0, 0.0, 0.8], np.float32)) / np.float32(0.14)
    bins = np.clip(np.rint((canon + 0.5) * 31) + rng.integers(-1, 2, canon.shape), 0, 31).reshape(T, 6).astype(np.uint8)
    scales = (np.array([0.57, 0.71, 0.41], np.float32) + 0.02 * rng.standard_normal((T, 3)).astype(np.float32))
    return pc, idx, bins, scales


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--min-log2", type=int, default=16)
    ap.add_argument("--max-log2", type=int, default=22)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check-max-log2", type=int, default=20, help="largest T for which rank 0 also runs the unsharded chain")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from cppf2_b200.pipeline import PoseVoter, VoteConfig
    from cppf2_b200.sharded import ShardedPoseVoter, shard_bounds
    cfg = VoteConfig(res=0.002)
    for lg in range(args.min_log2, args.max_log2 + 1, 2):
        T = 1 << lg
        pc, idx, bins, scales = make_inputs(T)
        lo, hi = shard_bounds(T, world, rank)
        pc_d = torch.from_numpy(pc).to(dev)
        idx_l, bins_l, sc_l = (torch.from_numpy(a[lo:hi]).to(dev) for a in (idx, bins, scales))
        sv = ShardedPoseVoter(T, pc.shape[0], device=dev)
        res = None
        for _ in range(2):
            res = sv.vote(pc_d, idx_l, cfg, sc_l, bins_l)
        mid = sv.stages.intermediates()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(args.reps):
            res = sv.vote(pc_d, idx_l, cfg, sc_l, bins_l)
        ev[1].record()
        torch.cuda.synchronize()
        ms = torch.tensor([ev[0].elapsed_time(ev[1]) / args.reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        parity = None
        if rank == 0 and lg <= args.check_max_log2:
            v = PoseVoter(T, pc.shape[0], device=dev)
            single = v.vote(pc_d, torch.from_numpy(idx).to(dev), cfg, pred_scales=torch.from_numpy(scales).to(dev),
                            bins=torch.from_numpy(bins).to(dev)).result()
            smid = v.intermediates()
            parity = dict(grid_bit_exact=bool(np.array_equal(mid["grid"], smid["grid"])),
                          kept_set_equal=bool(np.array_equal(mid["pairs_mask"], smid["pairs_mask"])),
                          centre_equal=bool(np.array_equal(res.t, single.t)),
                          bins_equal=bool(res.bin_up == single.bin_up and res.bin_right == single.bin_right),
                          R_max_abs_diff=float(np.abs(res.R - single.R).max()),
                          grid_checksum=int(mid["grid"].astype(np.uint64).sum()))
            del v
        if rank == 0:
            print(json.dumps({"config": "vote sweep (BASELINE configs[3])", "log2_T": lg, "tuples": T, "n_gpus": world,
                              "ms_per_vote": float(ms.item()), "tuples_per_sec": T / (float(ms.item()) * 1e-3),
                              "grid_cells": int(mid["grid"].size), "kept": int(res.kept), "parity_vs_unsharded": parity}), flush=True)
        del sv
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
