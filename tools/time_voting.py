"""Kernel-level timing of the voting path (CUDA events on the launching stream).  Scratch tool for
optimisation work; bench.py is the contract benchmark."""
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppf2_b200 import _lib, synth  # noqa: E402
from cppf2_b200.voting import angle_tables, idx_args, struct_tensor, stream_ptr  # noqa: E402
from cppf2_b200.pipeline import PoseVoter, VoteConfig  # noqa: E402


def time_fn(fn, iters=20, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    lib = _lib.load()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = []
    for n, extent_name in ((4096, "halfcyl"),):
        pc_h = synth.half_cylinder_cloud(n, seed=3)
        for T in (50000, 1 << 20):
            idx_h = synth.sample_tuples(n, T, 2, seed=11)
            tr_h = synth.noisy_center_targets(pc_h, idx_h, np.array([0, 0, 0.78]), seed=5)
            pc, idx, tr = torch.from_numpy(pc_h).to(dev), torch.from_numpy(idx_h).to(dev), torch.from_numpy(tr_h).to(dev)
            ct, st = angle_tables(180)
            geom = struct_tensor(_lib.GridGeom, dev)
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            grid = torch.empty(1 << 23, dtype=torch.int32, device=dev)
            s = stream_ptr()
            _lib.check(lib.cppf_cloud_bounds(pc.data_ptr(), n, 0.002, geom.data_ptr(), s))
            ip, i64, istr = idx_args(idx)

            import ctypes as C
            ex = lib.cppf_vote_center_ex
            ref = None
            for reps in (1, 2, 4, 8, 16, 32, 64, 0):
                def vote():
                    _lib.check(ex(pc.data_ptr(), n, ip, i64, istr, tr.data_ptr(), T, ct.data_ptr(), st.data_ptr(), 180,
                                  geom.data_ptr(), grid.data_ptr(), grid.numel(), 0, status.data_ptr(), int(reps == 0),
                                  max(reps, 1), 40 * 50 * 20, s))
                med, mn = time_fn(vote, flush=flush)
                g = grid[:40 * 50 * 20].clone()
                if ref is None:
                    ref = g
                same = bool(torch.equal(g, ref))
                out.append(dict(kernel="vote_center", variant=("smem" if reps == 0 else f"reps{reps}"), T=T, ms_med=round(med, 4),
                                ms_min=round(mn, 4), Mtuples_per_s=round(T / med / 1e3, 1), same_grid=same,
                                status=int(status.item())))
                print(json.dumps(out[-1]), flush=True)
    # whole chain at the reference's default size
    n, T = 4096, 50000
    pc_h = synth.half_cylinder_cloud(n, seed=3)
    idx_h = synth.sample_tuples(n, T, 5, seed=11)
    rng = np.random.default_rng(0)
    canon = (pc_h[idx_h[:, :2]].astype(np.float64) - np.array([0, 0, 0.8])) / 0.14
    bins = np.clip(np.rint((canon + 0.5) * 31) + rng.integers(-1, 2, canon.shape), 0, 31).reshape(T, 6).astype(np.uint8)
    scales = rng.standard_normal((T, 3)).astype(np.float32)
    voter = PoseVoter(T, n)
    pc, idx, b, sc = (torch.from_numpy(x).to(dev) for x in (pc_h, idx_h, bins, scales))
    cfg = VoteConfig()
    med, mn = time_fn(lambda: voter.vote(pc, idx, cfg, pred_scales=sc, bins=b), flush=flush)
    print(json.dumps(dict(kernel="pose_chain", T=T, ms_med=med, ms_min=mn, launches=voter.launches, result=str(voter.result().t))))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def time_shot():
    from cppf2_b200 import shot
    for n in (10000, 50000, 200000):
        pc = torch.from_numpy(synth.torus_cloud(n, 0.002)[0]).cuda()
        for fast in (False, True):
            med, mn = time_fn(lambda: shot.compute_device(pc, 0.02, 0.02, fast_math=fast), iters=5, warmup=2)
            print(json.dumps(dict(kernel="shot", n=n, fast=fast, ms_med=round(med, 3), pts_per_s=round(n / med * 1e3),
                                  alg_GBps=round(n * 1432 / med / 1e6, 1))), flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "shot":
    time_shot()
