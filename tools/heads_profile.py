"""Cycle breakdown of the tcgen05 heads kernel per warp role (debug counters written by chain_tc_kernel).

    python tools/heads_profile.py [T] [N]

Prints, per branch, the CUDA-event time of the forward and the mean over CTAs of: MMA-issuer total / wait for the
slot's operand / wait for weights; producer wait for a free ring stage; epilogue warp total / wait for MMAs /
time per action.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppf2_b200 import _lib, synth  # noqa: E402
from cppf2_b200.heads import BeyondCPPFDINO, BeyondCPPFSHOT  # noqa: E402

ACTIONS = ["LoadRows", "EncShotA", "EncShotB", "Gather", "CoordsB", "HiddenT", "HiddenS", "Out", "Final", "OutT"]


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2700
    lib = _lib.load()
    fn = lib.cppf_debug_heads_tc_profile
    fn.restype = None
    fn.argtypes = [C.c_void_p]
    dev = torch.device("cuda", 0)
    pc = torch.from_numpy(synth.half_cylinder_cloud(N, seed=1)).to(dev)
    n = pc.shape[0]
    idx = torch.from_numpy(np.random.default_rng(0).integers(0, n, (T, 5), dtype=np.int32)).to(dev)
    shot_feat = torch.rand((n, 352), device=dev)
    normal = torch.nn.functional.normalize(torch.randn((n, 3), device=dev), dim=-1)
    desc = torch.nn.functional.normalize(torch.randn((n, 1024), device=dev), dim=-1)
    prof = torch.zeros((148, 64), dtype=torch.int64, device=dev)
    for name, model, args in (("shot", BeyondCPPFSHOT(precision=1), (pc, idx, shot_feat, normal)),
                              ("dino", BeyondCPPFDINO(precision=1), (pc, desc, idx))):
        for _ in range(3):
            model(*args)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            model(*args)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        prof.zero_()
        fn(prof.data_ptr())
        model(*args)
        torch.cuda.synchronize()
        fn(None)
        p = prof.cpu().numpy().astype(np.float64)
        p = p[p[:, 8] > 0]
        m = p.mean(0)
        print(f"== {name}: T={T} N={n}  forward {ms * 1e3:.1f} us (both programs);  counters: mean over {p.shape[0]} CTAs, kcycles")
        for sl in range(2):
            o = m[40 + 8 * sl: 48 + 8 * sl]
            st = max(o[4], 1)
            print(f"  mma issuer s{sl}: total {o[0] / 1e3:8.1f}  wait_operand {o[1] / 1e3:8.1f}  wait_weights {o[2] / 1e3:8.1f}  issue {o[3] / 1e3:8.1f}  steps {o[4]:.0f} -> issue {o[3] / st:.0f} cycles/step")
        print(f"  producer   : total {m[4] / 1e3:8.1f}  wait_free_stage {m[5] / 1e3:8.1f}")
        for s in range(2):
            o = m[8 + 16 * s: 8 + 16 * s + 14]
            acts = "  ".join(f"{ACTIONS[k]} {o[2 + k] / 1e3:.1f}" for k in range(10) if o[2 + k] > 0)
            if o[13] > 0:
                acts += f"  GatherSum {o[13] / 1e3:.1f}"
            print(f"  epilogue s{s}: total {o[0] / 1e3:8.1f}  wait_mma {o[1] / 1e3:8.1f}  arrive {o[12] / 1e3:6.1f}  | {acts}")


if __name__ == "__main__":
    main()
