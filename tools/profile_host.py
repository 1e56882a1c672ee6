"""Host-side (Python/ctypes) cost of the public call: cProfile over PoseEstimator.estimate on the bench frame.

    python tools/profile_host.py [steps]
"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from cppf2_b200 import _lib  # noqa: E402
from cppf2_b200.estimator import Instance, PoseEstimator, build_models  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    torch.cuda.set_device(0)
    raw = bench.build_frame(0)
    cats = sorted({i["category"] for i in raw})
    models, cfgs = build_models(cats, precision=1)
    est = PoseEstimator(models, cfgs, num_pairs=bench.NUM_PAIRS, num_rots=bench.NUM_ROTS)
    insts = []
    for inst in raw:
        it = Instance(pc=torch.from_numpy(inst["pc"]).pin_memory(), category=inst["category"],
                      desc=torch.from_numpy(inst["desc"]).pin_memory(), point_idxs=None)
        it.cells_hint = est.voter.grid_cells_on_host(inst["pc"], inst["cfg"]["res"])
        insts.append(it)
    for _ in range(3):
        est.estimate(insts)
    torch.cuda.synchronize()
    # enqueue-only time with an empty queue (synchronise, enqueue one frame, read the clock before the GPU finishes):
    # what the host must get through per frame
    host = []
    for _ in range(steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pose_buf = torch.zeros((len(insts) * 2, est.pose_bytes), dtype=torch.uint8, device=est.device)
        est.enqueue(insts, pose_buf, staged=est.stage(insts))
        host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print(f"host enqueue (stage + enqueue, empty queue): median {1e3 * sorted(host)[len(host) // 2]:.2f} ms/frame, launches {est.launches}")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(steps):
        est.estimate(insts)
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(22)


if __name__ == "__main__":
    main()
