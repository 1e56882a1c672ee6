"""Debug driver of the whole-frame call: small frames first, one stage at a time (CPPF_FRAME_SYNC=1 names a failing stage),
compared against the per-instance path."""
import os
import sys

os.environ.setdefault("CPPF_FRAME_SYNC", "1")
import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppf2_b200 import synth
from cppf2_b200.estimator import Instance, PoseEstimator, build_models


def run(cats, T, n0, with_desc=True, idx_given=True):
    models, cfgs = build_models(sorted(set(cats)), precision=1, seed=5)
    est = PoseEstimator(models, cfgs, num_pairs=T, max_points=8000, seed=11)
    insts = []
    for i, cat in enumerate(cats):
        pc = synth.half_cylinder_cloud(n0 + 300 * i, seed=60 + i, jitter=0.0005)
        insts.append(Instance(pc=pc, category=cat, desc=synth.unit_descriptors(pc.shape[0], 1024, seed=70 + i) if with_desc else None,
                              point_idxs=synth.sample_tuples(pc.shape[0], T, 5, seed=80 + i) if idx_given else None))
    a = est.estimate(insts)
    torch.cuda.synchronize()
    print(f"frame ok: cats={cats} T={T} launches={est.launches}", [(x.branch, {b: (r.kept, r.status) for b, r in x.results.items()}) for x in a], flush=True)
    if idx_given:
        est.frame_call = False
        b = est.estimate(insts)
        for x, y in zip(a, b):
            for br in x.results:
                r, o = x.results[br], y.results[br]
                same = np.array_equal(r.t, o.t) and r.kept == o.kept and r.bin_up == o.bin_up and r.bin_right == o.bin_right and np.array_equal(r.scale, o.scale)
                print("   ", br, "same as per-instance path:", same, r.t, o.t, r.kept, o.kept, r.scale, o.scale, flush=True)


if __name__ == "__main__":
    run(["mug"], 4096, 1500)
    run(["mug"], 4096 + 128, 1500)          # odd tile count
    run(["mug", "laptop"], 20000, 2200)
    run(["mug", "laptop", "bowl"], 20000, 2200, with_desc=False)
    run(["can", "camera"], 16384, 1800, idx_given=False)
    run(["bottle", "bowl", "camera", "can", "laptop", "mug"], 50000, 2500)
    print("all ok")
