"""Small pass over every kernel family for compute-sanitizer (memcheck is 10-50x slower than a plain run):

    compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_smoke.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cppf2_b200 import shot, synth  # noqa: E402
from cppf2_b200.estimator import Instance, PoseEstimator, build_models  # noqa: E402
from cppf2_b200.pipeline import PoseVoter, VoteConfig  # noqa: E402


def main():
    torch.cuda.set_device(0)
    pc = synth.half_cylinder_cloud(600, seed=1)
    T = 700                                                  # not a multiple of 128: ragged last tile
    idx = synth.sample_tuples(pc.shape[0], T, 5, seed=2)
    desc, normals = shot.compute(pc, 0.02, 0.02)             # grid build, normals, LRF + neighbour list + integer histogram
    rng = np.random.default_rng(3)
    canon = (pc[idx[:, :2]].astype(np.float64) - np.array([0.0, 0.0, 0.8])) / 0.14
    bins = np.clip(np.rint((canon + 0.5) * 31) + rng.integers(-1, 2, canon.shape), 0, 31).reshape(T, 6).astype(np.uint8)
    scales = (np.array([0.57, 0.71, 0.41]) + 0.02 * rng.standard_normal((T, 3))).astype(np.float32)
    voter = PoseVoter(T, pc.shape[0])
    for opt in (False, True):                                # vote chain, with and without the refinement kernel
        res = voter.vote(pc, idx, VoteConfig(res=0.002, opt=opt), pred_scales=scales, bins=bins).result()
        assert np.isfinite(res.R).all()
    models, cfgs = build_models(["mug"], precision=1)        # both tensor-core programs (per point + per tuple), fused decode
    est = PoseEstimator(models, cfgs, num_pairs=T, max_points=pc.shape[0], opt=True)
    out = est.estimate([Instance(pc=pc, category="mug", desc=synth.unit_descriptors(pc.shape[0], 1024, seed=4), point_idxs=idx)])
    assert out[0] is not None and np.isfinite(out[0].RT).all()
    torch.cuda.synchronize()
    print("sanitize_smoke ok", out[0].branch, res.kept)


if __name__ == "__main__":
    main()
