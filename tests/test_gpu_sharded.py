"""Tuple-sharded voter on the GPU: the sharded stage sequence (world size 1, or NCCL when launched under torchrun by
tools/vote_sweep.py) must reproduce the oracle bit-exactly where the path is integer."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _inputs(T=8192, n=1500):
    from cppf2_b200 import synth
    pc = synth.half_cylinder_cloud(n, seed=1)
    idx = synth.sample_tuples(pc.shape[0], T, 5, seed=2)
    rng = np.random.default_rng(3)
    canon = (pc[idx[:, :2]].astype(np.float64) - np.array([0.0, 0.0, 0.8])) / 0.14
    bins = np.clip(np.rint((canon + 0.5) * 31) + rng.integers(-1, 2, canon.shape), 0, 31).reshape(T, 6).astype(np.uint8)
    scales = (np.array([0.57, 0.71, 0.41]) + 0.02 * rng.standard_normal((T, 3))).astype(np.float32)
    return pc, idx, bins, scales


def test_sharded_stage_sequence_matches_oracle(oracle):
    import torch
    from cppf2_b200.pipeline import VoteConfig
    from cppf2_b200.sharded import ShardedPoseVoter
    pc, idx, bins, scales = _inputs()
    sv = ShardedPoseVoter(idx.shape[0], pc.shape[0])
    dev = sv.stages.device
    res = sv.vote(torch.from_numpy(pc).to(dev), torch.from_numpy(idx).to(dev), VoteConfig(res=0.002),
                  torch.from_numpy(scales).to(dev), torch.from_numpy(bins).to(dev))
    mid = sv.stages.intermediates()
    o = oracle.instance_body(pc, idx, bins, scales, [0, 1, 0], [1, 0, 0], [0, 0, 1], 0.002)
    assert np.array_equal(mid["grid"], o["grid"])
    assert np.array_equal(mid["pairs_mask"], o["pairs_mask"])
    assert np.array_equal(res.t, o["T_est"])
    assert np.array_equal(mid["imp"][:pc.shape[0]], o["imp"])
    assert res.bin_up == o["bin_up"] and res.bin_right == o["bin_right"]
    np.testing.assert_allclose(res.R, o["R_est"], atol=1e-6)
    np.testing.assert_array_equal(res.scale, o["pred_scale"])
    np.testing.assert_allclose(res.loss, o["loss"], rtol=1e-5)


def test_rotation_parts_sum_to_whole():
    """cppf_rotation_hist_part over parts 0..g-1 adds up to the unsharded histogram (float64 bins, tolerance)."""
    import torch
    from cppf2_b200.pipeline import VoteConfig
    from cppf2_b200.sharded import CudaStages
    pc, idx, bins, scales = _inputs(T=4096)
    cfg = VoteConfig(res=0.002)
    st = CudaStages(idx.shape[0], pc.shape[0])
    dev = st.device
    pc_d, idx_d, bins_d = torch.from_numpy(pc).to(dev), torch.from_numpy(idx).to(dev), torch.from_numpy(bins).to(dev)
    tr, rot = st.decode_targets(pc_d, idx_d, bins_d, cfg)
    grid = st.vote_center(pc_d, idx_d, tr, cfg)
    st.argmax(grid, cfg)
    errs = st.errors(pc_d, idx_d, tr)
    st.select_and_mask(errs, idx_d, pc_d, cfg)
    whole = st.rotation_counts(pc_d, idx_d, rot, cfg, 0, 1).clone()
    parts = sum(st.rotation_counts(pc_d, idx_d, rot, cfg, p, 3).clone() for p in range(3))
    assert whole.sum().item() > 0
    torch.testing.assert_close(parts, whole, rtol=1e-9, atol=1e-9)
