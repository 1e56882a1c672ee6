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
    """World size 1: the sharded stage sequence (block-local mask, two-pass histogram scale median, loss from the scratch
    sum) is the single-GPU chain, and the oracle, bit for bit where the path is integer."""
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
    assert np.array_equal(mid["pairs_mask_local"], o["pairs_mask"]) and np.array_equal(sv.gather_mask(), o["pairs_mask"])
    assert np.array_equal(res.t, o["T_est"])
    assert np.array_equal(mid["imp"], o["imp"]) and res.kept == int(o["pairs_mask"].sum()) and res.status == 0
    assert res.bin_up == o["bin_up"] and res.bin_right == o["bin_right"]
    np.testing.assert_allclose(res.R, o["R_est"], atol=1e-6)
    np.testing.assert_array_equal(res.scale, o["pred_scale"])          # radix-histogram median == torch.median, exactly
    np.testing.assert_allclose(res.loss, o["loss"], rtol=1e-5)
    # the SHOT branch's reuse of the DINO scale (eval.py:308-310) skips the histogram passes
    res2 = sv.vote(torch.from_numpy(pc).to(dev), torch.from_numpy(idx).to(dev), VoteConfig(res=0.002), None,
                   torch.from_numpy(bins).to(dev), scale_override=np.float32([0.5, 0.6, 0.7]))
    np.testing.assert_array_equal(res2.scale, np.float32([0.5, 0.6, 0.7]))
    assert np.array_equal(res2.t, res.t) and res2.kept == res.kept


def test_histogram_scale_median_equals_sort_median():
    """cppf_scale_median_hist / _pick (two 16-bit radix passes per axis) against the sorted lower median, for even and odd
    counts, negative values and ties."""
    import torch
    from cppf2_b200 import _lib
    from cppf2_b200.voting import stream_ptr, struct_tensor
    lib = _lib.load()
    rng = np.random.default_rng(8)
    for m in (1, 2, 7, 1000, 5001):
        T = 3 * m + 5
        scales = rng.standard_normal((T, 3)).astype(np.float32)
        scales[::3, 1] = 0.25                                            # ties
        kept = np.sort(rng.choice(T, m, replace=False)).astype(np.int32)
        d_s, d_k = torch.from_numpy(scales).cuda(), torch.from_numpy(kept).cuda()
        cnt64 = torch.tensor([m], dtype=torch.int64, device="cuda")
        cnt32 = torch.tensor([m], dtype=torch.int32, device="cuda")
        sel = struct_tensor(_lib.ScaleSelect, d_s.device)
        hist = torch.empty(3 * 65536, dtype=torch.int32, device="cuda")
        out = torch.zeros(3, dtype=torch.float32, device="cuda")
        for p in (0, 1):
            _lib.check(lib.cppf_scale_median_hist(d_s.data_ptr(), d_k.data_ptr(), cnt64.data_ptr(), m, p, sel.data_ptr(),
                                                  hist.data_ptr(), stream_ptr()))
            assert int(hist.sum().item()) <= 3 * m
            _lib.check(lib.cppf_scale_median_pick(hist.data_ptr(), cnt32.data_ptr(), p, sel.data_ptr(), out.data_ptr(), stream_ptr()))
        want = np.sort(scales[kept], axis=0)[(m - 1) // 2]
        np.testing.assert_array_equal(out.cpu().numpy(), want)


def test_rotation_parts_sum_to_whole():
    """cppf_rotation_hist_part over parts 0..g-1 (a partition by TUPLE ID, so it does not depend on the order the atomic
    compaction left kept_list in) adds up to the unsharded histogram."""
    import ctypes as C
    import torch
    from cppf2_b200 import _lib
    from cppf2_b200.pipeline import VoteConfig
    from cppf2_b200.sharded import CudaStages
    from cppf2_b200.voting import angle_tables, cos_threshold, sphere_lut, sphere_points, stream_ptr
    pc, idx, bins, scales = _inputs(T=4096)
    cfg = VoteConfig(res=0.002)
    st = CudaStages(idx.shape[0], pc.shape[0])
    dev = st.device
    st.begin(pc, torch.from_numpy(idx).to(dev), torch.from_numpy(bins).to(dev), torch.from_numpy(scales).to(dev), cfg, 1)
    st.vote_center()
    st.argmax()
    st.select(st.errors())
    st.mask_local(True)
    st.after_mask(True)
    out = st.rotation_counts()              # the sphere bins (and, packed next to them, the scale histogram of exchange E4)
    whole = (out.parts[0] if hasattr(out, "parts") else out[0]).clone()
    v, lib = st.v, st.lib
    S = cfg.num_sphere
    thr = cos_threshold(cfg.angle_tol)
    ct, sn = angle_tables(cfg.num_rots, dev)
    lut, lut_g = sphere_lut(S, thr, dev)
    ip, i64, istr = st.ip

    def part(p, g, kept_list):
        counts = torch.zeros((2, S), dtype=torch.float64, device=dev)
        _lib.check(lib.cppf_rotation_hist_part(st.pc.data_ptr(), ip, i64, istr, v.targets_rot.data_ptr(), 3, (C.c_int * 2)(0, 2), 2,
                                               kept_list.data_ptr(), st._kept_ptr(), st.T_local, st._x.data_ptr(), v.summary.data_ptr(),
                                               float(cfg.imp_wt_margin), ct.data_ptr(), sn.data_ptr(), int(cfg.num_rots),
                                               sphere_points(S, dev).data_ptr(), S, thr, lib.cppf_sphere_band(S, thr),
                                               None if lut is None else lut.data_ptr(), lut_g, counts.data_ptr(), p, g, stream_ptr()))
        return counts
    assert whole.sum().item() > 0
    torch.testing.assert_close(sum(part(p, 3, v.kept_list) for p in range(3)), whole, rtol=0, atol=0)
    # a differently ordered kept list (what another rank's compaction would produce) gives every part the same bins
    kept = int(st._x[st.N].item())
    shuffled = v.kept_list.clone()
    shuffled[:kept] = v.kept_list[:kept][torch.randperm(kept, device=dev)]
    for p in range(3):
        torch.testing.assert_close(part(p, 3, shuffled), part(p, 3, v.kept_list), rtol=0, atol=0)
