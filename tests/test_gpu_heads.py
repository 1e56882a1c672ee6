"""CUDA heads (through the C ABI) against the reference golden outputs and the torch float32 reference.
float32 path: rtol 1e-4 / atol 2e-5 on logits and scale.  bf16 tensor-core path: against the bf16-emulated
float32-accumulate reference, atol 1e-3 x logit range."""
import numpy as np
import pytest

from tests import gates
import torch

pytestmark = pytest.mark.gpu

from cppf2_b200 import synth  # noqa: E402
from cppf2_b200.heads_spec import init_state_dict  # noqa: E402
from oracle.heads_torch import Ref  # noqa: E402


def make_inputs(n, t, seed):
    rng = np.random.default_rng(seed)
    pc = synth.half_cylinder_cloud(n, seed=seed)
    idx = synth.sample_tuples(n, t, 5, seed=seed + 1)
    shot = np.abs(rng.standard_normal((n, 352))).astype(np.float32)
    shot /= np.linalg.norm(shot, axis=-1, keepdims=True)
    shot[::17] = 0.0                                   # NaN rows scrubbed to zero by the caller (eval.py:215)
    normal = rng.standard_normal((n, 3)).astype(np.float32)
    normal /= np.linalg.norm(normal, axis=-1, keepdims=True)
    desc = synth.unit_descriptors(n, 1024, seed=seed + 2)
    return pc, idx, shot, normal, desc


def test_fp32_heads_match_reference_golden(golden):
    from cppf2_b200.heads import BeyondCPPFDINO, BeyondCPPFSHOT
    g = golden("heads")
    idx = g["idx"].astype(np.int64)
    m = BeyondCPPFSHOT(dict(num_more=3)).cuda().eval()
    m.load_state_dict(init_state_dict("shot", int(g["seed_shot"])))
    cls, scale = m(torch.from_numpy(g["pc"]).cuda(), torch.from_numpy(idx).cuda(),
                   torch.from_numpy(g["shot"].astype(np.float32)).cuda(), torch.from_numpy(g["normal"]).cuda())
    assert cls.shape == (96, 6, 32) and scale.shape == (96, 3) and cls.is_cuda
    np.testing.assert_allclose(cls.cpu().numpy(), g["cls_shot"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(scale.cpu().numpy(), g["scale_shot"], rtol=1e-4, atol=2e-5)
    d = BeyondCPPFDINO(dict(num_more=3)).cuda().eval()
    d.load_state_dict(init_state_dict("dino", int(g["seed_dino"])))
    desc = synth.unit_descriptors(g["pc"].shape[0], 1024, seed=int(g["desc_seed"]))
    cls, scale = d(g["pc"], desc, idx)                 # host inputs are accepted as well
    np.testing.assert_allclose(cls.cpu().numpy(), g["cls_dino"], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(scale.cpu().numpy(), g["scale_dino"], rtol=1e-4, atol=2e-5)


@pytest.mark.parametrize("n,t", [(1000, 5000), (37, 1), (4096, 50000)])
def test_fp32_heads_match_torch(n, t):
    from cppf2_b200.heads import BeyondCPPFDINO, BeyondCPPFSHOT
    pc, idx, shot, normal, desc = make_inputs(n, t, seed=n)
    for branch, cls_ in (("shot", BeyondCPPFSHOT), ("dino", BeyondCPPFDINO)):
        sd = init_state_dict(branch, 99)
        m = cls_(dict(num_more=3)).cuda()
        m.load_state_dict(sd)
        ref = Ref(branch, sd, device="cuda")
        tpc, tidx = torch.from_numpy(pc).cuda(), torch.from_numpy(idx).cuda()
        with torch.no_grad():
            if branch == "shot":
                got = m(tpc, tidx, torch.from_numpy(shot).cuda(), torch.from_numpy(normal).cuda())
                want = ref.forward_shot(tpc, tidx, torch.from_numpy(shot).cuda(), torch.from_numpy(normal).cuda())
            else:
                got = m(tpc, torch.from_numpy(desc).cuda(), tidx)
                want = ref.forward_dino(tpc, torch.from_numpy(desc).cuda(), tidx)
        for a, b in zip(got, want):
            np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-4, atol=5e-5)


def test_checkpoint_surface(tmp_path):
    from cppf2_b200.heads import BeyondCPPFSHOT
    sd = {k: torch.from_numpy(v) for k, v in init_state_dict("shot", 5).items()}
    path = tmp_path / "ckpts" / "shot" / "bottle-num_more-3" / "lightning_logs" / "version_0" / "checkpoints" / "last.ckpt"
    path.parent.mkdir(parents=True)
    torch.save({"state_dict": sd, "epoch": 100}, path)
    m = BeyondCPPFSHOT.load_from_checkpoint(path, cfg=dict(num_more=3)).cuda().eval()
    assert m.checkpoint == str(path)
    for k, v in m.state_dict().items():
        assert torch.equal(v, sd[k])
    missing = BeyondCPPFSHOT.load_from_checkpoint(tmp_path / "nope" / "a" / "b" / "last.ckpt", cfg=dict(num_more=3))
    assert missing.checkpoint is None
    with pytest.raises(ValueError):
        bad = dict(sd)
        bad["tuple_encoder.0.fc1.weight"] = torch.zeros(3, 3)
        m.load_state_dict(bad)


def _tc_available(model):
    from cppf2_b200 import _lib
    model._ensure(torch.device("cuda", torch.cuda.current_device()))
    return bool(_lib.load().cppf_heads_has_tc(model._handle))


@pytest.mark.parametrize("n,t", [(700, 128), (1000, 5000), (4096, 50000), (333, 77)])
def test_bf16_tensor_core_shot_head(n, t):
    """tcgen05 path vs the bf16-emulated reference (operands rounded to bf16, float32 accumulate): the only
    differences are accumulation order inside the tensor core and bf16 roundings that land on a tie."""
    from cppf2_b200.heads import BeyondCPPFSHOT
    pc, idx, shot, normal, desc = make_inputs(n, t, seed=n + 1)
    sd = init_state_dict("shot", 321)
    m = BeyondCPPFSHOT(dict(num_more=3), precision=1).cuda()
    m.load_state_dict(sd)
    if not _tc_available(m):
        pytest.skip("library built without the tcgen05 heads")
    tpc, tidx = torch.from_numpy(pc).cuda(), torch.from_numpy(idx).cuda()
    tshot, tnormal = torch.from_numpy(shot).cuda(), torch.from_numpy(normal).cuda()
    cls, scale = m(tpc, tidx, tshot, tnormal)
    torch.cuda.synchronize()
    with torch.no_grad():
        want_cls, want_scale = Ref("shot", sd, emulate_bf16=True, device="cuda").forward_shot(tpc, tidx, tshot, tnormal)
        f32_cls, f32_scale = Ref("shot", sd, device="cuda").forward_shot(tpc, tidx, tshot, tnormal)
    rng_cls = float(want_cls.max() - want_cls.min())
    rng_scale = float(want_scale.max() - want_scale.min()) + 1e-3
    err_cls = float((cls - want_cls).abs().max())
    err_scale = float((scale - want_scale).abs().max())
    e = ((cls - want_cls).abs() / rng_cls).flatten()
    qs = torch.quantile(e[:: max(1, e.numel() // 4000000)].float(), torch.tensor([0.5, 0.99, 0.999], device=e.device)).tolist()
    print(f"GATE heads SHOT T={t}: logits range {rng_cls:.3f}; |tc - bf16 ref| / range: mean {float(e.mean()):.2e} p50 {qs[0]:.2e} "
          f"p99 {qs[1]:.2e} p99.9 {qs[2]:.2e} max {float(e.max()):.2e}; |tc - fp32 ref| max {float((cls - f32_cls).abs().max()) / rng_cls:.2e}; "
          f"scale max err / range {err_scale / rng_scale:.2e}")
    # the gate (tests/gates.py): mean, 99.9th percentile and maximum of the error in units of the output range
    assert err_cls <= gates.HEADS_MAX * rng_cls + gates.HEADS_ABS_FLOOR and err_scale <= gates.HEADS_MAX * rng_scale + gates.HEADS_ABS_FLOOR
    assert float(e.mean()) <= gates.HEADS_MEAN and qs[2] <= gates.HEADS_P999
    # against float32: reported above, loosely bounded (about 1 % of the logit range with random-init weights)
    assert float((cls - f32_cls).abs().max()) < 0.05 * rng_cls
    agree = (cls.argmax(-1) == f32_cls.argmax(-1)).float().mean().item()
    assert agree > 0.97


@pytest.mark.parametrize("n,t", [(600, 300), (4096, 50000)])
def test_bf16_tensor_core_dino_head(n, t):
    from cppf2_b200.heads import BeyondCPPFDINO
    pc, idx, shot, normal, desc = make_inputs(n, t, seed=n + 2)
    sd = init_state_dict("dino", 654)
    m = BeyondCPPFDINO(dict(num_more=3), precision=1).cuda()
    m.load_state_dict(sd)
    if not _tc_available(m):
        pytest.skip("library built without the tcgen05 DINO head")
    tpc, tidx, tdesc = torch.from_numpy(pc).cuda(), torch.from_numpy(idx).cuda(), torch.from_numpy(desc).cuda()
    cls, scale = m(tpc, tdesc, tidx)
    torch.cuda.synchronize()
    with torch.no_grad():
        # the tensor-core path evaluates desc_pair_transform per point and slot (linearity, kActGatherSum): the bf16-emulated
        # reference with the same rounding points is the tight one; the as-written evaluation order is bounded too
        want_cls, want_scale = Ref("dino", sd, emulate_bf16=True, device="cuda").forward_dino(tpc, tdesc, tidx, hoist_pair=True)
        asw_cls, _ = Ref("dino", sd, emulate_bf16=True, device="cuda").forward_dino(tpc, tdesc, tidx)
        f32_cls, _ = Ref("dino", sd, device="cuda").forward_dino(tpc, tdesc, tidx)
        f32_hoist, _ = Ref("dino", sd, device="cuda").forward_dino(tpc, tdesc, tidx, hoist_pair=True)
    rng_cls = float(want_cls.max() - want_cls.min())
    rng_scale = float(want_scale.max() - want_scale.min()) + 1e-3
    err_cls, err_scale = float((cls - want_cls).abs().max()), float((scale - want_scale).abs().max())
    e = ((cls - want_cls).abs() / rng_cls).flatten()
    qs = torch.quantile(e[:: max(1, e.numel() // 4000000)].float(), torch.tensor([0.5, 0.99, 0.999], device=e.device)).tolist()
    print(f"GATE heads DINO T={t}: logits range {rng_cls:.3f}; |tc - bf16 ref| / range: mean {float(e.mean()):.2e} p50 {qs[0]:.2e} "
          f"p99 {qs[1]:.2e} p99.9 {qs[2]:.2e} max {float(e.max()):.2e}; vs bf16 ref as written: max {float((cls - asw_cls).abs().max()) / rng_cls:.2e} "
          f"mean {float((cls - asw_cls).abs().mean()) / rng_cls:.2e}; |tc - fp32 ref| max {float((cls - f32_cls).abs().max()) / rng_cls:.2e}; "
          f"scale max err / range {err_scale / rng_scale:.2e}")
    assert float((f32_hoist - f32_cls).abs().max()) <= 2e-5 * rng_cls + 1e-6     # the hoisted form is the same linear map
    assert err_cls <= gates.HEADS_MAX * rng_cls + gates.HEADS_ABS_FLOOR and err_scale <= gates.HEADS_MAX * rng_scale + gates.HEADS_ABS_FLOOR
    assert float(e.mean()) <= gates.HEADS_MEAN and qs[2] <= gates.HEADS_P999
    assert float((cls - asw_cls).abs().max()) <= 5e-3 * rng_cls + 1e-4 and float((cls - asw_cls).abs().mean()) <= 1e-3 * rng_cls
    assert float((cls - f32_cls).abs().max()) < 0.05 * rng_cls


@pytest.mark.parametrize("branch", ["shot", "dino"])
@pytest.mark.parametrize("n,t,inject", [(900, 5000, True), (4096, 50000, False), (333, 77, True)])
def test_fused_sampling_epilogue_equals_forward_then_sample_bins(branch, n, t, inject):
    """cppf_heads_forward_sampled (decode fused into the logits epilogue, eval.py:221-229) draws the bins that
    cppf_sample_bins draws from the logits of cppf_heads_forward with the same uniforms: same kernel, same
    accumulators; the running sum is sequential instead of a warp scan, so a draw may differ only when u*total
    sits within float32 rounding of a CDF step.  The scale head output must be identical."""
    from cppf2_b200 import _lib
    from cppf2_b200.heads import BeyondCPPFDINO, BeyondCPPFSHOT
    from cppf2_b200.voting import stream_ptr
    lib = _lib.load()
    pc, idx, shot, normal, desc = make_inputs(n, t, seed=n + 3)
    m = (BeyondCPPFSHOT if branch == "shot" else BeyondCPPFDINO)(dict(num_more=3), precision=1).cuda()
    m.load_state_dict(init_state_dict(branch, 77))
    if not _tc_available(m):
        pytest.skip("library built without the tcgen05 heads")
    tpc, tidx = torch.from_numpy(pc).cuda(), torch.from_numpy(idx).cuda()
    args = (tpc, tidx, torch.from_numpy(shot).cuda(), torch.from_numpy(normal).cuda()) if branch == "shot" else \
           (tpc, torch.from_numpy(desc).cuda(), tidx)
    u01 = torch.rand((t, 6), generator=torch.Generator().manual_seed(5)).cuda() if inject else None
    seed = 4242
    logits, scale = m(*args)
    want = torch.empty((t, 6), dtype=torch.uint8, device="cuda")
    _lib.check(lib.cppf_sample_bins(logits.data_ptr(), t, 32, None if u01 is None else u01.data_ptr(), seed, want.data_ptr(),
                                    stream_ptr()))
    bins, scale2 = m.forward_sampled(*args, u01=u01, seed=seed)
    torch.cuda.synchronize()
    assert bins.shape == (t, 6) and bins.dtype == torch.uint8
    assert torch.equal(scale, scale2)
    mismatch = (bins != want)
    frac = mismatch.float().mean().item()
    assert frac < 1e-4, f"{frac:.2e} of the draws differ"
    if mismatch.any() and inject:      # every differing draw sits on a CDF step
        p = torch.softmax(logits.double(), -1).cumsum(-1)
        gap = (p - u01.double()[..., None]).abs().min(-1)[0]
        assert (gap[mismatch] < 1e-5).all()
    # the draws follow the logits: the mean log-probability of the drawn bins beats a uniform draw
    logp = torch.log_softmax(logits, -1).gather(-1, bins.long()[..., None]).mean().item()
    assert logp > float(np.log(1 / 32)) - 1e-3
