"""The random-init ViT stand-in (cppf2_b200/backbone.py; reference dataset.py:61-80).  CPU: token geometry, seeding, the
position-table resampling.  GPU: the whole `forward(rgb, pts)` against the reference's post-processing restated with
torch.nn.functional.grid_sample (dataset.py:40-59), and a frame driven from RGB through `estimate_frame`."""
import numpy as np
import pytest
import torch

from cppf2_b200.backbone import DINOV2StandIn


def tiny(**kw):
    return DINOV2StandIn(stride=4, width=64, depth=2, heads=4, table=5, **kw).eval()


def test_patch_tokens_shape_and_seeding():
    rgb = torch.rand(3, 48, 64, generator=torch.Generator().manual_seed(1))
    a, b, c = tiny(seed=3), tiny(seed=3), tiny(seed=4)
    with torch.no_grad():
        ta, tb, tc = a.patch_tokens(rgb), b.patch_tokens(rgb), c.patch_tokens(rgb)
    assert ta.shape == (1, 64, 12, 16)                       # [1, width, H // stride, W // stride], dataset.py:69,77
    assert torch.equal(ta, tb) and not torch.allclose(ta, tc)
    assert torch.isfinite(ta).all()
    # the final LayerNorm makes every token zero-mean / unit-variance before its affine part (identity at init)
    tok = ta[0].permute(1, 2, 0).reshape(-1, 64)
    assert torch.allclose(tok.mean(1), torch.zeros(tok.shape[0]), atol=1e-5)
    assert not ta.is_contiguous()                            # a permuted view, like the reference's


def test_position_table_is_resampled_to_the_patch_grid():
    net = tiny()
    assert net._positions(5, 5).shape == (1, 26, 64)
    assert torch.equal(net._positions(5, 5), net.pos_embed)
    assert net._positions(12, 16).shape == (1, 1 + 12 * 16, 64)


def test_full_size_configuration_has_the_vit_l14_shape():
    net = DINOV2StandIn(depth=1)                             # one of the 24 blocks is enough to count
    per_block = sum(p.numel() for p in net.blocks[0].parameters())
    assert per_block == 4 * 1024 * 1024 + 2 * 4 * 1024 * 1024 + 4 * 1024 + 4096 + 1024 + 6 * 1024      # attention + MLP + norms + LayerScale
    assert net.pos_embed.shape == (1, 1 + 37 * 37, 1024) and net.patch_embed.weight.shape == (1024, 3, 14, 14)


@pytest.mark.gpu
def test_forward_equals_reference_post_processing():
    import torch.nn.functional as F
    net = DINOV2StandIn(stride=4, width=128, depth=2, heads=4, table=7, seed=5).cuda().eval()
    g = torch.Generator().manual_seed(2)
    rgb = torch.rand(3, 96, 128, generator=g).cuda()
    pts = torch.stack([torch.rand(500, generator=g) * 127, torch.rand(500, generator=g) * 95], -1).cuda()
    out = net(rgb, pts)
    assert out.shape == (500, 128) and out.is_cuda
    with torch.no_grad():
        raw = net.patch_tokens(rgb)
        h, w = raw.shape[-2:]
        kp = pts.clone()[None]                               # dataset.py:44-58
        kp[..., 0] = ((kp[..., 0] + 0.5) / w / 4) * 2 - 1
        kp[..., 1] = ((kp[..., 1] + 0.5) / h / 4) * 2 - 1
        ref = F.normalize(F.grid_sample(raw, kp.unsqueeze(-3), align_corners=False).squeeze(-2), dim=1)[0].T
    np.testing.assert_allclose(out.cpu().numpy(), ref.cpu().numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(np.linalg.norm(out.cpu().numpy(), axis=1), 1.0, atol=1e-5)


@pytest.mark.gpu
def test_frame_from_rgb_through_the_stand_in():
    from cppf2_b200 import synth
    from cppf2_b200.estimator import PoseEstimator, build_models
    frame = synth.synth_real275_frame(2, 3)
    cats = list(frame["cats"])
    models, cfgs = build_models(sorted(set(cats)), precision=1)
    est = PoseEstimator(models, cfgs, num_pairs=8192, seed=1)
    net = DINOV2StandIn(stride=8, depth=1, seed=1).cuda().eval()        # width 1024 as the heads expect; one block keeps it quick
    h, w = frame["depth"].shape
    rgb = torch.rand(3, h, w, generator=torch.Generator().manual_seed(0)).cuda()
    poses = est.estimate_frame(frame["depth"].astype(np.uint16), list(frame["masks"]), cats, synth.REAL275_K, desc_fn=net.desc_fn(rgb))
    assert len(poses) == len(cats) and any(p is not None for p in poses)
    for p in poses:
        if p is not None:
            assert np.isfinite(p.RT).all() and p.branch in ("dino", "shot")
