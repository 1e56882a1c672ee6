"""Instance cloud preparation through the C ABI: back-projection against the golden minted from the
reference's own backproject (utils/util.py:2586-2607), voxel down-sampling against the numpy restatement of Open3D's
binning with injected draws (parity unpinned: Open3D is not installable here), and the frame-level helper."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from cppf2_b200 import synth  # noqa: E402


def _mask_of(g):
    return np.unpackbits(g["mask"])[:int(np.prod(g["mask_shape"]))].reshape(g["mask_shape"]).astype(bool)


def test_backproject_matches_reference_golden(golden):
    from cppf2_b200 import cloud
    g = golden("backproject")
    mask = _mask_of(g)
    # drop-in signature on metric float depth (what eval.py:185 passes)
    pts, (rows, cols) = cloud.backproject(g["depth"] / 1000., g["K"], mask)
    assert np.array_equal(rows, g["rows"]) and np.array_equal(cols, g["cols"])
    got = pts.copy()
    got[:, 0] = -got[:, 0]
    got[:, 1] = -got[:, 1]
    got = got.astype(np.float32)
    # float32 depth input differs from the reference's float64 depth/1000 by < 1 float32 ulp
    np.testing.assert_allclose(got, g["pc"], rtol=3e-7, atol=0)
    # device form on the raw uint16 millimetres with the division inside the kernel: float64 like the reference
    pc, pix, count = cloud.backproject_device(torch.from_numpy(g["depth"].astype(np.uint16)).cuda(), g["K"],
                                              torch.from_numpy(mask).cuda(), 1000.0)
    n = int(count.item())
    assert n == g["pc"].shape[0]
    assert np.array_equal(pix[:n].cpu().numpy(), g["rows"].astype(np.int64) * mask.shape[1] + g["cols"])
    dev = pc[:n].cpu().numpy()
    exact = (dev == g["pc"]).mean()
    assert exact > 0.999 and np.abs(dev - g["pc"]).max() <= 2e-7      # a BLAS fma in the reference's 3x3 product may move 1 ulp


def test_backproject_edge_cases():
    from cppf2_b200 import cloud
    depth = torch.zeros((48, 64), dtype=torch.float32, device="cuda")
    mask = torch.ones((48, 64), dtype=torch.bool, device="cuda")
    pc, pix, count = cloud.backproject_device(depth, synth.REAL275_K, mask)
    assert int(count.item()) == 0                                       # no valid depth
    depth[5, 7] = 1.5
    depth[47, 63] = 0.75
    mask[47, 63] = False
    pc, pix, count = cloud.backproject_device(depth, synth.REAL275_K, mask)
    assert int(count.item()) == 1 and int(pix[0]) == 5 * 64 + 7
    assert abs(float(pc[0, 2]) - 1.5) < 1e-6


@pytest.mark.parametrize("n,res", [(20000, 0.002), (3000, 0.01), (1, 0.002), (200000, 0.002)])
def test_voxel_downsample_matches_restatement(oracle, n, res):
    from cppf2_b200 import cloud
    rng = np.random.default_rng(n)
    pc = synth.half_cylinder_cloud(n, seed=n % 97) if n > 1 else np.float32([[0.1, 0.2, 0.8]])
    pc = (pc + rng.normal(0, 0.0003, pc.shape)).astype(np.float32)
    prio = rng.random(pc.shape[0]).astype(np.float32)
    want = oracle.voxel_downsample(pc, res, prio)
    out, kept, count, side = cloud.voxel_downsample_device(torch.from_numpy(pc).cuda(), res, prio=torch.from_numpy(prio).cuda(),
                                                           side=torch.arange(pc.shape[0], dtype=torch.int32, device="cuda") * 3)
    m = int(count.item())
    got = kept[:m].cpu().numpy()
    assert np.array_equal(got, want)
    assert np.array_equal(out[:m].cpu().numpy(), pc[want])
    assert np.array_equal(side[:m].cpu().numpy(), want * 3)
    # without injected draws: same voxel partition (one point per voxel, every voxel represented), different members
    idx = cloud.downsample(pc, res, seed=5)
    p64 = pc.astype(np.float64)
    vox = np.floor((p64 - (p64.min(0) - 0.5 * res)) / res).astype(np.int64)
    assert len(np.unique(vox[idx], axis=0)) == len(idx) == len(np.unique(vox, axis=0))
    assert np.array_equal(cloud.downsample(pc, res, seed=5), idx)


def test_prepare_instance_clouds_frame():
    from cppf2_b200 import cloud
    from cppf2_b200.config import default_category_cfg
    frame = synth.synth_real275_frame(1, 4)
    depth = torch.from_numpy(frame["depth"].astype(np.float32)).cuda()
    masks = [torch.from_numpy(m).cuda() for m in frame["masks"]]
    res = [default_category_cfg(c)["res"] for c in frame["cats"]]
    out = cloud.prepare_instance_clouds(depth, masks, synth.REAL275_K, res, depth_div=1000.0, seed=3)
    assert len(out) == len(masks)
    for i, item in enumerate(out):
        n_pix = int((frame["masks"][i] & (frame["depth"] > 0)).sum())
        if n_pix < 50:
            assert item is None
            continue
        pc, pix = item
        assert 0 < pc.shape[0] <= min(n_pix, 50000) and pix.shape[0] == pc.shape[0]
        # every kept point is the back-projection of its pixel
        rows, cols = (pix // 640).cpu().numpy(), (pix % 640).cpu().numpy()
        z = frame["depth"][rows, cols] / 1000.0
        np.testing.assert_allclose(pc[:, 2].cpu().numpy(), z, rtol=1e-6)
        assert frame["masks"][i][rows, cols].all()


def test_interpolate_features_matches_reference_golden(golden):
    """dataset.py:40-59 (bilinear grid_sample, align_corners=False, zero padding, then F.normalize) against the output of
    the reference's own function; the token map is passed as the permuted view of the ViT layout, like the reference does."""
    from cppf2_b200 import cloud
    g = golden("interp_features")
    h, w, C = int(g["h"]), int(g["w"]), g["tokens"].shape[1]
    tokens = torch.from_numpy(g["tokens"]).cuda()
    raw = tokens.reshape(1, h, w, C).permute(0, 3, 1, 2)           # [1,C,h,w], channel stride 1
    pts = torch.from_numpy(g["pts"]).cuda()
    got = cloud.interpolate_features(raw, pts, strides=float(g["stride"]), normalize=True)
    assert got.shape == (1, C, pts.shape[1])
    np.testing.assert_allclose(got[0].T.cpu().numpy(), g["out"], rtol=1e-5, atol=2e-6)
    got_raw = cloud.interpolate_features(raw.contiguous(), pts, strides=float(g["stride"]), normalize=False)   # channel-major copy
    np.testing.assert_allclose(got_raw[0].T.cpu().numpy(), g["out_raw"], rtol=1e-5, atol=2e-6)
    assert np.all(got_raw[0].T.cpu().numpy()[3] == 0)               # key-point below the image: zero padding
