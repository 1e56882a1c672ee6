"""CUDA SHOT (through the C ABI) against the CPU oracle (PCL 1.9.1 semantics).  The reference pins
nothing for this path, so the gates are the tolerances of BASELINE.md section 4: normals <= 0.5 deg,
descriptors max-abs <= 1e-4 outside hard-boundary / ill-conditioned-frame cases (fraction reported)."""
import numpy as np
import pytest

from tests import gates
import torch

pytestmark = pytest.mark.gpu

from cppf2_b200 import synth  # noqa: E402


def angle_deg(a, b):
    c = np.clip(np.sum(a.astype(np.float64) * b, -1), -1, 1)
    return np.degrees(np.arccos(c))


def run_case(oracle, pc, r, fast, inject_normals=False):
    from cppf2_b200 import shot
    o_desc, o_normals = oracle.shot_compute(pc, r, r)
    o_desc, o_normals = o_desc.reshape(-1, 352), o_normals.reshape(-1, 3)
    desc, normals, rf = shot.compute_device(torch.from_numpy(pc).cuda(), r, r, fast_math=fast, want_rf=True,
                                            normals_in=o_normals if inject_normals else None)
    return desc.cpu().numpy(), normals.cpu().numpy(), rf.cpu().numpy(), o_desc, o_normals


def clouds(name):
    if name == "halfcyl":
        return synth.half_cylinder_cloud(4000, seed=7, jitter=0.001), 0.02
    return synth.torus_cloud(20000, res=0.002, seed=7)[0], 0.02


@pytest.mark.parametrize("cloud", ["halfcyl", "torus"])
def test_normals_match_oracle(oracle, cloud):
    pc, r = clouds(cloud)
    desc, normals, rf, o_desc, o_normals = run_case(oracle, pc, r, False)
    assert np.array_equal(np.isnan(normals), np.isnan(o_normals))        # neighbour SETS are bit-identical
    ok = ~np.isnan(o_normals).any(1)
    n_gpu, n_cpu = normals[ok], o_normals[ok].astype(np.float64)
    # the viewpoint flip is a sign test on (-p).n: where the normal is perpendicular to the view ray
    # (silhouette points) its sign is decided by rounding, so those points are compared up to sign
    view = np.abs(np.sum(n_cpu * pc[ok], -1)) / np.linalg.norm(pc[ok], axis=1)
    grazing = view < 2e-3
    err = angle_deg(n_gpu, n_cpu)
    err = np.where(grazing, np.minimum(err, 180.0 - err), err)
    print(f"GATE normals {cloud}: error vs oracle (deg): median {np.median(err):.4f} p99 {np.percentile(err, 99):.4f} "
          f"p99.9 {np.percentile(err, 99.9):.4f} max {err.max():.4f}; {int(grazing.sum())} grazing points compared up to sign")
    # float32 single-pass un-centred covariance: accumulation-order noise (PCL's own is ~0.05 deg).
    # 0.5 deg holds on the smooth surface; the thin torus (tube radius ~ support radius) has nearly equal
    # eigenvalues, which amplifies that noise -- there the bound is on the 99th percentile.
    if cloud == "halfcyl":
        assert err.max() < gates.NORMALS_MAX_DEG
    else:
        assert np.percentile(err, 99) < gates.NORMALS_TORUS_P99_DEG and err.max() < gates.NORMALS_TORUS_MAX_DEG
    assert np.all(np.sum(normals[ok] * (-pc[ok]), -1) >= -1e-6)            # flipped towards the origin
    np.testing.assert_allclose(np.linalg.norm(normals[ok], axis=1), 1.0, atol=1e-5)


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("cloud", ["halfcyl", "torus"])
def test_descriptor_stage_matches_oracle(oracle, cloud, fast):
    """LRF + histogram with the oracle's normals injected: everything downstream of the normals agrees to
    1e-4 on every row whose frame is defined.  Rows with an exact sign-vote tie are resolved by kd-tree
    search order in PCL (shot_lrf.hpp) -- either sign is acceptable there, and they are only counted."""
    pc, r = clouds(cloud)
    desc, normals, rf, o_desc, o_normals = run_case(oracle, pc, r, fast, inject_normals=True)
    assert np.array_equal(np.isnan(desc).any(1), np.isnan(o_desc).any(1))
    o_rf, margins = oracle.shot_lrf(pc, r)
    good = ~np.isnan(o_desc).any(1)
    tie = (margins == 0).any(1) & good
    defined = good & ~tie
    # frames: identical axes (to float rounding) wherever the votes are not tied
    rf_diff = np.abs(rf[defined] - o_rf[defined]).max(1)
    diff = np.abs(desc[defined] - o_desc[defined]).max(1)
    bad = diff > 1e-4
    print(f"{cloud} fast={fast}: {int(tie.sum())} tie rows ({tie.mean():.3%}); defined rows {int(defined.sum())}: "
          f"frame max diff {rf_diff.max():.2e}, desc median {np.median(diff):.2e}, p99.9 {np.percentile(diff, 99.9):.2e}, "
          f"rows > 1e-4: {int(bad.sum())}")
    assert tie.mean() < 0.05
    assert np.percentile(rf_diff, 99.9) < 1e-5
    assert bad.mean() < 1e-3 and np.median(diff) < 1e-6
    # tie rows: the descriptor equals the oracle's for one of the sign choices or not -- just require validity
    np.testing.assert_allclose(np.linalg.norm(desc[good], axis=1), 1.0, atol=1e-5)


def test_dense_patch_beyond_the_neighbour_list(oracle):
    """A patch sampled at radius/25 has ~1 900 neighbours per key-point, beyond the 768-entry shared-memory list of the
    scatter-matrix and descriptor kernels (shot.cu, kShotListCap): both fall back to sweeping the cells again.  Same gate
    as the descriptor-stage test, with the oracle's normals injected."""
    rng = np.random.default_rng(5)
    n, r = 6000, 0.02
    u = rng.uniform(-0.03, 0.03, (n, 2))
    z = 0.9 + 0.15 * u[:, 0] ** 2 * 30 - 0.1 * u[:, 1] ** 2 * 30 + rng.uniform(-2e-4, 2e-4, n)      # gently curved sheet
    pc = np.stack([u[:, 0], u[:, 1], z], -1).astype(np.float32)
    desc, normals, rf, o_desc, o_normals = run_case(oracle, pc, r, True, inject_normals=True)
    d2 = ((pc[:200, None, :] - pc[None, :, :]) ** 2).sum(-1)
    assert ((d2 < r * r).sum(1) > 768).mean() > 0.9          # the premise: most key-points overflow the list
    assert np.array_equal(np.isnan(desc).any(1), np.isnan(o_desc).any(1))
    o_rf, margins = oracle.shot_lrf(pc, r)
    defined = ~np.isnan(o_desc).any(1) & ~(margins == 0).any(1)
    assert defined.mean() > 0.9
    diff = np.abs(desc[defined] - o_desc[defined]).max(1)
    print(f"dense patch: defined rows {int(defined.sum())}, desc median {np.median(diff):.2e}, max {diff.max():.2e}")
    assert np.percentile(np.abs(rf[defined] - o_rf[defined]).max(1), 99.9) < 1e-5
    assert (diff > 1e-4).mean() < 1e-3 and np.median(diff) < 1e-6


@pytest.mark.parametrize("cloud", ["halfcyl", "torus"])
def test_end_to_end_descriptor_noise_is_reported(oracle, cloud):
    """With each side's own normals, the float32 covariance noise of the normals (<= 0.5 deg, present in PCL
    itself through its kd-tree accumulation order) moves the cosine bin of every neighbour.  Reported and
    loosely bounded; the tight gates are the two tests above."""
    pc, r = clouds(cloud)
    desc, normals, rf, o_desc, o_normals = run_case(oracle, pc, r, True)
    assert np.array_equal(np.isnan(desc).any(1), np.isnan(o_desc).any(1))
    _, margins = oracle.shot_lrf(pc, r)
    good = ~np.isnan(o_desc).any(1) & ~(margins == 0).any(1)
    diff = np.abs(desc[good] - o_desc[good]).max(1)
    cos = np.sum(desc[good] * o_desc[good], 1)
    print(f"{cloud}: end-to-end max-abs diff median {np.median(diff):.2e}, p99 {np.percentile(diff, 99):.2e}; "
          f"cosine similarity median {np.median(cos):.6f}, p1 {np.percentile(cos, 1):.6f}")
    assert np.median(cos) > 0.999 and np.percentile(cos, 1) > 0.97


def test_shot_api_layout_and_invalid_rows(oracle):
    from cppf2_b200 import shot
    rng = np.random.default_rng(0)
    cluster = rng.uniform(-0.005, 0.005, (40, 3)).astype(np.float32) + np.float32([0, 0, 1])
    lonely = np.float32([[0.5, 0, 1], [0.0, 0.5, 1], [-0.5, 0, 1]])
    pc = np.concatenate([cluster, lonely])
    out = shot.compute(pc, 0.02, 0.02)
    assert isinstance(out, list) and len(out) == 2
    desc, normal = out
    assert desc.dtype == np.float32 and desc.shape == (43 * 352,) and normal.shape == (43 * 3,)
    desc, normal = desc.reshape(-1, 352), normal.reshape(-1, 3)
    assert np.all(np.isnan(normal[-3:])) and np.all(np.isnan(desc[-3:]))
    assert not np.isnan(normal[:40]).any() and not np.isnan(desc[:40]).any()
    o_desc, o_normal = oracle.shot_compute(pc, 0.02, 0.02)
    assert np.array_equal(np.isnan(desc), np.isnan(o_desc.reshape(-1, 352)))
    n_only = shot.estimate_normal(pc, 0.02)
    assert np.array_equal(n_only, normal.reshape(-1), equal_nan=True)      # deterministic: same bits on every call
    again = shot.compute(pc, 0.02, 0.02)
    assert np.array_equal(again[0], out[0], equal_nan=True) and np.array_equal(again[1], out[1], equal_nan=True)
    with pytest.raises(NotImplementedError):
        shot.compute_color(pc, pc, 0.02, 0.02)
    # different radii for normals and descriptor (shot.cpp's defaults are 0.1 / 0.17)
    d2, n2 = shot.compute(pc, 0.01, 0.02)
    od2, on2 = oracle.shot_compute(pc, 0.01, 0.02)
    assert np.array_equal(np.isnan(n2), np.isnan(on2))
    # empty cloud
    d0, n0 = shot.compute(np.zeros((0, 3), np.float32), 0.02, 0.02)
    assert d0.shape == (0,) and n0.shape == (0,)
