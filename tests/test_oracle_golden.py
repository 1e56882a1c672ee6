"""CPU oracle (oracle/) versus the golden vectors minted from the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; the -m gpu tests then compare the CUDA
path with the oracle."""
import hashlib

import numpy as np
import pytest


@pytest.mark.parametrize("name", ["vote_center_halfcyl", "vote_center_halfcyl_r36"])
def test_vote_center_bit_exact(oracle, golden, name):
    g = golden(name)
    grid, world = oracle.vote_center(g["pc"], g["tr"], float(g["res"]), g["idx"], int(g["num_rots"]),
                                     tables=(g["cos_tab"], g["sin_tab"]))
    assert grid.shape == g["grid"].shape
    assert np.array_equal(grid, g["grid"].astype(np.int64)), f"{(grid != g['grid']).sum()} cells differ"
    assert np.array_equal(world, g["cand_world"])


def test_vote_center_example_cloud_bit_exact(oracle, golden):
    g = golden("vote_center_example")
    grid, world = oracle.vote_center(g["pc"], g["tr"], float(g["res"]), g["idx"], int(g["num_rots"]),
                                     tables=(g["cos_tab"], g["sin_tab"]))
    assert tuple(grid.shape) == tuple(g["grid_shape"])
    assert np.array_equal(grid.sum((1, 2)), g["grid_sum_x"])
    assert np.array_equal(grid.sum((0, 2)), g["grid_sum_y"])
    assert np.array_equal(grid.sum((0, 1)), g["grid_sum_z"])
    assert int(grid.argmax()) == int(g["argmax"]) and int(grid.max()) == int(g["peak"])
    digest = np.frombuffer(hashlib.sha256(np.ascontiguousarray(grid).tobytes()).digest(), np.uint8)
    assert np.array_equal(digest, g["grid_sha256"])
    assert np.array_equal(world, g["cand_world"])


def test_angle_tables_match_minting_host(oracle, golden):
    """The tables are inputs of both implementations; this only reports whether this host's torch
    reproduces the build container's Sleef results (it must for the golden end-to-end cases)."""
    g = golden("vote_center_halfcyl")
    ct, st = oracle.angle_tables(int(g["num_rots"]))
    assert np.array_equal(ct, g["cos_tab"]) and np.array_equal(st, g["sin_tab"])


def test_generate_target_pairs(oracle, golden):
    g = golden("targets")
    # call-site positional order (up, front, right), eval.py:237-240
    tr0, rot0 = oracle.generate_target_pairs(g["pairs"], g["up"], g["front"], g["right"])
    tr1, rot1 = oracle.generate_target_pairs(g["pairs"], g["up"], g["front"], g["right"], g["center"])
    # translation targets involve only +,-,*,/,sqrt -> bit-exact (NaN for the zero-length pair included)
    assert np.array_equal(tr0, g["tr0"], equal_nan=True)
    assert np.array_equal(tr1, g["tr1"], equal_nan=True)
    # arccos: libm here and there -> allow 1 ulp of float32
    for mine, ref in ((rot0, g["rot0"]), (rot1, g["rot1"])):
        np.testing.assert_allclose(mine, ref, rtol=2e-7, atol=0, equal_nan=True)


def test_vote_rotation_and_topk(oracle, golden):
    g = golden("rotation")
    R = int(g["num_rots"])
    tables = (g["cos_tab"], g["sin_tab"])
    up, mask = oracle.vote_rotation(g["pc"], g["theta"], g["idx"], R, tables=tables)
    assert np.array_equal(mask, g["mask"])
    # tan() is evaluated in double here and by Sleef tanf in the reference: candidates agree to ~1 ulp
    np.testing.assert_allclose(up[:16], g["up_head"], rtol=0, atol=3e-7)
    np.testing.assert_allclose(up.astype(np.float64).sum((0, 1)), g["up_sum"], rtol=0, atol=2e-3)
    wt_rows = np.repeat(g["wt"][mask], R)
    counts = oracle.sphere_counts(up.reshape(-1, 3), g["sphere"], float(g["angle_tol"]), wt_rows)
    # a candidate whose dot product sits within 2 ulp of the threshold may flip between the BLAS
    # accumulation order of the reference and the fma chain here: bound the difference by their weight
    thr = oracle.cos_threshold(float(g["angle_tol"]))
    dots = up.reshape(-1, 3).astype(np.float64) @ g["sphere"].astype(np.float64).T
    border = np.abs(dots - float(thr)) < 3e-7
    slack = (border * (1.0 / wt_rows)[:, None]).sum(0)
    diff = np.abs(counts - g["counts"].astype(np.float64))
    assert np.all(diff <= slack + 1e-4 * np.maximum(counts, 1.0)), f"max diff {diff.max()}"
    assert int(np.argmax(counts)) == int(g["best"])
    fused = oracle.rotation_counts(g["pc"], g["idx"], g["theta"], g["wt"], None, R, g["sphere"], float(g["angle_tol"]), tables)
    np.testing.assert_allclose(fused, counts, rtol=1e-12)
    dirs, cnts = oracle.get_topk_dir(up.reshape(-1, 3), g["sphere"], 100000, float(g["angle_tol"]), wt_rows, topk=1)
    assert np.array_equal(dirs[0], g["sphere"][int(g["best"])])


def test_fibonacci_sphere(oracle, golden):
    g = golden("rotation")
    assert np.array_equal(oracle.fibonacci_sphere(720), g["sphere"])


def test_percentile_semantics(oracle, golden):
    """eval.py:257 -- the kept set only depends on the floor(q)-th order statistic."""
    from cppf2_b200.hostmath import percentile_plan, lerp_f32
    g = golden("percentile")
    for c in "abcde":
        e = g["errs_" + c]
        assert str(g["thr_dtype_" + c]) == "float32"
        lo, gamma = percentile_plan(e.shape[0], 0.1)
        s = np.sort(e)
        thr = lerp_f32(s[lo], s[min(lo + 1, e.shape[0] - 1)], gamma)
        assert np.float32(thr) == np.float32(g["thr_" + c]), (c, thr, g["thr_" + c])
        assert int((e < thr).sum()) == int(g["kept_" + c])


def test_instance_body(oracle, golden):
    g = golden("instance")
    T = int(g["num_tuples"])
    out = oracle.instance_body(g["pc"], g["idx"].astype(np.int64), g["bins"], g["pred_scales"].astype(np.float32),
                               [0, 1, 0], [1, 0, 0], [0, 0, 1], 0.002)
    assert np.array_equal(out["pair_scale"][:512], g["pair_scale_head"])
    assert np.array_equal(out["targets_tr"][:512], g["targets_tr_head"])
    np.testing.assert_allclose(out["targets_rot"][:512], g["targets_rot_head"], rtol=2e-7)
    assert np.array_equal(out["grid"], g["grid"].astype(np.int64))
    assert np.array_equal(out["T_est"], g["T_est"])
    assert np.array_equal(out["back_errs"][:512], g["back_errs_head"])
    assert np.float32(out["thr"]) == np.float32(g["thr"])
    assert np.array_equal(out["pairs_mask"], np.unpackbits(g["pairs_mask"])[:T].astype(bool))
    np.testing.assert_array_equal(out["imp_pair_wt"], g["imp_pair_wt"])
    for k in ("counts_up", "counts_right"):
        np.testing.assert_allclose(out[k], g[k].astype(np.float64), rtol=2e-3, atol=150.0)
        assert int(np.argmax(out[k])) == int(np.argmax(g[k]))
    np.testing.assert_allclose(out["R_est"], g["R_est"], atol=1e-7)
    assert np.array_equal(out["pred_scale"], g["pred_scale"])
    np.testing.assert_allclose(out["loss"], float(g["loss_all"]), rtol=1e-6)


def test_oracle_backproject_matches_reference_golden(golden, oracle):
    """numpy restatement of utils/util.py:2586-2607 (+ eval.py:185-189) against the output of the reference's own function."""
    g = golden("backproject")
    mask = np.unpackbits(g["mask"])[:int(np.prod(g["mask_shape"]))].reshape(g["mask_shape"]).astype(bool)
    pts, (rows, cols) = oracle.backproject(g["depth"] / 1000., g["K"], mask)
    assert np.array_equal(pts, g["pc"]) and np.array_equal(rows, g["rows"]) and np.array_equal(cols, g["cols"])


def test_oracle_voxel_downsample_properties(oracle):
    """One representative per occupied voxel of the Open3D grid (origin = min - res/2), the injected draw decides which."""
    rng = np.random.default_rng(0)
    pc = (rng.random((5000, 3)) * 0.05).astype(np.float32)
    prio = rng.random(5000).astype(np.float32)
    idx = oracle.voxel_downsample(pc, 0.004, prio)
    p = pc.astype(np.float64)
    vox = np.floor((p - (p.min(0) - 0.002)) / 0.004).astype(np.int64)
    assert len(np.unique(vox[idx], axis=0)) == len(idx) == len(np.unique(vox, axis=0))
    for i in idx[:50]:     # the representative has the smallest draw of its voxel
        same = np.all(vox == vox[i], axis=1)
        assert prio[i] == prio[same].min()


def _grid_matches_example_golden(grid, g):
    import hashlib
    grid = np.ascontiguousarray(grid.astype(np.int64))
    assert tuple(grid.shape) == tuple(int(v) for v in g["grid_shape"])
    assert np.array_equal(grid.sum((1, 2)), g["grid_sum_x"]) and np.array_equal(grid.sum((0, 2)), g["grid_sum_y"])
    assert np.array_equal(grid.sum((0, 1)), g["grid_sum_z"])
    assert int(np.argmax(grid)) == int(g["argmax"]) and int(grid.max()) == int(g["peak"])
    sha = np.frombuffer(hashlib.sha256(grid.tobytes()).digest(), dtype=np.uint8)
    assert np.array_equal(sha, g["grid_sha256"]), "centre grid differs from the reference's grid on example_data"


def test_example_data_instance_body(oracle, golden):
    """BASELINE config 1: the SHOT-branch body of notebook cell 13 on the reference's example_data cloud (0.8 M-cell grid,
    T = 50 000), minted by the reference's own functions (oracle/make_golden.py::mint_example_instance)."""
    g = golden("example_instance")
    T = int(g["num_tuples"])
    out = oracle.instance_body(g["pc"], g["idx"].astype(np.int64), g["bins"], g["pred_scales"].astype(np.float32),
                               [0, 1, 0], [1, 0, 0], [0, 0, 1], 0.002)
    assert np.array_equal(out["targets_tr"][:512], g["targets_tr_head"])
    _grid_matches_example_golden(out["grid"], g)
    assert np.array_equal(out["T_est"], g["T_est"])
    assert np.float32(out["thr"]) == np.float32(g["thr"])
    assert np.array_equal(out["pairs_mask"], np.unpackbits(g["pairs_mask"])[:T].astype(bool))
    np.testing.assert_array_equal(out["imp_pair_wt"], g["imp_pair_wt"])
    for k in ("counts_up", "counts_right"):
        assert int(np.argmax(out[k])) == int(np.argmax(g[k]))
    np.testing.assert_allclose(out["R_est"], g["R_est"], atol=1e-7)
    assert np.array_equal(out["pred_scale"], g["pred_scale"])
    np.testing.assert_allclose(out["loss"], float(g["loss_all"]), rtol=1e-6)
