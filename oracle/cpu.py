"""ctypes front-end of the CPU oracle (oracle/cppf_oracle.c, oracle/shot_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by cppf2_b200.  Function names and argument order mirror the reference's operators
(train_dino.py:171,218; dataset.py:118; eval.py:37; src_shot/shot.cpp:45) so the parity tests read
like calls into the reference.  numpy in, numpy out.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")


def build(force: bool = False) -> str:
    """Compiles the oracle with the recipe in oracle/Makefile (gcc, -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("cppf_oracle.c", "shot_oracle.cpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_grid_geometry.argtypes = [_f32p, C.c_int64, C.c_float, _f32p, _f32p, _i64p]
        L.oracle_vote_center.restype = C.c_int64
        L.oracle_vote_center.argtypes = [_f32p, C.c_int64, _i64p, C.c_int64, _f32p, C.c_int64, C.c_float, _f32p, _f32p,
                                         C.c_int, _f32p, _i64p, _i64p]
        L.oracle_grid_argmax.restype = C.c_int64
        L.oracle_grid_argmax.argtypes = [_i64p, _i64p, _f32p, C.c_double, _f64p]
        L.oracle_generate_targets.argtypes = [_f32p, C.c_int64, _f64p, _f64p, _f32p, C.c_void_p]
        L.oracle_decode_pairs.argtypes = [_f32p, _i64p, C.c_int64, _u8p, C.c_int64, C.c_int, C.c_void_p, _f32p, _f32p]
        L.oracle_vote_rotation.restype = C.c_int64
        L.oracle_vote_rotation.argtypes = [_f32p, _i64p, C.c_int64, _f32p, C.c_int64, _f32p, _f32p, C.c_int, _f32p, _u8p]
        L.oracle_sphere_hist.argtypes = [_f32p, C.c_int64, C.c_void_p, C.c_void_p, _f32p, C.c_int, C.c_float, _f64p]
        L.oracle_rotation_hist.argtypes = [_f32p, _i64p, C.c_int64, _f32p, C.c_void_p, C.c_void_p, C.c_int64, _f32p, _f32p,
                                           C.c_int, _f32p, C.c_int, C.c_float, _f64p]
        L.oracle_fibonacci_sphere.argtypes = [C.c_int, _f32p]
        L.oracle_backvote_errors.argtypes = [_f32p, _f32p, C.c_int64, _f32p]
        if hasattr(L, "oracle_shot_compute"):
            L.oracle_shot_compute.restype = C.c_int
            L.oracle_shot_compute.argtypes = [_f32p, C.c_int64, C.c_double, C.c_double, C.c_void_p, _f32p, C.c_int]
        if hasattr(L, "oracle_shot_lrf"):
            L.oracle_shot_lrf.restype = C.c_int
            L.oracle_shot_lrf.argtypes = [_f32p, C.c_int64, C.c_double, _f32p, np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")]
        _lib = L
    return _lib


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _idx(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


# -------------------------------------------------------------------------------------------------
# host-side constants both implementations take as inputs
# -------------------------------------------------------------------------------------------------

def angle_tables(num_rots: int) -> Tuple[np.ndarray, np.ndarray]:
    """cos/sin of `arange(R)/R*2*pi`, evaluated by torch on the CPU exactly as train_dino.py:195-196 writes it."""
    import torch
    angles = torch.arange(num_rots).to(torch.float32) / num_rots * 2 * np.pi
    return torch.cos(angles).numpy().copy(), torch.sin(angles).numpy().copy()


def fibonacci_sphere(samples: int) -> np.ndarray:
    """utils/util.py:191-207, float64 math cast to float32 (eval.py:80).  Returns [samples,3] float32."""
    out = np.empty((samples, 3), np.float32)
    lib().oracle_fibonacci_sphere(int(samples), out)
    return out


def cos_threshold(angle_tol: float) -> np.float32:
    """float32(cos(2*angle_tol degrees)) -- the comparison in eval.py:45 happens in float32."""
    return np.float32(np.cos(2 * angle_tol / 180 * np.pi))


# -------------------------------------------------------------------------------------------------
# operators
# -------------------------------------------------------------------------------------------------

def grid_geometry(pc, res: float):
    pc = _f32(pc)
    lo, hi, gr = np.empty(3, np.float32), np.empty(3, np.float32), np.empty(3, np.int64)
    lib().oracle_grid_geometry(pc, pc.shape[0], np.float32(res), lo, hi, gr)
    return lo, hi, gr


def vote_center(pc, preds_tr, res, point_idxs, num_rots=36, tables=None):
    """train_dino.py:171-215 -> (grid int64 [gx,gy,gz], cand_world float64 [3])."""
    pc, tr, idx = _f32(pc), _f32(preds_tr), _idx(point_idxs)
    ct, st = tables if tables is not None else angle_tables(num_rots)
    lo, _, gr = grid_geometry(pc, res)
    grid = np.zeros(int(gr.prod()), np.int64)
    lib().oracle_vote_center(pc, pc.shape[0], idx, idx.shape[1], tr, idx.shape[0], np.float32(res), _f32(ct), _f32(st),
                             int(num_rots), lo, gr, grid)
    world = np.empty(3, np.float64)
    lib().oracle_grid_argmax(grid, gr, lo, float(res), world)
    return grid.reshape(*gr), world


def generate_target_pairs(point_pairs, up, right, front, center=np.zeros((3,)), want_rot=True):
    """dataset.py:118-135 for float32 `point_pairs` [T,2,3] -> (target_tr f32 [T,2], target_rot f32 [T,3])."""
    pairs = _f32(point_pairs)
    T = pairs.shape[0]
    axes = np.ascontiguousarray(np.stack([np.asarray(up), np.asarray(right), np.asarray(front)]), dtype=np.float64)
    tr = np.empty((T, 2), np.float32)
    rot = np.empty((T, 3), np.float32) if want_rot else None
    lib().oracle_generate_targets(pairs, T, axes, np.ascontiguousarray(center, dtype=np.float64), tr,
                                  rot.ctypes.data if rot is not None else None)
    return tr, rot


def decode_pairs(pc, point_idxs, bins, num_bins=32):
    """eval.py:228-235 with injected draws -> (pred_pairs [T,2,3], pred_pairs_scaled [T,2,3], scale [T])."""
    pc, idx = _f32(pc), _idx(point_idxs)
    bins = np.ascontiguousarray(bins, dtype=np.uint8)
    T = idx.shape[0]
    pred, scaled, s = np.empty((T, 2, 3), np.float32), np.empty((T, 2, 3), np.float32), np.empty(T, np.float32)
    lib().oracle_decode_pairs(pc, idx, idx.shape[1], bins, T, int(num_bins), pred.ctypes.data, scaled, s)
    return pred, scaled, s


def vote_rotation(pc, preds_rot, point_idxs, num_rots=36, tables=None):
    """train_dino.py:218-239 -> (up f32 [M',R,3], mask bool [M])."""
    pc, th, idx = _f32(pc), _f32(preds_rot), _idx(point_idxs)
    ct, st = tables if tables is not None else angle_tables(num_rots)
    M = idx.shape[0]
    up = np.empty((M, num_rots, 3), np.float32)
    mask = np.empty(M, np.uint8)
    lib().oracle_vote_rotation(pc, idx, idx.shape[1], th, M, _f32(ct), _f32(st), int(num_rots), up, mask)
    mask = mask.astype(bool)
    return up[mask], mask


def sphere_counts(pred, sphere_pts, angle_tol, wt=None) -> np.ndarray:
    """The histogram inside get_topk_dir (eval.py:37-46), float64 [S]."""
    pred, sph = _f32(pred).reshape(-1, 3), _f32(sphere_pts)
    counts = np.zeros(sph.shape[0], np.float64)
    w = None if wt is None else np.ascontiguousarray(np.asarray(wt, dtype=np.float64).reshape(-1))
    lib().oracle_sphere_hist(pred, pred.shape[0], None if w is None else w.ctypes.data, None, sph, sph.shape[0],
                             cos_threshold(angle_tol), counts)
    return counts


def get_topk_dir(pred, sphere_pts, bmm_size, angle_tol, wt=None, topk=1):
    """eval.py:37-51 -> (dirs [topk,3], counts f32 [topk]); ties resolve to the lowest bin index."""
    counts = sphere_counts(pred, sphere_pts, angle_tol, wt).astype(np.float32)
    order = np.argsort(-counts, kind="stable")[:topk]
    return np.asarray(sphere_pts)[order], counts[order]


def rotation_counts(pc, point_idxs, theta, wt_pair, keep, num_rots, sphere_pts, angle_tol, tables=None) -> np.ndarray:
    """Fused vote_rotation+get_topk_dir histogram over kept pairs (float64 [S]); no [M,R,3] materialised."""
    pc, idx, th, sph = _f32(pc), _idx(point_idxs), _f32(theta), _f32(sphere_pts)
    ct, st = tables if tables is not None else angle_tables(num_rots)
    counts = np.zeros(sph.shape[0], np.float64)
    w = None if wt_pair is None else np.ascontiguousarray(wt_pair, dtype=np.float64)
    k = None if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
    lib().oracle_rotation_hist(pc, idx, idx.shape[1], th, None if w is None else w.ctypes.data,
                               None if k is None else k.ctypes.data, idx.shape[0], _f32(ct), _f32(st), int(num_rots), sph,
                               sph.shape[0], cos_threshold(angle_tol), counts)
    return counts


def backvote_errors(targets_tr, targets_tr_back) -> np.ndarray:
    a, b = _f32(targets_tr), _f32(targets_tr_back)
    out = np.empty(a.shape[0], np.float32)
    lib().oracle_backvote_errors(a, b, a.shape[0], out)
    return out


def backvote_filter(back_errs, point_idxs, n_points: int, backproj_ratio=0.1, imp_wt_margin=0.01):
    """eval.py:257-275: 10th-percentile mask, per-point occurrence counts, pair weights (float64)."""
    thr = np.percentile(back_errs, backproj_ratio * 100)
    mask = back_errs < thr
    flat = np.asarray(point_idxs)[mask, :2].reshape(-1)
    imp = np.bincount(flat, minlength=n_points).astype(np.int64)
    imp_wt = imp / imp.max()
    pair_wt = imp_wt[np.asarray(point_idxs)[mask, :2]].sum(-1) + imp_wt_margin
    return thr, mask, imp, pair_wt


def assemble_rotation(preds_up, preds_right, up_axis, right_axis) -> np.ndarray:
    """eval.py:295-313: Gram-Schmidt of `right` against `up`, third column by cross product."""
    preds_up = np.array(preds_up, dtype=np.float32)
    preds_right = np.array(preds_right, dtype=np.float32)
    preds_right -= np.dot(preds_up, preds_right) * preds_up
    preds_right /= (np.linalg.norm(preds_right) + 1e-9)
    up_loc = int(np.where(np.asarray(up_axis))[0][0])
    right_loc = int(np.where(np.asarray(right_axis))[0][0])
    R = np.eye(3)
    R[:3, up_loc] = preds_up
    R[:3, right_loc] = preds_right
    other = list({0, 1, 2} - {up_loc, right_loc})[0]
    R[:3, other] = np.cross(R[:3, (other + 1) % 3], R[:3, (other + 2) % 3])
    return R


def lower_median(x: np.ndarray) -> np.ndarray:
    """torch.median(x, 0)[0] (eval.py:309): the lower of the two middle elements for even counts."""
    s = np.sort(np.asarray(x), axis=0)
    return s[(s.shape[0] - 1) // 2]


def instance_body(pc, point_idxs_all, bins, pred_scales, cfg_up, cfg_right, cfg_front, res, num_rots=180,
                  backproj_ratio=0.1, imp_wt_margin=0.01, angle_tol=1.0, num_bins=32, sym_y_only=False, scale_override=None):
    """One branch of the per-instance body, eval.py:225-313 and :358-363 with opt=False, draws injected.

    Returns a dict with every intermediate the CUDA pipeline is compared on.
    """
    pc = _f32(pc)
    idx = _idx(point_idxs_all)
    tables = angle_tables(num_rots)
    sphere = fibonacci_sphere(int(4 * np.pi / (angle_tol / 180 * np.pi)))
    pred, scaled, pair_scale = decode_pairs(pc, idx, bins, num_bins)
    # call-site order (eval.py:237-240): positional (up, front, right)
    tr, rot = generate_target_pairs(scaled, cfg_up, cfg_front, cfg_right)
    grid, T_est = vote_center(pc, tr, res, idx[:, :2], num_rots, tables)
    input_pairs = pc[idx[:, :2]]
    tr_back, _ = generate_target_pairs(input_pairs, cfg_up, cfg_front, cfg_right, T_est, want_rot=False)
    errs = backvote_errors(tr, tr_back)
    thr, mask, imp, pair_wt = backvote_filter(errs, idx, pc.shape[0], backproj_ratio, imp_wt_margin)
    wt_full = np.ones(idx.shape[0], np.float64)
    wt_full[mask] = pair_wt
    counts_up = rotation_counts(pc, idx, rot[:, 0], wt_full, mask, num_rots, sphere, angle_tol, tables)
    counts_right = rotation_counts(pc, idx, rot[:, 2], wt_full, mask, num_rots, sphere, angle_tol, tables)
    b_up = int(np.argmax(counts_up.astype(np.float32)))
    b_right = int(np.argmax(counts_right.astype(np.float32)))
    R_est = assemble_rotation(sphere[b_up], sphere[b_right], cfg_up, cfg_right)
    if scale_override is not None:      # eval.py:308-310: the SHOT branch keeps the DINO branch's scale
        pred_scale = np.asarray(scale_override, dtype=np.float32)
    else:
        pred_scale = lower_median(np.asarray(pred_scales, dtype=np.float32)[mask])
    scale_norm = np.linalg.norm(pred_scale)
    pc_canon = (pc - T_est) @ R_est / scale_norm
    loss = np.abs(pc_canon[idx[mask, :2]] - pred[mask])
    if sym_y_only:
        loss = loss[..., 1]
    loss = np.clip(loss, 0, 0.1).mean()
    return dict(pred_pairs=pred, pair_scale=pair_scale, targets_tr=tr, targets_rot=rot, grid=grid, T_est=T_est,
                back_errs=errs, thr=thr, pairs_mask=mask, imp=imp, imp_pair_wt=pair_wt, counts_up=counts_up,
                counts_right=counts_right, bin_up=b_up, bin_right=b_right, R_est=R_est, pred_scale=pred_scale,
                loss=float(loss))


# -------------------------------------------------------------------------------------------------
# SHOT (filled in by shot_oracle.cpp)
# -------------------------------------------------------------------------------------------------

def shot_compute(pc, normal_r: float, shot_r: float, threads: int = 1):
    """src_shot/shot.cpp:45-100 -> [desc f32 flat N*352, normals f32 flat N*3]; NaN rows as PCL leaves them."""
    pc = _f32(pc).reshape(-1, 3)
    n = pc.shape[0]
    desc = np.empty(n * 352, np.float32)
    normals = np.empty(n * 3, np.float32)
    rc = lib().oracle_shot_compute(pc, n, float(normal_r), float(shot_r), desc.ctypes.data, normals, int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_shot_compute failed with code {rc}")
    return [desc, normals]


def estimate_normal(pc, normal_r: float, threads: int = 1) -> np.ndarray:
    """src_shot/shot.cpp:12-42 -> normals f32 flat N*3."""
    pc = _f32(pc).reshape(-1, 3)
    normals = np.empty(pc.shape[0] * 3, np.float32)
    rc = lib().oracle_shot_compute(pc, pc.shape[0], float(normal_r), 0.0, None, normals, int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_shot_compute failed with code {rc}")
    return normals


def shot_lrf(pc, shot_r: float):
    """SHOT local reference frames [N,9] (rows x,y,z) and the sign-vote margins [N,2] (0 = tie)."""
    pc = _f32(pc).reshape(-1, 3)
    rf = np.empty((pc.shape[0], 9), np.float32)
    margins = np.zeros((pc.shape[0], 2), np.int32)
    lib().oracle_shot_lrf(pc, pc.shape[0], float(shot_r), rf, margins)
    return rf, margins


# ---- instance cloud preparation -- plain numpy restatements -------------------------------------
def backproject(depth_m, intrinsics, instance_mask):
    """utils/util.py:2586-2607 followed by the callers' un-flip of x and y (eval.py:185-189): float32 camera-frame
    points of the masked, valid-depth pixels in np.where order, and their (rows, cols).  Pinned by
    tests/golden/backproject.npz (minted from the reference's own function)."""
    kinv = np.linalg.inv(np.asarray(intrinsics, dtype=np.float64))
    final = np.logical_and(instance_mask, depth_m > 0)
    rows, cols = np.where(final)
    uv1 = np.stack([cols, rows, np.ones_like(cols)], 0).astype(np.float64)
    xyz = (kinv @ uv1).T
    z = np.asarray(depth_m, dtype=np.float64)[rows, cols]
    pts = xyz * z[:, None] / xyz[:, -1:]
    return pts.astype(np.float32), (rows, cols)


def voxel_downsample(pc, res: float, prio) -> np.ndarray:
    """utils/util.py:39-46 restated: Open3D 0.18 `voxel_down_sample_and_trace(res, min_bound, max_bound)` bins point p
    into voxel floor((p - (min_bound - res/2)) / res) (float64; VoxelDownSampleAndTrace in open3d/geometry/PointCloud.cpp),
    then one member per voxel is drawn.  Open3D is not installable here (parity unpinned for the binning; the draw is
    injected): the member with the smallest (prio, index) represents its voxel.  Returns kept indices, ascending
    (Open3D's own order is that of an unordered_map and carries no meaning)."""
    p = np.asarray(pc, dtype=np.float32).astype(np.float64)
    lo = p.min(0) - 0.5 * float(res)
    vox = np.floor((p - lo) / float(res)).astype(np.int64)
    _, inv = np.unique(vox, axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.lexsort((np.arange(p.shape[0]), np.asarray(prio, dtype=np.float32), inv))
    first = np.ones(order.shape[0], bool)
    first[1:] = inv[order][1:] != inv[order][:-1]
    return np.sort(order[first])
