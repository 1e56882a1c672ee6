"""Drop-in operators of the voting path, with the reference's names, argument order and return types.

    vote_center(pc, preds_tr, res, point_idxs, num_rots=36, vis=None)     train_dino.py:171-215
    vote_rotation(pc, preds_rot, point_idxs, num_rots=36)                 train_dino.py:218-239
    generate_target_pairs(point_pairs, up, right, front, center=0)        dataset.py:118-135
    get_topk_dir(pred, sphere_pts, bmm_size, angle_tol, wt=None, topk=1)  eval.py:37-51
    fibonacci_sphere(samples)                                             utils/util.py:191-207

Inputs may be CUDA tensors (what eval.py passes) or host numpy arrays (copied to the current
device).  All arithmetic runs in libcppf_b200.so on the current CUDA stream; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import BackvoteSummary, Center, GridGeom, check


# ---------------------------------------------------------------------------------------------------
# plumbing
# ---------------------------------------------------------------------------------------------------

def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_device(x, dtype: torch.dtype, device=None) -> torch.Tensor:
    """numpy / tensor -> contiguous tensor of `dtype` on the current CUDA device."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if isinstance(x, torch.Tensor):
        t = x
    else:
        t = torch.from_numpy(np.ascontiguousarray(x))
    if t.dtype != dtype or t.device != device:
        t = t.to(device=device, dtype=dtype, non_blocking=True)
    return t.contiguous()


def idx_args(idx: torch.Tensor) -> Tuple[int, int, int]:
    """(pointer, is_i64, row stride in elements) of a [T,>=2] index tensor whose columns are adjacent."""
    if idx.dtype not in (torch.int64, torch.int32):
        raise TypeError(f"tuple indices must be int64 or int32, got {idx.dtype}")
    if idx.dim() != 2 or idx.shape[1] < 2:
        raise ValueError(f"tuple indices must be [T,>=2], got {tuple(idx.shape)}")
    if idx.shape[0] > 1 and idx.stride(1) != 1:
        raise ValueError("tuple index columns must be adjacent in memory")
    stride = idx.stride(0) if idx.shape[0] > 1 else idx.shape[1]
    return idx.data_ptr(), int(idx.dtype == torch.int64), int(stride)


def device_index_tensor(point_idxs, device=None) -> torch.Tensor:
    """Keeps int32/int64 and row-strided CUDA views as they are; copies host arrays over."""
    if isinstance(point_idxs, torch.Tensor) and point_idxs.is_cuda and point_idxs.dtype in (torch.int64, torch.int32):
        if point_idxs.dim() == 2 and (point_idxs.shape[0] <= 1 or point_idxs.stride(1) == 1):
            return point_idxs
        return point_idxs.contiguous()
    arr = point_idxs.cpu().numpy() if isinstance(point_idxs, torch.Tensor) else np.asarray(point_idxs)
    dt = torch.int32 if arr.dtype == np.int32 else torch.int64
    return to_device(arr.astype(np.int32 if dt == torch.int32 else np.int64), dt, device)


@lru_cache(maxsize=16)
def _angle_tables_host(num_rots: int):
    # the literal reference expression (train_dino.py:195-196) evaluated on torch-CPU
    angles = torch.arange(num_rots).to(torch.float32) / num_rots * 2 * np.pi
    return torch.cos(angles).contiguous(), torch.sin(angles).contiguous()


_table_cache = {}


def angle_tables(num_rots: int, device=None):
    """cos/sin tables on the device; computed once per (R, device) by torch on the CPU."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (int(num_rots), str(device))
    if key not in _table_cache:
        c, s = _angle_tables_host(int(num_rots))
        _table_cache[key] = (c.to(device), s.to(device))
    return _table_cache[key]


def fibonacci_sphere(samples):
    """utils/util.py:191-207 -- host float64 math, list of (x, y, z) tuples like the reference."""
    points = []
    golden = math.pi * (3.0 - math.sqrt(5.0))
    for i in range(samples):
        y = 1 - (i / float(samples - 1)) * 2
        r = math.sqrt(1 - y * y)
        th = golden * i
        points.append((math.cos(th) * r, y, math.sin(th) * r))
    return points


_sphere_cache = {}


def sphere_points(samples: int, device=None) -> torch.Tensor:
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (int(samples), str(device))
    if key not in _sphere_cache:
        _sphere_cache[key] = torch.from_numpy(np.array(fibonacci_sphere(samples), dtype=np.float32)).to(device)
    return _sphere_cache[key]


_lut_cache = {}


def sphere_lut(samples: int, cos_thr: float, device=None):
    """(device table, G) of the cube-map lookup that narrows the sphere test to <= 4 lattice points per
    direction (cppf_sphere_lut_build, host-built once per (S, threshold, device)); (None, 0) when no cube
    resolution fits the 4-entry cells, in which case the kernels use the latitude band."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    key = (int(samples), float(cos_thr), str(device))
    if key not in _lut_cache:
        lib = _lib.load()
        sph = np.ascontiguousarray(np.array(fibonacci_sphere(samples), dtype=np.float32))
        entry = (None, 0)
        for g in (32, 48, 64, 96, 128):
            buf = np.empty(int(lib.cppf_sphere_lut_bytes(g)), dtype=np.uint8)
            rc = lib.cppf_sphere_lut_build(sph.ctypes.data, int(samples), float(cos_thr), g, buf.ctypes.data)
            if rc == 0:
                entry = (torch.from_numpy(buf).to(device), g)
                break
        _lut_cache[key] = entry
    return _lut_cache[key]


def cos_threshold(angle_tol: float) -> float:
    """float32(cos(2*angle_tol deg)): eval.py:45 compares in float32."""
    return float(np.float32(np.cos(2 * angle_tol / 180 * np.pi)))


def struct_tensor(ctype, device) -> torch.Tensor:
    """Zeroed device bytes large enough for one `ctype`."""
    return torch.zeros(C.sizeof(ctype), dtype=torch.uint8, device=device)


def read_struct(t: torch.Tensor, ctype):
    """Device struct -> host ctypes instance (synchronises the current stream)."""
    host = t.cpu().numpy().tobytes()
    return ctype.from_buffer_copy(host)


# ---------------------------------------------------------------------------------------------------
# vote_center
# ---------------------------------------------------------------------------------------------------

def vote_center(pc, preds_tr, res, point_idxs, num_rots=36, vis=None):
    """train_dino.py:171-215.  Returns (grid_obj int64 numpy [gx,gy,gz], cand_world float64 numpy [3])."""
    lib = _lib.load()
    pc = to_device(pc, torch.float32)
    dev = pc.device
    tr = to_device(preds_tr, torch.float32, dev)
    idx = device_index_tensor(point_idxs, dev)
    T = idx.shape[0]
    if tr.shape[0] != T:
        raise ValueError("preds_tr and point_idxs disagree on the number of tuples")
    ct, st = angle_tables(num_rots, dev)
    s = stream_ptr()
    geom_t = struct_tensor(GridGeom, dev)
    check(lib.cppf_cloud_bounds(pc.data_ptr(), pc.shape[0], float(res), geom_t.data_ptr(), s), "cppf_cloud_bounds")
    geom = read_struct(geom_t, GridGeom)  # the reference synchronises here as well (grid_res.long() -> zeros)
    shape = tuple(int(g) for g in geom.grid_res)
    cells = int(geom.cells)
    # room for up to 32 copies of the grid (<= 64 MB): the kernel spreads the vote peak over them
    copies = max(1, min(32, (1 << 24) // max(cells, 1)))
    grid = torch.empty(max(cells, 1) * copies, dtype=torch.int32, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    ip, i64, istr = idx_args(idx)
    check(lib.cppf_vote_center(pc.data_ptr(), pc.shape[0], ip, i64, istr, tr.data_ptr(), T, ct.data_ptr(), st.data_ptr(),
                               int(num_rots), geom_t.data_ptr(), grid.data_ptr(), grid.numel(), cells, 0, status.data_ptr(), s),
          "cppf_vote_center")
    center_t = struct_tensor(Center, dev)
    check(lib.cppf_grid_argmax(grid.data_ptr(), grid.numel(), geom_t.data_ptr(), float(res), status.data_ptr(), center_t.data_ptr(), s),
          "cppf_grid_argmax")
    grid64 = torch.empty(max(cells, 1), dtype=torch.int64, device=dev)
    check(lib.cppf_grid_to_i64(grid.data_ptr(), grid.numel(), geom_t.data_ptr(), grid64.data_ptr(), s), "cppf_grid_to_i64")
    grid_obj = grid64[:cells].cpu().numpy().reshape(shape)
    center = read_struct(center_t, Center)
    cand_world = np.array(list(center.world), dtype=np.float64)
    if vis is not None:  # debugging hook of the reference (visdom heatmaps); kept for signature parity
        try:
            vis.heatmap(grid_obj.max(0), win="33", opts=dict(title="front"))
        except Exception:
            pass
    return grid_obj, cand_world


# ---------------------------------------------------------------------------------------------------
# generate_target_pairs
# ---------------------------------------------------------------------------------------------------

def generate_target_pairs(point_pairs, up, right, front, center=np.zeros((3,))):
    """dataset.py:118-135.  float32 pairs [T,2,3] -> (target_tr f32 [T,2], target_rot f32 [T,3]) numpy."""
    lib = _lib.load()
    pairs = to_device(point_pairs, torch.float32)
    dev = pairs.device
    T = pairs.shape[0]
    tr = torch.empty((T, 2), dtype=torch.float32, device=dev)
    rot = torch.empty((T, 3), dtype=torch.float32, device=dev)
    ctr = to_device(np.asarray(center, dtype=np.float64).reshape(3), torch.float64, dev)
    axes = _lib.axes_array(up, right, front)
    check(lib.cppf_generate_targets(pairs.data_ptr(), T, axes, ctr.data_ptr(), tr.data_ptr(), rot.data_ptr(), stream_ptr()),
          "cppf_generate_targets")
    return tr.cpu().numpy(), rot.cpu().numpy()


# ---------------------------------------------------------------------------------------------------
# vote_rotation / get_topk_dir
# ---------------------------------------------------------------------------------------------------

def vote_rotation(pc, preds_rot, point_idxs, num_rots=36):
    """train_dino.py:218-239.  Returns (up f32 CUDA [M',R,3], mask bool CUDA [M])."""
    lib = _lib.load()
    pc = to_device(pc, torch.float32)
    dev = pc.device
    th = to_device(preds_rot, torch.float32, dev)
    idx = device_index_tensor(point_idxs, dev)
    M = idx.shape[0]
    ct, st = angle_tables(num_rots, dev)
    up = torch.empty((M, int(num_rots), 3), dtype=torch.float32, device=dev)
    mask = torch.empty(M, dtype=torch.uint8, device=dev)
    ip, i64, istr = idx_args(idx)
    check(lib.cppf_vote_rotation(pc.data_ptr(), ip, i64, istr, th.data_ptr(), M, ct.data_ptr(), st.data_ptr(), int(num_rots),
                                 up.data_ptr(), mask.data_ptr(), stream_ptr()), "cppf_vote_rotation")
    mask = mask.bool()
    return up[mask], mask


def sphere_counts(pred, sphere_pts, angle_tol, wt=None) -> torch.Tensor:
    """The float64 histogram inside get_topk_dir (eval.py:43-45), on the device."""
    lib = _lib.load()
    pred = to_device(pred, torch.float32).reshape(-1, 3)
    dev = pred.device
    sph = to_device(sphere_pts, torch.float32, dev)
    S = sph.shape[0]
    w = None
    if wt is not None:
        w = to_device(wt, torch.float64, dev).reshape(-1)
        if w.numel() != pred.shape[0]:
            raise ValueError("wt must have one entry per row of pred")
    thr = cos_threshold(angle_tol)
    band = lib.cppf_sphere_band(S, thr)
    counts = torch.zeros(S, dtype=torch.float64, device=dev)
    check(lib.cppf_sphere_hist(pred.data_ptr(), pred.shape[0], None if w is None else w.data_ptr(), sph.data_ptr(), S, thr,
                               band, counts.data_ptr(), stream_ptr()), "cppf_sphere_hist")
    return counts


def get_topk_dir(pred, sphere_pts, bmm_size, angle_tol, wt=None, topk=1):
    """eval.py:37-51.  Returns (dirs numpy [topk,3], counts numpy f32 [topk]).  `bmm_size` is accepted for
    signature parity; nothing is materialised, so no chunking is needed."""
    counts = sphere_counts(pred, sphere_pts, angle_tol, wt).to(torch.float32).cpu().numpy()
    order = np.argsort(-counts, kind="stable")[:topk]  # ties -> lowest bin index
    sph = sphere_pts.cpu().numpy() if isinstance(sphere_pts, torch.Tensor) else np.asarray(sphere_pts)
    return np.array(sph[order]), counts[order]
