"""Instance cloud preparation on the device -- the step in front of SHOT.

Drop-in mirrors of the reference's host helpers, same names and argument meaning:
    backproject(depth, intrinsics, instance_mask) -> (pts, idxs)        utils/util.py:2586-2607
    downsample(pc, res) -> indices                                      utils/util.py:39-46
and the device-resident forms the frame driver uses (`backproject_device`, `voxel_downsample_device`,
`prepare_instance_clouds`), which keep the cloud in HBM between the depth image and SHOT.
There is no CPU fallback: every function calls libcppf_b200.so.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check
from .voting import stream_ptr, to_device

MAX_POINTS = 50000          # eval.py:195
GRID_GUARD = 1000           # eval.py:200

_ws_cache = {}


def _workspace(n: int, device) -> torch.Tensor:
    need = int(_lib.load().cppf_cloud_workspace_bytes(int(n)))
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = _ws_cache[key] = torch.empty(need, dtype=torch.uint8, device=device)
    return ws


def _kinv(intrinsics) -> "np.ndarray":
    return np.ascontiguousarray(np.linalg.inv(np.asarray(intrinsics, dtype=np.float64)))


def backproject_device(depth: torch.Tensor, intrinsics, mask: torch.Tensor, depth_div: float = 1.0):
    """depth CUDA uint16/int16/float32 [H,W] (divided by `depth_div`: 1000 for REAL275 millimetres), mask CUDA bool/uint8
    [H,W] -> (pc f32 [H*W,3] capacity, pix i32 [H*W] capacity (row*W + col), count i64 [1] on the device).
    Stream-ordered, no synchronisation: the caller reads `count` when it needs the size."""
    lib = _lib.load()
    dev = depth.device
    H, W = depth.shape
    if depth.dtype in (torch.uint16, torch.int16):
        is_u16 = 1
    elif depth.dtype == torch.float32:
        is_u16 = 0
    else:
        raise TypeError(f"depth must be uint16 or float32, got {depth.dtype}")
    depth = depth.contiguous()
    m8 = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.to(torch.uint8).contiguous()
    pc = torch.empty((H * W, 3), dtype=torch.float32, device=dev)
    pix = torch.empty(H * W, dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = _workspace(H * W, dev)
    kinv = _kinv(intrinsics)
    check(lib.cppf_backproject(depth.data_ptr(), is_u16, float(depth_div), m8.data_ptr(), H, W,
                               kinv.ctypes.data_as(_lib._DP), pc.data_ptr(), pix.data_ptr(), count.data_ptr(),
                               ws.data_ptr(), ws.numel(), stream_ptr()), "cppf_backproject")
    return pc, pix, count


def voxel_downsample_device(pc: torch.Tensor, res: float, prio: Optional[torch.Tensor] = None, seed: int = 0,
                            side: Optional[torch.Tensor] = None):
    """pc CUDA f32 [n,3] -> (pc_out [n,3] capacity, kept_idx i32 [n] capacity, count i64 [1], side_out or None).
    `prio` f32 [n] in [0,1) injects the per-point draw (smallest wins in its voxel); else a counter RNG keyed by seed."""
    lib = _lib.load()
    pc = to_device(pc, torch.float32).reshape(-1, 3)
    dev, n = pc.device, pc.shape[0]
    out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
    kept = torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    count = torch.zeros(1, dtype=torch.int64, device=dev)
    side_out = None if side is None else torch.empty(max(n, 1), dtype=torch.int32, device=dev)
    pr = None if prio is None else to_device(prio, torch.float32, dev)
    ws = _workspace(n, dev)
    check(lib.cppf_voxel_downsample(pc.data_ptr(), n, float(res), None if pr is None else pr.data_ptr(), int(seed), out.data_ptr(),
                                    kept.data_ptr(), count.data_ptr(), None if side is None else side.data_ptr(),
                                    None if side_out is None else side_out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()),
          "cppf_voxel_downsample")
    return out, kept, count, side_out


def cap_points_device(pc: torch.Tensor, side: Optional[torch.Tensor], seed: int):
    """eval.py:195-198: more than 50 000 points -> 50 000 draws with replacement."""
    lib = _lib.load()
    n = pc.shape[0]
    if n <= MAX_POINTS:
        return pc, side
    sub = torch.empty(MAX_POINTS, dtype=torch.int32, device=pc.device)
    check(lib.cppf_sample_tuples(n, MAX_POINTS, 1, int(seed), sub.data_ptr(), stream_ptr()), "cppf_sample_tuples")
    out = torch.empty((MAX_POINTS, 3), dtype=torch.float32, device=pc.device)
    side_out = None if side is None else torch.empty(MAX_POINTS, dtype=torch.int32, device=pc.device)
    check(lib.cppf_gather_points(pc.data_ptr(), sub.data_ptr(), MAX_POINTS, out.data_ptr(), None if side is None else side.data_ptr(),
                                 None if side_out is None else side_out.data_ptr(), stream_ptr()), "cppf_gather_points")
    return out, side_out


def prepare_instance_clouds(depth: torch.Tensor, masks: Sequence[torch.Tensor], intrinsics, res: Sequence[float],
                            depth_div: float = 1000.0, seed: int = 0, min_points: int = 50
                            ) -> List[Optional[Tuple[torch.Tensor, torch.Tensor]]]:
    """eval.py:185-201 for every instance of a frame, on the device: back-project, voxel down-sample at the category's
    `res`, cap at 50 000 points, drop instances whose extent exceeds 1000 voxels.  Two host synchronisations per frame
    (the point counts after each stage), not per instance.  Returns per instance (pc f32 [N,3], pix i32 [N]) or None."""
    stage1 = [backproject_device(depth, intrinsics, m, depth_div) for m in masks]
    counts = torch.cat([c for _, _, c in stage1]).cpu().numpy()                  # sync 1
    stage2 = []
    for i, ((pc, pix, _), n) in enumerate(zip(stage1, counts)):
        if n < min_points:
            stage2.append(None)
            continue
        stage2.append(voxel_downsample_device(pc[:n], res[i], seed=seed * 1000003 + i, side=pix[:n]))
    live = [s for s in stage2 if s is not None]
    counts2 = torch.cat([s[2] for s in live]).cpu().numpy() if live else np.zeros(0, np.int64)   # sync 2
    out, k = [], 0
    for i, s in enumerate(stage2):
        if s is None:
            out.append(None)
            continue
        n = int(counts2[k])
        k += 1
        pc, pix = cap_points_device(s[0][:n], s[3][:n], seed=seed * 7919 + i)
        out.append((pc, pix))
    # the extent guard (eval.py:200) is evaluated by cppf_cloud_bounds inside the vote chain (CPPF_STATUS_GRID_GUARD);
    # the frame driver drops such instances when it reads the pose record
    return out


# ---- drop-in host signatures -----------------------------------------------------------------------------------------
def backproject(depth, intrinsics, instance_mask):
    """utils/util.py:2586-2607: returns (pts [n,3] with x and y negated, (rows, cols)) like the reference.  pts are the
    float32 values of the device kernel widened to float64: the callers' `.astype(np.float32)` (eval.py:189) is exact."""
    d = np.asarray(depth)
    dt = torch.from_numpy(np.ascontiguousarray(d, dtype=np.float32)).cuda()
    mt = torch.from_numpy(np.ascontiguousarray(instance_mask).astype(np.uint8)).cuda()
    pc, pix, count = backproject_device(dt, intrinsics, mt, 1.0)
    n = int(count.item())
    pts = pc[:n].cpu().numpy().astype(np.float64)
    pts[:, 0] = -pts[:, 0]
    pts[:, 1] = -pts[:, 1]
    p = pix[:n].cpu().numpy().astype(np.int64)
    return pts, (p // d.shape[1], p % d.shape[1])


def downsample(pc, res, prio=None, seed: int = 0):
    """utils/util.py:39-46: one random member per occupied voxel; returns the kept indices (ascending)."""
    _, kept, count, _ = voxel_downsample_device(torch.from_numpy(np.ascontiguousarray(pc, dtype=np.float32)).cuda(), res, prio=prio, seed=seed)
    return kept[:int(count.item())].cpu().numpy().astype(np.int64)


def interpolate_features(descriptors, pts, strides=8, normalize=True):
    """dataset.py:40-59: descriptors [1,C,h,w] (any strides: the reference passes the permuted ViT token view),
    pts [1,n,2] (x, y) -> [1,C,n] like the reference (a transposed view of the kernel's [n,C] rows; the caller's
    `[0].T` (dataset.py:79) then yields contiguous [n,C])."""
    lib = _lib.load()
    d = descriptors if isinstance(descriptors, torch.Tensor) else torch.from_numpy(np.asarray(descriptors))
    d = d.cuda().float() if not d.is_cuda else d.float()
    p = pts if isinstance(pts, torch.Tensor) else torch.from_numpy(np.asarray(pts))
    p = p.to(d.device, torch.float32).reshape(-1, 2).contiguous()
    if d.dim() != 4 or d.shape[0] != 1:
        raise ValueError("descriptors must be [1,C,h,w]")
    _, C, h, w = d.shape
    out = torch.empty((p.shape[0], C), dtype=torch.float32, device=d.device)
    check(lib.cppf_interpolate_features(d.data_ptr(), C, h, w, d.stride(1), d.stride(2), d.stride(3), p.data_ptr(), p.shape[0],
                                        float(strides), int(bool(normalize)), out.data_ptr(), stream_ptr()), "cppf_interpolate_features")
    return out.T.unsqueeze(0)
