"""Drop-in for the reference's native `shot` module (pybind11 + PCL, src_shot/shot.cpp:164-169).

    from cppf2_b200 import shot            # instead of: from src_shot.build import shot
    desc, normal = shot.compute(pc, cfg.res * 10, cfg.res * 10)        # eval.py:210
    desc.reshape(-1, 352); normal.reshape(-1, 3)                        # eval.py:211,214

Same argument meaning and return layout: float32 `pc [N,3]` in, a list of two FLAT float32 arrays out
([N*352], [N*3]); invalid rows are NaN (the callers scrub them, eval.py:215-216), not errors.  A numpy
input returns numpy arrays like the reference; a CUDA tensor input returns CUDA tensors so that a
device-resident pipeline never round-trips.  The work runs in libcppf_b200.so on the current stream.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import check
from .voting import stream_ptr, to_device

_ws_cache = {}


def _workspace(n: int, device) -> torch.Tensor:
    lib = _lib.load()
    need = int(lib.cppf_shot_workspace_bytes(n))
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)   # one scratch per stream: instances may overlap
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def compute_device(pc: torch.Tensor, normal_r: float, shot_r: float, fast_math: bool = True, want_rf: bool = False,
                   normals_in=None):
    """pc CUDA f32 [N,3] -> (desc CUDA f32 [N,352], normals CUDA f32 [N,3][, rf CUDA f32 [N,9]]).
    `normals_in` [N,3] skips the normal estimation and describes with the given normals."""
    lib = _lib.load()
    pc = to_device(pc, torch.float32).reshape(-1, 3)
    n = pc.shape[0]
    desc = torch.empty((n, 352), dtype=torch.float32, device=pc.device)
    normals = torch.empty((n, 3), dtype=torch.float32, device=pc.device)
    rf = torch.empty((n, 9), dtype=torch.float32, device=pc.device) if want_rf else None
    ws = _workspace(n, pc.device)
    nin = None if normals_in is None else to_device(normals_in, torch.float32, pc.device).reshape(-1, 3)
    check(lib.cppf_shot_compute_ex(pc.data_ptr(), n, float(normal_r), float(shot_r), desc.data_ptr(), normals.data_ptr(),
                                   None if rf is None else rf.data_ptr(), int(fast_math),
                                   None if nin is None else nin.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()),
          "cppf_shot_compute")
    return (desc, normals, rf) if want_rf else (desc, normals)


def compute(pc, normal_r: float = 0.1, shot_r: float = 0.17):
    """shot.cpp:45-100.  Returns [desc_flat f32 [N*352], normal_flat f32 [N*3]]."""
    on_device = isinstance(pc, torch.Tensor) and pc.is_cuda
    desc, normals = compute_device(pc, normal_r, shot_r)
    if on_device:
        return [desc.reshape(-1), normals.reshape(-1)]
    return [desc.reshape(-1).cpu().numpy(), normals.reshape(-1).cpu().numpy()]


def estimate_normal(pc, normal_r: float = 0.1):
    """shot.cpp:12-42.  Returns normal_flat f32 [N*3]."""
    lib = _lib.load()
    on_device = isinstance(pc, torch.Tensor) and pc.is_cuda
    pc_t = to_device(pc, torch.float32).reshape(-1, 3)
    n = pc_t.shape[0]
    normals = torch.empty((n, 3), dtype=torch.float32, device=pc_t.device)
    ws = _workspace(n, pc_t.device)
    check(lib.cppf_estimate_normal(pc_t.data_ptr(), n, float(normal_r), normals.data_ptr(), ws.data_ptr(), ws.numel(),
                                   stream_ptr()), "cppf_estimate_normal")
    return normals.reshape(-1) if on_device else normals.reshape(-1).cpu().numpy()


def compute_color(pc, pc_color, normal_r: float = 0.1, shot_r: float = 0.17):
    """shot.cpp:102-161 (SHOT1344).  No call site exists anywhere in the reference; kept as a named stub."""
    raise NotImplementedError("compute_color (SHOT1344) is not on the CPPF++ inference path and is not implemented")
