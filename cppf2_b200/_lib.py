"""ctypes binding of libcppf_b200.so (the C ABI declared in include/cppf_b200.h).

There is no fallback: if the shared library is missing or a call returns an error code, a Python
exception is raised.  PyTorch is only used by the callers for device memory and streams; nothing
torch-typed crosses this boundary (plain pointers and sizes).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CPPF_B200_LIB") or os.path.join(_HERE, "libcppf_b200.so")   # override: experiment builds only
CSRC = os.path.join(_HERE, "csrc")

CPPF_STATUS_GRID_OVERFLOW = 1
CPPF_STATUS_GRID_GUARD = 2
CPPF_STATUS_EMPTY = 4
CPPF_STATUS_REFINED = 8


class CppfError(RuntimeError):
    pass


class GridGeom(C.Structure):
    _fields_ = [("lo", C.c_float * 3), ("hi", C.c_float * 3), ("res", C.c_float), ("flags", C.c_uint32),
                ("grid_res", C.c_int64 * 3), ("cells", C.c_int64)]


class Center(C.Structure):
    _fields_ = [("world", C.c_double * 3), ("cell", C.c_int64 * 3), ("linear", C.c_int64), ("votes", C.c_uint32),
                ("status", C.c_uint32), ("cells", C.c_int64)]


class BackvoteSummary(C.Structure):
    _fields_ = [("threshold", C.c_float), ("s_lo", C.c_float), ("s_hi", C.c_float), ("imp_max", C.c_int32),
                ("kept", C.c_int64)]


class Pose(C.Structure):
    _fields_ = [("R", C.c_double * 9), ("t", C.c_double * 3), ("scale", C.c_float * 3), ("scale_norm", C.c_float),
                ("loss", C.c_double), ("bin_up", C.c_int32), ("bin_right", C.c_int32), ("count_up", C.c_float),
                ("count_right", C.c_float), ("kept", C.c_int64), ("status", C.c_uint32), ("grid_cells", C.c_uint32)]


class ScaleSelect(C.Structure):
    _fields_ = [("prefix", C.c_uint32 * 3), ("pad", C.c_uint32), ("k", C.c_uint64 * 3)]


class VoteParams(C.Structure):
    _fields_ = [("res", C.c_double), ("num_rots", C.c_int), ("num_bins", C.c_int), ("sphere_bins", C.c_int),
                ("cos_thr", C.c_float), ("band", C.c_int), ("lut_g", C.c_int), ("up_loc", C.c_int), ("right_loc", C.c_int),
                ("loss_y_only", C.c_int), ("pad0", C.c_int), ("lut", C.c_void_p), ("cos_tab", C.c_void_p), ("sin_tab", C.c_void_p),
                ("sphere", C.c_void_p), ("axes", C.c_double * 9), ("imp_margin", C.c_double), ("rank_lo", C.c_int64),
                ("gamma", C.c_float), ("refine_iters", C.c_int), ("refine_lr", C.c_float), ("pad", C.c_int)]


class VoteBuffers(C.Structure):
    _fields_ = [("grid", C.c_void_p), ("grid_capacity", C.c_int64), ("geom", C.c_void_p), ("center", C.c_void_p),
                ("summary", C.c_void_p), ("status", C.c_void_p), ("targets_tr", C.c_void_p), ("targets_rot", C.c_void_p),
                ("errs", C.c_void_p), ("keep", C.c_void_p), ("kept_list", C.c_void_p), ("imp", C.c_void_p),
                ("counts", C.c_void_p), ("ws_backvote", C.c_void_p), ("ws_backvote_bytes", C.c_int64),
                ("ws_pose", C.c_void_p), ("ws_pose_bytes", C.c_int64)]


class InstanceIO(C.Structure):
    _fields_ = [("pc", C.c_void_p), ("n", C.c_int64), ("idx", C.c_void_p), ("idx_is_i64", C.c_int), ("pad0", C.c_int),
                ("idx_stride", C.c_int64), ("T", C.c_int64), ("dino_desc", C.c_void_p), ("heads_dino", C.c_void_p),
                ("heads_shot", C.c_void_p), ("normal_r", C.c_float), ("shot_r", C.c_float), ("shot_desc", C.c_void_p),
                ("normals", C.c_void_p), ("ws_shot", C.c_void_p), ("ws_shot_bytes", C.c_int64), ("bins", C.c_void_p),
                ("scales", C.c_void_p), ("ws_heads", C.c_void_p), ("ws_heads_bytes", C.c_int64), ("seed_dino", C.c_uint64),
                ("seed_shot", C.c_uint64), ("cells_hint", C.c_int64), ("pose_dino", C.c_void_p), ("pose_shot", C.c_void_p),
                ("idx_draw", C.c_void_p), ("seed_idx", C.c_uint64)]


class Frame(C.Structure):
    _fields_ = [("n_instances", C.c_int), ("mode", C.c_int), ("io", C.c_void_p), ("params", C.c_void_p), ("buffers", C.c_void_p),
                ("shared", C.c_void_p), ("heads_dino_any", C.c_void_p), ("heads_shot_any", C.c_void_p), ("table_host", C.c_void_p),
                ("table_dev", C.c_void_p), ("capacity_instances", C.c_int), ("replicas_max", C.c_int),
                ("capacity_tuples", C.c_int64), ("capacity_points", C.c_int64), ("stage_events", C.c_void_p)]


FRAME_FILL, FRAME_COPY, FRAME_LAUNCH, FRAME_ALL = 1, 2, 4, 7
FRAME_MAX_INSTANCES = 16
FRAME_STAGES = ("sample", "shot", "heads", "center", "backvote", "rotation", "pose")      # between the 8 stage events


P, I, I64, F, D, U64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_uint64
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)

# name -> (restype, argtypes); mirrors include/cppf_b200.h one to one
SIGNATURES = {
    "cppf_error_string": (C.c_char_p, [I]),
    "cppf_version": (I, []),
    "cppf_device_info": (I, [_IP, C.POINTER(I64), _IP, _IP]),
    "cppf_cloud_bounds": (I, [P, I64, F, P, P]),
    "cppf_vote_center": (I, [P, I64, P, I, I64, P, I64, P, P, I, P, P, I64, I64, I, P, P]),
    "cppf_vote_center_ex": (I, [P, I64, P, I, I64, P, I64, P, P, I, P, P, I64, I, P, I, I, I64, P]),
    "cppf_vote_center_smem_cells": (I64, []),
    "cppf_grid_argmax": (I, [P, I64, P, D, P, P, P]),
    "cppf_grid_to_i64": (I, [P, I64, P, P, P]),
    "cppf_sample_tuples": (I, [I64, I64, I, U64, P, P]),
    "cppf_sample_bins": (I, [P, I64, I, P, U64, P, P]),
    "cppf_decode_targets": (I, [P, P, I, I64, P, I64, I, _DP, P, P, P, P, P]),
    "cppf_generate_targets": (I, [P, I64, _DP, P, P, P, P]),
    "cppf_backvote_workspace_bytes": (I64, [I64, I64]),
    "cppf_backvote_filter": (I, [P, I64, P, I, I64, P, I64, _DP, P, I64, F, P, P, P, P, P, P, I64, P]),
    "cppf_backvote_errors": (I, [P, P, I, I64, P, I64, P, P, P]),
    "cppf_backvote_select": (I, [P, I64, I64, F, P, P, I64, P]),
    "cppf_backvote_mask": (I, [P, P, I, I64, I64, I64, P, P, P, P, I, P]),
    "cppf_backvote_imp_max": (I, [P, I64, P, P]),
    "cppf_vote_rotation": (I, [P, P, I, I64, P, I64, P, P, I, P, P, P]),
    "cppf_sphere_hist": (I, [P, I64, P, P, I, F, I, P, P]),
    "cppf_sphere_band": (I, [I, F]),
    "cppf_rotation_hist": (I, [P, P, I, I64, P, I64, _IP, I, P, P, I64, P, P, D, P, P, I, P, I, F, I, P, I, P, P]),
    "cppf_rotation_hist_part": (I, [P, P, I, I64, P, I64, _IP, I, P, P, I64, P, P, D, P, P, I, P, I, F, I, P, I, P, I, I, P]),
    "cppf_sphere_lut_bytes": (I64, [I]),
    "cppf_sphere_lut_build": (I, [P, I, F, I, P]),
    "cppf_cloud_workspace_bytes": (I64, [I64]),
    "cppf_backproject": (I, [P, I, D, P, I, I, _DP, P, P, P, P, I64, P]),
    "cppf_voxel_downsample": (I, [P, I64, D, P, U64, P, P, P, P, P, P, I64, P]),
    "cppf_gather_points": (I, [P, P, I64, P, P, P, P]),
    "cppf_interpolate_features": (I, [P, I, I, I, I64, I64, I64, P, I64, F, I, P, P]),
    "cppf_vote_chain": (I, [P, I64, P, I, I64, I64, P, P, P, I64, P, P, P, P]),
    "cppf_instance_pose": (I, [P, P, P, P]),
    "cppf_frame_table_bytes": (I64, []),
    "cppf_frame_heads_workspace_bytes": (I64, [P, P, I64]),
    "cppf_frame_pose": (I, [P, P]),
    "cppf_pose_workspace_bytes": (I64, [I64]),
    "cppf_scale_median_hist": (I, [P, P, P, I64, I, P, P, P]),
    "cppf_scale_median_pick": (I, [P, P, I, P, P, P]),
    "cppf_pose_finalize": (I, [P, P, I, I64, P, I, P, P, P, P, P, I, P, I, I, I, P, P, P, I64, P]),
    "cppf_pose_finalize_refine": (I, [P, P, I, I64, P, I, P, P, P, P, P, I, P, I, I, I, P, I, F, I64, P, P, I64, P]),
    "cppf_shot_workspace_bytes": (I64, [I64]),
    "cppf_shot_compute": (I, [P, I64, F, F, P, P, P, I64, P]),
    "cppf_estimate_normal": (I, [P, I64, F, P, P, I64, P]),
    "cppf_shot_compute_ex": (I, [P, I64, F, F, P, P, P, I, P, P, I64, P]),
    "cppf_shot_compute_color": (I, [P, P, I64, F, F, P, P]),
    "cppf_heads_create": (I, [I, I, C.POINTER(C.c_float), I64, C.POINTER(P)]),
    "cppf_heads_destroy": (I, [P]),
    "cppf_heads_has_tc": (I, [P]),
    "cppf_heads_workspace_bytes": (I64, [P, I64, I64, I]),
    "cppf_heads_forward": (I, [P, I, P, I64, P, I, I64, I64, P, P, P, P, P, I64, P]),
    "cppf_heads_forward_sampled": (I, [P, I, P, I64, P, I, I64, I64, P, P, P, U64, P, P, P, I64, P]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compiles every CUDA source for sm_100a into libcppf_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"] + ([] if verbose else ["-s"])
    subprocess.run(cmd, check=True)
    if not os.path.exists(LIB_PATH):
        raise CppfError(f"build finished but {LIB_PATH} is missing")
    return LIB_PATH


def load() -> C.CDLL:
    """Loads the library; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CppfError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the cppf2_b200 hot path)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int, what: str = "") -> None:
    if code != 0:
        msg = load().cppf_error_string(code).decode()
        raise CppfError(f"{what or 'libcppf_b200'} failed: {msg} (code {code})")


def axes_array(up, right, front):
    """9 doubles: (up, right, front) in the positional order of dataset.py:118."""
    vals = [float(v) for ax in (up, right, front) for v in ax]
    return (C.c_double * 9)(*vals)
