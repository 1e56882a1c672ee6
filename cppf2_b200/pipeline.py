"""Per-instance pose voting, the body of the reference's instance loop (eval.py:219-313, 358-372;
demo.py:168-300) as one stream-ordered chain of libcppf_b200 kernels.

    decode draws -> vote targets -> centre Hough vote -> arg-max -> back-vote filter (exact percentile)
      -> importance weights -> rotation votes (up, right) -> top-1 directions -> R, t, scale, branch loss

Nothing returns to the host between the stages: the grid geometry, voted centre, percentile threshold,
kept list and histogram bins all stay in HBM, and the caller reads one 200-byte `cppf_pose` record at
the end.  The reference crosses the host/device boundary ~22 times per (instance, branch) on the same
path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import BackvoteSummary, Center, GridGeom, Pose, VoteBuffers, VoteParams, check
from .hostmath import percentile_plan
from .voting import (angle_tables, cos_threshold, device_index_tensor, idx_args, read_struct, sphere_lut, sphere_points,
                     stream_ptr, struct_tensor, to_device)


@dataclass
class VoteConfig:
    """The knobs eval.py's main() exposes for this path (eval.py:54-65) plus the per-category cfg keys."""
    res: float = 0.002
    up: Sequence[int] = (0, 1, 0)
    right: Sequence[int] = (1, 0, 0)
    front: Sequence[int] = (0, 0, 1)
    num_rots: int = 180
    angle_tol: float = 1.0
    backproj_ratio: float = 0.1
    imp_wt_margin: float = 0.01
    num_bins: int = 32
    loss_y_only: bool = False       # can / bottle / bowl: symmetric about y (eval.py:360-361)
    opt: bool = False               # online refinement (eval.py:62 `opt`, :319-355); the reference's default is True
    opt_iters: int = 100            # eval.py:326
    opt_lr: float = 1e-2            # eval.py:324

    @property
    def num_sphere(self) -> int:
        return int(4 * np.pi / (self.angle_tol / 180 * np.pi))   # eval.py:79


@dataclass
class PoseResult:
    R: np.ndarray
    t: np.ndarray
    scale: np.ndarray
    scale_norm: float
    loss: float
    kept: int
    bin_up: int
    bin_right: int
    count_up: float
    count_right: float
    status: int
    grid_cells: int = 0          # gx*gy*gz of the centre grid (what to grow the grid buffer to on CPPF_STATUS_GRID_OVERFLOW)

    @property
    def RT(self) -> np.ndarray:
        """4x4 with R*||scale|| and t, the layout eval.py:369-371 stores in pred_RTs."""
        out = np.eye(4)
        out[:3, :3] = self.R * self.scale_norm
        out[:3, 3] = self.t
        return out

    @property
    def unit_scale(self) -> np.ndarray:
        return self.scale / self.scale_norm


class PoseVoter:
    """Owns the device buffers of one in-flight (instance, branch) vote and replays the kernel chain.

    Buffers are sized for `max_tuples` / `max_points` / `grid_capacity` once and reused, so a steady
    stream of instances performs no allocation (and the chain is CUDA-graph capturable).
    """

    def __init__(self, max_tuples: int = 50000, max_points: int = 50000, grid_capacity: int = 1 << 23, device=None):
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_tuples = int(max_tuples)
        self.max_points = int(max_points)
        d = self.device
        T = self.max_tuples
        self.grid = torch.empty(int(grid_capacity), dtype=torch.int32, device=d)
        self.geom = struct_tensor(GridGeom, d)
        self.center = struct_tensor(Center, d)
        self.summary = struct_tensor(BackvoteSummary, d)
        self.pose = struct_tensor(Pose, d)
        self.status = torch.zeros(1, dtype=torch.int32, device=d)
        self.bins = torch.empty((T, 6), dtype=torch.uint8, device=d)
        self.targets_tr = torch.empty((T, 2), dtype=torch.float32, device=d)
        self.targets_rot = torch.empty((T, 3), dtype=torch.float32, device=d)
        self.errs = torch.empty(T, dtype=torch.float32, device=d)
        self.keep = torch.empty(T, dtype=torch.uint8, device=d)
        self.kept_list = torch.empty(T, dtype=torch.int32, device=d)
        self.imp = torch.empty(self.max_points, dtype=torch.int32, device=d)
        self.counts = torch.empty((2, 1), dtype=torch.float64, device=d)
        self.ws_backvote = torch.empty(int(self.lib.cppf_backvote_workspace_bytes(T, self.max_points)), dtype=torch.uint8, device=d)
        self.ws_pose = torch.empty(int(self.lib.cppf_pose_workspace_bytes(T)), dtype=torch.uint8, device=d)
        self.launches = 0   # kernels + memset nodes enqueued by the last vote() call
        self._params_cache = {}
        self._buffers = None

    def _vote_buffers(self, S: int) -> VoteBuffers:
        if self.counts.shape != (2, S):
            self.counts = torch.empty((2, S), dtype=torch.float64, device=self.device)
            self._buffers = None
        key = (self.grid.data_ptr(), self.counts.data_ptr(), self.errs.data_ptr())
        if self._buffers is None or self._buffers[0] != key:
            b = VoteBuffers(grid=self.grid.data_ptr(), grid_capacity=self.grid.numel(), geom=self.geom.data_ptr(),
                            center=self.center.data_ptr(), summary=self.summary.data_ptr(), status=self.status.data_ptr(),
                            targets_tr=self.targets_tr.data_ptr(), targets_rot=self.targets_rot.data_ptr(),
                            errs=self.errs.data_ptr(), keep=self.keep.data_ptr(), kept_list=self.kept_list.data_ptr(),
                            imp=self.imp.data_ptr(), counts=self.counts.data_ptr(), ws_backvote=self.ws_backvote.data_ptr(),
                            ws_backvote_bytes=self.ws_backvote.numel(), ws_pose=self.ws_pose.data_ptr(),
                            ws_pose_bytes=self.ws_pose.numel())
            self._buffers = (key, b)
        return self._buffers[1]

    def _vote_params(self, cfg: "VoteConfig", T: int) -> VoteParams:
        key = (cfg.res, tuple(cfg.up), tuple(cfg.right), tuple(cfg.front), cfg.num_rots, cfg.angle_tol, cfg.backproj_ratio,
               cfg.imp_wt_margin, cfg.num_bins, cfg.loss_y_only, T, cfg.opt, cfg.opt_iters, cfg.opt_lr)
        p = self._params_cache.get(key)
        if p is None:
            S = cfg.num_sphere
            thr = cos_threshold(cfg.angle_tol)
            ct, st = angle_tables(int(cfg.num_rots), self.device)
            sphere = sphere_points(S, self.device)
            lut, lut_g = sphere_lut(S, thr, self.device)
            rank_lo, gamma = percentile_plan(T, cfg.backproj_ratio)
            p = VoteParams(res=float(cfg.res), num_rots=int(cfg.num_rots), num_bins=int(cfg.num_bins), sphere_bins=S, cos_thr=thr,
                           band=self.lib.cppf_sphere_band(S, thr), lut_g=lut_g,
                           up_loc=int(np.where(np.asarray(cfg.up))[0][0]), right_loc=int(np.where(np.asarray(cfg.right))[0][0]),
                           loss_y_only=int(cfg.loss_y_only), lut=None if lut is None else lut.data_ptr(), cos_tab=ct.data_ptr(),
                           sin_tab=st.data_ptr(), sphere=sphere.data_ptr(),
                           axes=_lib.axes_array(cfg.up, cfg.front, cfg.right),   # call-site order, eval.py:237-240
                           imp_margin=float(cfg.imp_wt_margin), rank_lo=int(rank_lo), gamma=float(gamma),
                           refine_iters=int(cfg.opt_iters) if cfg.opt else 0, refine_lr=float(cfg.opt_lr))
            self._params_cache[key] = p
        return p

    def vote_bins(self, pc: torch.Tensor, idx: torch.Tensor, cfg: "VoteConfig", bins: torch.Tensor, pred_scales, scale_override,
                  cells_hint: Optional[int], pose_out: Optional[torch.Tensor]):
        """The whole chain from the drawn bins to the pose record in ONE host call (cppf_vote_chain): device tensors only,
        same kernels and buffers as vote()."""
        T, N = idx.shape[0], pc.shape[0]
        self._ensure(T, N, cells_hint)
        ip, i64, istr = idx_args(idx)
        params, bufs = self._vote_params(cfg, T), self._vote_buffers(cfg.num_sphere)
        check(self.lib.cppf_vote_chain(pc.data_ptr(), N, ip, i64, istr, T, bins.data_ptr(),
                                       None if pred_scales is None else pred_scales.data_ptr(),
                                       None if scale_override is None else scale_override.data_ptr(), int(cells_hint or 0),
                                       C.byref(params), C.byref(bufs), (self.pose if pose_out is None else pose_out).data_ptr(),
                                       stream_ptr()), "cppf_vote_chain")
        # kernels + memset nodes of the chain: the back-vote selection is one launch up to 2^17 tuples, seven above
        self.launches = (18 if T <= (1 << 17) else 24) + (1 if cfg.opt else 0)
        self._live = (pc, idx, bins, pred_scales, scale_override)
        self._T = T
        return self

    # -- helpers -------------------------------------------------------------------------------------
    def _ensure(self, T: int, N: int, cells_hint: Optional[int]):
        if T > self.max_tuples or N > self.max_points:
            self.__init__(max(T, self.max_tuples), max(N, self.max_points), self.grid.numel(), self.device)
        if cells_hint is not None and cells_hint > self.grid.numel():
            self.grid = torch.empty(int(cells_hint), dtype=torch.int32, device=self.device)

    @staticmethod
    def grid_cells_on_host(pc_host: np.ndarray, res: float) -> int:
        """gx*gy*gz with the same float32 arithmetic as the bounds kernel (train_dino.py:172-173)."""
        pc32 = np.asarray(pc_host, dtype=np.float32)
        ext = pc32.max(0) - pc32.min(0)
        g = (ext / np.float32(res)).astype(np.int64) + 1
        return int(g.prod())

    # -- the chain -----------------------------------------------------------------------------------
    def vote(self, pc, point_idxs_all, cfg: VoteConfig, pred_scales=None, *, bins=None, logits=None, u01=None,
             seed: int = 0, scale_override=None, cells_hint: Optional[int] = None, pose_out: Optional[torch.Tensor] = None):
        """Enqueues the whole chain on the current stream; returns self (read with .result()).

        Exactly one of `bins` (uint8 [T,6] injected multinomial draws) or `logits` (f32 [T,6,num_bins])
        must be given.  With logits, `u01` (f32 [T,6]) injects the uniforms, else a counter-based
        generator keyed by `seed` is used.  `pred_scales` f32 [T,3] is the scale head output;
        `scale_override` (3 floats, device or host) reproduces the reference's reuse of the DINO-branch
        scale in the SHOT branch (eval.py:308-310).  `pose_out` (uint8 CUDA tensor of sizeof(cppf_pose)
        bytes) receives the pose record instead of the voter's own slot, so that many votes can be queued
        before anything is read back.
        """
        lib = self.lib
        if isinstance(pc, np.ndarray) and cells_hint is None:
            cells_hint = self.grid_cells_on_host(pc, cfg.res)
        pc = to_device(pc, torch.float32, self.device)
        idx = device_index_tensor(point_idxs_all, self.device)
        T, N = idx.shape[0], pc.shape[0]
        if (bins is None) == (logits is None):
            raise ValueError("give exactly one of bins / logits")
        self._ensure(T, N, cells_hint)
        if logits is not None:                         # decode (eval.py:225-229): softmax + one multinomial draw
            lg = to_device(logits, torch.float32, self.device)
            u = None if u01 is None else to_device(u01, torch.float32, self.device)
            check(lib.cppf_sample_bins(lg.data_ptr(), T, cfg.num_bins, None if u is None else u.data_ptr(), int(seed),
                                       self.bins.data_ptr(), stream_ptr()), "cppf_sample_bins")
            bins_t = self.bins[:T]
        else:
            bins_t = to_device(bins, torch.uint8, self.device)
        so = None
        if scale_override is not None:   # the reference reuses the DINO-branch scale in the SHOT branch (eval.py:308-310)
            so = to_device(np.asarray(scale_override, dtype=np.float32) if not isinstance(scale_override, torch.Tensor)
                           else scale_override, torch.float32, self.device)
        ps = None if pred_scales is None else to_device(pred_scales, torch.float32, self.device)
        if ps is None and so is None:
            raise ValueError("pred_scales or scale_override is required")
        self.vote_bins(pc, idx, cfg, bins_t, ps, so, cells_hint, pose_out)
        self.launches += 0 if logits is None else 1
        return self

    @staticmethod
    def scale_ptr_of(pose_tensor: torch.Tensor) -> torch.Tensor:
        """View of the 3 scale floats inside a device pose record (to chain as `scale_override`)."""
        off = Pose.scale.offset
        return pose_tensor[off:off + 12].view(torch.float32)

    @staticmethod
    def parse(pose_bytes, extra_status: int = 0) -> "PoseResult":
        p = Pose.from_buffer_copy(bytes(pose_bytes))
        return PoseResult(R=np.array(list(p.R)).reshape(3, 3), t=np.array(list(p.t)), scale=np.array(list(p.scale), np.float32),
                          scale_norm=float(p.scale_norm), loss=float(p.loss), kept=int(p.kept), bin_up=int(p.bin_up),
                          bin_right=int(p.bin_right), count_up=float(p.count_up), count_right=float(p.count_right),
                          status=int(p.status) | int(extra_status), grid_cells=int(p.grid_cells))

    def result(self) -> PoseResult:
        """Synchronises the current stream and reads the pose record (152 bytes D2H)."""
        p = read_struct(self.pose, Pose)
        status = int(self.status.item()) | int(p.status)
        return PoseResult(R=np.array(list(p.R)).reshape(3, 3), t=np.array(list(p.t)), scale=np.array(list(p.scale), np.float32),
                          scale_norm=float(p.scale_norm), loss=float(p.loss), kept=int(p.kept), bin_up=int(p.bin_up),
                          bin_right=int(p.bin_right), count_up=float(p.count_up), count_right=float(p.count_right),
                          status=status, grid_cells=int(p.grid_cells))

    # -- debugging / parity access -------------------------------------------------------------------
    def intermediates(self) -> dict:
        """Host copies of every intermediate of the last vote() (for the parity tests)."""
        T = self._T
        geom = read_struct(self.geom, GridGeom)
        shape = tuple(int(g) for g in geom.grid_res)
        cells = int(geom.cells)
        center = read_struct(self.center, Center)
        summ = read_struct(self.summary, BackvoteSummary)
        kept = int(summ.kept)
        return dict(grid=self.grid[:cells].cpu().numpy().astype(np.int64).reshape(shape),
                    T_est=np.array(list(center.world)), targets_tr=self.targets_tr[:T].cpu().numpy(),
                    targets_rot=self.targets_rot[:T].cpu().numpy(), back_errs=self.errs[:T].cpu().numpy(),
                    thr=float(summ.threshold), pairs_mask=self.keep[:T].cpu().numpy().astype(bool),
                    kept_list=np.sort(self.kept_list[:kept].cpu().numpy()), imp=self.imp.cpu().numpy(),
                    imp_max=int(summ.imp_max), counts_up=self.counts[0].cpu().numpy(), counts_right=self.counts[1].cpu().numpy(),
                    bins=self.bins[:T].cpu().numpy())
