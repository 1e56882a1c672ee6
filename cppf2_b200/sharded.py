"""Tuple-sharded pose voting: the T tuples of one (instance, branch) split over the ranks of a process group.

SURVEY.md section 8e / BASELINE config 4.  Every tuple's votes are independent given the cloud (the grid
corners are functions of `pc` only, train_dino.py:172-173), so rank r takes the contiguous block
[r*T/g, (r+1)*T/g) of `point_idxs_all` with the whole cloud resident, and the path has three exchange steps:

  1. centre grid      all_reduce(SUM) of the uint32 counters   -- integer sum, order independent => bit-exact
  2. back-vote data   all_gather of the per-tuple float32 errors, tuple indices, draws, scales and rotation
                      targets (54 B/tuple); the exact percentile selection, kept list and importance counts
                      (eval.py:257-275) are then computed replicated and are identical on every rank
  3. sphere bins      all_reduce(SUM) of the 2 x 720 float64 rotation bins, every rank having voted the kept
                      pairs congruent to its rank modulo g

The expensive stages (heads, multinomial decode, targets, the T*R centre votes, the M*R rotation candidates)
are sharded; selection and pose assembly are cheap and replicated, so all ranks end with the same pose.

The orchestration is written against a small stage interface so that the collectives can be exercised on CPU
(`gloo`, world_size 2) with a test-side implementation of the stages; `CudaStages` is the product
implementation (libcppf_b200 kernels on the current stream, NCCL collectives) and is the only one this
package contains -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import BackvoteSummary, Center, GridGeom, Pose, check
from .hostmath import percentile_plan
from .pipeline import PoseResult, PoseVoter, VoteConfig
from .voting import (angle_tables, cos_threshold, device_index_tensor, idx_args, read_struct, sphere_lut, sphere_points,
                     stream_ptr, to_device)


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous block of rank `rank`; requires equal blocks (T = 50 000 divides by 1, 2, 4, 8)."""
    if n_items % world != 0:
        raise ValueError(f"{n_items} tuples do not split evenly over {world} ranks")
    per = n_items // world
    return rank * per, (rank + 1) * per


class ShardedVote:
    """Backend-agnostic orchestration of one sharded (instance, branch) vote."""

    def __init__(self, stages, group=None):
        self.stages = stages
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # -- collectives -----------------------------------------------------------------------------------
    def _all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def _all_gather(self, local: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return local
        out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out

    # -- the chain -------------------------------------------------------------------------------------
    def vote(self, pc, idx_local, cfg: VoteConfig, pred_scales_local, bins_local, scale_override=None):
        """`idx_local` [T/g,5], `bins_local` u8 [T/g,6], `pred_scales_local` f32 [T/g,3]: this rank's block.
        Returns whatever `stages.finalize` returns (a PoseResult for CudaStages)."""
        st = self.stages
        tr_l, rot_l = st.decode_targets(pc, idx_local, bins_local, cfg)
        grid = st.vote_center(pc, idx_local, tr_l, cfg)                 # this rank's partial grid
        self._all_reduce(grid)                                          # exchange step 1
        st.argmax(grid, cfg)
        errs_l = st.errors(pc, idx_local, tr_l)
        errs = self._all_gather(errs_l)                                 # exchange step 2
        idx = self._all_gather(idx_local)
        bins = self._all_gather(bins_local)
        scales = self._all_gather(pred_scales_local)
        rot = self._all_gather(rot_l)
        st.select_and_mask(errs, idx, pc, cfg)
        counts = st.rotation_counts(pc, idx, rot, cfg, self.rank, self.world)
        self._all_reduce(counts)                                        # exchange step 3
        return st.finalize(pc, idx, bins, scales, counts, cfg, scale_override)


class CudaStages:
    """The stages as libcppf_b200 kernels; buffers sized for the GLOBAL tuple count, reused across votes."""

    def __init__(self, max_tuples_global: int = 50000, max_points: int = 50000, grid_capacity: int = 1 << 23, device=None):
        self.v = PoseVoter(max_tuples_global, max_points, grid_capacity, device)
        self.lib = self.v.lib
        self.device = self.v.device

    def decode_targets(self, pc, idx_local, bins_local, cfg):
        v, lib = self.v, self.lib
        self.pc = to_device(pc, torch.float32, self.device)
        idx = device_index_tensor(idx_local, self.device)
        bins = to_device(bins_local, torch.uint8, self.device)
        T = idx.shape[0]
        v._ensure(T, self.pc.shape[0], None)
        ip, i64, istr = idx_args(idx)
        self.axes = _lib.axes_array(cfg.up, cfg.front, cfg.right)     # call-site order, eval.py:237-240
        tr, rot = v.targets_tr[:T], v.targets_rot[:T]
        check(lib.cppf_decode_targets(self.pc.data_ptr(), ip, i64, istr, bins.data_ptr(), T, cfg.num_bins, self.axes,
                                      tr.data_ptr(), rot.data_ptr(), None, None, stream_ptr()), "cppf_decode_targets")
        return tr, rot

    def vote_center(self, pc, idx_local, tr_l, cfg):
        v, lib, s = self.v, self.lib, stream_ptr()
        idx = device_index_tensor(idx_local, self.device)
        ip, i64, istr = idx_args(idx)
        N = self.pc.shape[0]
        ct, st = angle_tables(cfg.num_rots, self.device)
        check(lib.cppf_cloud_bounds(self.pc.data_ptr(), N, float(cfg.res), v.geom.data_ptr(), s), "cppf_cloud_bounds")
        geom = read_struct(v.geom, GridGeom)          # the all-reduce needs the cell count on the host
        self.cells = int(geom.cells)
        if self.cells > v.grid.numel():
            v.grid = torch.empty(self.cells, dtype=torch.int32, device=self.device)
        v.status.zero_()
        check(lib.cppf_vote_center(self.pc.data_ptr(), N, ip, i64, istr, tr_l.data_ptr(), idx.shape[0], ct.data_ptr(),
                                   st.data_ptr(), int(cfg.num_rots), v.geom.data_ptr(), v.grid.data_ptr(), v.grid.numel(),
                                   self.cells, 0, v.status.data_ptr(), s), "cppf_vote_center")
        return v.grid[:self.cells]

    def argmax(self, grid, cfg):
        v = self.v
        check(self.lib.cppf_grid_argmax(grid.data_ptr(), v.geom.data_ptr(), float(cfg.res), v.center.data_ptr(), stream_ptr()),
              "cppf_grid_argmax")

    def errors(self, pc, idx_local, tr_l):
        v = self.v
        idx = device_index_tensor(idx_local, self.device)
        ip, i64, istr = idx_args(idx)
        T = idx.shape[0]
        errs = v.errs[:T]
        check(self.lib.cppf_backvote_errors(self.pc.data_ptr(), ip, i64, istr, tr_l.data_ptr(), T, v.center.data_ptr(),
                                            errs.data_ptr(), stream_ptr()), "cppf_backvote_errors")
        return errs

    def select_and_mask(self, errs, idx, pc, cfg):
        v, lib, s = self.v, self.lib, stream_ptr()
        T, N = errs.shape[0], self.pc.shape[0]
        v._ensure(T, N, None)
        ip, i64, istr = idx_args(idx)
        rank_lo, gamma = percentile_plan(T, cfg.backproj_ratio)
        check(lib.cppf_backvote_select(errs.data_ptr(), T, rank_lo, float(gamma), v.summary.data_ptr(),
                                       v.ws_backvote.data_ptr(), v.ws_backvote.numel(), s), "cppf_backvote_select")
        check(lib.cppf_backvote_mask(errs.data_ptr(), ip, i64, istr, T, N, v.summary.data_ptr(), v.keep.data_ptr(),
                                     v.kept_list.data_ptr(), v.imp.data_ptr(), 1, s), "cppf_backvote_mask")
        check(lib.cppf_backvote_imp_max(v.imp.data_ptr(), N, v.summary.data_ptr(), s), "cppf_backvote_imp_max")
        self._errs, self._T = errs, T

    def rotation_counts(self, pc, idx, rot, cfg, part, n_parts):
        v, lib, s = self.v, self.lib, stream_ptr()
        S = cfg.num_sphere
        if v.counts.shape != (2, S):
            v.counts = torch.empty((2, S), dtype=torch.float64, device=self.device)
        v.counts.zero_()
        ip, i64, istr = idx_args(idx)
        ct, st = angle_tables(cfg.num_rots, self.device)
        sphere = sphere_points(S, self.device)
        thr = cos_threshold(cfg.angle_tol)
        cols = (C.c_int * 2)(0, 2)
        lut, lut_g = sphere_lut(S, thr, self.device)
        kept_count_ptr = v.summary.data_ptr() + BackvoteSummary.kept.offset
        check(lib.cppf_rotation_hist_part(self.pc.data_ptr(), ip, i64, istr, rot.data_ptr(), 3, cols, 2,
                                          v.kept_list.data_ptr(), kept_count_ptr, idx.shape[0], v.imp.data_ptr(),
                                          v.summary.data_ptr(), float(cfg.imp_wt_margin), ct.data_ptr(), st.data_ptr(),
                                          int(cfg.num_rots), sphere.data_ptr(), S, thr, lib.cppf_sphere_band(S, thr),
                                          None if lut is None else lut.data_ptr(), lut_g, v.counts.data_ptr(), int(part), int(n_parts), s), "cppf_rotation_hist_part")
        return v.counts

    def finalize(self, pc, idx, bins, scales, counts, cfg, scale_override=None) -> PoseResult:
        v, lib, s = self.v, self.lib, stream_ptr()
        S = cfg.num_sphere
        ip, i64, istr = idx_args(idx)
        so = None
        if scale_override is not None:
            so = to_device(np.asarray(scale_override, dtype=np.float32) if not isinstance(scale_override, torch.Tensor)
                           else scale_override, torch.float32, self.device)
        up_loc = int(np.where(np.asarray(cfg.up))[0][0])
        right_loc = int(np.where(np.asarray(cfg.right))[0][0])
        sphere = sphere_points(S, self.device)
        check(lib.cppf_pose_finalize(self.pc.data_ptr(), ip, i64, istr, bins.data_ptr(), cfg.num_bins, scales.data_ptr(),
                                     v.kept_list.data_ptr(), v.summary.data_ptr(), counts.data_ptr(), sphere.data_ptr(), S,
                                     v.center.data_ptr(), up_loc, right_loc, int(cfg.loss_y_only),
                                     None if so is None else so.data_ptr(), v.pose.data_ptr(), v.ws_pose.data_ptr(),
                                     v.ws_pose.numel(), s), "cppf_pose_finalize")
        v._T = idx.shape[0]
        self._live = (idx, bins, scales, so)
        return v.result()

    def intermediates(self) -> dict:
        """Host copies of the replicated intermediates (grid after the all-reduce, kept set, bins)."""
        v = self.v
        geom = read_struct(v.geom, GridGeom)
        shape = tuple(int(g) for g in geom.grid_res)
        summ = read_struct(v.summary, BackvoteSummary)
        center = read_struct(v.center, Center)
        T = self._T
        return dict(grid=v.grid[:int(geom.cells)].cpu().numpy().astype(np.int64).reshape(shape),
                    T_est=np.array(list(center.world)), pairs_mask=v.keep[:T].cpu().numpy().astype(bool),
                    back_errs=self._errs.cpu().numpy(), imp=v.imp.cpu().numpy(), imp_max=int(summ.imp_max),
                    counts_up=v.counts[0].cpu().numpy(), counts_right=v.counts[1].cpu().numpy())


class ShardedPoseVoter(ShardedVote):
    """Product form: CUDA stages + the process group's collectives (NCCL on the GPU box)."""

    def __init__(self, max_tuples_global: int = 50000, max_points: int = 50000, grid_capacity: int = 1 << 23, group=None,
                 device=None):
        super().__init__(CudaStages(max_tuples_global, max_points, grid_capacity, device), group)
