"""Tuple-sharded pose voting: the T tuples of one (instance, branch) split over the ranks of a process group.

SURVEY.md section 8e / BASELINE config 4 / the north star's multi-GPU design.  Every tuple's votes are independent given
the cloud (the grid corners are functions of `pc` only, train_dino.py:172-173), so rank r takes the contiguous block
[r*T/g, (r+1)*T/g) of `point_idxs_all` (eval.py:207) with the whole cloud resident and runs heads, decode, targets, centre
votes, back-vote errors, mask, rotation votes and the loss terms on ITS block only.  Nothing per-tuple is replicated and
nothing but the 4-byte back-vote error crosses the links per tuple.  The exchange steps, one per stage (eval.py:219-313):

  E1  centre grid        all_reduce(SUM) of the uint32 counters [cells]          integer sum => bit-exact for any g
  E2  back-vote errors   all_gather of the float32 errors (4 B/tuple)            every rank then selects the same exact
                                                                                 order statistic (np.percentile, eval.py:257)
  E3  importance + scale all_reduce(SUM) of one int32 buffer: per-point occurrence counts [n] (eval.py:260-266), the kept
                         count, and pass 0 of the scale median's radix histogram [3 x 65536] (eval.py:309)
  E4  sphere bins        the 2 x 720 float64 rotation bins (eval.py:277-293) and pass 1 of the int32 scale histogram: two dtypes,
                         so their byte images cross in ONE all_gather (0.8 MB per rank) and every rank adds them in rank order
                         (every per-CTA bin contribution is a multiple of 2^-32, so the float64 sums are exact below 2^21 per
                         bin and do not depend on g)
  E5  branch loss        all_reduce(SUM) of the sum of the clipped L1 terms (eval.py:358-363; the count follows from the
                         kept count E3 delivered)

Only E1 is on the critical path of the tuples/s metric; E3-E5 move < 1 MB and are latency-bound.  No host read-back
happens between a launch and a collective: the grid all-reduce is sized by the cell count the caller knows from the host
copy of the cloud (`cells_hint`), the rest by T and n.

The orchestration is written against a small stage interface so that the collectives can be exercised on CPU (`gloo`,
world_size 2) with a test-side implementation of the stages; `CudaStages` is the product implementation (libcppf_b200
kernels on the current stream, NCCL collectives) and is the only one this package contains -- there is no CPU fallback.
The online refinement (eval.py:319-355) is not sharded: `opt=True` raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import BackvoteSummary, Center, GridGeom, Pose, ScaleSelect, check
from .hostmath import percentile_plan
from .pipeline import PoseResult, PoseVoter, VoteConfig
from .voting import (angle_tables, cos_threshold, device_index_tensor, idx_args, read_struct, sphere_lut, sphere_points,
                     stream_ptr, struct_tensor, to_device)

SCALE_DIGITS = 1 << 16
_SPLITMIX_GAMMA = 0x9E3779B97F4A7C15


class Packed:
    """Payloads of different dtypes that live side by side in ONE byte buffer (`parts` are typed views into `buffer`, each
    starting at a 16-byte boundary): an exchange step moves the buffer as it is, without staging copies."""

    def __init__(self, buffer: torch.Tensor, parts):
        self.buffer, self.parts = buffer, list(parts)


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous block of rank `rank`; requires equal blocks (T = 50 000 divides by 1, 2, 4, 8)."""
    if n_items % world != 0:
        raise ValueError(f"{n_items} tuples do not split evenly over {world} ranks")
    per = n_items // world
    return rank * per, (rank + 1) * per


def shard_seed(seed: int, first_tuple: int) -> int:
    """Seed under which a rank whose block starts at global tuple `first_tuple` draws exactly the uniforms the unsharded
    call draws for those tuples: the generator is counter-based, u(seed, ctr) = mix(seed + G*(ctr+1)) with ctr = 6*tuple +
    coordinate (csrc/common.cuh uniform_from_counter), so shifting the counter by 6*first_tuple is a shift of the seed."""
    return (int(seed) + _SPLITMIX_GAMMA * 6 * int(first_tuple)) & 0xFFFFFFFFFFFFFFFF


class ShardedVote:
    """Backend-agnostic orchestration of one sharded (instance, branch) vote: stages + five exchange steps."""

    def __init__(self, stages, group=None):
        self.stages = stages
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_collectives = 0           # collectives issued by the last vote()
        self.timing = None               # optional list: (label, start event, end event) per exchange step (CUDA only)
        self.stage_marks = None          # optional list: (label, event) at the stage boundaries of vote() (CUDA only)

    def _mark(self, label: str):
        """Optional per-stage timing (CUDA only): `stage_marks` = list that receives (label, event) at the stage boundaries."""
        if self.stage_marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.stage_marks.append((label, e))

    # -- collectives -----------------------------------------------------------------------------------
    def _exchange(self, label: str, fn):
        if self.world == 1:
            return
        if self.timing is not None:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            self.timing.append((label, a, b))
        else:
            fn()
        self.n_collectives += 1

    def _all_reduce(self, label: str, tensors: List[torch.Tensor]):
        """One exchange step: the SUM over the ranks of every tensor of the list, in place.  One dtype: one all_reduce (a
        single flat buffer is what the stages hand over).  Mixed dtypes (E4: float64 sphere bins + int32 histogram): NCCL
        reduces one dtype per call, so the byte images travel in ONE all_gather instead and every rank adds the g images
        in rank order -- a fixed summation order, so the float64 bins are identical on every rank by construction."""
        packed = tensors if isinstance(tensors, Packed) else None
        tensors = [t for t in (packed.parts if packed is not None else tensors) if t is not None and t.numel() > 0]
        if not tensors or self.world == 1:
            return
        if len({t.dtype for t in tensors}) == 1:
            def run():
                for t in tensors:
                    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            if len(tensors) > 1:
                raise ValueError("same-dtype payloads of one exchange step must share one flat buffer")
            self._exchange(label, run)
            return
        segs, off = [], 0
        for t in tensors:
            nb = t.numel() * t.element_size()
            segs.append((off, nb))
            off += (nb + 15) // 16 * 16
        if packed is not None and packed.buffer.numel() == off:
            flat = packed.buffer                                  # the payloads already sit in one byte buffer
        else:
            flat = torch.zeros(off, dtype=torch.uint8, device=tensors[0].device)
            for t, (o, nb) in zip(tensors, segs):
                flat[o:o + nb].copy_(t.contiguous().reshape(-1).view(torch.uint8))
        gathered = torch.empty(self.world * off, dtype=torch.uint8, device=flat.device)
        self._exchange(label, lambda: dist.all_gather_into_tensor(gathered, flat, group=self.group))
        out = gathered.view(self.world, off)
        for t, (o, nb) in zip(tensors, segs):
            # rows of `out` start at multiples of 16 bytes: the typed view needs no copy; sum over the ranks in rank order
            parts = out[:, o:o + nb].view(t.dtype).reshape((self.world,) + tuple(t.shape))
            torch.sum(parts, 0, dtype=t.dtype, out=t)

    def _all_gather(self, label: str, local: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.world == 1:
            return local
        if out is None:
            out = torch.empty((self.world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        self._exchange(label, lambda: dist.all_gather_into_tensor(out, local.contiguous(), group=self.group))
        return out

    # -- the chain -------------------------------------------------------------------------------------
    def vote(self, pc, idx_local, cfg: VoteConfig, pred_scales_local, bins_local, scale_override=None, cells_hint=None,
             lazy: bool = False):
        """`idx_local` [T/g,>=2], `bins_local` u8 [T/g,6], `pred_scales_local` f32 [T/g,3]: this rank's block of one
        (instance, branch).  `scale_override` (3 floats) reproduces the SHOT branch's reuse of the DINO scale (eval.py:308).
        Returns the pose (identical on every rank); with `lazy` the un-synchronised handle of `stages.finish`, whose
        `.result()` reads it back -- a caller that streams votes resolves the handles later and the GPU never idles between
        two votes."""
        if getattr(cfg, "opt", False):
            raise NotImplementedError("the online refinement (eval.py:319-355) is not tuple-sharded; run it with opt=False")
        st = self.stages
        self.n_collectives = 0
        mark = self._mark
        mark("start")
        st.begin(pc, idx_local, bins_local, pred_scales_local, cfg, self.world, cells_hint)   # decode, targets, bounds
        grid = st.vote_center()                                          # this rank's partial grid, flat [cells]
        mark("decode+centre vote")
        self._all_reduce("E1 grid", [grid])
        st.argmax()                                                      # replicated: same grid everywhere
        errs_all = self._all_gather("E2 errors", st.errors(), st.errs_all_buffer(self.world))
        mark("E1+argmax+errors+E2")
        st.select(errs_all)                                              # replicated exact order statistic -> threshold
        own_scale = scale_override is None
        x_imp = st.mask_local(own_scale)                                 # int32 [n | kept | scale hist pass 0]
        mark("select+mask")
        self._all_reduce("E3 imp+scale0", [x_imp])
        st.after_mask(own_scale)                                         # imp_max; scale pick 0, local hist pass 1
        x_counts = st.rotation_counts()                                  # [counts f64 [2,S]] + [scale hist pass 1]
        mark("E3+rotation vote")
        self._all_reduce("E4 bins+scale1", x_counts)
        x_loss = st.pose_local(scale_override)                           # directions, scale, sum of the local loss terms
        self._all_reduce("E5 loss", [x_loss])
        mark("E4+pose+E5")
        out = st.finish()
        return out if lazy or not hasattr(out, "result") else out.result()


class CudaStages:
    """The stages as libcppf_b200 kernels on the current stream.  Per-tuple buffers are sized for this rank's BLOCK; only
    the gathered back-vote errors (4 B/tuple) are sized for the global tuple count."""

    def __init__(self, max_tuples_local: int = 50000, max_points: int = 50000, grid_capacity: int = 1 << 23, device=None):
        self.v = PoseVoter(max_tuples_local, max_points, grid_capacity, device)
        self.lib = self.v.lib
        self.device = self.v.device
        d = self.device
        self.sel = struct_tensor(ScaleSelect, d)
        self.scale_dev = torch.zeros(3, dtype=torch.float32, device=d)
        self.loss_x = torch.zeros(1, dtype=torch.float64, device=d)
        self.xbuf = None                 # int32 exchange buffer [n | kept | 3*65536]
        self.x4 = None                   # byte buffer of exchange E4: [sphere bins f64 [2,S] | scale histogram pass 1 i32 [3*65536]]
        self.hist1 = None
        self.errs_all = None
        self._cells = {}                 # grid cells per cloud (data_ptr, n, res) when no hint was given: one read-back, cached

    # -- stage 1: decode + targets + centre votes ------------------------------------------------------
    def begin(self, pc, idx_local, bins_local, scales_local, cfg, world, cells_hint=None):
        v, lib, s = self.v, self.lib, stream_ptr()
        if isinstance(pc, np.ndarray) and cells_hint is None:
            cells_hint = PoseVoter.grid_cells_on_host(pc, cfg.res)
        self.pc = to_device(pc, torch.float32, self.device)
        self.idx = device_index_tensor(idx_local, self.device)
        self.bins = to_device(bins_local, torch.uint8, self.device)
        self.scales = None if scales_local is None else to_device(scales_local, torch.float32, self.device)
        self.cfg, self.world = cfg, world
        T, N = self.idx.shape[0], self.pc.shape[0]
        self.T_local, self.N = T, N
        v._ensure(T, N, cells_hint)
        v._T = T
        ip, i64, istr = idx_args(self.idx)
        self.ip = (ip, i64, istr)
        self.axes = _lib.axes_array(cfg.up, cfg.front, cfg.right)     # call-site order, eval.py:237-240
        check(lib.cppf_decode_targets(self.pc.data_ptr(), ip, i64, istr, self.bins.data_ptr(), T, cfg.num_bins, self.axes,
                                      v.targets_tr.data_ptr(), v.targets_rot.data_ptr(), None, None, s), "cppf_decode_targets")
        check(lib.cppf_cloud_bounds(self.pc.data_ptr(), N, float(cfg.res), v.geom.data_ptr(), s), "cppf_cloud_bounds")
        if cells_hint is None:           # device cloud without a hint: one read-back per distinct cloud, then cached
            key = (self.pc.data_ptr(), N, float(cfg.res))
            if key not in self._cells:
                self._cells[key] = int(read_struct(v.geom, GridGeom).cells)
            cells_hint = self._cells[key]
            v._ensure(T, N, cells_hint)
        self.cells = int(cells_hint)
        need = N + 1 + 3 * SCALE_DIGITS
        if self.xbuf is None or self.xbuf.numel() < need:
            self.xbuf = torch.zeros(need, dtype=torch.int32, device=self.device)

    def vote_center(self):
        v, lib, s = self.v, self.lib, stream_ptr()
        cfg = self.cfg
        ip, i64, istr = self.ip
        ct, st = angle_tables(cfg.num_rots, self.device)
        v.status.zero_()
        check(lib.cppf_vote_center(self.pc.data_ptr(), self.N, ip, i64, istr, v.targets_tr.data_ptr(), self.T_local, ct.data_ptr(),
                                   st.data_ptr(), int(cfg.num_rots), v.geom.data_ptr(), v.grid.data_ptr(), v.grid.numel(),
                                   self.cells, 0, v.status.data_ptr(), s), "cppf_vote_center")
        return v.grid[:self.cells]

    def argmax(self):
        v = self.v
        check(self.lib.cppf_grid_argmax(v.grid.data_ptr(), v.grid.numel(), v.geom.data_ptr(), float(self.cfg.res),
                                        v.status.data_ptr(), v.center.data_ptr(), stream_ptr()), "cppf_grid_argmax")

    # -- stage 2: back-vote filter -----------------------------------------------------------------------
    def errors(self):
        v = self.v
        ip, i64, istr = self.ip
        check(self.lib.cppf_backvote_errors(self.pc.data_ptr(), ip, i64, istr, v.targets_tr.data_ptr(), self.T_local,
                                            v.center.data_ptr(), v.errs.data_ptr(), stream_ptr()), "cppf_backvote_errors")
        return v.errs[:self.T_local]

    def errs_all_buffer(self, world):
        T = self.T_local * world
        if world > 1 and (self.errs_all is None or self.errs_all.numel() != T):
            self.errs_all = torch.empty(T, dtype=torch.float32, device=self.device)
        return self.errs_all

    def select(self, errs_all):
        v, lib = self.v, self.lib
        T = errs_all.shape[0]
        rank_lo, gamma = percentile_plan(T, self.cfg.backproj_ratio)
        check(lib.cppf_backvote_select(errs_all.data_ptr(), T, rank_lo, float(gamma), v.summary.data_ptr(),
                                       v.ws_backvote.data_ptr(), v.ws_backvote.numel(), stream_ptr()), "cppf_backvote_select")
        self._errs_all = errs_all

    def _kept_ptr(self):
        return self.v.summary.data_ptr() + BackvoteSummary.kept.offset

    def mask_local(self, own_scale: bool):
        """keep / kept_list / occurrence counts of this rank's block; returns the int32 exchange buffer."""
        v, lib, s = self.v, self.lib, stream_ptr()
        N, T = self.N, self.T_local
        ip, i64, istr = self.ip
        x = self.xbuf[:N + 1 + (3 * SCALE_DIGITS if own_scale else 0)]
        imp = x[:N]
        # the kernel zeroes imp [n] and the kept counter (zero_outputs = 1)
        check(lib.cppf_backvote_mask(v.errs.data_ptr(), ip, i64, istr, T, N, v.summary.data_ptr(), v.keep.data_ptr(),
                                     v.kept_list.data_ptr(), imp.data_ptr(), 1, s), "cppf_backvote_mask")
        off = BackvoteSummary.kept.offset
        x[N:N + 1].copy_(v.summary[off:off + 4].view(torch.int32))          # local kept count (< 2^31), little endian low word
        if own_scale:
            self.sel.zero_()
            check(lib.cppf_scale_median_hist(self.scales.data_ptr(), v.kept_list.data_ptr(), self._kept_ptr(), T, 0,
                                             self.sel.data_ptr(), x[N + 1:].data_ptr(), s), "cppf_scale_median_hist")
        self._x = x
        return x

    def after_mask(self, own_scale: bool):
        v, lib, s = self.v, self.lib, stream_ptr()
        N, x = self.N, self._x
        check(lib.cppf_backvote_imp_max(x.data_ptr(), N, v.summary.data_ptr(), s), "cppf_backvote_imp_max")
        self._hist1 = None
        if own_scale:
            self._e4_buffers(self.cfg.num_sphere)
            check(lib.cppf_scale_median_pick(x[N + 1:].data_ptr(), x[N:].data_ptr(), 0, self.sel.data_ptr(), None, s),
                  "cppf_scale_median_pick")
            check(lib.cppf_scale_median_hist(self.scales.data_ptr(), v.kept_list.data_ptr(), self._kept_ptr(), self.T_local, 1,
                                             self.sel.data_ptr(), self.hist1.data_ptr(), s), "cppf_scale_median_hist")
            self._hist1 = self.hist1

    # -- stage 3: rotation votes, pose, loss ---------------------------------------------------------------
    def _e4_buffers(self, S: int):
        """Sphere bins and the pass-1 scale histogram as typed views of one byte buffer (exchange E4 moves it as it is)."""
        cb = 2 * S * 8
        cb_pad = (cb + 15) // 16 * 16
        if self.x4 is None or self.x4.numel() != cb_pad + 3 * SCALE_DIGITS * 4:
            self.x4 = torch.zeros(cb_pad + 3 * SCALE_DIGITS * 4, dtype=torch.uint8, device=self.device)
            self.v.counts = self.x4[:cb].view(torch.float64).view(2, S)
            self.v._buffers = None
            self.hist1 = self.x4[cb_pad:].view(torch.int32)

    def rotation_counts(self):
        v, lib, s = self.v, self.lib, stream_ptr()
        cfg = self.cfg
        S = cfg.num_sphere
        self._e4_buffers(S)
        v.counts.zero_()
        ip, i64, istr = self.ip
        ct, st = angle_tables(cfg.num_rots, self.device)
        sphere = sphere_points(S, self.device)
        thr = cos_threshold(cfg.angle_tol)
        cols = (C.c_int * 2)(0, 2)
        lut, lut_g = sphere_lut(S, thr, self.device)
        # this rank's kept tuples (its own kept_list) with the GLOBAL importance counts of the exchange buffer
        check(lib.cppf_rotation_hist(self.pc.data_ptr(), ip, i64, istr, v.targets_rot.data_ptr(), 3, cols, 2,
                                     v.kept_list.data_ptr(), self._kept_ptr(), self.T_local, self._x.data_ptr(),
                                     v.summary.data_ptr(), float(cfg.imp_wt_margin), ct.data_ptr(), st.data_ptr(),
                                     int(cfg.num_rots), sphere.data_ptr(), S, thr, lib.cppf_sphere_band(S, thr),
                                     None if lut is None else lut.data_ptr(), lut_g, v.counts.data_ptr(), s), "cppf_rotation_hist")
        if self._hist1 is None:
            return [v.counts]
        return Packed(self.x4, [v.counts, self._hist1])

    def pose_local(self, scale_override=None):
        v, lib, s = self.v, self.lib, stream_ptr()
        cfg = self.cfg
        S = cfg.num_sphere
        N, x = self.N, self._x
        if scale_override is None:
            check(lib.cppf_scale_median_pick(self.hist1.data_ptr(), x[N:].data_ptr(), 1, self.sel.data_ptr(),
                                             self.scale_dev.data_ptr(), s), "cppf_scale_median_pick")
            so = self.scale_dev
        else:
            so = to_device(np.asarray(scale_override, dtype=np.float32) if not isinstance(scale_override, torch.Tensor)
                           else scale_override, torch.float32, self.device)
        self._so = so
        ip, i64, istr = self.ip
        up_loc = int(np.where(np.asarray(cfg.up))[0][0])
        right_loc = int(np.where(np.asarray(cfg.right))[0][0])
        sphere = sphere_points(S, self.device)
        check(lib.cppf_pose_finalize(self.pc.data_ptr(), ip, i64, istr, self.bins.data_ptr(), cfg.num_bins, None,
                                     v.kept_list.data_ptr(), v.summary.data_ptr(), v.counts.data_ptr(), sphere.data_ptr(), S,
                                     v.center.data_ptr(), up_loc, right_loc, int(cfg.loss_y_only), so.data_ptr(),
                                     v.pose.data_ptr(), v.ws_pose.data_ptr(), v.ws_pose.numel(), s), "cppf_pose_finalize")
        # this rank's loss terms: the kernel left their float64 sum in the head of its scratch (PoseScratch.loss_sum); the
        # count of the mean is 2 * kept * (1 | 3) with the global kept count
        self.loss_x.copy_(v.ws_pose[:8].view(torch.float64))
        return self.loss_x

    def finish(self) -> "PendingShardedPose":
        """Packs the pose record, the status word, the all-reduced loss sum and the global kept count into ONE fresh device
        tensor; `.result()` reads it back (the only host synchronisation of a vote).  A caller that streams votes keeps
        the handle and resolves it later: the next vote's kernels are queued without waiting for this one."""
        v = self.v
        N = self.N
        pack = torch.cat([v.pose, v.status.view(torch.uint8), self.loss_x.view(torch.uint8), self._x[N:N + 1].view(torch.uint8)])
        return PendingShardedPose(pack, v.pose.numel(), bool(self.cfg.loss_y_only))

    def intermediates(self) -> dict:
        """Host copies: the all-reduced grid and importance counts, this rank's block of the kept mask and errors, the
        all-reduced sphere bins."""
        v = self.v
        geom = read_struct(v.geom, GridGeom)
        shape = tuple(int(g) for g in geom.grid_res)
        summ = read_struct(v.summary, BackvoteSummary)
        center = read_struct(v.center, Center)
        T = self.T_local
        return dict(grid=v.grid[:int(geom.cells)].cpu().numpy().astype(np.int64).reshape(shape),
                    T_est=np.array(list(center.world)), pairs_mask_local=v.keep[:T].cpu().numpy().astype(bool),
                    back_errs_local=v.errs[:T].cpu().numpy(), thr=float(summ.threshold),
                    imp=self._x[:self.N].cpu().numpy(), imp_max=int(summ.imp_max), kept=int(self._x[self.N].item()),
                    counts_up=v.counts[0].cpu().numpy(), counts_right=v.counts[1].cpu().numpy(),
                    scale=self._so.cpu().numpy())


class PendingShardedPose:
    """Result handle of one sharded vote (CudaStages.finish)."""

    def __init__(self, pack: torch.Tensor, pose_bytes: int, loss_y_only: bool):
        self._pack, self._nb, self._y = pack, pose_bytes, loss_y_only
        self._out = None

    def result(self) -> PoseResult:
        if self._out is None:
            pack, nb = self._pack.cpu().numpy(), self._nb
            status = int(pack[nb:nb + 4].view(np.int32)[0])
            loss_sum = float(pack[nb + 4:nb + 12].view(np.float64)[0])
            kept = int(pack[nb + 12:nb + 16].view(np.int32)[0])
            r = PoseVoter.parse(pack[:nb].tobytes(), extra_status=status)
            r.kept = kept
            cnt = 2.0 * kept * (1.0 if self._y else 3.0)
            r.loss = loss_sum / cnt if cnt > 0 else float("inf")
            if kept > 0:
                r.status &= ~_lib.CPPF_STATUS_EMPTY
            self._out, self._pack = r, None
        return self._out


class ShardedPoseVoter(ShardedVote):
    """Product form: CUDA stages + the process group's collectives (NCCL on the GPU box)."""

    def __init__(self, max_tuples_local: int = 50000, max_points: int = 50000, grid_capacity: int = 1 << 23, group=None,
                 device=None):
        super().__init__(CudaStages(max_tuples_local, max_points, grid_capacity, device), group)

    def gather_mask(self) -> np.ndarray:
        """The global kept mask [T] on the host (parity checks only: the path itself never gathers it)."""
        local = self.stages.v.keep[:self.stages.T_local]
        return self._all_gather("mask (debug)", local).cpu().numpy().astype(bool)

    def vote_with_heads(self, model, pc, idx_local, cfg: VoteConfig, first_tuple: int, seed: int = 0, desc=None, shot_feat=None,
                        normal=None, scale_override=None, cells_hint=None, lazy: bool = False):
        """eval.py:219-313 for this rank's block of tuples, heads included: `model` (BeyondCPPFSHOT with shot_feat / normal,
        or BeyondCPPFDINO with desc; precision 1) runs on idx_local with the decode fused in, drawing for its tuples the
        uniforms the unsharded call would draw for them (shard_seed), then the sharded vote."""
        s = shard_seed(seed, first_tuple)
        self._mark("heads begin")
        if model.branch == "dino":
            bins, scales = model.forward_sampled(pc, desc, idx_local, seed=s)
        else:
            bins, scales = model.forward_sampled(pc, idx_local, shot_feat, normal, seed=s)
        self._live = (bins, scales)
        return self.vote(pc, idx_local, cfg, scales, bins, scale_override=scale_override, cells_hint=cells_hint, lazy=lazy)
