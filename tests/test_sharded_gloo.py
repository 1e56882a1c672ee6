"""World-size-2 and -3 `gloo` tests of the tuple-sharded vote orchestration (cppf2_b200/sharded.py) on CPU.

The collectives and the sharding arithmetic are the product's; the stage kernels need a GPU, so this test
plugs an oracle-backed implementation of the stage interface into `ShardedVote` (test infrastructure only)
and checks that two ranks, each holding half of the tuples and never seeing the other half (only the five exchange
steps cross: partial grid, 4-byte errors, importance counts + scale histogram, sphere bins + scale histogram, loss sum),
end with exactly the single-process result: bit-identical centre grid, voted centre, threshold, kept set, importance
counts, scale median and rotation bins' arg-max; float bins within 1e-9.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class OracleStages:
    """The stage interface of cppf2_b200.sharded.ShardedVote on top of the CPU oracle (torch CPU tensors cross the
    collectives).  Every per-tuple stage sees this rank's block only, exactly like CudaStages."""

    def __init__(self):
        from oracle import cpu
        self.o = cpu
        self.out = {}

    def begin(self, pc, idx_local, bins_local, scales_local, cfg, world, cells_hint=None):
        o = self.o
        self.cfg = cfg
        self.pc = np.ascontiguousarray(pc, dtype=np.float32)
        self.idx = idx_local.numpy()
        self.scales = scales_local.numpy()
        self.pred_l, scaled, _ = o.decode_pairs(self.pc, self.idx, bins_local.numpy(), cfg.num_bins)
        self.tr, self.rot = o.generate_target_pairs(scaled, cfg.up, cfg.front, cfg.right)

    def vote_center(self):
        grid, _ = self.o.vote_center(self.pc, self.tr, self.cfg.res, self.idx[:, :2], self.cfg.num_rots)
        self.shape = grid.shape
        self.grid = torch.from_numpy(np.ascontiguousarray(grid.reshape(-1)))
        return self.grid

    def argmax(self):
        o = self.o
        lo, _, gr = o.grid_geometry(self.pc, self.cfg.res)
        world = np.empty(3, np.float64)
        g = np.ascontiguousarray(self.grid.numpy(), dtype=np.int64)
        o.lib().oracle_grid_argmax(g, gr, lo, float(self.cfg.res), world)
        self.T_est = world
        self.out["grid"] = g.reshape(self.shape).copy()
        self.out["T_est"] = world.copy()

    def errors(self):
        o, cfg = self.o, self.cfg
        tr_back, _ = o.generate_target_pairs(self.pc[self.idx[:, :2]], cfg.up, cfg.front, cfg.right, self.T_est, want_rot=False)
        self.errs = o.backvote_errors(self.tr, tr_back)
        return torch.from_numpy(self.errs)

    def errs_all_buffer(self, world):
        return None

    def select(self, errs_all):
        self.thr = np.percentile(errs_all.numpy(), self.cfg.backproj_ratio * 100)      # eval.py:257 on the GLOBAL errors
        self.out["thr"] = float(self.thr)

    def mask_local(self, own_scale):
        n = self.pc.shape[0]
        self.mask = self.errs < self.thr
        imp = np.bincount(self.idx[self.mask, :2].reshape(-1), minlength=n).astype(np.int32)
        hist = np.zeros(3 * 65536, np.int32)
        self.keys = _float_keys(self.scales[self.mask])                                  # [M_local, 3] uint32
        for a in range(3):
            np.add.at(hist, a * 65536 + (self.keys[:, a] >> 16), 1)
        self.x = torch.from_numpy(np.concatenate([imp, np.array([self.mask.sum()], np.int32), hist]))
        self.out["pairs_mask_local"] = self.mask.copy()
        return self.x

    def _pick(self, hist, k):
        cum = np.cumsum(hist.astype(np.int64))
        d = int(np.searchsorted(cum, k, side="right"))
        return d, k - (cum[d - 1] if d > 0 else 0)

    def after_mask(self, own_scale):
        n = self.pc.shape[0]
        x = self.x.numpy()
        self.imp, self.kept = x[:n].astype(np.int64), int(x[n])
        self.out.update(imp=self.imp.copy(), kept=self.kept)
        self.hi, self.k1 = [], []
        h1 = np.zeros(3 * 65536, np.int32)
        for a in range(3):
            d, k = self._pick(x[n + 1 + a * 65536: n + 1 + (a + 1) * 65536], (self.kept - 1) // 2)
            self.hi.append(d)
            self.k1.append(k)
            sel = (self.keys[:, a] >> 16) == d
            np.add.at(h1, a * 65536 + (self.keys[sel, a] & 0xffff), 1)
        self.h1 = torch.from_numpy(h1)

    def rotation_counts(self):
        o, cfg = self.o, self.cfg
        imp_wt = self.imp / self.imp.max()
        wt = np.ones(self.idx.shape[0], np.float64)
        wt[self.mask] = imp_wt[self.idx[self.mask, :2]].sum(-1) + cfg.imp_wt_margin          # eval.py:274-275, global counts
        sphere = o.fibonacci_sphere(cfg.num_sphere)
        keep = self.mask.astype(np.uint8)
        c_up = o.rotation_counts(self.pc, self.idx, self.rot[:, 0], wt, keep, cfg.num_rots, sphere, cfg.angle_tol)
        c_right = o.rotation_counts(self.pc, self.idx, self.rot[:, 2], wt, keep, cfg.num_rots, sphere, cfg.angle_tol)
        self.counts = torch.from_numpy(np.stack([c_up, c_right]))
        return [self.counts, self.h1]

    def pose_local(self, scale_override=None):
        o, cfg = self.o, self.cfg
        c = self.counts.numpy()
        sphere = o.fibonacci_sphere(cfg.num_sphere)
        b_up, b_right = int(np.argmax(c[0].astype(np.float32))), int(np.argmax(c[1].astype(np.float32)))
        R = o.assemble_rotation(sphere[b_up], sphere[b_right], cfg.up, cfg.right)
        h1 = self.h1.numpy()
        keys = [np.uint32((self.hi[a] << 16) | self._pick(h1[a * 65536:(a + 1) * 65536], self.k1[a])[0]) for a in range(3)]
        scale = _keys_to_float(np.array(keys, np.uint32))
        canon = (self.pc - self.T_est) @ R / np.linalg.norm(scale)
        terms = np.clip(np.abs(canon[self.idx[self.mask, :2]] - self.pred_l[self.mask]), 0, 0.1)
        self.loss_sum = torch.tensor([terms.sum()], dtype=torch.float64)
        self.out.update(counts_up=c[0].copy(), counts_right=c[1].copy(), bin_up=b_up, bin_right=b_right, R_est=R, pred_scale=scale)
        return self.loss_sum

    def finish(self):
        self.out["loss"] = float(self.loss_sum[0] / (2 * self.kept * 3))
        return self.out


def _float_keys(x):
    """Order-preserving uint32 image of float32 values (csrc/common.cuh float_to_key)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)


def _keys_to_float(k):
    u = np.where(k & 0x80000000, k & 0x7fffffff, ~k).astype(np.uint32)
    return u.view(np.float32)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, payload, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cppf2_b200.pipeline import VoteConfig
        from cppf2_b200.sharded import ShardedVote, shard_bounds
        pc, idx, bins, scales = payload
        cfg = VoteConfig(res=0.002)
        lo, hi = shard_bounds(idx.shape[0], world, rank)
        stages = OracleStages()
        sv = ShardedVote(stages)
        assert sv.world == world and sv.rank == rank
        out = sv.vote(pc, torch.from_numpy(idx[lo:hi]), cfg, torch.from_numpy(scales[lo:hi]), torch.from_numpy(bins[lo:hi]))
        assert sv.n_collectives == 5, "one exchange per stage: grid, errors, imp+scale, bins+scale, loss"
        ret[rank] = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in out.items()}
    finally:
        dist.destroy_process_group()


def _inputs(T=6000, n=1200):
    from cppf2_b200 import synth
    pc = synth.half_cylinder_cloud(n, seed=1)
    idx = synth.sample_tuples(pc.shape[0], T, 5, seed=2)
    rng = np.random.default_rng(3)
    canon = (pc[idx[:, :2]].astype(np.float64) - np.array([0.0, 0.0, 0.8])) / 0.14
    bins = np.clip(np.rint((canon + 0.5) * 31) + rng.integers(-1, 2, canon.shape), 0, 31).reshape(T, 6).astype(np.uint8)
    scales = (np.array([0.57, 0.71, 0.41]) + 0.02 * rng.standard_normal((T, 3))).astype(np.float32)
    return pc, idx, bins, scales


def test_shard_bounds():
    from cppf2_b200.sharded import shard_bounds
    assert [shard_bounds(50000, 8, r) for r in (0, 7)] == [(0, 6250), (43750, 50000)]
    with pytest.raises(ValueError):
        shard_bounds(50001, 2, 0)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world", [2, 3])          # 3: blocks that are not a power-of-two share of the tuples
def test_multi_rank_vote_equals_single_process(oracle, world):
    pc, idx, bins, scales = _inputs()
    ref = oracle.instance_body(pc, idx, bins, scales, [0, 1, 0], [1, 0, 0], [0, 0, 1], 0.002)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), (pc, idx, bins, scales), ret), nprocs=world, join=True)
    assert set(ret.keys()) == set(range(world))
    # each rank masked only its own block; together the blocks are the single-process kept set
    assert np.array_equal(np.concatenate([ret[r]["pairs_mask_local"] for r in range(world)]), ref["pairs_mask"])
    for r in range(world):
        out = ret[r]
        assert np.array_equal(out["grid"], ref["grid"]), "all-reduced grid differs from the single-process grid"
        assert np.array_equal(out["T_est"], ref["T_est"])
        assert out["thr"] == float(ref["thr"]) and out["kept"] == int(ref["pairs_mask"].sum())
        assert np.array_equal(out["imp"], ref["imp"])
        np.testing.assert_allclose(out["loss"], ref["loss"], rtol=1e-12)
        np.testing.assert_allclose(out["counts_up"], ref["counts_up"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["counts_right"], ref["counts_right"], rtol=1e-9, atol=1e-9)
        assert out["bin_up"] == ref["bin_up"] and out["bin_right"] == ref["bin_right"]
        np.testing.assert_allclose(out["R_est"], ref["R_est"], atol=1e-12)
        np.testing.assert_array_equal(out["pred_scale"], ref["pred_scale"])
    # the partial grids really were partial: each rank alone does not reproduce the full grid
    half, _ = oracle.vote_center(pc, oracle.generate_target_pairs(oracle.decode_pairs(pc, idx[:3000], bins[:3000])[1],
                                                                 [0, 1, 0], [0, 0, 1], [1, 0, 0])[0], 0.002, idx[:3000, :2], 180)
    assert half.sum() < ref["grid"].sum()
