"""World-size-2 `gloo` test of the tuple-sharded vote orchestration (cppf2_b200/sharded.py) on CPU.

The collectives and the sharding arithmetic are the product's; the stage kernels need a GPU, so this test
plugs an oracle-backed implementation of the stage interface into `ShardedVote` (test infrastructure only)
and checks that two ranks, each holding half of the tuples, end with exactly the single-process result:
bit-identical centre grid, voted centre, kept set and rotation bins' arg-max; float bins within 1e-9.
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
    """The stage interface of cppf2_b200.sharded.ShardedVote on top of the CPU oracle (torch CPU tensors in/out)."""

    def __init__(self):
        from oracle import cpu
        self.o = cpu
        self.out = {}

    def decode_targets(self, pc, idx_local, bins_local, cfg):
        o = self.o
        self.pc = np.ascontiguousarray(pc, dtype=np.float32)
        idx = idx_local.numpy()
        self.pred_l, scaled, _ = o.decode_pairs(self.pc, idx, bins_local.numpy(), cfg.num_bins)
        tr, rot = o.generate_target_pairs(scaled, cfg.up, cfg.front, cfg.right)
        return torch.from_numpy(tr), torch.from_numpy(rot)

    def vote_center(self, pc, idx_local, tr_l, cfg):
        grid, _ = self.o.vote_center(self.pc, tr_l.numpy(), cfg.res, idx_local.numpy()[:, :2], cfg.num_rots)
        self.shape = grid.shape
        return torch.from_numpy(np.ascontiguousarray(grid.reshape(-1)))

    def argmax(self, grid, cfg):
        o = self.o
        lo, _, gr = o.grid_geometry(self.pc, cfg.res)
        world = np.empty(3, np.float64)
        g = np.ascontiguousarray(grid.numpy(), dtype=np.int64)
        o.lib().oracle_grid_argmax(g, gr, lo, float(cfg.res), world)
        self.T_est = world
        self.out["grid"] = g.reshape(self.shape).copy()
        self.out["T_est"] = world.copy()

    def errors(self, pc, idx_local, tr_l):
        o = self.o
        pairs = self.pc[idx_local.numpy()[:, :2]]
        tr_back, _ = o.generate_target_pairs(pairs, self._cfg.up, self._cfg.front, self._cfg.right, self.T_est, want_rot=False)
        return torch.from_numpy(o.backvote_errors(tr_l.numpy(), tr_back))

    def select_and_mask(self, errs, idx, pc, cfg):
        thr, mask, imp, pair_wt = self.o.backvote_filter(errs.numpy(), idx.numpy(), self.pc.shape[0], cfg.backproj_ratio,
                                                         cfg.imp_wt_margin)
        self.mask, self.pair_wt = mask, pair_wt
        self.out.update(pairs_mask=mask.copy(), imp=imp.copy(), thr=float(thr))

    def rotation_counts(self, pc, idx, rot, cfg, part, n_parts):
        o = self.o
        idx_np, rot_np = idx.numpy(), rot.numpy()
        kept = np.nonzero(self.mask)[0]
        mine = np.zeros(idx_np.shape[0], np.uint8)
        mine[kept[part::n_parts]] = 1
        wt = np.ones(idx_np.shape[0], np.float64)
        wt[self.mask] = self.pair_wt
        sphere = o.fibonacci_sphere(cfg.num_sphere)
        c_up = o.rotation_counts(self.pc, idx_np, rot_np[:, 0], wt, mine, cfg.num_rots, sphere, cfg.angle_tol)
        c_right = o.rotation_counts(self.pc, idx_np, rot_np[:, 2], wt, mine, cfg.num_rots, sphere, cfg.angle_tol)
        return torch.from_numpy(np.stack([c_up, c_right]))

    def finalize(self, pc, idx, bins, scales, counts, cfg, scale_override=None):
        o = self.o
        c = counts.numpy()
        sphere = o.fibonacci_sphere(cfg.num_sphere)
        b_up, b_right = int(np.argmax(c[0].astype(np.float32))), int(np.argmax(c[1].astype(np.float32)))
        R = o.assemble_rotation(sphere[b_up], sphere[b_right], cfg.up, cfg.right)
        scale = o.lower_median(scales.numpy()[self.mask])
        self.out.update(counts_up=c[0].copy(), counts_right=c[1].copy(), bin_up=b_up, bin_right=b_right, R_est=R, pred_scale=scale)
        return self.out


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
        stages._cfg = cfg
        sv = ShardedVote(stages)
        assert sv.world == world and sv.rank == rank
        out = sv.vote(pc, torch.from_numpy(idx[lo:hi]), cfg, torch.from_numpy(scales[lo:hi]), torch.from_numpy(bins[lo:hi]))
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
def test_two_rank_vote_equals_single_process(oracle):
    pc, idx, bins, scales = _inputs()
    ref = oracle.instance_body(pc, idx, bins, scales, [0, 1, 0], [1, 0, 0], [0, 0, 1], 0.002)
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), (pc, idx, bins, scales), ret), nprocs=world, join=True)
    assert set(ret.keys()) == {0, 1}
    for r in range(world):
        out = ret[r]
        assert np.array_equal(out["grid"], ref["grid"]), "all-reduced grid differs from the single-process grid"
        assert np.array_equal(out["T_est"], ref["T_est"])
        assert np.array_equal(out["pairs_mask"], ref["pairs_mask"])
        assert np.array_equal(out["imp"], ref["imp"])
        np.testing.assert_allclose(out["counts_up"], ref["counts_up"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["counts_right"], ref["counts_right"], rtol=1e-9, atol=1e-9)
        assert out["bin_up"] == ref["bin_up"] and out["bin_right"] == ref["bin_right"]
        np.testing.assert_allclose(out["R_est"], ref["R_est"], atol=1e-12)
        np.testing.assert_array_equal(out["pred_scale"], ref["pred_scale"])
    # the partial grids really were partial: each rank alone does not reproduce the full grid
    half, _ = oracle.vote_center(pc, oracle.generate_target_pairs(oracle.decode_pairs(pc, idx[:3000], bins[:3000])[1],
                                                                 [0, 1, 0], [0, 0, 1], [1, 0, 0])[0], 0.002, idx[:3000, :2], 180)
    assert half.sum() < ref["grid"].sum()
