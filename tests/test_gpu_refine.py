"""Online pose refinement (eval.py:319-355, `opt=True`) on the device against the torch-CPU restatement of the reference
loop (oracle/refine_torch.py: the reference's own lines with lietorch.SO3 restated -- parity unpinned for that part).

Tolerance (BASELINE north star): 1 mm on t and 0.1 degrees on R.  Both sides run float32 per-element arithmetic; Adam on
an L1 objective moves by about lr per step whatever the gradient's magnitude, so a sign that flips for one row of one
step (reduction order: float64 sums here, cuBLAS/atomics in the reference) shifts the trajectory by a fraction of a
micrometre -- measured differences are printed."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _instance(seed, n=1500, T=8192):
    from cppf2_b200 import synth
    pc = synth.half_cylinder_cloud(n, seed=seed)
    idx = synth.sample_tuples(pc.shape[0], T, 5, seed=seed + 1)
    rng = np.random.default_rng(seed + 2)
    canon = (pc[idx[:, :2]].astype(np.float64) - np.array([0.0, 0.0, 0.8])) / 0.14
    bins = np.clip(np.rint((canon + 0.5) * 31) + rng.integers(-1, 2, canon.shape), 0, 31).reshape(T, 6).astype(np.uint8)
    scales = (np.array([0.57, 0.71, 0.41]) + 0.02 * rng.standard_normal((T, 3))).astype(np.float32)
    return pc, idx, bins, scales


def _angle_deg(Ra, Rb):
    """Rotation angle between two rotations; the chordal form ||Ra - Rb||_F / sqrt(2) resolves small angles that
    arccos((trace - 1) / 2) cannot (float32 entries leave ~0.03 degrees of noise there)."""
    chord = float(np.linalg.norm(Ra - Rb)) / np.sqrt(2.0)
    return float(np.degrees(2.0 * np.arcsin(min(1.0, chord / 2.0))))


@pytest.mark.parametrize("seed,y_only", [(1, False), (5, True), (9, False)])
def test_refined_pose_matches_the_reference_loop(seed, y_only):
    from cppf2_b200 import _lib
    from cppf2_b200.pipeline import PoseVoter, VoteConfig
    from oracle import cpu as oracle
    from oracle.refine_torch import final_loss, refine_pose
    pc, idx, bins, scales = _instance(seed)
    T = idx.shape[0]
    voter = PoseVoter(T, pc.shape[0])
    base = voter.vote(pc, idx, VoteConfig(res=0.002, loss_y_only=y_only), pred_scales=scales, bins=bins).result()
    mask = voter.intermediates()["pairs_mask"]
    assert not base.status & _lib.CPPF_STATUS_REFINED
    got = voter.vote(pc, idx, VoteConfig(res=0.002, loss_y_only=y_only, opt=True), pred_scales=scales, bins=bins).result()
    assert got.status & _lib.CPPF_STATUS_REFINED and got.kept == base.kept
    pred, scaled, _ = oracle.decode_pairs(pc, idx[:, :2], bins)
    want_t, want_R = refine_pose(pc, idx[mask][:, :2], scaled[mask], base.t, base.R, y_only)
    dt = float(np.abs(got.t - want_t.astype(np.float64)).max())
    dR = _angle_deg(got.R, want_R.astype(np.float64))
    moved_t = float(np.linalg.norm(want_t - base.t))
    moved_R = _angle_deg(base.R, want_R.astype(np.float64))
    print(f"seed {seed}: refinement moved t by {moved_t * 1e3:.3f} mm and R by {moved_R:.3f} deg; "
          f"device vs reference loop: dt {dt * 1e3:.5f} mm, dR {dR:.5f} deg")
    assert dt < 1e-3 and dR < 0.1
    # tighter in practice: two float32 implementations of the same 100 steps
    assert dt < 1e-4 and dR < 0.01
    # the pose record holds float32 values (the reference reads opt_trans / the matrix back as float32)
    assert np.array_equal(got.t, got.t.astype(np.float32).astype(np.float64))
    want_loss = final_loss(pc, idx[mask][:, :2], pred[mask], want_t, want_R, np.float32(base.scale_norm), y_only)
    assert abs(got.loss - want_loss) < 1e-4
    assert got.scale_norm == base.scale_norm and np.array_equal(got.scale, base.scale)


def test_refinement_lowers_the_unclipped_objective():
    from cppf2_b200.pipeline import PoseVoter, VoteConfig
    from oracle import cpu as oracle
    pc, idx, bins, scales = _instance(3)
    T = idx.shape[0]
    voter = PoseVoter(T, pc.shape[0])
    base = voter.vote(pc, idx, VoteConfig(res=0.002), pred_scales=scales, bins=bins).result()
    mask = voter.intermediates()["pairs_mask"]
    got = voter.vote(pc, idx, VoteConfig(res=0.002, opt=True), pred_scales=scales, bins=bins).result()
    _, scaled, _ = oracle.decode_pairs(pc, idx[:, :2], bins)

    def objective(r):
        return float(np.abs(((pc - r.t) @ r.R)[idx[mask][:, :2]] - scaled[mask]).mean())

    assert objective(got) < objective(base)


def test_estimator_opt_flag_runs_both_paths_alike():
    """opt=True through the public call: the whole-frame call (batched refinement kernel), the one-call instance path and the
    step-by-step Python sequence agree."""
    import os
    from cppf2_b200 import synth
    from cppf2_b200.estimator import Instance, PoseEstimator, build_models
    pc = synth.half_cylinder_cloud(1200, seed=2)
    desc = synth.unit_descriptors(pc.shape[0], 1024, seed=4)
    idx = synth.sample_tuples(pc.shape[0], 4096, 5, seed=3)
    models, cfgs = build_models(["mug"], precision=1)
    outs = []
    for frame_call, one_call in (("1", "1"), ("0", "1"), ("0", "0")):
        os.environ["CPPF_FRAME_CALL"], os.environ["CPPF_ONE_CALL"] = frame_call, one_call
        try:
            est = PoseEstimator(models, cfgs, num_pairs=4096, max_points=pc.shape[0], opt=True, seed=7)
            assert est.frame_call == (frame_call == "1")
            outs.append(est.estimate([Instance(pc=pc, category="mug", desc=desc, point_idxs=idx)])[0])
        finally:
            os.environ.pop("CPPF_ONE_CALL", None)
            os.environ.pop("CPPF_FRAME_CALL", None)
    assert all(o is not None for o in outs)
    for a, b in ((outs[0], outs[2]), (outs[1], outs[2])):
        for br in a.results:
            ra, rb = a.results[br], b.results[br]
            assert ra.status & 8 and rb.status & 8 and ra.kept == rb.kept                 # CPPF_STATUS_REFINED on both paths
            for r in (ra, rb):
                assert np.isfinite(r.R).all() and np.isfinite(r.t).all() and np.isfinite(r.loss)
                assert np.allclose(r.R @ r.R.T, np.eye(3), atol=1e-5)                      # Q(q) R_est stays a rotation
            # Same kernels and inputs on both paths.  The kept list leaves the atomic compaction in a different order from launch to
            # launch; the float64 row sums make a step independent of that order except for a last-bit rounding, and 100 Adam steps
            # on the predictions of RANDOM-INIT heads (an ill-conditioned objective) can amplify such a bit: reported, not asserted.
            print(f"{br}: one-call vs step-by-step after refinement: max |dR| {np.abs(ra.R - rb.R).max():.2e}, max |dt| {np.abs(ra.t - rb.t).max():.2e}")
