"""Oracle-side checks of the pose refinement restatement (oracle/refine_torch.py; eval.py:319-355).  lietorch is absent
here, so its part is parity-unpinned: these tests pin the restatement's internal consistency -- the tangent-space gradient
it returns is the derivative along left perturbations exp(xi) X, and the loop lowers the objective it optimises."""
import numpy as np
import torch

from oracle.refine_torch import SO3Matrix, refine_pose


def _quat_mul(a, b):   # (x, y, z, w)
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def _rot(q):
    return SO3Matrix.apply(torch.from_numpy(np.asarray(q, dtype=np.float64))).numpy()


def test_matrix_is_the_rotation_of_the_normalised_quaternion():
    q = np.array([0.3, -0.2, 0.5, 1.0])
    Q = _rot(q)
    assert np.allclose(Q @ Q.T, np.eye(3), atol=1e-12) and np.isclose(np.linalg.det(Q), 1.0)
    assert np.allclose(_rot(3.7 * q), Q, atol=1e-12)            # the raw data is normalised first
    axis = q[:3] / np.linalg.norm(q[:3])
    assert np.allclose(Q @ axis, axis, atol=1e-12)               # rotation about the vector part


def test_matrix_matches_scipy_scalar_last_quaternions():
    """lietorch stores an SO3 element as (x, y, z, w) (`delta_rot = [0, 0, 0, 1]` is the identity, eval.py:323); scipy's
    Rotation uses the same scalar-last layout, so it is an independent reading of `SO3.InitFromVec(q).matrix()` for unit q."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(4)
    assert np.array_equal(_rot([0.0, 0.0, 0.0, 1.0]), np.eye(3))
    for _ in range(20):
        q = rng.standard_normal(4)
        np.testing.assert_allclose(_rot(q), Rotation.from_quat(q / np.linalg.norm(q)).as_matrix(), atol=1e-12)


def test_gradient_is_the_left_perturbation_derivative():
    rng = np.random.default_rng(0)
    q = np.array([0.1, -0.25, 0.4, 1.0])
    A = rng.standard_normal((3, 3))
    qt = torch.tensor(q, dtype=torch.float64, requires_grad=True)
    (SO3Matrix.apply(qt) * torch.from_numpy(A)).sum().backward()
    g = qt.grad.numpy()
    assert g[3] == 0.0                                           # embedded tangent gradient: fourth slot empty
    eps = 1e-6
    for k in range(3):
        xi = np.zeros(3)
        xi[k] = eps
        dq = np.concatenate([np.sin(eps / 2) * xi / eps, [np.cos(eps / 2)]])     # exp(xi) as a quaternion
        up = (_rot(_quat_mul(dq, q / np.linalg.norm(q))) * A).sum()
        dq[:3] *= -1
        dn = (_rot(_quat_mul(dq, q / np.linalg.norm(q))) * A).sum()
        assert np.isclose((up - dn) / (2 * eps), g[k], rtol=1e-6, atol=1e-8)


def test_refinement_lowers_the_objective_on_a_perturbed_pose():
    rng = np.random.default_rng(1)
    n, m = 400, 600
    canon = rng.uniform(-0.1, 0.1, (n, 3)).astype(np.float32)
    ang = 0.7
    R_true = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t_true = np.array([0.05, -0.02, 0.9])
    pc = (canon @ R_true.T + t_true).astype(np.float32)
    pair_idx = rng.integers(0, n, (m, 2))
    target = canon[pair_idx] + rng.normal(0, 1e-3, (m, 2, 3)).astype(np.float32)
    da = 0.05
    R0 = R_true @ np.array([[1, 0, 0], [0, np.cos(da), -np.sin(da)], [0, np.sin(da), np.cos(da)]])
    t0 = t_true + np.array([0.004, -0.003, 0.005])

    def objective(t, R):
        return float(np.abs(((pc - t) @ R)[pair_idx] - target).mean())

    before = objective(t0, R0)
    t1, R1 = refine_pose(pc, pair_idx, target, t0, R0, y_only=False)
    after = objective(t1.astype(np.float64), R1.astype(np.float64))
    assert t1.dtype == np.float32 and R1.dtype == np.float32
    assert after < 0.5 * before, (before, after)
    assert np.allclose(R1.astype(np.float64) @ R1.T.astype(np.float64), np.eye(3), atol=1e-5)
