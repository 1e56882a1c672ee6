"""TEST INFRASTRUCTURE ONLY.  The reference's online pose refinement (eval.py:319-355, `opt=True`, its default) on
torch-CPU: the loop body is the reference's own lines -- torch.optim.Adam over (translation, quaternion), L1 distance
between the canonicalised kept pairs and the scaled predictions, `delta_rot.grad / 180 * np.pi` -- with ONE substitution:
`lietorch.SO3.InitFromVec(q).matrix()[:3, :3]` (lietorch==0.2, environment.yml:110, absent from this image and not
vendored under /root/reference) is restated by `SO3Matrix` below.

PARITY UNPINNED for the lietorch part: nothing in the reference pins its values.  What is restated (lietorch 0.2,
`lietorch/include/so3.h`, `lietorch/group_ops.py`, `lietorch/groups.py::matrix`):
  * the group element built from raw data normalises the quaternion (x, y, z, w);
  * `matrix()` acts the element on the rows of the identity: column i of the result is q * e_i, Eigen's
    `v + w * (2 u x v) + u x (2 u x v)` with u = (x, y, z);
  * the gradient with respect to a group element is returned in the TANGENT space of a left perturbation,
    d/d(xi) L(exp(xi) X) at xi = 0 -- for `act` that is sum_i (X p_i) x dL/d(X p_i) -- stored in the first three of
    the four embedding slots, the fourth being zero.  Adam therefore moves x, y, z of the raw quaternion and never w.
The torch parts (autograd of abs / mean / matmul / indexing, torch.optim.Adam) are the real thing.
"""
import numpy as np
import torch


class SO3Matrix(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q):
        qn = q / torch.linalg.vector_norm(q)
        u, w = qn[:3], qn[3]
        eye = torch.eye(3, dtype=q.dtype)
        cols = []
        for i in range(3):                                       # Eigen's order of operations, column by column
            v = eye[i]
            uv = torch.linalg.cross(u, v)
            uv = uv + uv
            cols.append(v + w * uv + torch.linalg.cross(u, uv))
        Q = torch.stack(cols, dim=1)
        ctx.save_for_backward(Q)
        return Q

    @staticmethod
    def backward(ctx, G):
        (Q,) = ctx.saved_tensors
        g = torch.zeros(3, dtype=G.dtype)
        for i in range(3):
            g = g + torch.linalg.cross(Q[:, i], G[:, i])
        return torch.cat([g, torch.zeros(1, dtype=G.dtype)])


def refine_pose(pc, pair_idx, pred_pairs_scaled, T_est, R_est, y_only, iters=100, lr=1e-2):
    """eval.py:319-355.  pc [N,3] f32, pair_idx [M,2] (the kept tuples' first two points), pred_pairs_scaled [M,2,3] f32
    (pred_pairs_scaled[pairs_mask]), T_est [3] / R_est [3,3] float64 as voted.  Returns (T_est f32 [3], R_est f32 [3,3])."""
    from torch import optim
    pc_t = torch.from_numpy(np.ascontiguousarray(pc)).float()
    idx_t = torch.from_numpy(np.ascontiguousarray(pair_idx)).long()
    target = torch.from_numpy(np.ascontiguousarray(pred_pairs_scaled)).float()
    R0 = torch.from_numpy(np.ascontiguousarray(R_est)).float()
    # a hundred steps of ~20 tiny ops: intra-op threads only add hand-off cost here (4.2 s with 16 threads, 0.13 s with one)
    n_threads = torch.get_num_threads()
    torch.set_num_threads(1)
    with torch.enable_grad():
        opt_trans = torch.nn.Parameter(torch.from_numpy(np.ascontiguousarray(T_est)).float(), requires_grad=True)
        delta_rot = torch.tensor([0, 0, 0, 1.], requires_grad=True)
        opt = optim.Adam([opt_trans, delta_rot], lr=lr)
        for _ in range(iters):
            opt.zero_grad()
            rot = SO3Matrix.apply(delta_rot) @ R0
            pc_canon = (pc_t - opt_trans) @ rot
            loss = torch.abs(pc_canon[idx_t] - target)
            if y_only:
                loss = loss[..., 1]
            loss = loss.mean()
            loss.backward()
            delta_rot.grad = delta_rot.grad / 180 * np.pi
            opt.step()
    T_out = opt_trans.detach().numpy()
    R_out = (SO3Matrix.apply(delta_rot.detach()) @ R0).numpy()
    torch.set_num_threads(n_threads)
    return T_out, R_out


def final_loss(pc, pair_idx, pred_pairs, T_est, R_est, scale_norm, y_only):
    """eval.py:358-363 with the refined (float32) pose."""
    pc_canon = (pc - T_est) @ R_est / scale_norm
    loss = np.abs(pc_canon[pair_idx] - pred_pairs)
    if y_only:
        loss = loss[..., 1]
    return float(np.clip(loss, 0, 0.1).mean())
