"""Seeded synthetic inputs shaped like the reference's workloads.

Nothing here is on the timed path; these generators only manufacture the
clouds / frames / tuples that tests, golden vectors and bench.py feed to it.
All generators are numpy-only and deterministic given the seed.
"""
from __future__ import annotations

import numpy as np

# REAL275 camera used by the reference evaluation loop (eval.py:82)
REAL275_K = np.array([[591.0125, 0.0, 322.525], [0.0, 590.16775, 244.11084], [0.0, 0.0, 1.0]])

# metric diagonal ranges per category (dataset.py:165-172), keyed by name
CATEGORY_DIAG = {
    "can": (0.128, 0.18),
    "bottle": (0.16, 0.25),
    "bowl": (0.1851, 0.26),
    "camera": (0.1430, 0.28),
    "laptop": (0.3862, 0.58),
    "mug": (0.1501, 0.1995),
}
CATEGORY_RES = {"can": 0.002, "bottle": 0.002, "bowl": 0.002, "camera": 0.002, "laptop": 0.01, "mug": 0.002}
REAL275_CATEGORIES = ["bottle", "bowl", "camera", "can", "laptop", "mug"]


def half_cylinder_cloud(n: int = 4096, radius: float = 0.04, height: float = 0.10, z0: float = 0.8,
                        seed: int = 3, jitter: float = 0.0) -> np.ndarray:
    """Camera-facing half cylinder (config 4 cloud): grid about 40x50x20 cells at res 2 mm."""
    rng = np.random.default_rng(seed)
    phi = rng.uniform(0.0, np.pi, n)
    h = rng.uniform(-0.5 * height, 0.5 * height, n)
    r = radius + (rng.uniform(-jitter, jitter, n) if jitter > 0 else 0.0)
    pts = np.stack([r * np.cos(phi), h, z0 - r * np.sin(phi)], -1)
    return pts.astype(np.float32)


def torus_cloud(n: int, res: float = 0.002, z0: float = 1.0, seed: int = 7):
    """Closed smooth surface whose area is n*res^2 (config 3: SHOT sweep).

    Returns (points f32 [n,3], outward normals f32 [n,3]).  Major radius = 4 x minor radius.
    """
    rng = np.random.default_rng(seed)
    area = n * res * res
    r_minor = np.sqrt(area / (4.0 * np.pi * np.pi * 4.0))
    r_major = 4.0 * r_minor
    # rejection-sample the poloidal angle so that the density is uniform in area
    u = rng.uniform(0, 2 * np.pi, 2 * n + 64)
    acc = rng.uniform(0, 1, u.shape[0]) < (r_major + r_minor * np.cos(u)) / (r_major + r_minor)
    u = u[acc][:n]
    while u.shape[0] < n:  # pragma: no cover - vanishingly unlikely
        extra = rng.uniform(0, 2 * np.pi, n)
        u = np.concatenate([u, extra])[:n]
    v = rng.uniform(0, 2 * np.pi, n)
    nrm = np.stack([np.cos(u) * np.cos(v), np.cos(u) * np.sin(v), np.sin(u)], -1)
    ctr = np.stack([r_major * np.cos(v), r_major * np.sin(v), np.zeros(n)], -1)
    pts = ctr + r_minor * nrm + nrm * rng.uniform(-res / 4, res / 4, (n, 1))
    pts[:, 2] += z0
    return pts.astype(np.float32), nrm.astype(np.float32)


def sample_tuples(n_points: int, n_tuples: int, arity: int = 5, seed: int = 11) -> np.ndarray:
    """Tuple indices with replacement, like np.random.randint(0,N,(T,K)) (eval.py:207)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_points, (n_tuples, arity), dtype=np.int64)


def noisy_center_targets(pc: np.ndarray, idx: np.ndarray, center: np.ndarray, sigma: float = 0.002,
                         seed: int = 5) -> np.ndarray:
    """(proj_len, dist2o) of each pair w.r.t. `center` plus N(0,sigma) so that a peak exists (config 4)."""
    rng = np.random.default_rng(seed)
    a = pc[idx[:, 0]].astype(np.float64)
    b = pc[idx[:, 1]].astype(np.float64)
    d = a - b
    u = d / (np.linalg.norm(d, axis=-1, keepdims=True) + 1e-7)
    proj = np.sum((a - center) * u, -1)
    dist = np.linalg.norm((a - center) - proj[:, None] * u, axis=-1)
    tr = np.stack([proj, dist], -1) + rng.normal(0, sigma, (idx.shape[0], 2))
    return tr.astype(np.float32)


def noisy_axis_angles(pc: np.ndarray, idx: np.ndarray, axis: np.ndarray, sigma_deg: float = 2.0,
                      seed: int = 6) -> np.ndarray:
    """Angle between each pair direction and `axis`, plus N(0,sigma) (config 4 rotation stage)."""
    rng = np.random.default_rng(seed)
    d = pc[idx[:, 0]].astype(np.float64) - pc[idx[:, 1]].astype(np.float64)
    u = d / (np.linalg.norm(d, axis=-1, keepdims=True) + 1e-7)
    ang = np.arccos(np.clip(u @ axis, -1, 1)) + np.deg2rad(sigma_deg) * rng.normal(0, 1, idx.shape[0])
    return ang.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# REAL275-shaped synthetic frames (config 2 / 5)
# ---------------------------------------------------------------------------------------------

def _rot_yaw_elev(yaw: float, elev: float) -> np.ndarray:
    cy, sy = np.cos(yaw), np.sin(yaw)
    ce, se = np.cos(elev), np.sin(elev)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rx = np.array([[1, 0, 0], [0, ce, -se], [0, se, ce]])
    return rx @ ry


def _primitive_surface(cat: str, diag: float, rng: np.random.Generator, n: int) -> np.ndarray:
    """Dense surface samples of an analytic stand-in for a category, object frame, y up, diagonal=diag."""
    if cat in ("bottle", "can", "mug"):
        ratio = {"bottle": 3.0, "can": 1.6, "mug": 1.1}[cat]  # height / diameter
        d = diag / np.sqrt(2 + ratio * ratio)
        r, h = 0.5 * d, ratio * d
        k = rng.uniform(0, 1, n)
        side = k < 0.8
        phi = rng.uniform(0, 2 * np.pi, n)
        y = rng.uniform(-0.5 * h, 0.5 * h, n)
        rad = np.where(side, r, r * np.sqrt(rng.uniform(0, 1, n)))
        y = np.where(side, y, np.where(rng.uniform(0, 1, n) < 0.5, 0.5 * h, -0.5 * h))
        pts = np.stack([rad * np.cos(phi), y, rad * np.sin(phi)], -1)
        if cat == "mug":  # handle: thin box on +x
            m = n // 8
            hb = np.stack([r + rng.uniform(0, 0.5 * r, m), rng.uniform(-0.25 * h, 0.25 * h, m),
                           rng.uniform(-0.08 * r, 0.08 * r, m)], -1)
            pts = np.concatenate([pts, hb])
        return pts
    if cat == "bowl":
        r = diag / np.sqrt(4 + 4 + 1.0)  # bbox 2r x r x 2r
        phi = rng.uniform(0, 2 * np.pi, n)
        ct = rng.uniform(0, 1, n)  # lower hemisphere cap
        st = np.sqrt(1 - ct * ct)
        return np.stack([r * st * np.cos(phi), 0.5 * r - r * ct, r * st * np.sin(phi)], -1)
    # camera / laptop: boxes
    dims = {"camera": np.array([1.0, 0.7, 0.6]), "laptop": np.array([1.0, 0.7, 0.75])}[cat]
    dims = dims / np.linalg.norm(dims) * diag
    if cat == "laptop":  # base slab + tilted screen slab
        m = n // 2
        base = np.stack([rng.uniform(-.5, .5, m) * dims[0], np.full(m, -0.5 * dims[1]),
                         rng.uniform(-.5, .5, m) * dims[2]], -1)
        t = rng.uniform(0, 1, n - m)
        screen = np.stack([rng.uniform(-.5, .5, n - m) * dims[0], -0.5 * dims[1] + t * dims[1],
                           -0.5 * dims[2] - 0.15 * t * dims[2]], -1)
        return np.concatenate([base, screen])
    face = rng.integers(0, 6, n)
    uvw = rng.uniform(-0.5, 0.5, (n, 3))
    ax = face // 2
    uvw[np.arange(n), ax] = np.where(face % 2 == 0, -0.5, 0.5)
    return uvw * dims


def synth_real275_frame(frame_id: int = 0, n_instances: int = 6, height: int = 480, width: int = 640,
                        noise_mm: float = 0.0, K: np.ndarray = REAL275_K):
    """z-buffered 640x480 uint16 depth (mm) with `n_instances` analytic objects.

    Poses follow dataset.py:216-226: yaw U(0,2pi), elevation U(10,80) deg, z U(0.6,2.0).  x/y are drawn
    inside the view frustum (the reference's U(-0.3,0.3) at z>=0.6 is inside it as well).
    Returns dict(depth uint16 [H,W], masks bool [n,H,W], cats list[str], RTs [n,4,4], diags [n]).
    """
    rng = np.random.default_rng(1000 + frame_id)
    depth = np.zeros((height, width), np.float64)
    owner = np.full((height, width), -1, np.int64)
    cats, RTs, diags = [], [], []
    for i in range(n_instances):
        cat = REAL275_CATEGORIES[i % len(REAL275_CATEGORIES)]
        lo, hi = CATEGORY_DIAG[cat]
        diag = rng.uniform(lo, hi)
        R = _rot_yaw_elev(rng.uniform(0, 2 * np.pi), np.deg2rad(rng.uniform(10, 80)))
        z = rng.uniform(0.6, 2.0)
        # spread instances over a 3x2 lattice in the image so that masks rarely occlude
        cx = (i % 3 + 0.5) / 3.0 * width + rng.uniform(-30, 30)
        cy = (i // 3 + 0.5) / 2.0 * height + rng.uniform(-30, 30)
        t = np.array([(cx - K[0, 2]) / K[0, 0] * z, (cy - K[1, 2]) / K[1, 1] * z, z])
        # enough surface samples to fill every covered pixel a few times over
        px_area = (diag * K[0, 0] / z) ** 2
        n_s = int(min(max(12 * px_area, 40000), 1_500_000))
        P = _primitive_surface(cat, diag, rng, n_s) @ R.T + t
        u = np.round(P[:, 0] / P[:, 2] * K[0, 0] + K[0, 2]).astype(np.int64)
        v = np.round(P[:, 1] / P[:, 2] * K[1, 1] + K[1, 2]).astype(np.int64)
        ok = (u >= 0) & (u < width) & (v >= 0) & (v < height) & (P[:, 2] > 0)
        u, v, zz = u[ok], v[ok], P[ok, 2]
        order = np.argsort(-zz)  # far first so that the nearest sample wins the pixel
        u, v, zz = u[order], v[order], zz[order]
        inst_depth = np.zeros((height, width))
        inst_depth[v, u] = zz
        win = (inst_depth > 0) & ((depth == 0) | (inst_depth < depth))
        depth[win] = inst_depth[win]
        owner[win] = i
        RT = np.eye(4)
        RT[:3, :3] = R
        RT[:3, 3] = t
        cats.append(cat)
        RTs.append(RT)
        diags.append(diag)
    if noise_mm > 0:
        depth[depth > 0] += rng.uniform(-noise_mm, noise_mm, int((depth > 0).sum())) * 1e-3
    depth_mm = np.round(depth * 1000.0).astype(np.uint16)
    masks = np.stack([owner == i for i in range(n_instances)])
    return dict(depth=depth_mm, masks=masks, cats=cats, RTs=np.stack(RTs), diags=np.array(diags))


def unit_descriptors(n: int, dim: int = 1024, seed: int = 0) -> np.ndarray:
    """Stand-in for DINOv2 key-point descriptors (backbone out of scope): seeded, L2-normalised."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, dim)).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return d
