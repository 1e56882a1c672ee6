"""Independent float64 numpy implementation of the SHOT352 spec, written
separately from oracle/shot_oracle.cpp and used only to cross-check it (the reference pins nothing for
this path).  Brute-force neighbours, centred float64 covariance, numpy eigh; small clouds only."""
import numpy as np

RAD_45, RAD_90, RAD_135, RAD_7_8 = np.pi / 4, np.pi / 2, 3 * np.pi / 4, 7 * np.pi / 8


def neighbours(pc32, i, r):
    d = pc32 - pc32[i]
    d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]      # float32, FLANN order
    idx = np.nonzero(d2 < np.float32(r * r))[0]
    return idx, d2[idx]


def normal(pc32, i, r):
    idx, _ = neighbours(pc32, i, r)
    if len(idx) < 3:
        return np.full(3, np.nan)
    q = pc32[idx].astype(np.float64)
    c = np.cov(q.T, bias=True)
    w, v = np.linalg.eigh(c)
    n = v[:, 0]
    if np.dot(-pc32[i].astype(np.float64), n) < 0:
        n = -n
    return n


def lrf(pc32, i, r):
    idx, d2 = neighbours(pc32, i, r)
    keep = ~np.all(pc32[idx] == pc32[i], axis=1)
    idx, d2 = idx[keep], d2[keep]
    if len(idx) < 5:
        return None
    v = (pc32[idx] - pc32[i]).astype(np.float64)
    w = r - np.sqrt(d2.astype(np.float64))
    m = (v * w[:, None]).T @ v / w.sum()
    ev, evec = np.linalg.eigh(m)
    x, z = evec[:, 2].copy(), evec[:, 0].copy()
    info = {}
    for name, ax in (("x", x), ("z", z)):
        plus = int((v @ ax >= 0).sum())
        s = 2 * plus - len(idx)
        info[name] = s
        if s < 0:
            ax *= -1
        # s == 0 (tie) depends on the search order in PCL: callers skip such points
    y = np.cross(z, x)
    info["gap"] = (ev[1] / ev[2], ev[0] / ev[1])
    return np.stack([x, y, z]), info


def shot352(pc32, normals, i, r, rf):
    idx, d2 = neighbours(pc32, i, r)
    if len(idx) < 5:
        return np.full(352, np.nan), 0.0
    h = np.zeros(352)
    margin = np.inf        # distance of the closest neighbour to any hard bin boundary (relative)
    fx, fy, fz = rf
    for k, dd in zip(idx, d2):
        n = normals[k]
        if not np.all(np.isfinite(n)):
            continue
        cosd = float(np.clip(np.dot(n, fz), -1, 1))
        b = (1 + cosd) * 10 / 2
        delta = (pc32[k] - pc32[i]).astype(np.float64)
        d = float(np.sqrt(np.float64(dd)))
        if d < 1e-15:
            continue
        xf, yf, zf = float(delta @ fx), float(delta @ fy), float(delta @ fz)
        bit4 = 1 if (yf > 0 or (yf == 0 and xf < 0)) else 0
        bit3 = (1 - bit4) if (xf > 0 or (xf == 0 and yf > 0)) else bit4
        di = (bit4 << 4) + (bit3 << 3)
        if xf * yf > 0 or xf == 0:
            di += 0 if abs(xf) >= abs(yf) else 4
        else:
            di += 4 if abs(xf) > abs(yf) else 0
        di += 1 if zf > 0 else 0
        di += 2 if d > r / 2 else 0
        margin = min(margin, abs(d - r / 2) / r, abs(zf) / r, abs(xf) / r, abs(yf) / r, abs(abs(xf) - abs(yf)) / r,
                     abs((b + 0.5) - np.floor(b + 0.5) - 0.0) / 10 if False else np.inf)
        frac = (b + 0.5) - np.floor(b + 0.5)
        margin = min(margin, min(frac, 1 - frac) / 10)
        step = int(np.floor(b + 0.5))
        vol = di * 11
        b -= step
        w = 1 - abs(b)
        if b > 0:
            h[vol + ((step + 1) % 10)] += b
        else:
            h[vol + ((step - 1 + 10) % 10)] += -b
        if d > r / 2:
            rho = (d - 3 * r / 4) / (r / 2)
            if d > 3 * r / 4:
                w += 1 - rho
            else:
                w += 1 + rho
                h[(di - 2) * 11 + step] -= rho
        else:
            rho = (d - r / 4) / (r / 2)
            if d < r / 4:
                w += 1 + rho
            else:
                w += 1 - rho
                h[(di + 2) * 11 + step] += rho
        inc = np.arccos(np.clip(zf / d, -1, 1))
        if inc > RAD_90 or (abs(inc - RAD_90) < 1e-30 and zf <= 0):
            iota = (inc - RAD_135) / RAD_90
            if inc > RAD_135:
                w += 1 - iota
            else:
                w += 1 + iota
                h[(di + 1) * 11 + step] -= iota
        else:
            iota = (inc - RAD_45) / RAD_90
            if inc < RAD_45:
                w += 1 + iota
            else:
                w += 1 - iota
                h[(di - 1) * 11 + step] += iota
        if yf != 0 or xf != 0:
            az = np.arctan2(yf, xf)
            sel = di >> 2
            a = (az - (-RAD_7_8 + RAD_45 * sel)) / RAD_45
            a = max(-0.5, min(a, 0.5))
            if a > 0:
                w += 1 - a
                h[((di + 4) % 32) * 11 + step] += a
            else:
                w += 1 + a
                h[((di - 4 + 32) % 32) * 11 + step] -= a
        h[vol + step] += w
    return h / np.sqrt((h * h).sum()), margin
