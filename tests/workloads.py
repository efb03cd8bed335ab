"""Seeded synthetic workloads shared by tests, golden generation and bench.py (no oracle import here)."""
import math

import numpy as np


def gui_stroke_imprints(cpu_or_none, pts):
    """painty_gui mouse-move expansion (DigitalCanvas.cxx:107-123) evaluated with the given CPU checker's
    Catmull-Rom (used to cross-check the product's pb_expand_stroke(mode=1))."""
    c = cpu_or_none
    cx, cy, th = [], [], []
    path = []
    for p in pts:
        path.append(np.array(p, dtype=np.float64))
        if len(path) >= 2:
            n = len(path)
            p0, p1, p2 = path[max(0, n - 3)], path[max(0, n - 2)], path[n - 1]
            dist = math.sqrt((p2[0] - p1[0]) * (p2[0] - p1[0]) + (p2[1] - p1[1]) * (p2[1] - p1[1]))
            for pd in range(1, int(dist) + 1):
                t = pd / dist
                d = c.catmull_rom(p0, p1, p2, p2, t, True)
                q = c.catmull_rom(p0, p1, p2, p2, t)
                cx.append(q[0])
                cy.append(q[1])
                th.append(math.atan2(d[1], d[0]))
    return np.array(cx), np.array(cy), np.array(th)


def km_random_planes(rows, cols, seed=42, edge_cases=False, wet=True):
    """SURVEY.md §8d config 5 planes. AoS f64: K,S,R0 [rows,cols,3], V [rows,cols].
    roofline variant: every pixel wet, V~U(.05,.95), K log-U[1e-3,4.32], S log-U[1e-3,1.21], R0~U(.02,.98).
    edge_cases: 10 % V==0, K,S log-uniform down to 1e-9, 1 % S==0, 0.1 % K==0 (NaN in the reference)."""
    rng = np.random.default_rng(seed)
    n = rows * cols
    lo = 1e-9 if edge_cases else 1e-3
    K = np.exp(rng.uniform(np.log(lo), np.log(4.32), (n, 3)))
    S = np.exp(rng.uniform(np.log(lo), np.log(1.21), (n, 3)))
    V = rng.uniform(0.05, 0.95, n)
    R0 = rng.uniform(0.02, 0.98, (n, 3))
    if edge_cases:
        V[rng.random(n) < 0.10] = 0.0
        S[rng.random((n, 3)) < 0.01] = 0.0
        K[rng.random((n, 3)) < 0.001] = 0.0
        V[rng.random(n) < 0.02] *= 40.0  # thick layers: coth -> 1 branch
    return K.reshape(rows, cols, 3), S.reshape(rows, cols, 3), V.reshape(rows, cols), R0.reshape(rows, cols, 3)


def sbr_strokes(rows, cols, n_strokes, seed=1234, sizes=(80, 60, 30, 20), safe_radius=None, palette=None,
                max_points=20, min_points=5):
    """Synthetic sbr_painter-shaped stroke list: n_strokes split over the brush
    sizes (image px at a 1024-wide image, scaled to the canvas), radius ~ U[0.35,0.5]*size*scale snapped to an
    OOB-free radius, 5..20 control points spaced 0.25*radius with a smooth heading random walk, paint = one of 5
    random convex mixes of the palette, grouped by colour like PictureTargetSbrPainter.cxx:366-378.
    Returns list of dict(radius, K, S, path[n,2])."""
    rng = np.random.default_rng(seed)
    scale = cols / 1024.0
    if palette is None:
        pk = np.exp(rng.uniform(np.log(1e-2), np.log(2.0), (14, 3)))
        ps = np.exp(rng.uniform(np.log(1e-2), np.log(1.0), (14, 3)))
    else:
        pk, ps = palette
    mixes = []
    for _ in range(5):
        w = rng.dirichlet(np.ones(len(pk)))
        mixes.append(((w[:, None] * pk).sum(0), (w[:, None] * ps).sum(0)))
    strokes = []
    per = [n_strokes // len(sizes)] * len(sizes)
    per[0] += n_strokes - sum(per)
    for size, cnt in zip(sizes, per):
        batch = []
        for _ in range(cnt):
            r = rng.uniform(0.35, 0.5) * size * scale
            r = max(1.0, float(r))
            if safe_radius is not None:
                r = float(safe_radius(round(r)))
            npts = int(rng.integers(min_points, max_points + 1))
            step = 0.25 * r
            p = np.array([rng.uniform(0, cols), rng.uniform(0, rows)])
            ang = rng.uniform(0, 2 * np.pi)
            pts = [p.copy()]
            for _k in range(npts - 1):
                ang += 0.3 * rng.normal() * 0.5
                p = p + step * np.array([np.cos(ang), np.sin(ang)])
                pts.append(p.copy())
            ci = int(rng.integers(0, 5))
            batch.append((ci, dict(radius=r, K=mixes[ci][0].copy(), S=mixes[ci][1].copy(), path=np.array(pts))))
        batch.sort(key=lambda t: t[0])  # grouped by colour index (std::map order)
        strokes.extend(b for _, b in batch)
    return strokes
