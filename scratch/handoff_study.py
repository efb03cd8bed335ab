"""Host study for the on-chip record hand-off: of the canvas pixels an imprint touches, how many were
written one imprint earlier by the same thread / warp / CTA / another CTA of the stroke's cluster, and how many are new?
Uses the library's own hit finder (pb_imprint_hits, exact mode) and the kernel's cell -> (CTA, thread) numbering
(capi.cu: register_footprint tile order, imprint.cu: contiguous chunks per CTA, slot = tid + k * bd). CPU only.
usage: handoff_study.py [R1,R2,..]"""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
from painty_b200 import api, assets
lib = api.lib()
rows, cols = 2160, 3840
P32 = C.POINTER(C.c_int32)

def shape_for(n_active):  # imprint.cu: imprint_plan, latency policy
    if n_active <= 256: return 1, 256
    if n_active <= 4096: return 8, 256
    if n_active <= 8192: return 16, 256
    return 16, 512

def hits(cx, cy, th, wr, mx, my):
    n = len(mx)
    nh = np.zeros(n, np.int32); px = np.zeros(2 * n, np.int32); py = np.zeros(2 * n, np.int32)
    rc = lib.pb_imprint_hits(C.c_double(cx), C.c_double(cy), C.c_double(th), wr, rows, cols, C.c_int64(n), mx.ctypes.data_as(P32),
                             my.ctypes.data_as(P32), 0, C.c_double(0.0), -1, nh.ctypes.data_as(P32), px.ctypes.data_as(P32), py.ctypes.data_as(P32))
    assert rc == 0
    return nh, px.reshape(n, 2), py.reshape(n, 2)

radii = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "30,64,112,151".split(","))]
print("%-6s %-8s %-9s %-7s | %% of an imprint's interactions whose pixel was written one imprint earlier by ..." % ("r", "cells", "cluster", "dtheta"))
print("%-6s %-8s %-9s %-7s | %8s %8s %8s %10s %8s" % ("", "", "", "", "thread", "warp", "CTA", "other CTA", "nobody"))
for r in radii:
    r = assets.snap_to_safe_radius(r)
    fp = assets.baked_footprint(r)
    side = fp.shape[0]; wr = (side - 1) // 2
    ys, xs = np.nonzero(fp > 0)
    key = ((ys >> 2).astype(np.int64) << 40) | ((xs >> 3).astype(np.int64) << 20) | ((ys & 3) << 3) | (xs & 7)
    o = np.argsort(key, kind="stable")
    mx, my = xs[o].astype(np.int32), ys[o].astype(np.int32)
    n = len(mx); csize, bd = shape_for(n)
    per_cta = (n + csize - 1) // csize
    j = np.arange(n)
    cta = j // per_cta; tid = (j % per_cta) % bd
    owner_thread = cta * bd + tid; owner_warp = cta * (bd // 32) + tid // 32
    for dth in (0.0, 0.01, 0.04):
        acc = np.zeros(5); tot = 0
        th, x, y = 0.6, 900.3, 800.7
        prev = None
        for i in range(12):
            nh, px, py = hits(x, y, th, wr, mx, my)
            cur = {}
            for q in range(2):
                sel = nh > q
                pix = py[sel, q].astype(np.int64) * cols + px[sel, q]
                for pxl, t, w, c in zip(pix, owner_thread[sel], owner_warp[sel], cta[sel]):
                    cur[int(pxl)] = (int(t), int(w), int(c))
            if prev is not None:
                for pxl, (t, w, c) in cur.items():
                    pw = prev.get(pxl)
                    if pw is None: acc[4] += 1
                    elif pw[0] == t: acc[0] += 1
                    elif pw[1] == w: acc[1] += 1
                    elif pw[2] == c: acc[2] += 1
                    else: acc[3] += 1
                tot += len(cur)
            prev = cur
            x += np.cos(th); y += np.sin(th); th += dth
        p = 100 * acc / tot
        print("%-6g %-8d %2d x %-4d %-7g | %8.1f %8.1f %8.1f %10.1f %8.1f" % (r, n, csize, bd, dth, p[0], p[1], p[2], p[3], p[4]))
