"""Model of config 4 (strong scaling, one 16K^2 canvas, 60k strokes) with the library's planner. CPU only."""
import sys, os, numpy as np
sys.path.insert(0, '.')
from painty_b200 import api, assets
from tests.workloads import sbr_strokes
nst = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
rows = cols = 16384
pk, ps_ = assets.palette("lindemeier_measured")
k = 3840.0 / cols
strokes = sbr_strokes(rows, cols, nst, seed=1234, sizes=(80 * k, 60 * k, 30 * k, 20 * k), safe_radius=assets.snap_to_safe_radius, palette=(pk, ps_))
F, M, R, xs, ys = [], [], [], [], []
first = 0
for s in strokes:
    cx, cy, th = api.expand_stroke(s["path"], mode=0)
    F.append(first); M.append(len(cx)); R.append(s["radius"]); first += len(cx); xs.append(cx); ys.append(cy)
cx, cy = np.concatenate(xs), np.concatenate(ys)
F, M, R = np.array(F, np.int64), np.array(M, np.int64), np.array(R, float)
n = len(F)
geo = {}
def g(r):
    kk = int(np.ceil(r))
    if kk not in geo:
        gg = assets.footprint_geometry(float(r)); geo[kk] = (gg[3], int((assets.baked_footprint(float(r)) > 0).sum()))
    return geo[kk]
side = np.array([g(r)[0] for r in R], np.int32); NA = np.array([g(r)[1] for r in R])
cls = np.where(NA <= 256, 1, np.where(NA <= 4096, 16, 17))
cost = 5.2 + 0.81e-3 * NA
base = None
for world in (1, 2, 4, 8):
    rpb = (rows + world - 1) // world
    ex = np.zeros(n, np.int32); remote = np.zeros(n, np.uint8)
    for s in range(n):
        a, m = F[s], M[s]
        if m == 0: continue
        ex[s] = min(min(max(int(cy[a]), 0), rows - 1) // rpb, world - 1)
        mm = (side[s] - 1) // 2 + R[s] + 2.0
        lo = max(0, int(np.floor(cy[a:a+m].min() - mm))); hi = min(rows - 1, int(np.ceil(cy[a:a+m].max() + mm)))
        remote[s] = lo < ex[s] * rpb or hi > min((ex[s] + 1) * rpb, rows) - 1
    run = np.zeros(n, np.int32); slots = [[] for _ in range(world)]; last = [None] * world
    for s in range(n):
        p = ex[s]
        if last[p] != cls[s]: slots[p].append(9 if cls[s] != 1 else 148); last[p] = cls[s]
        run[s] = len(slots[p]) - 1
    out = []
    for single in (remote, None):
        _, mk = api.plan_claim_order(rows, cols, F, M, side, R, cx, cy, ex, run, cost, slots, 64, True, single=single, return_makespan=True)
        out.append(mk * 1e-6)
    if base is None: base = out[1]
    print("world %d: straddlers %.1f%%  model makespan: one-segment straddlers %.2f s (%.2fx), segmented %.2f s (%.2fx)" % (
        world, 100 * remote.mean(), out[0], base / out[0], out[1], base / out[1]))
