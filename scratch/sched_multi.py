"""Model study of the band-sharded (multi-GPU) schedule with the library's own planner (CPU only).
usage: sched_multi.py WORLD"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from painty_b200 import assets, api
world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = 10000 * world; rows, cols = 2160 * world, 3840; rpb = 2160
_, rec, cx, cy, th, radii = bench.build_workload(n, rows=rows)
R = rec["radius"].astype(float); M = rec["n_imprints"].astype(np.int64); F = rec["first_imprint"].astype(np.int64)
geo = {}
def g(r):
    k = int(np.ceil(r))
    if k not in geo:
        gg = assets.footprint_geometry(float(r)); geo[k] = (gg[3], int((assets.baked_footprint(float(r)) > 0).sum()))
    return geo[k]
side = np.array([g(r)[0] for r in R], np.int32); NA = np.array([g(r)[1] for r in R])
cls = np.where(NA <= 256, 1, np.where(NA <= 4096, 16, 17))
ex = np.zeros(n, np.int32); remote = np.zeros(n, np.uint8)
for s in range(n):
    a, m = F[s], M[s]
    if m == 0: continue
    y0 = min(max(int(cy[a]), 0), rows - 1); ex[s] = min(y0 // rpb, world - 1)
    mm = (side[s] - 1) // 2 + R[s] + 2.0
    lo = max(0, int(np.floor(cy[a:a+m].min() - mm))); hi = min(rows - 1, int(np.ceil(cy[a:a+m].max() + mm)))
    b0 = ex[s] * rpb; b1 = min(b0 + rpb, rows) - 1
    remote[s] = lo < b0 or hi > b1
print("world", world, "strokes", n, "straddlers %.1f%%" % (100 * remote.mean()), "share of imprints %.1f%%" % (100 * M[remote > 0].sum() / M.sum()))
run = np.zeros(n, np.int32); slots = [[] for _ in range(world)]; last = [None] * world
for s in range(n):
    p = ex[s]
    if last[p] != cls[s]: slots[p].append(9 if cls[s] != 1 else 148); last[p] = cls[s]
    run[s] = len(slots[p]) - 1
cost = np.interp(NA, [146, 1107, 5081, 15200, 20319, 27507], [6.3, 4.7, 4.8, 5.8, 7.6, 13.0])  # measured r02, us per imprint
for name, single, c in (("straddlers as one segment (today)", remote, cost),
                        ("straddlers segmented like the rest (ideal)", None, cost),
                        ("one segment + 30% slower straddlers", remote, cost * np.where(remote > 0, 1.3, 1.0)),
                        ("whole strokes everywhere", np.ones(n, np.uint8), cost)):
    _, mk = api.plan_claim_order(rows, cols, F, M, side, R, cx, cy, ex, run, c, slots, 64, True, single=single, return_makespan=True)
    print("%-45s model makespan %.2f s" % (name, mk * 1e-6))
