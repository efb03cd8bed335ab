"""Model of config 4 (strong scaling, one 16K^2 canvas) with the library's planner and the round-2 cost curve: one batch of
N strokes of the benchmark's generator (benchmarks/config4_16k.py), 1/2/4/8 GPUs. CPU only.
usage: sched_config4.py [N_STROKES=100000]"""
import sys, os, importlib.util, numpy as np
sys.path.insert(0, '.')
from painty_b200 import api, assets
spec = importlib.util.spec_from_file_location("c4", "benchmarks/config4_16k.py"); c4 = importlib.util.module_from_spec(spec); spec.loader.exec_module(c4)
nst = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
rows = cols = 16384
R, K, S, first, nv, path = c4.make_strokes(nst, rows, cols)
rec, cx, cy, th = c4.batch_records(R, K, S, first, nv, path, 0, nst)
F, M = rec["first_imprint"].astype(np.int64), rec["n_imprints"].astype(np.int64)
n = len(F)
geo = {}
def g(r):
    kk = int(np.ceil(r))
    if kk not in geo:
        gg = assets.footprint_geometry(float(r)); geo[kk] = (gg[3], int((assets.baked_footprint(float(r)) > 0).sum()))
    return geo[kk]
side = np.array([g(r)[0] for r in R], np.int32); NA = np.array([g(r)[1] for r in R])
cls = np.where(NA <= 256, 1, np.where(NA <= 4096, 16, 17))
cost = np.interp(NA, [146, 1107, 5081, 15200, 20319, 27507], [6.3, 4.7, 4.8, 5.8, 7.6, 13.0])  # measured r02, us per imprint
print("strokes", n, "imprints", int(M.sum()), "serial sum %.1f s" % ((cost * M).sum() * 1e-6))
base = None
y0 = np.array([min(max(int(cy[a]), 0), rows - 1) if m else 0 for a, m in zip(F, M)])
ymin = np.minimum.reduceat(cy, np.minimum(F, len(cy) - 1)); ymax = np.maximum.reduceat(cy, np.minimum(F, len(cy) - 1))
for world in (1, 2, 4, 8):
    rpb = (rows + world - 1) // world
    ex = np.minimum(y0 // rpb, world - 1).astype(np.int32)
    mm = (side - 1) // 2 + R + 2.0
    remote = ((ymin - mm < ex * rpb) | (ymax + mm > np.minimum((ex + 1) * rpb, rows) - 1)).astype(np.uint8)
    run = np.zeros(n, np.int32); slots = [[] for _ in range(world)]; last = [None] * world
    for s in range(n):
        p = ex[s]
        if last[p] != cls[s]: slots[p].append(9 if cls[s] != 1 else 148); last[p] = cls[s]
        run[s] = len(slots[p]) - 1
    _, mk = api.plan_claim_order(rows, cols, F, M, side, R, cx, cy, ex, run, cost, slots, 64, True, return_makespan=True)
    mk *= 1e-6
    if base is None: base = mk
    per_rank = [(cost * M)[ex == r].sum() * 1e-6 / 9 for r in range(world)]
    print("world %d: straddlers %.1f%%  model makespan %.2f s (%.2fx)   slot-bound per rank (serial/9 slots): max %.2f s" % (
        world, 100 * remote.mean(), mk, base / mk, max(per_rank)))
