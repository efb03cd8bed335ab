"""What-if study of the single-GPU 10k-stroke schedule with the library's own planner (CPU only): model makespan for
different slot counts and per-imprint cost curves. usage: sched_whatif.py [N_STROKES]"""
import sys, numpy as np
sys.path.insert(0, '.')
import bench
from painty_b200 import assets, api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rows, cols = 2160, 3840
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
ex = np.zeros(n, np.int32)
print("imprints per class:", {c: int(M[cls == c].sum()) for c in (1, 16, 17)}, "strokes per class", {c: int((cls == c).sum()) for c in (1, 16, 17)})
def runs_for(slot16, slot17):
    run = np.zeros(n, np.int32); slots = [[]]; last = None
    for s in range(n):
        if last != cls[s]:
            slots[0].append({1: 148, 16: slot16, 17: slot17}[cls[s]]); last = cls[s]
        run[s] = len(slots[0]) - 1
    return run, slots
# measured r02 (scratch/imprint_micro.py, 45 degrees): active cells -> us per imprint
xs = [146, 1107, 5081, 15200, 20319, 27507]; ys = [6.3, 4.7, 4.8, 5.8, 7.6, 13.0]
cost_meas = np.interp(NA, xs, ys)
print("serial sum of imprint costs %.2f s" % ((cost_meas * M).sum() * 1e-6))
for name, cost, s16, s17 in (("measured r02 curve, 9/9 slots", cost_meas, 9, 9),
                             ("measured, 18 slots for class 16", cost_meas, 18, 9),
                             ("measured, 36 slots for class 16", cost_meas, 36, 9),
                             ("measured, inf slots", cost_meas, 1000, 1000),
                             ("flat 4.7 us, 9/9", np.full(n, 4.7), 9, 9),
                             ("flat 4.7 us, inf", np.full(n, 4.7), 1000, 1000),
                             ("flat 3.0 us, 9/9", np.full(n, 3.0), 9, 9),
                             ("measured but max 7.6 (fix the 4-cells-per-thread cliff)", np.minimum(cost_meas, 7.6), 9, 9),
                             ("class17 on 8-CTA clusters at 1.5x cost, 18 slots", cost_meas * np.where(cls == 17, 1.5, 1.0), 9, 18)):
    run, slots = runs_for(s16, s17)
    _, mk = api.plan_claim_order(rows, cols, F, M, side, R, cx, cy, ex, run, cost, slots, 64, True, return_makespan=True)
    print("%-60s model makespan %.2f s" % (name, mk * 1e-6))
