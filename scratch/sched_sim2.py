"""What-if: a 'fast lane' of G cooperating clusters for strokes with little slack (two in-order queues)."""
import sys, heapq, numpy as np
sys.path.insert(0, '.')
import bench
from painty_b200 import assets, api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
rows, cols = 2160, 3840
rec, cx, cy, th, radii = bench.build_workload(n, rows=rows)
def tmodel(r): return np.interp(r, [11, 30, 64, 112, 151, 200], [5.3, 6.0, 10, 16, 27, 45]) * 1e-6
def tG(r, G, ovh=3e-6): return tmodel(r) if G == 1 else tmodel(r / np.sqrt(G)) + ovh
box = np.zeros((n, 4), np.int32); alw = np.zeros((n, 4), np.int32)
R = rec["radius"].astype(float); M = rec["n_imprints"].astype(int)
for s in range(n):
    a, m = int(rec["first_imprint"][s]), int(M[s]); r = R[s]
    side = assets.footprint_geometry(r)[3]; wr = (side - 1) // 2
    def reg(mg):
        mm = wr + mg + 2
        return (max(0, int(np.floor(cx[a:a+m].min() - mm))), max(0, int(np.floor(cy[a:a+m].min() - mm))),
                min(cols - 1, int(np.ceil(cx[a:a+m].max() + mm))), min(rows - 1, int(np.ceil(cy[a:a+m].max() + mm))))
    box[s] = reg(0); alw[s] = reg(r)
offs, preds_flat = api.plan_dependencies(rows, cols, box, alw)
preds = [preds_flat[offs[s]:offs[s+1]] for s in range(n)]
succs = [[] for _ in range(n)]
for s in range(n):
    for p in preds[s]: succs[int(p)].append(s)
d1 = M * np.array([tmodel(r) for r in R])
def levels(dur):
    tl = np.zeros(n); bl = np.zeros(n)
    for s in range(n): tl[s] = max((tl[p] + dur[p] for p in preds[s]), default=0.0)
    for s in range(n - 1, -1, -1): bl[s] = dur[s] + max((bl[q] for q in succs[s]), default=0.0)
    return tl, bl
tl, bl = levels(d1); CP = (tl + bl).max()
print("serial %.2f  CP %.2f" % (d1.sum(), CP))
def sim(lane, dur, nfast, nslow):
    """lane[s] in {0 slow, 1 fast}; each lane is an in-order queue served by its own slots."""
    free = [[0.0] * max(nslow, 1), [0.0] * max(nfast, 1)]
    for f in free: heapq.heapify(f)
    fin = np.zeros(n)
    for s in range(n):
        f = free[lane[s]]; t0 = heapq.heappop(f)
        start = max([t0] + [fin[p] for p in preds[s]]); fin[s] = start + dur[s]; heapq.heappush(f, fin[s])
    return fin.max()
print("baseline 9 slots: %.2f" % sim(np.zeros(n, int), d1, 0, 9))
for G, nfast in ((4, 1), (2, 1), (2, 2), (4, 2), (3, 1), (3, 2)):
    nslow = 9 - G * nfast
    dG = M * np.array([tG(r, G) for r in R])
    for theta in (0.5, 0.7, 0.8, 0.9):
        for rmin in (60, 90):
            lane = ((tl + bl >= theta * CP) & (R >= rmin)).astype(int)
            dur = np.where(lane == 1, dG, d1)
            t2, b2 = levels(dur)
            print("G=%d fast=%d slow=%d theta=%.1f rmin=%d: fast strokes %d  new CP %.2f  makespan %.2f" % (
                G, nfast, nslow, theta, rmin, lane.sum(), (t2 + b2).max(), sim(lane, dur, nfast, nslow)))
print("---- per-pass uniform G (separate launches)")
def sim_range(a, b, dur, slots):
    free = [0.0] * slots; heapq.heapify(free); fin = {}
    for s in range(a, b):
        t0 = heapq.heappop(free)
        start = max([t0] + [fin[int(p)] for p in preds[s] if int(p) >= a]); fin[s] = start + dur[s]; heapq.heappush(free, fin[s])
    return max(fin.values())
q = n // 4
for ps in range(4):
    a, b = ps * q, (ps + 1) * q
    print("pass", ps, "radius %.0f..%.0f" % (R[a:b].min(), R[a:b].max()), "serial %.2f" % d1[a:b].sum())
    for G in (1, 2, 3, 4):
        for ovh in (3e-6, 2e-6):
            dG = M * np.array([tG(r, G, ovh) for r in R])
            print("   G=%d ovh=%.0fus slots=%d makespan %.2f" % (G, ovh * 1e6, 9 // G, sim_range(a, b, dG, 9 // G)))
