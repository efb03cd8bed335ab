"""What-if: dependencies tracked per stroke SEGMENT (progress flags) instead of per whole stroke."""
import sys, heapq, numpy as np
sys.path.insert(0, '.')
import bench
from painty_b200 import assets, api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
seg = int(sys.argv[2]) if len(sys.argv) > 2 else 32
rows, cols = 2160, 3840
rec, cx, cy, th, radii = bench.build_workload(n, rows=rows)
def tmodel(r): return np.interp(r, [11, 30, 64, 112, 151, 200], [5.3, 6.0, 10, 16, 27, 45]) * 1e-6
R = rec["radius"].astype(float); M = rec["n_imprints"].astype(int); F = rec["first_imprint"].astype(int)
owner = []; sdur = []; boxes = []; alws = []
first_seg = np.zeros(n + 1, int)
for s in range(n):
    r = R[s]; side = assets.footprint_geometry(r)[3]; wr = (side - 1) // 2
    ns = max(1, (M[s] + seg - 1) // seg); first_seg[s + 1] = first_seg[s] + ns
    for k in range(ns):
        a = F[s] + k * seg; m = min(seg, M[s] - k * seg)
        def reg(mg):
            mm = wr + mg + 2
            return (max(0, int(np.floor(cx[a:a+m].min() - mm))), max(0, int(np.floor(cy[a:a+m].min() - mm))),
                    min(cols - 1, int(np.ceil(cx[a:a+m].max() + mm))), min(rows - 1, int(np.ceil(cy[a:a+m].max() + mm))))
        boxes.append(reg(0)); alws.append(reg(r)); owner.append(s); sdur.append(m * tmodel(r))
boxes = np.array(boxes, np.int32); alws = np.array(alws, np.int32); owner = np.array(owner); sdur = np.array(sdur)
offs, pf = api.plan_dependencies(rows, cols, boxes, alws)
print("segments", len(owner), "edges", len(pf), "cross-stroke edges", int((owner[pf] != np.repeat(owner, np.diff(offs))).sum()))
def sim(slots, ovh=0.0):
    free = [0.0] * slots; heapq.heapify(free); fin = np.zeros(len(owner))
    for s in range(n):
        t = heapq.heappop(free)
        for g in range(first_seg[s], first_seg[s + 1]):
            ps = pf[offs[g]:offs[g + 1]]; ps = ps[owner[ps] != s]
            if len(ps): t = max(t, fin[ps].max())
            t += sdur[g] + ovh; fin[g] = t
        heapq.heappush(free, t)
    return max(free)
for sl in (9, 18, 148): print("seg", seg, "slots", sl, "makespan %.2f" % sim(sl), " with 2us/segment overhead %.2f" % sim(sl, 2e-6))
