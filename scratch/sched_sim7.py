"""What-if: list scheduling with critical-path priority (bottom level) instead of first-ready."""
import sys, heapq, numpy as np
sys.path.insert(0, '.')
import bench
from painty_b200 import assets, api
n = 10000; seg = 64; rows, cols = 2160, 3840
rec, cx, cy, th, radii = bench.build_workload(n, rows=rows)
def tmodel(r): return np.interp(r, [11, 30, 64, 112, 151, 200], [5.3, 6.0, 10, 16, 27, 45]) * 1e-6
R = rec["radius"].astype(float); M = rec["n_imprints"].astype(int)
side = np.array([assets.footprint_geometry(float(r))[3] for r in R], np.int32)
sf, sl, so, ps, pn = api.plan_segments(rows, cols, rec["first_imprint"], M, side, R, cx, cy, seg, True)
nseg = np.diff(sf); nsegs = int(sf[-1])
owner = np.repeat(np.arange(n), nseg)
true = np.array([tmodel(r) for r in R])
segdur = np.array([max(min(sl[owner[g]], M[owner[g]] - (g - sf[owner[g]]) * sl[owner[g]]), 0) * true[owner[g]] for g in range(nsegs)])
# bottom level per segment: longest path to the end through (next own segment) and (waiting segments)
succ = [[] for _ in range(nsegs)]
for g in range(nsegs):
    for i in range(so[g], so[g + 1]):
        succ[sf[ps[i]] + pn[i] - 1].append(g)
bl = np.zeros(nsegs)
for g in range(nsegs - 1, -1, -1):
    b = 0.0
    if g + 1 < sf[owner[g] + 1]: b = bl[g + 1]
    for q in succ[g]: b = max(b, bl[q])
    bl[g] = segdur[g] + b
prio = bl[sf[:-1]]   # stroke priority = bottom level of its first segment
print("critical path %.2f" % bl.max())
pred_strokes = [set(int(x) for x in ps[so[sf[s]]:so[sf[s + 1]]]) for s in range(n)]
def ready(s, k, prog):
    g = sf[s] + k
    return all(prog[ps[i]] >= pn[i] for i in range(so[g], so[g + 1]))
def simulate(a, b, slots, window, use_prio):
    prog = np.zeros(n, np.int64); prog[:a] = 1 << 30
    claimed = np.zeros(n, bool); claimed[:a] = True
    state = [None] * slots; ev = []; t = 0.0; pending = list(range(a, b)); cnt = 0
    def try_start(i):
        nonlocal cnt
        st = state[i]
        if st is None:
            pick = None; c = 0; best = -1
            for s in pending:
                if c >= window: break
                c += 1
                if all(claimed[p] for p in pred_strokes[s]) and ready(s, 0, prog):
                    if not use_prio: pick = s; break
                    if prio[s] > best: best = prio[s]; pick = s
            if pick is None: return
            pending.remove(pick); claimed[pick] = True; cnt += 1; state[i] = [pick, 0, False]; st = state[i]
        if not st[2] and ready(st[0], st[1], prog):
            st[2] = True; heapq.heappush(ev, (t + segdur[sf[st[0]] + st[1]], i))
    for i in range(slots): try_start(i)
    while ev:
        t, i = heapq.heappop(ev)
        s, k, _ = state[i]; k += 1
        if k >= nseg[s]: prog[s] = 1 << 30; state[i] = None
        else: prog[s] = k; state[i] = [s, k, False]
        for j in range(slots):
            if state[j] is None or not state[j][2]: try_start(j)
    assert cnt == b - a
    return t
h = n // 2
for window, up in ((64, False), (64, True), (256, True), (1024, True)):
    print("window %d prio %s: big %.2f small %.2f" % (window, up, simulate(0, h, 9, window, up), simulate(h, n, 9, window, up)))
