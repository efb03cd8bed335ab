"""Check the C++ claim-order planner: replay its order under strict in-order claiming with the 'true' cost model."""
import sys, time, heapq, numpy as np
sys.path.insert(0, '.')
import bench
from painty_b200 import assets, api
n = 10000; seg = 64; rows, cols = 2160, 3840
rec, cx, cy, th, radii = bench.build_workload(n, rows=rows)
def tmodel(r): return np.interp(r, [11, 30, 64, 112, 151, 200], [5.3, 6.0, 10, 16, 27, 45]) * 1e-6
R = rec["radius"].astype(float); M = rec["n_imprints"].astype(int)
geo = [assets.footprint_geometry(float(r)) for r in R]
side = np.array([g[3] for g in geo], np.int32)
fp = {}
def nact(r):
    k = int(np.ceil(r))
    if k not in fp: fp[k] = int((assets.baked_footprint(float(r)) > 0).sum())
    return fp[k]
NA = np.array([nact(r) for r in R])
cls = np.where(NA <= 256, 1, np.where(NA <= 4096, 16, 17))
run = np.zeros(n, np.int32); j = 0
for s in range(1, n):
    if cls[s] != cls[s - 1]: j += 1
    run[s] = j
nruns = j + 1
print("runs", nruns, [int((run == q).sum()) for q in range(nruns)])
sf, sl, so, ps, pn = api.plan_segments(rows, cols, rec["first_imprint"], M, side, R, cx, cy, seg, True)
t = time.time()
order = api.plan_claim_order(rows, cols, rec["first_imprint"], M, side, R, cx, cy, np.zeros(n, np.int32), run, 5.2 + 0.81e-3 * NA, [[9] * nruns], seg, True)
print("C++ planner (segments + claim order) %.3fs" % (time.time() - t))
assert sorted(order) == list(range(n))
pos = np.empty(n, int); pos[order] = np.arange(n)
owner = np.repeat(np.arange(n), np.diff(sf))
assert all(pos[ps[i]] < pos[owner[g]] for g in range(len(owner)) for i in range(so[g], so[g + 1])), "not topological"
assert all(run[order[i]] <= run[order[i + 1]] for i in range(n - 1)), "runs out of sequence"
nseg = np.diff(sf)
def ready(s, k, prog):
    g = sf[s] + k
    return all(prog[ps[i]] >= pn[i] for i in range(so[g], so[g + 1]))
true = np.array([tmodel(r) for r in R])
def replay(order_run, slots):
    prog_done = 1 << 30
    state = [None] * slots; ev = []; t = 0.0; pos = 0
    def dur(s, k): return max(min(sl[s], M[s] - k * sl[s]), 0) * true[s]
    def try_start(i):
        nonlocal pos
        st = state[i]
        if st is None:
            if pos >= len(order_run): return
            state[i] = [order_run[pos], 0, False]; pos += 1; st = state[i]
        if not st[2] and ready(st[0], st[1], prog):
            st[2] = True; heapq.heappush(ev, (t + dur(st[0], st[1]), i))
    for i in range(slots): try_start(i)
    while ev:
        t, i = heapq.heappop(ev)
        s, k, _ = state[i]; k += 1
        if k >= nseg[s]: prog[s] = prog_done; state[i] = None
        else: prog[s] = k; state[i] = [s, k, False]
        for j in range(slots):
            if state[j] is None or not state[j][2]: try_start(j)
    assert pos == len(order_run) and all(x is None for x in state), "deadlock in replay"
    return t
for name, od in (("submission", np.arange(n)), ("planned", order)):
    prog = np.zeros(n, np.int64); tot = 0.0
    for q in range(nruns):
        tot += replay([int(s) for s in od if run[s] == q], 9)
    print(name, "order: %.2f s" % tot)
