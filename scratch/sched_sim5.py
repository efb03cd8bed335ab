"""What-if: host-side list scheduling produces the CLAIM ORDER of the queue (a topological order); device claims in that order."""
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
nseg = np.diff(sf)
pred_strokes = [set(int(x) for x in ps[so[sf[s]]:so[sf[s + 1]]]) for s in range(n)]
def ready(s, k, prog):
    g = sf[s] + k
    return all(prog[ps[i]] >= pn[i] for i in range(so[g], so[g + 1]))
def simulate(a, b, slots, per_imprint, order=None, window=64):
    """order given: claim strictly in that order. order None: ready-first look-ahead among strokes whose predecessor
    strokes are all claimed; returns (makespan, claim order)."""
    prog = np.zeros(n, np.int64); prog[:a] = 1 << 30
    claimed = np.zeros(n, bool); claimed[:a] = True
    state = [None] * slots; ev = []; t = 0.0; out = []
    pos = 0; pending = list(range(a, b))
    def dur(s, k): return max(min(sl[s], M[s] - k * sl[s]), 0) * per_imprint[s]
    def try_start(i):
        nonlocal pos
        st = state[i]
        if st is None:
            pick = None
            if order is not None:
                if pos < len(order): pick = order[pos]; pos += 1
            else:
                c = 0
                for s in pending:
                    if c >= window: break
                    c += 1
                    if all(claimed[p] for p in pred_strokes[s]) and ready(s, 0, prog): pick = s; break
                if pick is None and pending and i == 0:
                    pick = pending[0]      # slot 0 never idles on a non-ready head (keeps the order total)
                if pick is not None: pending.remove(pick)
            if pick is None: return
            claimed[pick] = True; out.append(pick); state[i] = [pick, 0, False]; st = state[i]
        if not st[2] and ready(st[0], st[1], prog):
            st[2] = True; heapq.heappush(ev, (t + dur(st[0], st[1]), i))
    for i in range(slots): try_start(i)
    while ev:
        t, i = heapq.heappop(ev)
        s, k, _ = state[i]; k += 1
        if k >= nseg[s]: prog[s] = 1 << 30; state[i] = None
        else: prog[s] = k; state[i] = [s, k, False]
        for j in range(slots):
            if state[j] is None or not state[j][2]: try_start(j)
    assert len(out) == b - a, (len(out), b - a)
    return t, out
true = np.array([tmodel(r) for r in R])
h = n // 2
rng = np.random.default_rng(0)
for a, b, name in ((0, h, "big"), (h, n, "small")):
    t0, _ = simulate(a, b, 9, true, order=list(range(a, b)))
    t1, od = simulate(a, b, 9, true)
    t2, _ = simulate(a, b, 9, true, order=od)
    noisy = true * rng.uniform(0.7, 1.3, n)
    _, od_n = simulate(a, b, 9, noisy)
    t3, _ = simulate(a, b, 9, true, order=od_n)
    lin = (4.5e-6 + 0.8e-9 * np.array([assets.footprint_geometry(float(r))[3] ** 2 * 0.145 for r in R]))  # crude linear-in-cells model
    _, od_l = simulate(a, b, 9, lin)
    t4, _ = simulate(a, b, 9, true, order=od_l)
    print("%s: submission order %.2f | dynamic look-ahead %.2f | replay of its order %.2f | order from +-30%% noisy model %.2f | order from crude model %.2f" % (name, t0, t1, t2, t3, t4))
