"""Event-driven what-if on top of segment dataflow: in-order ticket queue vs ready-first look-ahead claiming."""
import sys, heapq, numpy as np
sys.path.insert(0, '.')
import bench
from painty_b200 import assets, api
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
seg = 64
rows, cols = 2160, 3840
rec, cx, cy, th, radii = bench.build_workload(n, rows=rows)
def tmodel(r): return np.interp(r, [11, 30, 64, 112, 151, 200], [5.3, 6.0, 10, 16, 27, 45]) * 1e-6
R = rec["radius"].astype(float); M = rec["n_imprints"].astype(int)
side = np.array([assets.footprint_geometry(float(r))[3] for r in R], np.int32)
sf, sl, so, ps, pn = api.plan_segments(rows, cols, rec["first_imprint"], M, side, R, cx, cy, seg, True)
cls = np.array([1 if a <= 256 else (16 if a <= 4096 else 17) for a in [assets.footprint_geometry(float(r))[4] if len(assets.footprint_geometry(float(r))) > 4 else 99999 for r in R]])
def segdur(s, k):
    m = min(sl[s], M[s] - k * sl[s]); return max(m, 0) * tmodel(R[s])
def ready(s, k, prog):
    g = sf[s] + k
    return all(prog[ps[i]] >= pn[i] for i in range(so[g], so[g + 1]))
def run(a, b, slots, window, inorder_slots=1):
    """strokes [a,b) as one launch. window=1: strict in-order claiming."""
    prog = np.zeros(n, int); prog[:a] = 1 << 30
    claimed = np.zeros(n, bool); head = a
    state = [None] * slots   # (stroke, seg, running?)
    ev = []; t = 0.0
    nseg = np.diff(sf)
    def try_start(i):
        nonlocal head
        st = state[i]
        if st is None:
            while head < b and claimed[head]: head += 1
            if head >= b: return
            pick = None
            if window > 1 and i >= inorder_slots:
                c = 0; s = head
                while s < b and c < window:
                    if not claimed[s]:
                        c += 1
                        if ready(s, 0, prog): pick = s; break
                    s += 1
                if pick is None: return
            else:
                pick = head
            claimed[pick] = True; state[i] = [pick, 0, False]; st = state[i]
        if not st[2] and ready(st[0], st[1], prog):
            st[2] = True; heapq.heappush(ev, (t + segdur(st[0], st[1]), i))
    for i in range(slots): try_start(i)
    while ev:
        t, i = heapq.heappop(ev)
        s, k, _ = state[i]
        k += 1
        if k >= nseg[s]: prog[s] = 1 << 30; state[i] = None
        else: prog[s] = k; state[i] = [s, k, False]
        for j in range(slots):
            if state[j] is None or not state[j][2]: try_start(j)
    assert claimed[a:b].all() and all(x is None for x in state)
    return t
# launches: runs of equal class; here passes 0+1 (class 17) and 2+3 (class 16) -> approximate by halves
h = n // 2
for window, io in ((1, 1), (8, 1), (32, 1), (128, 1), (32, 0)):
    ta = run(0, h, 9, window, io); tb = run(h, n, 9, window, io)
    print("window %3d inorder_slots %d: big %.2f small %.2f total %.2f" % (window, io, ta, tb, ta + tb))
print("---- groups of G clusters per stroke in the big launch (segments, look-ahead 32)")
base_tmodel = tmodel
for G, ovh in ((1, 0), (2, 3e-6), (2, 1.5e-6), (3, 3e-6), (3, 1.5e-6)):
    def tmodel(r, G=G, ovh=ovh): return base_tmodel(r) if G == 1 else base_tmodel(r / np.sqrt(G)) + ovh
    for window in (1, 32):
        print("G=%d ovh=%.1fus window=%d: big %.2f" % (G, ovh * 1e6, window, run(0, h, 9 // G, window, 1)))
