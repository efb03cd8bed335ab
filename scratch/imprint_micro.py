"""Latency per imprint of a single N-imprint stroke: python scratch/imprint_micro.py R1,R2,.. [N] [THETA] (GPU).
Kernel time only: best of 4 repetitions of the imprint launch between two events on the context stream."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from painty_b200 import api, assets
ctx = api.Context(0, api.F32)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
rows, cols = 2160, 3840
cv = api.Canvas(ctx, rows, cols)
radii = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "11,30,64,112,150".split(","))]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.79
for r in radii:
    r = assets.snap_to_safe_radius(r)
    br = api.FootprintBrush(ctx, r)
    br.dip(([.3, .2, .1], [.2, .4, .3]))
    cx = np.linspace(600, 600 + n, n); cy = np.linspace(700, 700 + 0.3 * n, n); th = np.full(n, theta)
    best = 1e30
    for rep in range(5):
        cv.clear(); br.updateSnapshot(cv); ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); br.imprint_batch(cv, cx, cy, th); e1.record(stream); ctx.synchronize()
        if rep: best = min(best, e0.elapsed_time(e1))
    v, a = br.counters()
    g = assets.footprint_geometry(r)
    print(f"r={r} side={g[3]} cells/imprint={a/(5*n):.0f}  {best:.2f} ms  {best*1e3/n:.2f} us/imprint  {a/5/best/1e3:.1f} M active px/s", flush=True)
