"""Where one imprint's time goes: SM cycle stamps inside imprint_kernel (pb_fbrush_enable_trace) for a single stroke.
python scratch/imprint_trace.py R1,R2,.. [N] [THETA]  (GPU). Prints, per radius, the mean cycles between the stamps of the
first thread (owns cells) and of the last thread of the CTA (a ring thread): start -> ring done -> pickup/deposit done ->
next list built -> barrier passed, averaged over imprints 8..N."""
import ctypes as C, sys, numpy as np
sys.path.insert(0, '.')
from painty_b200 import api, assets
ctx = api.Context(0, api.F32)
rows, cols = 2160, 3840
cv = api.Canvas(ctx, rows, cols)
radii = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "30,64,112,129,151".split(","))]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 200
theta = float(sys.argv[3]) if len(sys.argv) > 3 else 0.79
lib = api.lib()
for r in radii:
    r = assets.snap_to_safe_radius(r)
    br = api.FootprintBrush(ctx, r)
    br.dip(([.3, .2, .1], [.2, .4, .3]))
    cx = np.linspace(600, 600 + n, n); cy = np.linspace(700, 700 + 0.3 * n, n); th = np.full(n, theta)
    cv.clear(); br.updateSnapshot(cv)
    br.imprint_batch(cv, cx, cy, th)  # warm
    cv.clear(); br.updateSnapshot(cv)
    api._chk(lib.pb_fbrush_enable_trace(br.h, 1))
    br.imprint_batch(cv, cx, cy, th)
    out = np.zeros((256, 2, 8), dtype=np.uint64)
    api._chk(lib.pb_fbrush_read_trace(br.h, out.ctypes.data_as(C.c_void_p)))
    t = out[8:min(n, 256)].astype(np.int64)
    names = ["ring", "process", "build_list", "barrier"]
    for who, label in ((0, "thread 0 (cells)"), (1, "last thread (ring)")):
        d = np.diff(t[:, who, :5], axis=1).mean(axis=0)
        per = np.diff(t[:, who, 0]).mean()
        print("r=%5.1f %-20s " % (r, label) + "  ".join("%s %6.0f" % (nm, v) for nm, v in zip(names, d)) + "   | per imprint %6.0f cycles" % per, flush=True)
