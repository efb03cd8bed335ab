"""One texture-brush batch of the config-3 kind (8K canvas, canvas-pattern substrate, dictionary textures) for ncu captures
of texture_kernel: python scratch/texture_only.py [N_STROKES=2000]  (GPU)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from painty_b200 import api, assets
from tests.workloads import sbr_strokes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
rows, cols = 4320, 7680
ctx = api.Context(0, api.F32)
pk, ps = assets.palette("lindemeier_measured")
strokes = sbr_strokes(rows, cols, n, seed=4321, palette=(pk, ps))
tex = assets.brush_textures()
dic = api.TextureBrushDictionary([t[1] for t in tex], [t[2] for t in tex], [t[3].shape[0] for t in tex], [t[3].shape[1] for t in tex])
cv = api.Canvas(ctx, rows, cols)
cv.setBackground(assets.canvas_pattern(rows, cols))
tb = api.TextureBrush(ctx)
ids = [tb.addTexture(t[3]) for t in tex]
rng = np.random.default_rng(1)
rec = np.zeros(n, dtype=api.TSTROKE_DTYPE); verts = []; first = 0
for i, s in enumerate(strokes):
    cand = dic.lookup(s["path"], 2.0 * s["radius"])[2]
    rec[i] = (s["radius"], s["K"], s["S"], 0.05, first, len(s["path"]), ids[int(cand[int(rng.integers(0, len(cand)))])])
    first += len(s["path"]); verts.append(s["path"])
verts = np.concatenate(verts)
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0 = tb.counters()
    e0.record(stream); tb.stroke_batch(cv, rec, verts); e1.record(stream); ctx.synchronize()
    px = tb.counters() - p0
    print("rep %d: %d strokes %.2f ms, %d stroke-px, %.2f G stroke-px/s" % (rep, n, e0.elapsed_time(e1), px, px / e0.elapsed_time(e1) / 1e6), flush=True)
