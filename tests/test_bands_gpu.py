"""Band sharding on the device (both bands on one GPU): routed texture strokes on band canvases + band compose
reproduce the rows of the single-canvas render bit for bit (FP64 mode) / within 1e-4 (FP32)."""
import numpy as np
import pytest

from painty_b200 import bands

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", [0, 1])
def test_texture_bands_equal_full_canvas(ctx32, ctx64, prec):
    import torch

    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols, world = 300, 260, 3
    r = np.random.default_rng(12)
    strokes = []
    for _ in range(30):
        m = int(r.integers(2, 7))
        p0 = r.uniform(-10, [cols + 10, rows + 10])
        strokes.append(dict(radius=float(r.uniform(3, 25)), K=r.uniform(.05, 1.5, 3), S=r.uniform(.05, 1, 3),
                            scale=float(r.uniform(.05, 1)), path=p0 + np.cumsum(r.normal(0, 10, (m, 2)), axis=0)))
    R0 = np.random.default_rng(3).uniform(0.3, 1.0, (rows, cols, 3))

    def batch(indices):
        rec = np.zeros(len(indices), dtype=api.TSTROKE_DTYPE)
        verts, first = [], 0
        # the fuzzy setRadius state depends on every stroke of the list: resolve it before routing
        radius, eff = 0.0, []
        for s in strokes:
            if not abs(radius - s["radius"]) < 0.5:
                radius = s["radius"]
            eff.append(radius)
        for j, i in enumerate(indices):
            s = strokes[i]
            rec[j] = (eff[i], s["K"], s["S"], s["scale"], first, len(s["path"]), 0)
            first += len(s["path"])
            verts.append(s["path"])
        return rec, np.concatenate(verts)

    full = api.Canvas(ctx, rows, cols)
    full.setBackground(R0)
    tb = api.TextureBrush(ctx)
    rec, verts = batch(list(range(len(strokes))))
    tb.stroke_batch(full, rec, verts)
    want_state, want_R = full.download("KSV"), full.compose()

    routed = bands.route_texture_strokes(strokes, rows, world)
    dt = torch.float64 if prec else torch.float32
    parts = []
    for rank, (b, e) in enumerate(bands.band_ranges(rows, world)):
        cv = api.Canvas(ctx, rows, cols, band=(b, e, 0))
        cv.setBackground(R0[b:e])
        tbr = api.TextureBrush(ctx)
        rec, verts = batch(routed[rank])
        tbr.stroke_batch(cv, rec, verts)
        st = cv.download("KSV")
        for k in "KSV":
            assert np.array_equal(st[k], want_state[k][b:e]), (rank, k)
        out = torch.empty((3, (e - b) * cols), dtype=dt, device="cuda")
        torch.cuda.synchronize()
        cv.compose_band_device(out.data_ptr(), (e - b) * cols)
        ctx.synchronize()
        parts.append(out)
    img = torch.cat([p.reshape(3, -1, cols) for p in parts], dim=1).permute(1, 2, 0).cpu().numpy().astype(np.float64)
    assert np.abs(img - want_R).max() <= (1e-12 if prec else 1e-6)


@pytest.mark.parametrize("cols", [260, 261])
@pytest.mark.parametrize("prec", [0, 1])
def test_compose_gather_assembles_the_image(ctx32, ctx64, prec, cols):
    """pb_canvas_compose_gather: band canvases compose straight into the assembled image(s) at their rows — equal to
    Renderer::compose of the whole canvas bit for bit; cols = 261 takes the unaligned (scalar store) path."""
    from painty_b200 import api
    from tests.workloads import km_random_planes

    ctx = [ctx32, ctx64][prec]
    rows, world = 301, 3
    K, S, V, R0 = km_random_planes(rows, cols, seed=21, edge_cases=False)
    V[::5] = 0.0
    full = api.Canvas(ctx, rows, cols)
    full.setBackground(R0)
    full.upload_layer(K, S, V)
    want = full.compose()
    img_a, img_b = api.BandImage(ctx, rows, cols), api.BandImage(ctx, rows, cols)
    assert img_a.plane_stride_bytes == img_b.plane_stride_bytes
    for b, e in bands.band_ranges(rows, world):
        cv = api.Canvas(ctx, rows, cols, band=(b, e, 0))
        cv.setBackground(R0[b:e])
        cv.upload_layer(K[b:e], S[b:e], V[b:e])
        api.compose_gather(cv, [img_a, img_b] if b else [img_a])  # band 0 only reaches image a
        if b == 0:
            api.compose_gather(cv, [img_b])
    ctx.synchronize()
    assert np.array_equal(img_a.download(), want)
    assert np.array_equal(img_b.download(), want)
