"""GPU parity of the Kubelka-Munk compose path through the C ABI vs the CPU oracle.
Tolerances (BASELINE.json north_star): reflectance within 1e-4 absolute in FP32 mode, 1e-10 in FP64."""
import numpy as np
import pytest

from tests.workloads import km_random_planes

pytestmark = pytest.mark.gpu
TOL = {0: 1e-4, 1: 1e-10}


def both(request_ctx32, request_ctx64):
    return [request_ctx32, request_ctx64]


def _cmp(a, b, tol):
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs (K == 0 policy: reproduce the reference)"
    m = ~np.isnan(b)
    err = np.abs(a[m] - b[m]).max() if m.any() else 0.0
    assert err <= tol, err
    return err


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("shape", [(64, 64), (37, 53), (1, 1), (3, 1), (240, 321)])
def test_layer_compose_matches_oracle(ctx32, ctx64, port, prec, shape):
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    K, S, V, R0 = km_random_planes(shape[0], shape[1], seed=shape[0] * 7 + shape[1], edge_cases=True)
    layer = api.PaintLayer(ctx, *shape)
    layer.upload(K, S, V)
    R = api.Renderer().compose(layer, R0)
    _cmp(R, port.compose(K, S, V, R0), TOL[prec])
    # composeOnto is the same in place; a wrongly sized R0 is replaced by ones (PaintLayer.hxx:82-87)
    R2 = layer.composeOnto(R0.copy())
    _cmp(R2, port.compose_onto(K, S, V, R0), TOL[prec])
    R3 = layer.composeOnto(np.zeros((2, 2, 3)))
    _cmp(R3, port.compose_onto(K, S, V, np.ones_like(R0)), TOL[prec])


@pytest.mark.parametrize("prec", [0, 1])
def test_compose_golden_fixture(ctx32, ctx64, golden, prec):
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    layer = api.PaintLayer(ctx, 64, 64)
    layer.upload(golden["km_K"], golden["km_S"], golden["km_V"])
    _cmp(api.Renderer().compose(layer, golden["km_R0"]), golden["km_R"], TOL[prec])


@pytest.mark.parametrize("prec", [0, 1])
def test_palette_extremes(ctx32, ctx64, port, prec):
    """K,S spanning the measured palette's 1e-9 .. 4.32 range, thin and very thick layers."""
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    vals = np.array([1e-9, 1e-6, 1e-3, 0.01, 0.3, 1.21, 4.32])
    ds = np.array([1e-13, 1e-6, 1e-3, 0.05, 1.0, 7.0, 60.0, 400.0])
    K, S, D, R = np.meshgrid(vals, vals, ds, [0.0, 0.02, 0.5, 1.0], indexing="ij")
    n = K.size
    Kp = np.repeat(K.reshape(n, 1), 3, 1).reshape(1, n, 3)
    Sp = np.repeat(S.reshape(n, 1), 3, 1).reshape(1, n, 3)
    Rp = np.repeat(R.reshape(n, 1), 3, 1).reshape(1, n, 3)
    Vp = D.reshape(1, n)
    layer = api.PaintLayer(ctx, 1, n)
    layer.upload(Kp, Sp, Vp)
    _cmp(api.Renderer().compose(layer, Rp), port.compose(Kp, Sp, Vp, Rp), TOL[prec])


@pytest.mark.parametrize("prec", [0, 1])
def test_canvas_compose_dry_clear(ctx32, ctx64, port, prec):
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols = 75, 101
    K, S, V, R0 = km_random_planes(rows, cols, seed=3, edge_cases=False)
    cv, cc = api.Canvas(ctx, rows, cols), port.canvas(rows, cols)
    st = cv.download()
    assert (st["R0"] == 1).all() and (st["V"] == 0).all() and (st["h"] == 0).all()  # Canvas ctor = clear()
    cv.setBackground(R0)
    cc.set_background(R0)
    cv.upload_layer(K, S, V)
    cc.set_layer(K, S, V)
    _cmp(cv.compose(), cc.compose(), TOL[prec])
    cv.dryCanvas()
    cc.dry()
    a, b = cv.download(), cc.get()
    _cmp(a["R0"], b["R0"], TOL[prec])
    assert np.abs(a["h"] - b["h"]).max() <= (1e-6 if prec == 0 else 0)
    assert (a["K"] == 0).all() and (a["S"] == 0).all() and (a["V"] == 0).all()
    _cmp(cv.compose(), cc.compose(), TOL[prec])  # everything dry -> R0
    # second wet layer on top of the dried one, then clear
    K2, S2, V2, _ = km_random_planes(rows, cols, seed=4)
    cv.upload_layer(K2, S2, V2)
    cc.set_layer(K2, S2, V2)
    _cmp(cv.compose(), cc.compose(), TOL[prec])  # the north-star budget holds across dryCanvas as well
    cv.clear()
    assert (cv.compose() == 1).all()


@pytest.mark.parametrize("prec", [0, 1])
@pytest.mark.parametrize("layers", [1, 2, 4, 8, 11])
def test_stacked_layers(ctx32, ctx64, port, prec, layers):
    """config 5: L stacked layers in one pass == L successive composeOnto calls of the reference."""
    import torch

    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols = 33, 47
    n = rows * cols
    dt = torch.float64 if prec else torch.float32
    R = None
    Ks, Ss, Vs, keep = [], [], [], []
    for l in range(layers):
        K, S, V, R0 = km_random_planes(rows, cols, seed=100 + l)
        R = R0 if R is None else R
        R = port.compose_onto(K, S, V, R)
        if l == 0:
            r0_first = R0
        t = torch.tensor(np.ascontiguousarray(np.concatenate([K.reshape(n, 3).T, S.reshape(n, 3).T, V.reshape(1, n)])), dtype=dt,
                         device="cuda").contiguous()
        keep.append(t)
        Ks.append([t[i].data_ptr() for i in range(3)])
        Ss.append([t[3 + i].data_ptr() for i in range(3)])
        Vs.append(t[6].data_ptr())
    r0 = torch.tensor(np.ascontiguousarray(r0_first.reshape(n, 3).T), dtype=dt, device="cuda").contiguous()
    out = torch.empty_like(r0)
    torch.cuda.synchronize()
    ctx.km_compose_stacked_planes(n, Ks, Ss, Vs, [r0[i].data_ptr() for i in range(3)], [out[i].data_ptr() for i in range(3)])
    ctx.synchronize()
    got = out.cpu().numpy().T.reshape(rows, cols, 3).astype(np.float64)
    _cmp(got, R, TOL[prec] * (1 if prec else layers))


def test_full_size_properties(ctx32):
    """4K canvas (BASELINE config): size-independent properties instead of a full CPU run:
    dry pixels return R0 bit-exactly, thick opaque layers converge to R_inf = a - b, and compose is
    pixel-local (a shuffled copy of the planes gives the shuffled result)."""
    import torch

    rows, cols = 2160, 3840
    n = rows * cols
    g = torch.Generator(device="cuda").manual_seed(1)
    pl = torch.rand((10, n), device="cuda", generator=g) * 0.9 + 0.05
    pl[6, ::3] = 0.0  # every third pixel dry
    pl[6, 1::3] = 500.0  # every third pixel opaque
    out = torch.empty((3, n), device="cuda")
    torch.cuda.synchronize()
    p = [pl[i].data_ptr() for i in range(10)]
    ctx32.km_compose_planes(n, p[0:3], p[3:6], p[6], p[7:10], [out[i].data_ptr() for i in range(3)])
    ctx32.synchronize()
    assert torch.equal(out[:, ::3], pl[7:10, ::3])
    a = 1.0 + pl[0:3, 1::3].double() / pl[3:6, 1::3].double()
    rinf = a - torch.sqrt(a * a - 1.0)
    assert (out[:, 1::3].double() - rinf).abs().max().item() < 1e-5
    perm = torch.randperm(n, device="cuda", generator=g)
    pl2 = pl[:, perm].contiguous()
    out2 = torch.empty_like(out)
    torch.cuda.synchronize()
    p2 = [pl2[i].data_ptr() for i in range(10)]
    ctx32.km_compose_planes(n, p2[0:3], p2[3:6], p2[6], p2[7:10], [out2[i].data_ptr() for i in range(3)])
    ctx32.synchronize()
    assert torch.equal(out2, out[:, perm])


@pytest.mark.parametrize("prec", [0, 1])
def test_band_canvas_compose(ctx32, ctx64, port, prec):
    from painty_b200 import api
    import torch

    ctx = [ctx32, ctx64][prec]
    rows, cols = 40, 37
    K, S, V, R0 = km_random_planes(rows, cols, seed=8)
    want = port.compose(K, S, V, R0)
    dt = torch.float64 if prec else torch.float32
    for (b, e, halo) in [(0, 13, 3), (13, 29, 5), (29, 40, 2)]:
        cv = api.Canvas(ctx, rows, cols, band=(b, e, halo))
        lo, hi = cv.store_first, cv.store_first + cv.store_rows
        assert lo == max(0, b - halo) and hi == min(rows, e + halo)
        cv.setBackground(R0[lo:hi])
        cv.upload_layer(K[lo:hi], S[lo:hi], V[lo:hi])
        out = torch.empty((3, (e - b) * cols), dtype=dt, device="cuda")
        torch.cuda.synchronize()
        cv.compose_band_device(out.data_ptr(), (e - b) * cols)
        ctx.synchronize()
        got = out.cpu().numpy().T.reshape(e - b, cols, 3).astype(np.float64)
        _cmp(got, want[b:e], TOL[prec])


@pytest.mark.parametrize("prec", [0, 1])
def test_display_epilogue(ctx32, ctx64, port, prec):
    """compose fused with rgb2srgb + quantisation (GUI qRgb path, imSave 8/16-bit BGR path). FP64 mode is exact;
    in FP32 mode a value within 1e-6 of a quantisation step may land on the neighbouring code (<= 1 LSB, rare)."""
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols = 90, 131
    K, S, V, R0 = km_random_planes(rows, cols, seed=15)
    cv = api.Canvas(ctx, rows, cols)
    cv.setBackground(R0)
    cv.upload_layer(K, S, V)
    want_rgb = port.compose(K, S, V, R0)
    for got, want in ((cv.compose_qrgb32(), port.qrgb32(want_rgb)), (cv.compose_bgr(8), port.bgr(want_rgb, 8)),
                      (cv.compose_bgr(16), port.bgr(want_rgb, 16)), (cv.compose_bgr(16, srgb=False), port.bgr(want_rgb, 16, False))):
        if got.dtype == np.uint32:
            got = np.stack([(got >> s) & 0xFF for s in (16, 8, 0)], -1).astype(np.int64)
            want = np.stack([(want >> s) & 0xFF for s in (16, 8, 0)], -1).astype(np.int64)
        diff = np.abs(got.astype(np.int64) - want.astype(np.int64))
        if prec:
            assert diff.max() == 0
        else:
            lsb = 1 if got.max() <= 255 else 8  # 16 bit: 1e-4 of 65535 = 6.5 codes
            assert diff.max() <= lsb and (diff > 0).mean() < (0.01 if lsb == 1 else 1.0)


@pytest.mark.parametrize("prec", [0, 1])
def test_render_relighting(ctx32, ctx64, port, prec):
    """Renderer::render: compose + Beckmann/Cook-Torrance relighting with the wet thickness as height
    field, BORDER_REFLECT stencil at the edges."""
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols = 67, 91
    K, S, V, R0 = km_random_planes(rows, cols, seed=21)
    V = V * 3.0
    V[10:20, 30:50] = 0.0  # flat dry patch
    cv, cvo = api.Canvas(ctx, rows, cols), port.canvas(rows, cols)
    cv.setBackground(R0)
    cvo.set_background(R0)
    cv.upload_layer(K, S, V)
    cvo.set_layer(K, S, V)
    got, want = api.Renderer().render(cv), cvo.render()
    assert np.abs(got - want).max() <= (1e-10 if prec else 2e-4)


@pytest.mark.parametrize("prec", [0, 1])
def test_lab_scaled_readback_prep(ctx32, ctx64, port, prec):
    """pb_canvas_compose_lab_scaled == ScaledMat(convertColor(compose(canvas), rgb_2_CIELab), rows, cols)
    (PictureTargetSbrPainter.cxx:334-341): FP64 within 1e-10 (Lab units), FP32 within 1e-4 of reflectance scaled to the
    Lab range (x100) — and the same-size case is a plain conversion."""
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols = 150, 202
    K, S, V, R0 = km_random_planes(rows, cols, seed=11, edge_cases=False)
    V[::3, ::2] = 0.0
    cv, cc = api.Canvas(ctx, rows, cols), port.canvas(rows, cols)
    cv.setBackground(R0)
    cc.set_background(R0)
    cv.upload_layer(K, S, V)
    cc.set_layer(K, S, V)
    want_rgb = cc.compose()
    tol = 1e-10 if prec else 1e-2
    for (orows, ocols) in ((75, 101), (64, 90), (rows, cols), (200, 260)):
        got = cv.compose_lab_scaled(orows, ocols)
        want = port.lab_scaled(want_rgb, orows, ocols)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= tol, (orows, ocols, float(np.abs(got - want).max()))
