"""GPU parity of the texture-brush deposit (smudge off) through the C ABI vs oracle / reference fixture."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = {0: 1e-4, 1: 1e-10}


@pytest.mark.parametrize("prec", [0, 1])
def test_texture_brush_test_stroke_fixture(ctx32, ctx64, golden, prec):
    """renderer/test/src/TextureBrushTest.cxx:16-40 stroke + a crossing one; fixture from the reference."""
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    cv = api.Canvas(ctx, 768, 1024)
    tb = api.TextureBrush(ctx)
    tb.setRadius(40.0)
    tb.dip(([.2, .3, .4], [.1, .23, .14]))
    tb.paintStroke([(50, 250), (400, 250), (650, 250)], cv)
    tb.dip(([.5, .1, .2], [.3, .2, .5]))
    tb.setRadius(25.0)
    tb.paintStroke([(300.5, 50.2), (350.1, 200.7), (330.3, 400.9), (420.0, 600.5)], cv)
    R = cv.compose()
    st = cv.download("V")
    assert (st["V"] > 0).sum() == int(golden["tex_wet"])
    assert np.abs(R[200:300, 300:400] - golden["tex_R_crop"]).max() <= TOL[prec]
    assert np.abs(R.sum(axis=(0, 2)) - golden["tex_R_colsum"]).max() <= TOL[prec] * 768 * 3
    if prec:
        assert np.array_equal(st["V"][200:300, 300:400], golden["tex_V_crop"])
        assert st["V"].sum() == float(golden["tex_sumV"])


@pytest.mark.parametrize("prec", [0, 1])
def test_texture_batch_matches_sequential_oracle(ctx32, ctx64, port, prec):
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols = 260, 330
    r = np.random.default_rng(31)
    cv, cvo = api.Canvas(ctx, rows, cols), port.canvas(rows, cols)
    tb, tbo = api.TextureBrush(ctx), port.texture_brush()
    n = 40
    rec = np.zeros(n, dtype=api.TSTROKE_DTYPE)
    verts = []
    first = 0
    for i in range(n):
        K, S = r.uniform(0.05, 1.5, 3), r.uniform(0.05, 1.0, 3)
        rad = float(r.uniform(3, 30))
        scale = float(r.uniform(0.05, 1.0))
        m = int(r.integers(1, 9))  # a 1-vertex stroke is a no-op (TextureBrush.hxx:53-55)
        p0 = r.uniform(-20, [cols + 20, rows + 20])
        path = p0 + np.cumsum(r.normal(0, 12, (m, 2)), axis=0)
        tbo.set_radius(rad)
        tbo.dip(K, S)
        tbo.set_thickness_scale(scale)
        tbo.paint_stroke(cvo, path)
        rec[i] = (rad, K, S, scale, first, m, 0)
        first += m
        verts.append(path)
    tb.stroke_batch(cv, rec, np.concatenate(verts))
    a, b = cv.download("KSV"), cvo.get()
    if prec:
        for k in "KSV":
            assert np.array_equal(a[k], b[k]), k
    assert np.abs(cv.compose() - cvo.compose()).max() <= TOL[prec]
    import ctypes as C

    assert tb.counters() == port.fn("tbrush_pixels", C.c_uint64, [C.c_void_p])(tbo.h)
