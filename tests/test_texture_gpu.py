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


@pytest.mark.parametrize("prec", [0, 1])
def test_smudge_matches_oracle(ctx32, ctx64, port, prec):
    """renderer/Smudge.hxx on the device: the pickup windows persist across strokes, are re-created by a radius
    change, smudge can be switched off and on between strokes, strokes run off the canvas."""
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    rows, cols = 300, 400
    cv, cvo = api.Canvas(ctx, rows, cols), port.canvas(rows, cols)
    bg = np.random.default_rng(6).uniform(0.3, 1.0, (rows, cols, 3))
    cv.setBackground(bg)
    cvo.set_background(bg)
    tb, tbo = api.TextureBrush(ctx), port.texture_brush()
    tb.enableSmudge(True)
    tbo.enable_smudge(True)

    def both(fn_dev, fn_cpu):
        fn_dev(tb)
        fn_cpu(tbo)

    script = [
        (20.0, ([.2, .3, .4], [.1, .23, .14]), 0.8, [(30, 150), (200, 140), (350, 180)], True),
        (20.0, ([.5, .1, .2], [.3, .2, .5]), 0.8, [(150.5, 20.2), (180.1, 120.7), (170.3, 200.9), (220.0, 280.5)], True),
        (11.0, ([.1, .6, .2], [.3, .2, .1]), 0.4, [(50.5, 220.2), (380.1, 100.7)], True),
        (11.0, ([.3, .3, .3], [.2, .2, .2]), 0.4, [(10, 10), (100, 60)], False),
        (11.2, ([.6, .1, .1], [.1, .2, .3]), 1.0, [(390, 290), (300, 200), (320, 100)], True),
        (7.0, ([.2, .2, .6], [.3, .1, .2]), 0.0, [(100, 100), (200, 100)], True),  # deposits nothing: smudge state untouched
        (7.0, ([.2, .2, .6], [.3, .1, .2]), 0.9, [(-30, 50), (60, 80), (420, 40)], True),
    ]
    for radius, (K, S), scale, path, smudge in script:
        tb.enableSmudge(smudge)
        tbo.enable_smudge(smudge)
        tb.setRadius(radius)
        tbo.set_radius(radius)
        tb.dip((K, S))
        tbo.dip(K, S)
        tb.setThicknessScale(scale)
        tbo.set_thickness_scale(scale)
        tb.paintStroke(path, cv)
        tbo.paint_stroke(cvo, path)
    a, b = cv.download("KSV"), cvo.get()
    if prec:
        for k in "KSV":
            assert np.abs(a[k] - b[k]).max() <= 1e-12, k
    assert np.abs(cv.compose() - cvo.compose()).max() <= TOL[prec]


@pytest.mark.parametrize("prec", [0, 1])
def test_texture_brush_test_with_smudge_fixture(ctx32, ctx64, golden, prec):
    """The reference's TextureBrushTest as written (smudge on), against the fixture produced by the reference."""
    from painty_b200 import api

    ctx = [ctx32, ctx64][prec]
    cv = api.Canvas(ctx, 768, 1024)
    tb = api.TextureBrush(ctx)
    tb.enableSmudge(True)
    tb.dip(([.2, .3, .4], [.1, .23, .14]))
    tb.setRadius(40.0)
    tb.paintStroke([(50, 250), (400, 250), (650, 250)], cv)
    R = cv.compose()
    st = cv.download("V")
    assert (st["V"] > 0).sum() == int(golden["texs_wet"])
    assert np.abs(R[200:300, 300:400] - golden["texs_R_crop"]).max() <= TOL[prec]
    assert np.abs(R.sum(axis=(0, 2)) - golden["texs_R_colsum"]).max() <= TOL[prec] * 768 * 3
    assert np.abs(st["V"][200:300, 300:400] - golden["texs_V_crop"]).max() <= (1e-12 if prec else 1e-5)


@pytest.mark.parametrize("prec", [0, 1])
def test_per_stroke_dictionary_textures_match_oracle(ctx32, ctx64, port, prec):
    """Config-3 semantics: every stroke samples the thickness texture the brush-texture dictionary picked for it
    (TextureBrushDictionary.cxx:25-79) on a canvas-pattern substrate (CanvasGpu.cxx:27-40). CPU side: the reference's
    TextureBrush::paintStroke with that texture installed as the brush's thickness map (one CPU brush per texture),
    strokes in submission order."""
    from painty_b200 import api, assets

    ctx = [ctx32, ctx64][prec]
    rows, cols = 240, 320
    tex = assets.brush_textures()
    dic = api.TextureBrushDictionary([t[1] for t in tex], [t[2] for t in tex], [t[3].shape[0] for t in tex],
                                     [t[3].shape[1] for t in tex])
    r = np.random.default_rng(77)
    R0 = assets.canvas_pattern(rows, cols)
    cv, cvo = api.Canvas(ctx, rows, cols), port.canvas(rows, cols)
    cv.setBackground(R0)
    cvo.set_background(R0)
    tb = api.TextureBrush(ctx)
    ids, cpu_brushes = {}, {}
    state = 0.0  # the ONE brush's radius: setRadius only acts on a change >= 0.5 (TextureBrush.hxx:33-41)
    tb.setRadius(state)

    def cpu_set_radius(tbo, radius):  # force a CPU brush (one per texture) to exactly the single brush's effective radius
        tbo.set_radius(radius + 1000.0)
        tbo.set_radius(radius)

    n = 30
    rec = np.zeros(n, dtype=api.TSTROKE_DTYPE)
    verts, first, groups = [], 0, set()
    for i in range(n):
        K, S = r.uniform(0.05, 1.5, 3), r.uniform(0.05, 1.0, 3)
        rad = float(r.uniform(3, 130))  # brush sizes 6 .. 260 reach every size class of the dictionary
        m = int(r.integers(2, 9))
        p0 = r.uniform(0, [cols, rows])
        path = p0 + np.cumsum(r.normal(0, rad * 0.6, (m, 2)), axis=0)
        i0, i1, cand = dic.lookup(path, 2.0 * rad)
        groups.add((i0, i1))
        pick = int(cand[int(r.integers(0, len(cand)))])  # the reference draws with std::random_device; we record the pick
        if pick not in ids:
            ids[pick] = tb.addTexture(tex[pick][3])
            cpu_brushes[pick] = port.texture_brush(tex[pick][3])
        tbo = cpu_brushes[pick]
        if not abs(state - rad) < 0.5:
            state = rad
        cpu_set_radius(tbo, state)
        tbo.dip(K, S)
        tbo.set_thickness_scale(0.05)
        tbo.paint_stroke(cvo, path)
        rec[i] = (rad, K, S, 0.05, first, m, ids[pick])
        first += m
        verts.append(path)
    assert len(groups) >= 3  # the strokes exercise several (size, length) classes
    tb.stroke_batch(cv, rec, np.concatenate(verts))
    a, b = cv.download("KSV"), cvo.get()
    if prec:
        for k in "KSV":
            assert np.array_equal(a[k], b[k]), k
    assert np.abs(cv.compose() - cvo.compose()).max() <= TOL[prec]
    # single-stroke API with a selected texture == batch entry
    cv2_, cvo2 = api.Canvas(ctx, rows, cols), port.canvas(rows, cols)
    some = next(iter(ids))
    tb.selectTexture(ids[some])
    tb.setRadius(12.0)
    tb.dip(([.2, .3, .4], [.1, .23, .14]))
    tb.setThicknessScale(0.3)
    tb.paintStroke([(20, 30), (150, 120), (300, 200)], cv2_)
    tbo = cpu_brushes[some]
    cpu_set_radius(tbo, 12.0)
    tbo.dip([.2, .3, .4], [.1, .23, .14])
    tbo.set_thickness_scale(0.3)
    tbo.paint_stroke(cvo2, [(20, 30), (150, 120), (300, 200)])
    assert np.abs(cv2_.compose() - cvo2.compose()).max() <= TOL[prec]
    with pytest.raises(Exception):
        tb.selectTexture(10_000)
