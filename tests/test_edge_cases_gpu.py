"""Edge cases through the C ABI on the device: empty and degenerate inputs, ragged strokes, re-used handles."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_empty_and_degenerate_inputs(ctx32, port):
    from painty_b200 import api

    cv = api.Canvas(ctx32, 40, 50)
    br = api.FootprintBrush(ctx32, 6.0)
    br.dip(([.3, .2, .1], [.2, .4, .3]))
    br.imprint_batch(cv, [], [], [])  # no imprints
    br.stroke_batch(cv, np.zeros(0, dtype=api.STROKE_DTYPE), [], [], [])  # no strokes
    assert (cv.download("V")["V"] == 0).all()
    # strokes without imprints between real ones; a path with a single point expands to nothing
    assert len(api.expand_stroke([(3.0, 4.0)])[0]) == 0
    rec = np.zeros(3, dtype=api.STROKE_DTYPE)
    cx, cy, th = api.expand_stroke([(5.0, 5.0), (30.0, 20.0)])
    rec[0] = (6.0, [.3, .2, .1], [.2, .4, .3], 0, 0)
    rec[1] = (6.0, [.3, .2, .1], [.2, .4, .3], 0, len(cx))
    rec[2] = (6.0, [.1, .2, .3], [.2, .4, .3], len(cx), 0)
    br.stroke_batch(cv, rec, cx, cy, th)
    cvo, bro = port.canvas(40, 50), port.footprint_brush(6.0)
    bro.dip([.3, .2, .1], [.2, .4, .3])
    bro.imprint_batch(cvo, cx, cy, th)
    assert np.abs(cv.compose() - cvo.compose()).max() < 1e-4
    # imprints completely off the canvas are no-ops
    before = cv.download("V")["V"].copy()
    br.imprint_batch(cv, [-500.0, 5000.0], [-500.0, 20.0], [0.1, 0.2])
    assert np.array_equal(cv.download("V")["V"], before)
    # texture strokes: empty batch, single vertex, zero thickness scale
    tb = api.TextureBrush(ctx32)
    tb.setRadius(5.0)
    tb.stroke_batch(cv, np.zeros(0, dtype=api.TSTROKE_DTYPE), np.zeros((0, 2)))
    tb.paintStroke([(10.0, 10.0)], cv)
    tb.setThicknessScale(0.0)
    tb.paintStroke([(10.0, 10.0), (30.0, 30.0)], cv)
    assert np.array_equal(cv.download("V")["V"], before)
    assert tb.counters() == 0


def test_tiny_canvases(ctx64, port):
    from painty_b200 import api

    for rows, cols in [(1, 1), (1, 7), (5, 1), (3, 3)]:
        cv, cvo = api.Canvas(ctx64, rows, cols), port.canvas(rows, cols)
        br, bro = api.FootprintBrush(ctx64, 4.0), port.footprint_brush(4.0)
        br.dip(([.3, .2, .1], [.2, .4, .3]))
        bro.dip([.3, .2, .1], [.2, .4, .3])
        cx, cy, th = np.array([0.4, 1.2, 0.7]), np.array([0.6, 0.1, 2.2]), np.array([0.3, 1.1, -2.0])
        br.imprint_batch(cv, cx, cy, th)
        bro.imprint_batch(cvo, cx, cy, th)
        a, b = cv.download("KSV"), cvo.get()
        for k in "KSV":
            assert np.array_equal(a[k], b[k]), (rows, cols, k)
        assert np.abs(cv.compose() - cvo.compose()).max() < 1e-12


def test_errors_are_reported_not_swallowed(ctx32):
    from painty_b200 import api

    cv = api.Canvas(ctx32, 16, 16)
    br = api.FootprintBrush(ctx32, 4.0)
    rec = np.zeros(1, dtype=api.STROKE_DTYPE)
    rec[0] = (77.0, [.1, .1, .1], [.1, .1, .1], 0, 1)  # no footprint registered for radius 77
    with pytest.raises(api.PaintyError):
        br.stroke_batch(cv, rec, [1.0], [1.0], [0.0])
    rec[0] = (4.0, [.1, .1, .1], [.1, .1, .1], 0, 5)  # imprint range beyond the arrays
    with pytest.raises(api.PaintyError):
        br.stroke_batch(cv, rec, [1.0], [1.0], [0.0])
    with pytest.raises(api.PaintyError):
        api.Canvas(ctx32, 70000, 70000)  # >= 2^31 pixels (the reference's int32 index limit)


def test_remaining_api_surface(ctx64, port):
    """PaintLayer::copyTo / set, brush rate accessors, clean(), public updateSnapshot(canvas), snapshot toggling."""
    import ctypes as C

    from painty_b200 import api

    rows, cols = 48, 64
    rng = np.random.default_rng(3)
    K, S, V = rng.uniform(0, 1, (rows, cols, 3)), rng.uniform(0, 1, (rows, cols, 3)), rng.uniform(0, 0.5, (rows, cols))
    a, b = api.PaintLayer(ctx64, rows, cols), api.PaintLayer(ctx64, 2, 2)
    a.upload(K, S, V)
    a.copyTo(b)  # reallocates the destination like PaintLayer.hxx:104-107
    assert (b.getRows(), b.getCols()) == (rows, cols)
    for x, y in zip(b.download(), (K, S, V)):
        assert np.array_equal(x, y)
    a.clear()
    assert all((x == 0).all() for x in a.download())

    br, bro = api.FootprintBrush(ctx64, 6.0), port.footprint_brush(6.0)
    assert (br.getPickupRate(), br.getDepositionRate(), br.getUseSnapshotBuffer()) == (0.9, 0.05, True)  # :477-495
    br.setPickupRate(0.5)
    br.setDepositionRate(0.2)
    bro.set_rates(0.5, 0.2)
    assert (br.getPickupRate(), br.getDepositionRate()) == (0.5, 0.2)
    cv, cvo = api.Canvas(ctx64, rows, cols), port.canvas(rows, cols)
    cv.upload_layer(K, S, V)
    cvo.set_layer(K, S, V)
    br.updateSnapshot(cv)  # FootprintBrush::updateSnapshot(canvas) :168-172 == what the first imprint would do
    cx, cy, th = np.linspace(10, 50, 30), np.linspace(12, 30, 30), np.linspace(0, 1.5, 30)
    for obj, canvas in ((br, cv), (bro, cvo)):
        obj.dip(([.3, .2, .1], [.2, .4, .3])) if obj is br else obj.dip([.3, .2, .1], [.2, .4, .3])
        obj.imprint_batch(canvas, cx, cy, th)
    br.clean()  # :160-166, keeps the paint
    assert all((x == 0).all() for x in br.getPickupMap())
    port.fn("fbrush_dip", None, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)])  # (clean on the port = dip with the same paint)
    bro.dip([.3, .2, .1], [.2, .4, .3])
    br.imprint_batch(cv, cx[::-1].copy(), cy[::-1].copy(), th)
    bro.imprint_batch(cvo, cx[::-1].copy(), cy[::-1].copy(), th)
    got, want = cv.download("KSV"), cvo.get()
    for k in "KSV":
        assert np.array_equal(got[k], want[k]), k
