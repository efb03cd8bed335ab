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
