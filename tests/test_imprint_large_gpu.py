"""Oracle parity of the imprint engine at the radii the benchmark runs (r = 64 ... 151, and one footprint beyond 200):
the launch classes with 256- and 512-thread CTAs, more than two cells per thread, and the per-CTA pickup scratch in
global memory. Each case is a short stroke batch of overlapping strokes at oblique, changing angles, one of them
overhanging the left/top border, followed by a continuation without dip (the brush keeps paint and pickup map),
compared with the sequential CPU oracle: FP64 mode bit for bit on canvas K/S/V, pickup map and snapshot buffer; FP32
mode within 1e-4 on reflectance (FootprintBrush.hxx:73-143, 278-319, 349-431)."""
import ctypes as C

import numpy as np
import pytest

from painty_b200 import assets

pytestmark = pytest.mark.gpu


def _strokes_for(radius, rows, cols, seed):
    """Six overlapping strokes (30..45 imprints, ~1 px steps like FootprintBrush::paintStroke, headings that turn along
    the stroke), stroke 1 starts outside the top-left corner, stroke 4 runs off the bottom-right."""
    r = np.random.default_rng(seed)
    out = []
    starts = [(0.45 * cols, 0.40 * rows), (-0.15 * radius, 0.10 * radius), (0.50 * cols, 0.55 * rows), (0.40 * cols, 0.45 * rows),
              (cols - 0.6 * radius, rows - 0.4 * radius), (0.55 * cols, 0.35 * rows)]
    for k, (x0, y0) in enumerate(starts):
        n = int(r.integers(30, 46))
        heading = r.uniform(-np.pi, np.pi) if k != 1 else 0.6
        turn = np.cumsum(r.normal(0.0, 0.03, n))
        th = heading + turn
        step = r.uniform(0.8, 1.3, n)
        cx = x0 + np.cumsum(step * np.cos(th))
        cy = y0 + np.cumsum(step * np.sin(th))
        K, S = r.uniform(0.05, 1.5, 3), r.uniform(0.05, 1.0, 3)
        out.append(dict(K=K, S=S, cx=cx, cy=cy, th=th))
    return out


def _oracle_snapshot(port, bro, rows, cols):
    Ks, Ss, Vs = np.empty((rows, cols, 3)), np.empty((rows, cols, 3)), np.empty((rows, cols))
    PD = C.POINTER(C.c_double)
    port.fn("fbrush_get_snapshot", None, [C.c_void_p, PD, PD, PD])(bro.h, Ks.ctypes.data_as(PD), Ss.ctypes.data_as(PD),
                                                                  Vs.ctypes.data_as(PD))
    return Ks, Ss, Vs


def _run_case(ctx, port, radius, prec, seed, rows=None, cols=None, n_strokes=6):
    from painty_b200 import api

    radius = assets.snap_to_safe_radius(radius)
    rows = rows or int(5.2 * radius)
    cols = cols or int(6.0 * radius)
    strokes = _strokes_for(radius, rows, cols, seed)[:n_strokes]
    cvo, cv = port.canvas(rows, cols), api.Canvas(ctx, rows, cols)
    bro, br = port.footprint_brush(radius), api.FootprintBrush(ctx, radius)
    br.enable_visited_count(True)
    rec = np.zeros(len(strokes), dtype=api.STROKE_DTYPE)
    first = 0
    for i, s in enumerate(strokes):
        bro.dip(s["K"], s["S"])
        bro.set_radius(radius)
        bro.imprint_batch(cvo, s["cx"], s["cy"], s["th"])
        rec[i] = (radius, s["K"], s["S"], first, len(s["cx"]))
        first += len(s["cx"])
    cat = lambda k: np.concatenate([s[k] for s in strokes])
    br.stroke_batch(cv, rec, cat("cx"), cat("cy"), cat("th"))
    # continuation without dip: same paint, the pickup map of the last stroke carries on (imprint_batch path)
    s0 = strokes[0]
    ccx, ccy, cth = s0["cx"][::-1][:12].copy() + 3.3, s0["cy"][::-1][:12].copy() - 2.1, s0["th"][:12] + 0.5
    bro.imprint_batch(cvo, ccx, ccy, cth)
    br.imprint_batch(cv, ccx, ccy, cth)
    a, b = cv.download("KSV"), cvo.get()
    Rg, Ro = cv.compose(), cvo.compose()
    assert br.counters() == bro.counters()
    if prec:
        for k in "KSV":
            assert np.array_equal(a[k], b[k]), "canvas %s differs at r=%g" % (k, radius)
        for x, y, name in zip(br.getPickupMap(), bro.pickup_map(), "KSV"):
            assert np.array_equal(x, y), "pickup map %s differs at r=%g" % (name, radius)
        for x, y, name in zip(br.getSnapshot(cv), _oracle_snapshot(port, bro, rows, cols), "KSV"):
            assert np.array_equal(x, y), "snapshot %s differs at r=%g" % (name, radius)
        assert float(np.abs(Rg - Ro).max()) <= 1e-10
    else:
        assert np.array_equal(a["V"] > 0, b["V"] > 0)
        assert float(np.abs(Rg - Ro).max()) <= 1e-4
    return float(np.abs(Rg - Ro).max())


@pytest.mark.parametrize("radius", [64, 100, 112, 129, 151])
@pytest.mark.parametrize("prec", [0, 1])
def test_bench_radii_match_oracle(ctx32, ctx64, port, radius, prec):
    """r = 64 (16 x 256 threads), 100 / 112 / 129 / 151 (16 x 512 threads, 1.5 ... 3.4 cells per thread): the launch
    classes that carry > 90 % of the benchmarked step."""
    _run_case([ctx32, ctx64][prec], port, radius, prec, seed=1000 + radius)


@pytest.mark.parametrize("prec", [0, 1])
def test_footprint_beyond_shared_memory_matches_oracle(ctx32, ctx64, port, prec):
    """r = 240: ~70 000 active cells; in FP64 the per-CTA slice of the pickup map no longer fits 200 KB of shared memory and
    lives in the global scratch buffer."""
    _run_case([ctx32, ctx64][prec], port, 240, prec, seed=7, rows=900, cols=1100, n_strokes=3)


def test_axis_aligned_and_diagonal_angles_fp64(ctx64, port):
    """theta = 0, pi/2, pi/4 exactly and steps of several pixels: rounding ties of the rotated coordinates (round half away
    from zero) and jumps of the snapshot ring larger than one pixel."""
    from painty_b200 import api

    radius, rows, cols = assets.snap_to_safe_radius(100), 520, 640
    cvo, cv = port.canvas(rows, cols), api.Canvas(ctx64, rows, cols)
    bro, br = port.footprint_brush(radius), api.FootprintBrush(ctx64, radius)
    n = 16
    for k, (th, dx, dy) in enumerate([(0.0, 3.0, 0.0), (np.pi / 2, 0.0, 2.5), (np.pi / 4, 5.0, 5.0), (-np.pi / 4, 40.0, -30.0),
                                       (np.pi, -1.0, 0.5)]):
        cx = 260.0 + dx * np.arange(n) + 7 * k
        cy = 250.0 + dy * np.arange(n) - 5 * k
        if k == 3:  # integer centres, huge steps (the ring of one imprint crosses the previous imprint's pixels)
            cx, cy = np.round(cx), np.round(cy)
        t = np.full(n, th)
        K, S = [.3 + .1 * k, .2, .1], [.2, .4, .3 + .05 * k]
        bro.dip(K, S)
        br.dip((K, S))
        bro.imprint_batch(cvo, cx, cy, t)
        br.imprint_batch(cv, cx, cy, t)
    a, b = cv.download("KSV"), cvo.get()
    for k in "KSV":
        assert np.array_equal(a[k], b[k]), k
    for x, y in zip(br.getPickupMap(), bro.pickup_map()):
        assert np.array_equal(x, y)
    for x, y in zip(br.getSnapshot(cv), _oracle_snapshot(port, bro, rows, cols)):
        assert np.array_equal(x, y)
