"""No-GPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/painty_b200.h declares, runs its host-side f64 calls bit-exactly, and refuses to create a
context without a CUDA device (there is no CPU fallback)."""
import ctypes
import math
import os
import re
import subprocess

import numpy as np
import pytest

from tests.workloads import gui_stroke_imprints

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "painty_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = declared_symbols()
    assert len(names) > 50
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_library_contains_sm100a_code_only(built_lib):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_host_scalars_match_oracle(built_lib, port):
    from painty_b200 import api

    rng = np.random.default_rng(3)
    for _ in range(200):
        K, S, R0 = rng.uniform(0, 2, 3), rng.uniform(0, 1.2, 3), rng.uniform(0, 1, 3)
        d = float(rng.choice([0.0, 1e-13, rng.uniform(0, 3), 60.0]))
        if rng.random() < 0.1:
            S[0] = 0.0
        assert np.array_equal(api.ComputeReflectance(K, S, R0, d), port.compute_reflectance(K, S, R0, d), equal_nan=True)
    K, S = api.ComputeScatteringAndAbsorption([.2, .05, .4], [.6, .3, .7])
    K2, S2 = port.compute_scattering_absorption([.2, .05, .4], [.6, .3, .7])
    assert np.array_equal(K, K2) and np.array_equal(S, S2)
    with pytest.raises(ValueError):  # std::invalid_argument in the reference (KubelkaMunk.hxx:98)
        api.ComputeScatteringAndAbsorption([.7, .05, .4], [.6, .3, .7])


def test_mixing_calls(built_lib):
    from painty_b200 import api, assets

    pk, ps = assets.palette("lindemeier_measured")
    assert pk.shape == (14, 3)
    w = np.random.default_rng(0).uniform(0, 1, 14)
    K, S = api.mixSinglePaint(pk, ps, w)
    norm = 1.0 / w.sum()
    eK = np.zeros(3)
    for l in range(14):  # mixer/src/PaintMixer.cxx:348-353, same accumulation order
        eK += norm * w[l] * pk[l]
    assert np.array_equal(K, eK)
    with pytest.raises(ValueError):
        api.mixSinglePaint(pk, ps, w[:5])
    tk, ts = assets.palette("thinning_medium")
    K2, S2 = api.mixed(K, S, 1.0, tk[0], ts[0], 0.5)
    assert np.array_equal(K2, ((1.0 * K) + (0.5 * tk[0])) * (1.0 / 1.5))


def test_expand_stroke_matches_reference_loops(built_lib, port):
    from painty_b200 import api

    pts = [(100.3, 200.7), (400.9, 260.2), (700.1, 180.4)]
    cx, cy, th = api.expand_stroke(pts, mode=1)
    ex, ey, eth = gui_stroke_imprints(port, pts)
    assert len(cx) == 615
    assert np.array_equal(cx, ex) and np.array_equal(cy, ey) and np.array_equal(th, eth)
    # library form (FootprintBrush.hxx:251-267) with p_pre = path[0] on the first segment
    pts = np.array([(10.5, 20.25), (60.0, 40.5), (90.75, 90.0), (140.0, 95.5)])
    cx, cy, th = api.expand_stroke(pts, mode=0)
    e = []
    for i in range(len(pts) - 1):
        a, b, c, d = pts[max(i - 1, 0)], pts[i], pts[i + 1], pts[min(i + 2, len(pts) - 1)]
        dist = float(np.sqrt(((c - b) * (c - b)).sum()))
        for pd in range(1, int(dist) + 1):
            t = pd / dist
            q, dr = port.catmull_rom(a, b, c, d, t), port.catmull_rom(a, b, c, d, t, True)
            e.append((q[0], q[1], math.atan2(dr[1], dr[0])))  # libm atan2 like the reference (numpy differs by ulps)
    e = np.array(e)
    assert np.array_equal(cx, e[:, 0]) and np.array_equal(cy, e[:, 1]) and np.array_equal(th, e[:, 2])
    assert len(api.expand_stroke([(1.0, 2.0)], mode=0)[0]) == 0


def test_no_cpu_fallback(built_lib):
    import torch

    from painty_b200 import api

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(api.PaintyError):
        api.Context(0, api.F32)


def test_dataflow_planner_orders_every_conflicting_pair(built_lib):
    """The host planner behind the stroke batches: for random box / ring rectangles, every pair whose box meets the
    other's allowed region must be ordered through the (transitive) predecessor lists; rings that only overlap each
    other need not be. Pure host code — runs without a GPU."""
    from painty_b200 import api

    rng = np.random.default_rng(17)
    rows, cols, n = 1500, 2000, 400
    box, allowed = [], []
    for _ in range(n):
        x, y = int(rng.integers(-50, cols)), int(rng.integers(-50, rows))
        w, h, m = int(rng.integers(5, 300)), int(rng.integers(5, 300)), int(rng.integers(0, 150))
        box.append((x, y, x + w, y + h))
        allowed.append((x - m, y - m, x + w + m, y + h + m))
    off, preds = api.plan_dependencies(rows, cols, box, allowed)
    assert off[0] == 0 and off[-1] == len(preds) and all(0 <= p < i for i in range(n) for p in preds[off[i]:off[i + 1]])
    # transitive closure of "must finish before"
    before = [set() for _ in range(n)]
    for i in range(n):
        for p in preds[off[i]:off[i + 1]]:
            before[i].add(int(p))
            before[i] |= before[int(p)]

    def clip(r):
        return (max(r[0], 0), max(r[1], 0), min(r[2], cols - 1), min(r[3], rows - 1))

    def meets(a, b):
        a, b = clip(a), clip(b)
        return a[0] <= a[2] and a[1] <= a[3] and b[0] <= b[2] and b[1] <= b[3] and a[0] <= b[2] and b[0] <= a[2] and a[1] <= b[3] and b[1] <= a[3]

    conflicts = independent = 0
    for i in range(n):
        for j in range(i):
            if meets(box[i], allowed[j]) or meets(allowed[i], box[j]):
                conflicts += 1
                assert j in before[i], (j, i)
            elif j not in before[i]:
                independent += 1
    assert conflicts > 1000 and independent > 1000  # the plan really leaves parallelism


@pytest.mark.parametrize("use_snapshot", [True, False])
def test_segment_plan_orders_every_conflicting_segment_pair(built_lib, use_snapshot):
    """The footprint brush tracks dependencies per SEGMENT of a stroke (progress counters instead of one completion
    flag). For a random overlapping stroke list: every pair of segments of two different strokes whose regions
    conflict (box of one meets the allowed region of the other) must be ordered, earlier stroke first, through the
    waits (stroke, segments needed) + the in-order execution of a stroke's own segments. Pure host code."""
    from painty_b200 import api, assets

    rng = np.random.default_rng(5)
    rows, cols, n, seg = 1200, 1600, 90, 16
    first, count, side, radius, cx, cy = [], [], [], [], [], []
    for _ in range(n):
        m = int(rng.integers(1, 300))
        r = float(rng.uniform(8, 40))
        x, y, a = rng.uniform(-20, cols + 20), rng.uniform(-20, rows + 20), rng.uniform(0, 2 * np.pi)
        first.append(len(cx)); count.append(m); radius.append(r); side.append(assets.footprint_geometry(r)[3])
        for _ in range(m):
            cx.append(x); cy.append(y)
            a += rng.normal(0, 0.05); x += np.cos(a); y += np.sin(a)
    first.append(len(cx)); count.append(0); radius.append(10.0); side.append(assets.footprint_geometry(10.0)[3])  # empty stroke
    n += 1
    seg_first, seg_len, seg_off, ps, pn = api.plan_segments(rows, cols, first, count, side, radius, cx, cy, seg, use_snapshot)
    n_seg = int(seg_first[-1])
    assert len(seg_off) == n_seg + 1 and seg_off[0] == 0 and seg_off[-1] == len(ps)
    owner = np.repeat(np.arange(n), np.diff(seg_first))
    for s in range(n):
        assert seg_first[s + 1] - seg_first[s] == max(1, -(-count[s] // seg_len[s]))
    # whole strokes (segment_length 0) give one segment each
    sf0 = api.plan_segments(rows, cols, first, count, side, radius, cx, cy, 0, True)[0]
    assert list(sf0) == list(range(n + 1))

    cxa, cya = np.asarray(cx), np.asarray(cy)

    def regions(s, k):
        a = first[s] + k * seg_len[s]
        m = min(seg_len[s], count[s] - k * seg_len[s])
        if m <= 0:
            return None, None
        half = (side[s] - 1) // 2
        out = []
        for margin in (0.0, radius[s]):
            mm = half + margin + 2.0
            out.append((max(0, int(np.floor(cxa[a:a + m].min() - mm))), max(0, int(np.floor(cya[a:a + m].min() - mm))),
                        min(cols - 1, int(np.ceil(cxa[a:a + m].max() + mm))), min(rows - 1, int(np.ceil(cya[a:a + m].max() + mm)))))
        return out

    def meets(a, b):
        return (a is not None and b is not None and a[0] <= a[2] and a[1] <= a[3] and b[0] <= b[2] and b[1] <= b[3]
                and a[0] <= b[2] and b[0] <= a[2] and a[1] <= b[3] and b[1] <= a[3])

    box, alw = zip(*[regions(int(owner[g]), g - int(seg_first[owner[g]])) for g in range(n_seg)])
    # reach[g] = bitset of the segments that are guaranteed complete before segment g starts
    reach = [0] * n_seg
    for g in range(n_seg):
        s = int(owner[g])
        r = 0
        if g > seg_first[s]:
            r |= reach[g - 1] | (1 << (g - 1))
        for i in range(seg_off[g], seg_off[g + 1]):
            p, need = int(ps[i]), int(pn[i])
            assert p < s and 1 <= need <= seg_first[p + 1] - seg_first[p]
            q = int(seg_first[p]) + need - 1
            r |= reach[q] | (1 << q)
        reach[g] = r
    conflicts = free = 0
    for g in range(n_seg):
        for q in range(int(seg_first[owner[g]])):  # segments of earlier strokes
            # snapshot buffer on: box of one meets the allowed region of the other; off: the boxes meet
            if (meets(box[g], alw[q]) or meets(alw[g], box[q])) if use_snapshot else meets(box[g], box[q]):
                conflicts += 1
                assert (reach[g] >> q) & 1, (q, g)
            elif not (reach[g] >> q) & 1:
                free += 1
    assert conflicts > 1000 and free > 1000
    # finer than stroke level: some stroke starts before an earlier stroke it depends on has finished
    partial = 0
    for s in range(n):
        g0 = int(seg_first[s])
        for i in range(seg_off[g0], seg_off[g0 + 1]):
            partial += int(pn[i]) < seg_first[ps[i] + 1] - seg_first[ps[i]]
    assert partial > 0


def test_claim_order_is_topological_and_keeps_runs(built_lib):
    """The device pops strokes strictly in the host-planned claim order and blocks on the dataflow waits; that is
    deadlock free only if the order is a topological order of the segment-level graph (and keeps a pool's launches
    in sequence). Two pools (GPUs), several launches each, random overlapping strokes. Pure host code."""
    from painty_b200 import api, assets

    rng = np.random.default_rng(11)
    rows, cols, n, seg = 1400, 1800, 300, 16
    first, count, side, radius, cx, cy, pool = [], [], [], [], [], [], []
    for i in range(n):
        m = int(rng.integers(0, 200))
        r = float(rng.uniform(6, 45)) if (i // 50) % 2 == 0 else float(rng.uniform(3, 8))
        x, y, a = rng.uniform(0, cols), rng.uniform(0, rows), rng.uniform(0, 2 * np.pi)
        first.append(len(cx)); count.append(m); radius.append(r); side.append(assets.footprint_geometry(r)[3])
        pool.append(0 if y < rows / 2 else 1)
        for _ in range(m):
            cx.append(x); cy.append(y)
            a += rng.normal(0, 0.05); x += np.cos(a); y += np.sin(a)
    # launches: consecutive strokes of a pool with the same size class
    run, slots = np.zeros(n, np.int32), [[], []]
    last_cls = [None, None]
    for i in range(n):
        c = radius[i] >= 8
        if last_cls[pool[i]] != c:
            slots[pool[i]].append(3 if c else 7)
            last_cls[pool[i]] = c
        run[i] = len(slots[pool[i]]) - 1
    cost = 5.0 + 0.002 * np.asarray(side, dtype=np.float64) ** 2
    order = api.plan_claim_order(rows, cols, first, count, side, radius, cx, cy, pool, run, cost, slots, seg, True)
    assert sorted(order) == list(range(n))
    pos = np.empty(n, dtype=np.int64)
    pos[order] = np.arange(n)
    seg_first, seg_len, seg_off, ps, pn = api.plan_segments(rows, cols, first, count, side, radius, cx, cy, seg, True)
    owner = np.repeat(np.arange(n), np.diff(seg_first))
    moved = int((order != np.arange(n)).sum())
    assert moved > 0  # the planner does reorder something on this list
    for g in range(len(owner)):
        for i in range(seg_off[g], seg_off[g + 1]):
            assert pos[ps[i]] < pos[owner[g]]
    for p in (0, 1):
        seq = [int(run[s]) for s in order if pool[s] == p]
        assert seq == sorted(seq)


def test_null_handles_are_errors_not_crashes(built_lib):
    """A binding that passes a NULL handle gets status 1 and a message from pb_last_error(), never a segfault."""
    import ctypes as C

    from painty_b200 import api

    lib = api.lib()
    lib.pb_last_error.restype = C.c_char_p
    null = C.c_void_p(None)
    out = (C.c_double * 3)()
    calls = [
        ("pb_canvas_clear", (null,)),
        ("pb_canvas_dry", (null,)),
        ("pb_canvas_compose", (null, out)),
        ("pb_layer_clear", (null,)),
        ("pb_fbrush_set_pickup_rate", (null, C.c_double(0.5))),
        ("pb_fbrush_set_use_snapshot", (null, 1)),
        ("pb_fbrush_clean", (null,)),
        ("pb_fbrush_dip", (null, out, out)),
        ("pb_tbrush_dip", (null, out, out)),
        ("pb_tbrush_enable_smudge", (null, 1)),
        ("pb_canvas_stored_rows", (null, None, None)),
    ]
    for name, args in calls:
        rc = getattr(lib, name)(*args)
        assert rc == 1, name
        assert b"null handle" in lib.pb_last_error(), name
    # destroying NULL is a no-op, like free()
    for name in ("pb_canvas_destroy", "pb_layer_destroy", "pb_fbrush_destroy", "pb_tbrush_destroy", "pb_context_destroy"):
        assert getattr(lib, name)(null) == 0, name


def _ring_items(lib, box, allowed, prev=None, pitch=240):
    import ctypes as C

    b, a = (C.c_int32 * 4)(*box), (C.c_int32 * 4)(*allowed)
    pb = pa = None
    if prev is not None:
        pb, pa = (C.c_int32 * 4)(*prev[0]), (C.c_int32 * 4)(*prev[1])
    n = C.c_int64(0)
    dummy = (C.c_int32 * 1)()
    assert lib.pb_ring_rects(b, a, pb, pa, pitch, C.c_int64(0), dummy, dummy, C.byref(n)) == 0
    rows = (C.c_int32 * max(n.value, 1))()
    words = (C.c_int32 * max(n.value, 1))()
    assert lib.pb_ring_rects(b, a, pb, pa, pitch, C.c_int64(n.value), rows, words, C.byref(n)) == 0
    return list(zip(rows[:n.value], words[:n.value]))


def _ring_pixels(box, allowed):
    """FootprintBrush.hxx:298-316: allowed box minus open interior, as a set of (row, col)."""
    tlx, tly, brx, bry = box
    ax0, ay0, ax1, ay1 = allowed
    return {(row, col) for row in range(ay0, ay1 + 1) for col in range(ax0, ax1 + 1)
            if not (tly < row < bry and tlx < col < brx)}


def _covered(items, pitch):
    """Pixels (row, col) of the item's own row that the 4 bytes of each flat dirty-map word cover."""
    out = set()
    for row, w in items:
        for b in range(4):
            col = 4 * w + b - row * pitch
            if 0 <= col < pitch:
                out.add((row, col))
    return out


def _geom(cx, cy, wr, rad, rows, cols):
    box = (int(cx - wr), int(cy - wr), int(cx + wr), int(cy + wr))
    allowed = (max(int(cx - wr - rad), 0), max(int(cy - wr - rad), 0), min(int(cx + wr + rad), cols - 1), min(int(cy + wr + rad), rows - 1))
    return box, allowed


def test_ring_enumeration_covers_the_reference_ring(built_lib):
    """The snapshot ring pass reads the flat dirty map in aligned 4-pixel words, rectangle by rectangle
    (imprint_geom.hpp: ring_rects). Full pass: the words must cover every pixel of the reference's ring (allowed box minus
    open interior) and stay inside the allowed rows; only a thin margin of interior pixels may be read. The same code
    the device runs is evaluated here on the host (pb_ring_rects), for random and degenerate geometries and for canvas
    widths that are not a multiple of 4 (rows then start at any word alignment)."""
    from painty_b200 import api

    lib = api.lib()
    rng = np.random.default_rng(3)
    cases = []
    for _ in range(200):
        rows, cols = 180, int(rng.choice([240, 241, 243, 250]))
        cases.append(_geom(rng.uniform(-60, 260), rng.uniform(-60, 200), int(rng.integers(0, 40)), int(rng.integers(0, 40)), rows, cols) + (cols,))
    cases += [((5, 5, 5, 5), (0, 0, 20, 20), 240), ((5, 5, 6, 6), (3, 3, 9, 9), 240), ((0, 0, 100, 100), (10, 10, 50, 50), 240),
              ((10, 10, 50, 50), (10, 10, 50, 50), 240), ((-30, -30, 10, 10), (0, 0, 40, 40), 240), ((10, 10, 20, 20), (30, 30, 20, 20), 240)]
    total = interior_read = 0
    for box, allowed, pitch in cases:
        items = _ring_items(lib, box, allowed, None, pitch)
        assert len(items) - len(set(items)) <= 2 * (box[3] - box[1] + 1) + 4  # only left/right words of narrow interiors repeat
        want = _ring_pixels(box, allowed)
        got = _covered(items, pitch)
        assert want <= got, (box, allowed, pitch)
        assert all(allowed[1] <= row <= allowed[3] for row, _ in items)
        total += len(items)
        interior_read += len(got - want)
    assert total > 10000 and interior_read < 12 * total  # <= a few bytes per word outside the ring


def test_incremental_ring_covers_what_entered_the_ring(built_lib):
    """Incremental pass: given the previous imprint's geometry the words must cover ring(now) minus ring(previous) — the
    only pixels that can be dirty when the previous pass left its ring clean and the previous imprint touched nothing
    outside its open interior — for small steps, large jumps and jumps beyond the whole box."""
    from painty_b200 import api

    lib = api.lib()
    rng = np.random.default_rng(5)
    rows, cols = 200, 260
    total_full = total_inc = 0
    for it in range(250):
        wr, rad = int(rng.integers(1, 30)), int(rng.integers(0, 30))
        cx, cy = rng.uniform(-40, cols + 40), rng.uniform(-40, rows + 40)
        step = [1.5, 1.5, 6.0, 40.0, 400.0][it % 5]
        nx, ny = cx + rng.uniform(-step, step), cy + rng.uniform(-step, step)
        prev, now = _geom(cx, cy, wr, rad, rows, cols), _geom(nx, ny, wr, rad, rows, cols)
        items = _ring_items(lib, now[0], now[1], prev, cols)
        want = _ring_pixels(*now) - _ring_pixels(*prev)
        got = _covered(items, cols)
        assert want <= got, (prev, now)
        assert all(now[1][1] <= row <= now[1][3] for row, _ in items)
        total_inc += len(items)
        total_full += len(_ring_items(lib, now[0], now[1], None, cols))
    assert total_inc < total_full  # and it is cheaper than rescanning everything


def _forward_hits(cx, cy, theta, wr, rows, cols):
    """Brute-force restatement of the reference's loop (FootprintBrush.hxx:88-114): every (row, col) of the footprint box
    -> (canvas pixel, map cell), in row-major order. Returns {cell: [(px, py), ...]}."""
    c, s = math.cos(-theta), math.sin(-theta)
    r = np.arange(-wr, wr + 1)
    row, col = np.meshgrid(r, r, indexing="ij")
    rc = col * c - row * s
    rr = col * s + row * c
    half_away = lambda x: np.sign(x) * np.floor(np.abs(x) + 0.5)  # std::round
    mx, my = half_away(rc + wr).astype(np.int64), half_away(rr + wr).astype(np.int64)
    px, py = np.trunc(col + cx).astype(np.int64), np.trunc(row + cy).astype(np.int64)
    ok = (py >= 0) & (px >= 0) & (px < cols) & (py < rows)
    out = {}
    for a, b, x, y in zip(mx[ok], my[ok], px[ok], py[ok]):
        out.setdefault((int(a), int(b)), []).append((int(x), int(y)))
    return out


def test_hit_finder_matches_the_forward_mapping(built_lib):
    """The imprint kernel inverts the reference's pixel -> cell mapping: per cell, the <= 2 canvas pixels whose rotated and
    rounded position is that cell. Both device paths are evaluated on the host (pb_imprint_hits): the exact f64 test must
    reproduce the brute-force forward mapping for every cell, and the single-precision test must either agree or defer
    (n = -1) — never decide differently — over random centres and angles, exact diagonals and integer centres."""
    import ctypes as C

    from painty_b200 import api

    lib = api.lib()
    rng = np.random.default_rng(11)
    I32 = C.POINTER(C.c_int32)
    cases = [(rng.uniform(-30, 330), rng.uniform(-30, 250), rng.uniform(-np.pi, np.pi), int(rng.choice([8, 21, 45, 91]))) for _ in range(24)]
    cases += [(100.0, 80.0, 0.0, 21), (100.5, 80.5, np.pi / 2, 21), (100.0, 80.0, np.pi / 4, 45), (-3.25, -7.75, 0.3, 21),
              (12.0, 9.0, -np.pi / 4, 21), (100.25, 80.0, np.pi, 21), (299.999999, 219.0000001, 1.1, 21)]
    deferred = decided = 0
    for cx, cy, theta, wr in cases:
        rows, cols = 220, 300
        side = 2 * wr + 1
        my, mx = [a.ravel().astype(np.int32) for a in np.meshgrid(np.arange(side), np.arange(side), indexing="ij")]
        keep = np.hypot(np.abs(mx - wr) + 0.5, np.abs(my - wr) + 0.5) <= wr - 2  # the cells a compact footprint can hold
        mx, my = np.ascontiguousarray(mx[keep]), np.ascontiguousarray(my[keep])
        n = len(mx)
        want = _forward_hits(cx, cy, theta, wr, rows, cols)
        res = {}
        for mode in (0, 1):
            nh, px, py = np.zeros(n, np.int32), np.zeros(2 * n, np.int32), np.zeros(2 * n, np.int32)
            eps = 1e-6 * wr + 2e-5
            assert lib.pb_imprint_hits(C.c_double(cx), C.c_double(cy), C.c_double(theta), wr, rows, cols, C.c_int64(n),
                                       mx.ctypes.data_as(I32), my.ctypes.data_as(I32), mode, C.c_double(eps), -1,
                                       nh.ctypes.data_as(I32), px.ctypes.data_as(I32), py.ctypes.data_as(I32)) == 0
            res[mode] = (nh, px.reshape(n, 2), py.reshape(n, 2))
        for i in range(n):
            w = want.get((int(mx[i]), int(my[i])), [])
            nh, px, py = res[0]
            assert [(int(px[i, j]), int(py[i, j])) for j in range(nh[i])] == w, ("exact", cx, cy, theta, wr, mx[i], my[i])
            nh, px, py = res[1]
            if nh[i] < 0:
                deferred += 1
            else:
                decided += 1
                assert [(int(px[i, j]), int(py[i, j])) for j in range(nh[i])] == w, ("fast", cx, cy, theta, wr, mx[i], my[i])
    assert decided > 50 * max(deferred, 1) or deferred < 0.03 * decided  # the fast path decides almost everything


def _emulate_queues(n, seg_first, seg_off, ps, pn, order, pool, run, slots, rng, sub=None, sub_share=0.25):
    """Discrete-event emulation of the device protocol: per pool (GPU) the launches (runs) follow each other; within a
    launch `slots` clusters pop strokes strictly in claim order and block on (stroke, segments needed) waits; segment
    durations are random. With `sub` (0/1 per stroke) every run is TWO concurrent launches on its pool — sub-queue 1 (the
    straddlers' launch) gets max(1, sub_share * slots) of the run's clusters, sub-queue 0 the rest — and the next run starts
    when both have drained (the stream join of the library). Returns the number of strokes that finished (== n unless the
    protocol deadlocks)."""
    import heapq

    nseg = np.diff(seg_first)
    subs = np.zeros(n, np.int32) if sub is None else np.asarray(sub, np.int32)
    n_sub = 1 if sub is None else 2
    queues = {}
    for s in order:
        queues.setdefault((int(pool[s]), int(run[s]), int(subs[s])), []).append(int(s))
    progress = np.zeros(n, dtype=np.int64)
    cur_run = {p: 0 for p in range(len(slots))}
    head = {k: 0 for k in queues}
    left = {k: len(v) for k, v in queues.items()}

    def fresh(p):
        if cur_run[p] >= len(slots[p]):
            return [[] for _ in range(n_sub)]
        total = max(slots[p][cur_run[p]], 1)
        if n_sub == 1:
            return [[None] * total]
        views = max(1, int(sub_share * total))
        return [[None] * max(total - views, 1), [None] * views]

    state = {p: fresh(p) for p in range(len(slots))}
    events, now, done = [], 0.0, 0

    def ready(s, k):
        g = seg_first[s] + k
        return all(progress[ps[i]] >= pn[i] for i in range(seg_off[g], seg_off[g + 1]))

    def advance(p):
        while cur_run[p] < len(slots[p]) and all(left.get((p, cur_run[p], q), 0) == 0 for q in range(n_sub)):
            cur_run[p] += 1
            state[p] = fresh(p)
        if cur_run[p] >= len(slots[p]):
            return
        for q in range(n_sub):
            key = (p, cur_run[p], q)
            if key not in queues:
                continue
            for i in range(len(state[p][q])):
                st = state[p][q][i]
                if st is None and head[key] < len(queues[key]):
                    st = state[p][q][i] = [queues[key][head[key]], 0, False]
                    head[key] += 1
                if st is not None and not st[2] and ready(st[0], st[1]):
                    st[2] = True
                    heapq.heappush(events, (now + float(rng.uniform(0.2, 3.0)), p, q, i))

    for p in range(len(slots)):
        advance(p)
    while events:
        now, p, q, i = heapq.heappop(events)
        s, k, _ = state[p][q][i]
        k += 1
        if k >= nseg[s]:
            progress[s] = 1 << 40
            state[p][q][i] = None
            left[(p, cur_run[p], q)] -= 1
            done += 1
        else:
            progress[s] = k
            state[p][q][i] = [s, k, False]
        for r in range(len(slots)):
            advance(r)
    return done


def test_claim_order_protocol_never_deadlocks_on_eight_gpus(built_lib):
    """The band-sharded bench workload shape on 8 pools (GPUs): executor = band of the first imprint, strokes whose
    region leaves their band are single-segment (as the multi-GPU path plans them), launches split by footprint class.
    The planned claim order must let the in-order queues drain under arbitrary timing. Pure host code."""
    from painty_b200 import api, assets

    rng = np.random.default_rng(23)
    world, rpb, cols = 8, 300, 700
    rows = world * rpb
    n = 1200
    first, count, side, radius, cx, cy = [], [], [], [], [], []
    for i in range(n):
        m = int(rng.integers(1, 260))
        r = float(rng.choice([6.0, 14.0, 30.0, 45.0])) if i % 3 else float(rng.uniform(5, 45))
        x, y, a = rng.uniform(0, cols), rng.uniform(0, rows), rng.uniform(0, 2 * np.pi)
        first.append(len(cx)); count.append(m); radius.append(r); side.append(assets.footprint_geometry(r)[3])
        for _ in range(m):
            cx.append(x); cy.append(y)
            a += rng.normal(0, 0.06); x += np.cos(a); y += np.sin(a)
    cxa, cya = np.asarray(cx), np.asarray(cy)
    pool = np.zeros(n, np.int32)
    single = np.zeros(n, np.uint8)
    for s in range(n):
        a, m = first[s], count[s]
        pool[s] = min(min(max(int(cya[a]), 0), rows - 1) // rpb, world - 1)
        mm = (side[s] - 1) // 2 + radius[s] + 2.0
        lo = max(0, int(np.floor(cya[a:a + m].min() - mm)))
        hi = min(rows - 1, int(np.ceil(cya[a:a + m].max() + mm)))
        single[s] = lo < pool[s] * rpb or hi > min((pool[s] + 1) * rpb, rows) - 1
    assert 0.05 < single.mean() < 0.9
    run, slots, last = np.zeros(n, np.int32), [[] for _ in range(world)], [None] * world
    for s in range(n):
        cls = 0 if radius[s] < 10 else (1 if radius[s] < 35 else 2)
        if last[pool[s]] != cls:
            slots[pool[s]].append(9 if cls else 20)
            last[pool[s]] = cls
        run[s] = len(slots[pool[s]]) - 1
    cost = 5.2 + 0.002 * np.asarray(side, dtype=np.float64) ** 2
    order, makespan = api.plan_claim_order(rows, cols, first, count, side, radius, cx, cy, pool, run, cost, slots, 32, True,
                                           single=single, return_makespan=True)
    assert sorted(order) == list(range(n)) and makespan > 0
    seg_first, seg_len, seg_off, ps, pn = api.plan_segments(rows, cols, first, count, side, radius, cx, cy, 32, True, single=single)
    assert all(seg_first[s + 1] - seg_first[s] == 1 for s in range(n) if single[s])
    for trial in range(3):
        assert _emulate_queues(n, seg_first, seg_off, ps, pn, order, pool, run, slots, np.random.default_rng(trial)) == n
    # the check has teeth: reversing the queues makes later strokes wait for strokes stuck behind them
    assert _emulate_queues(n, seg_first, seg_off, ps, pn, order[::-1], pool, run[::-1] * 0, [[2]] * world, np.random.default_rng(0)) < n


def test_paired_launches_drain_in_claim_order(built_lib):
    """The multi-GPU protocol as it runs today: straddling strokes are segmented like the rest and run in their OWN launch,
    concurrent with the launch of the in-band strokes of the same run (two in-order queues per GPU and run, clusters split
    between them, the next run starts when both have drained). Both launches pop in the order of one global claim sequence;
    the queues must drain under arbitrary timing and for any split of the clusters. Pure host code."""
    from painty_b200 import api, assets

    rng = np.random.default_rng(5)
    world, rpb, cols = 4, 320, 640
    rows = world * rpb
    n = 900
    first, count, side, radius, cx, cy = [], [], [], [], [], []
    for i in range(n):
        m = int(rng.integers(1, 300))
        r = float(rng.choice([6.0, 14.0, 30.0, 45.0]))
        x, y, a = rng.uniform(0, cols), rng.uniform(0, rows), rng.uniform(0, 2 * np.pi)
        first.append(len(cx)); count.append(m); radius.append(r); side.append(assets.footprint_geometry(r)[3])
        for _ in range(m):
            cx.append(x); cy.append(y)
            a += rng.normal(0, 0.06); x += np.cos(a); y += np.sin(a)
    cya = np.asarray(cy)
    pool, straddles = np.zeros(n, np.int32), np.zeros(n, np.int32)
    for s in range(n):
        a, m = first[s], count[s]
        pool[s] = min(min(max(int(cya[a]), 0), rows - 1) // rpb, world - 1)
        mm = (side[s] - 1) // 2 + radius[s] + 2.0
        lo = max(0, int(np.floor(cya[a:a + m].min() - mm)))
        hi = min(rows - 1, int(np.ceil(cya[a:a + m].max() + mm)))
        straddles[s] = lo < pool[s] * rpb or hi > min((pool[s] + 1) * rpb, rows) - 1
    assert 0.05 < straddles.mean() < 0.9
    run, slots, last = np.zeros(n, np.int32), [[] for _ in range(world)], [None] * world
    for s in range(n):
        cls = 0 if radius[s] < 35 else 1
        if last[pool[s]] != cls:
            slots[pool[s]].append(15 if cls == 0 else 7)
            last[pool[s]] = cls
        run[s] = len(slots[pool[s]]) - 1
    cost = 5.2 + 0.002 * np.asarray(side, dtype=np.float64) ** 2
    order = api.plan_claim_order(rows, cols, first, count, side, radius, cx, cy, pool, run, cost, slots, 32, True)
    assert sorted(order) == list(range(n))
    seg_first, seg_len, seg_off, ps, pn = api.plan_segments(rows, cols, first, count, side, radius, cx, cy, 32, True)
    assert any(seg_first[s + 1] - seg_first[s] > 1 for s in range(n) if straddles[s])  # straddlers are segmented too
    for trial, share in enumerate((0.15, 0.3, 0.6)):
        assert _emulate_queues(n, seg_first, seg_off, ps, pn, order, pool, run, slots, np.random.default_rng(trial), sub=straddles,
                               sub_share=share) == n
    # teeth: popping the queues back to front strands later strokes behind earlier ones
    assert _emulate_queues(n, seg_first, seg_off, ps, pn, order[::-1], pool, run * 0, [[4]] * world, np.random.default_rng(0),
                           sub=straddles) < n


def _dictionary_restatement(tex):
    """Plain-Python restatement of TextureBrushDictionary::createBrushTexturesFromFolder / lookup
    (renderer/src/TextureBrushDictionary.cxx:25-69, 81-164) for the test: returns lookup(path, brush_size)."""
    sizes = sorted(set(t[1] for t in tex))
    groups = []
    for s in sizes:
        lens = sorted(set(t[2] for t in tex if t[1] == s))
        groups.append([[i for i, t in enumerate(tex) if t[1] == s and t[2] == l] for l in lens])
    avg_sizes = [sum(tex[i][3].shape[0] for g in gs for i in g) * (1.0 / sum(len(g) for g in gs)) for gs in groups]
    avg_len = [[sum(tex[i][3].shape[1] for i in g) * (1.0 / len(g)) for g in gs] for gs in groups]

    def lookup(path, brush_size):
        length = sum(math.hypot(path[i][0] - path[i + 1][0], path[i][1] - path[i + 1][1]) for i in range(len(path) - 1))
        i0, i1 = 0, 1
        mr = avg_sizes[0]
        for i in range(len(avg_sizes)):
            d = abs(avg_sizes[i] - brush_size)
            if d < mr:
                mr, i0 = d, i
        ml = avg_len[i0][0]
        for i in range(min(len(avg_sizes), len(avg_len[i0]))):
            d = abs(avg_len[i0][i] - length)
            if d < ml:
                ml, i1 = d, i
        return i0, i1, groups[i0][i1]

    return lookup


def test_texture_dictionary_lookup_rule(built_lib):
    """pb_texdict_* == the restatement above on the shipped 236 textures (5 size keys x 5 length keys), incl. the
    reference's quirk that the running minima start from VALUES (so small brushes / short strokes keep the defaults)."""
    from painty_b200 import api, assets

    tex = assets.brush_textures()
    assert len(tex) == 236 and sorted(set(t[1] for t in tex)) == [1, 2, 4, 10, 14]
    for _, _, _, m in tex[::17]:
        assert m.dtype == np.float64 and m.min() == 0.0 and abs(m.max() - 1.0) < 1e-15  # cv::normalize(NORM_MINMAX)
    dic = api.TextureBrushDictionary([t[1] for t in tex], [t[2] for t in tex], [t[3].shape[0] for t in tex],
                                     [t[3].shape[1] for t in tex])
    want = _dictionary_restatement(tex)
    r = np.random.default_rng(5)
    seen = set()
    for _ in range(300):
        n = int(r.integers(2, 20))
        path = r.uniform(0, 4000, 2) + np.cumsum(r.normal(0, r.uniform(1, 120), (n, 2)), axis=0)
        size = float(r.uniform(1, 400))
        i0, i1, cand = dic.lookup(path, size)
        w0, w1, wc = want(path, size)
        assert (i0, i1) == (w0, w1) and list(cand) == wc
        seen.add((i0, i1))
    assert len(seen) >= 10
    # defaults survive when nothing beats the seeds: i0 = 0, i1 = 1
    assert dic.lookup([(0, 0), (1e9, 0)], 1e9)[:2] == (0, 1)


def test_canvas_pattern_asset():
    """canvas_patterns/0.png as CanvasGpu::clear prepares it (CanvasGpu.cxx:27-40): linear RGB, float32, LANCZOS4."""
    from painty_b200 import assets

    p = assets.canvas_pattern(2048, 2048)
    assert p.shape == (2048, 2048, 3) and p.dtype == np.float64
    lin198 = float(np.float32(assets.srgb_to_linear(198 / 255.0)))
    assert p.min() == lin198 and p.max() == 1.0  # the shipped pattern's sRGB values span 198..255
    q = assets.canvas_pattern(270, 480)
    assert q.shape == (270, 480, 3) and 0.5 < q.min() and q.max() < 1.2


def test_lanczos4_taps_reproduce_cv2_resize(built_lib, port):
    """The resize behind pb_canvas_compose_lab_scaled: the library's tap tables + the two separable passes (restated here in
    numpy with the kernel's operation order) == cv2.resize(INTER_LANCZOS4) bit for bit on f64 images, for down- and
    up-scaling incl. the sbr painter's canvas -> target-image factors; and the CIELab conversion of the port is sane."""
    import cv2

    from painty_b200 import api

    r = np.random.default_rng(1)
    for (R, C, oR, oC) in ((37, 53, 20, 31), (64, 48, 32, 24), (50, 70, 23, 33), (30, 30, 45, 50), (216, 384, 108, 192), (90, 160, 77, 102)):
        src = r.uniform(-100, 100, (R, C, 3))
        xo, xa = api.lanczos4_taps(C, oC)
        yo, ya = api.lanczos4_taps(R, oR)
        tmp = None
        for j in range(8):
            p = src[:, np.clip(xo - 3 + j, 0, C - 1), :] * xa[:, j].astype(np.float64)[None, :, None]
            tmp = p if tmp is None else tmp + p
        out = None
        for k in range(8):
            p = tmp[np.clip(yo - 3 + k, 0, R - 1)] * ya[:, k].astype(np.float64)[:, None, None]
            out = p if out is None else out + p
        assert np.array_equal(out, cv2.resize(src, (oC, oR), interpolation=cv2.INTER_LANCZOS4)), (R, C, oR, oC)
    lab = port.rgb2lab(np.array([[1.0, 1.0, 1.0], [0.0, 0.0, 0.0], [0.2, 0.5, 0.1], [0.001, 0.002, 0.0005]]))
    assert abs(lab[0, 0] - 100.0) < 1e-3 and abs(lab[0, 1]) < 1e-2 and abs(lab[0, 2]) < 1e-2  # D65 white
    assert np.abs(lab[1]).max() < 1e-12
    assert 0 < lab[3, 0] < 3  # linear branch of f
