"""N > 1 host logic on CPU: band partitioning, stroke routing, wave levels, and the gather of band results with a
world_size-2 gloo process group. Texture-stroke routing is checked for exactness against the CPU oracle: rendering
each band separately with only its routed strokes must reproduce the single-canvas result on the band's rows."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from painty_b200 import bands


def test_band_ranges():
    assert bands.band_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert bands.band_ranges(16384, 8)[3] == (6144, 8192)
    for rows, w in [(2160, 8), (7, 8), (1, 2)]:
        b = bands.band_ranges(rows, w)
        assert b[0][0] == 0 and b[-1][1] == rows and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_stroke_levels_are_conflict_free():
    rng = np.random.default_rng(0)
    regs = []
    for _ in range(300):
        x, y, w, h = rng.integers(0, 900), rng.integers(0, 700), rng.integers(5, 200), rng.integers(5, 200)
        regs.append((int(x), int(y), int(min(x + w, 999)), int(min(y + h, 799))))
    lv = bands.stroke_levels(regs, 800, 1000, tile=32)
    assert lv.min() >= 1
    for i in range(len(regs)):
        for j in range(i):
            a, b = regs[i], regs[j]
            overlap = a[0] <= b[2] and b[0] <= a[2] and a[1] <= b[3] and b[1] <= a[3]
            if overlap:
                assert lv[i] > lv[j]  # an overlapping later stroke always runs in a later wave


def test_footprint_routing_flags_straddlers():
    radii = [30.0, 30.0, 30.0]
    cys = [np.linspace(100, 140, 40), np.linspace(480, 520, 40), np.linspace(900, 950, 50)]
    owner, straddle = bands.route_footprint_strokes(radii, cys, 1000, 2)
    assert owner.tolist() == [0, 0, 1]
    assert straddle.tolist() == [False, True, False]


def test_texture_routing_is_exact_per_band(port):
    """Each band rendered on its own canvas with only the routed strokes == the rows of the full render."""
    rows, cols, world = 240, 200, 3
    r = np.random.default_rng(4)
    strokes = []
    for _ in range(25):
        m = int(r.integers(1, 7))
        p0 = r.uniform(-10, [cols + 10, rows + 10])
        strokes.append(dict(radius=float(r.uniform(3, 25)), K=r.uniform(.05, 1.5, 3), S=r.uniform(.05, 1, 3),
                            scale=float(r.uniform(.05, 1)), path=p0 + np.cumsum(r.normal(0, 10, (m, 2)), axis=0)))

    def render(indices):
        cv, tb = port.canvas(rows, cols), port.texture_brush()
        radius_state = None
        for i, s in enumerate(strokes):
            # the brush's fuzzy radius state evolves with EVERY stroke of the list, routed or not
            tb.set_radius(s["radius"])
            if i in indices:
                tb.dip(s["K"], s["S"])
                tb.set_thickness_scale(s["scale"])
                tb.paint_stroke(cv, s["path"])
        return cv.get()

    full = render(set(range(len(strokes))))
    routed = bands.route_texture_strokes(strokes, rows, world)
    assert sum(len(x) for x in routed) < world * len(strokes)  # routing really prunes
    for rank, (b, e) in enumerate(bands.band_ranges(rows, world)):
        part = render(set(routed[rank]))
        for k in "KSV":
            assert np.array_equal(part[k][b:e], full[k][b:e]), (rank, k)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gather_worker(rank, world, port_no, rows, cols, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.arange(3 * rows * cols, dtype=torch.float32).reshape(3, rows, cols)
    b, e = bands.band_ranges(rows, world)[rank]
    mine = full[:, b:e].reshape(3, -1).contiguous()
    out = bands.gather_bands(mine, rows, cols, world, dist)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("rows", [8, 7])
def test_gather_bands_gloo_world2(rows):
    world, cols = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port_no, rows, cols, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
