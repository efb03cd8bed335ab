"""Two GPUs, one canvas: footprint strokes on a band-sharded canvas (peer memory over NVLink, cross-GPU stroke
dependencies) must reproduce the single-GPU render of the same stroke list bit for bit. Needs >= 2 GPUs
(`gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`); skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload(rows, cols, n):
    from painty_b200 import api, assets
    from tests.workloads import sbr_strokes

    strokes = sbr_strokes(rows, cols, n, seed=5, sizes=(60, 40, 30, 20), safe_radius=assets.snap_to_safe_radius)
    rec = np.zeros(len(strokes), dtype=api.STROKE_DTYPE)
    xs, ys, ts, first = [], [], [], 0
    for i, s in enumerate(strokes):
        cx, cy, th = api.expand_stroke(s["path"], mode=0)
        rec[i] = (s["radius"], s["K"], s["S"], first, len(cx))
        first += len(cx)
        xs.append(cx), ys.append(cy), ts.append(th)
    return rec, np.concatenate(xs), np.concatenate(ys), np.concatenate(ts), sorted(set(float(s["radius"]) for s in strokes))


def _worker(rank, world, port, prec, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from painty_b200 import api
    from painty_b200.dist import DistCanvas

    rows, cols = 600, 500
    rec, cx, cy, th, radii = _workload(rows, cols, 120)
    ctx = api.Context(rank, prec)
    dc = DistCanvas(ctx, rows, cols, dist)
    br = api.FootprintBrush(ctx, radii[0])
    for r in radii:
        br.register_radius(r)
    dc.stroke_batch(br, rec, cx, cy, th)
    dc.stroke_batch(br, rec[::-1].copy(), cx, cy, th)  # a second batch on the painted canvas (new epoch, stale snapshot)
    st = dc.canvas.download("KSV")
    ok = True
    # final assembly: compose with the peer-store gather epilogue, once to every rank and once to rank 0 only
    dc.attach_image(root=None)
    dc.compose_gather()
    dc.finish_gather()
    image_all = dc.download_image()
    dc.attach_image(root=0)
    dc.compose_gather()
    dc.finish_gather()
    image_root = dc.download_image() if rank == 0 else None
    if rank == 0:  # reference: the same two batches on one GPU
        full = api.Canvas(ctx, rows, cols)
        b1 = api.FootprintBrush(ctx, radii[0])
        for r in radii:
            b1.register_radius(r)
        b1.stroke_batch(full, rec, cx, cy, th)
        b1.stroke_batch(full, rec[::-1].copy(), cx, cy, th)
        want = full.download("KSV")
        gathered = [None] * world
        dist.gather_object((dc.row_begin, dc.row_end, st), gathered, dst=0)
        for b, e, part in gathered:
            for k in "KSV":
                ok = ok and np.array_equal(part[k], want[k][b:e])
        ok = ok and float(want["V"].sum()) > 0
        want_R = full.compose()
        ok = ok and np.array_equal(image_root, want_R)
        images = [None] * world
        dist.gather_object(image_all, images, dst=0)
        for im in images:
            ok = ok and np.array_equal(im, want_R)
    else:
        dist.gather_object((dc.row_begin, dc.row_end, st), None, dst=0)
        dist.gather_object(image_all, None, dst=0)
    dc.close()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("prec", [0, 1])
def test_band_canvas_equals_single_gpu(built_lib, prec, world):
    """600 x 500 canvas in `world` bands (75 rows each at 8 GPUs: the staging window of a stroke segment then spans up
    to 3-4 bands)."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, prec, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]
