#!/usr/bin/env python
"""SURVEY.md §8d config 4 — one 16K x 16K canvas, synthetic sbr-shaped footprint strokes, row bands over 1/2/4/8 GPUs
(STRONG scaling: the same stroke list on every GPU count), final reflectance assembled with an NCCL all_gather.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      benchmarks/config4_16k.py [--strokes 100000] [--size 16384]

Stroke radii are those of the 4K config (26..151 px, the brush sizes of sbr_config.json at a 3840-wide canvas) spread
over the 16K x 16K sheet. 1M strokes (213 M imprints) exceed what this Python harness can generate and hold
comfortably; the default is 100 000 strokes, throughput is reported in stroke-pixels/s so runs are comparable.
A check value (sum of the reflectance image) is printed: it must be identical for every N.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from painty_b200 import api, assets, bands  # noqa: E402
from painty_b200.dist import DistCanvas  # noqa: E402
from tests.workloads import sbr_strokes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strokes", type=int, default=100000)
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows = cols = args.size
    pk, ps = assets.palette("lindemeier_measured")
    # sizes chosen so that radius = U[0.35,0.5] * size * (cols/1024) reproduces the 4K config's 26..151 px
    k = 3840.0 / cols
    strokes = sbr_strokes(rows, cols, args.strokes, seed=1234, sizes=(80 * k, 60 * k, 30 * k, 20 * k),
                          safe_radius=assets.snap_to_safe_radius, palette=(pk, ps))
    rec = np.zeros(len(strokes), dtype=api.STROKE_DTYPE)
    xs, ys, ts, first = [], [], [], 0
    for i, s in enumerate(strokes):
        cx, cy, th = api.expand_stroke(s["path"], mode=0)
        rec[i] = (s["radius"], s["K"], s["S"], first, len(cx))
        first += len(cx)
        xs.append(cx), ys.append(cy), ts.append(th)
    cx, cy, th = np.concatenate(xs), np.concatenate(ys), np.concatenate(ts)
    radii = sorted(set(float(s["radius"]) for s in strokes))
    ctx = api.Context(local, api.F32)
    dc = DistCanvas(ctx, rows, cols, dist)
    br = api.FootprintBrush(ctx, radii[0])
    for r in radii:
        br.register_radius(r)
    dc.attach(br)
    n_px = dc.canvas.store_rows * cols
    d_R = torch.empty((3, n_px), dtype=torch.float32, device="cuda")
    br.enable_visited_count(True)
    best, visited = 1e9, 0
    for it in range(args.reps + 1):
        dc.canvas.clear()
        br.updateSnapshot(dc.canvas)
        ctx.synchronize()
        dist.barrier()
        v0 = br.counters()[0]
        t0 = time.perf_counter()
        dc.stroke_batch(br, rec, cx, cy, th)
        dc.canvas.compose_device(d_R.data_ptr(), n_px)
        ctx.synchronize()
        img = bands.gather_bands(d_R, rows, cols, world, dist) if world > 1 else d_R.reshape(3, rows, cols)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it == 0:
            tv = torch.tensor([br.counters()[0] - v0], dtype=torch.int64, device="cuda")
            dist.all_reduce(tv)
            visited = int(tv.item())
            br.enable_visited_count(False)
        else:
            best = min(best, float(dt.item()))
        check = float(img.double().sum().item())
        del img
    if rank == 0:
        print(json.dumps(dict(config="16K config 4", canvas=[rows, cols], strokes=len(rec), imprints=int(len(cx)), n_gpus=world,
                              seconds=best, stroke_pixels=visited, stroke_pixels_per_s=visited / best, checksum_R=check)), flush=True)
    dc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
