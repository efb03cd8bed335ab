#!/usr/bin/env python
"""SURVEY.md §8d config 1 (painty_gui default stroke) and config 3 (8K canvas, texture-brush strokes), GPU vs the
CPU reference on the same inputs, with parity of the result checked in the same run.

  python benchmarks/gui_and_texture.py [--tex-strokes 2000] [--tex-cpu-strokes 40]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import cpu as ocpu  # noqa: E402  (CPU baseline + parity check only)
from painty_b200 import api, assets  # noqa: E402
from tests.workloads import sbr_strokes  # noqa: E402


def gui_config(ctx, cpu, rows, cols):
    K, S = api.ComputeScatteringAndAbsorption([.2, .05, .4], [.6, .3, .7])
    cx, cy, th = api.expand_stroke([(100.3, 200.7), (400.9, 260.2), (700.1, 180.4)], mode=1)
    cvo = cpu.canvas(rows, cols)
    bro = cpu.footprint_brush(30.0)
    bro.dip(K, S)
    t_cpu = bro.imprint_batch(cvo, cx, cy, th)
    t0 = time.perf_counter()
    want = cvo.compose()
    t_cpu_compose = time.perf_counter() - t0
    cv = api.Canvas(ctx, rows, cols)
    br = api.FootprintBrush(ctx, 30.0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    out = np.empty((rows, cols, 3))
    best = 1e9
    for _ in range(5):
        cv.clear()
        br.updateSnapshot(cv)
        br.dip((K, S))
        ctx.synchronize()
        t0 = time.perf_counter()
        br.imprint_batch(cv, cx, cy, th)
        cv.compose(out)
        best = min(best, time.perf_counter() - t0)
    err = float(np.abs(out - want).max())
    return dict(config="gui %dx%d, 615 imprints r=30 + compose" % (cols, rows), gpu_ms_e2e=best * 1e3, cpu_ms=(t_cpu + t_cpu_compose) * 1e3,
                speedup=(t_cpu + t_cpu_compose) / best, max_abs_err=err)


def texture_config(ctx, cpu, n_strokes, n_cpu):
    rows, cols = 4320, 7680
    pk, ps = assets.palette("lindemeier_measured")
    strokes = sbr_strokes(rows, cols, n_strokes, seed=4321, palette=(pk, ps))
    rec = np.zeros(len(strokes), dtype=api.TSTROKE_DTYPE)
    verts, first = [], 0
    for i, s in enumerate(strokes):
        rec[i] = (s["radius"], s["K"], s["S"], 0.05, first, len(s["path"]), 0)  # thicknessScale 0.05 like sbr_config.json
        first += len(s["path"])
        verts.append(s["path"])
    verts = np.concatenate(verts)
    cv = api.Canvas(ctx, rows, cols)
    tb = api.TextureBrush(ctx)
    stream = torch.cuda.ExternalStream(ctx.stream)
    best = 1e9
    for _ in range(3):
        cv.clear()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0 = tb.counters()
        e0.record(stream)
        tb.stroke_batch(cv, rec, verts)
        e1.record(stream)
        ctx.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
        pixels = tb.counters() - p0
    # CPU: the first n_cpu strokes on a fresh canvas (single thread, like the reference); parity on the same prefix
    cvo = cpu.canvas(rows, cols)
    tbo = cpu.texture_brush()
    t_cpu = 0.0
    for s in strokes[:n_cpu]:
        tbo.set_radius(s["radius"])
        tbo.dip(s["K"], s["S"])
        tbo.set_thickness_scale(0.05)
        t_cpu += tbo.paint_stroke(cvo, s["path"])
    cv2 = api.Canvas(ctx, rows, cols)
    tb2 = api.TextureBrush(ctx)
    nv = int(rec["first_vertex"][n_cpu]) if n_cpu < len(rec) else len(verts)
    tb2.stroke_batch(cv2, rec[:n_cpu], verts[:nv])
    err = float(np.abs(cv2.compose() - cvo.compose()).max())
    cpu_px = int(ocpu.C.c_uint64(cpu.fn("tbrush_pixels", ocpu.C.c_uint64, [ocpu.C.c_void_p])(tbo.h)).value) if cpu.kind == "port" else None
    return dict(config="8K 7680x4320, %d texture strokes" % n_strokes, gpu_s=best, stroke_pixels=int(pixels), gpu_stroke_px_per_s=pixels / best,
                cpu_strokes=n_cpu, cpu_s=t_cpu, cpu_ms_per_stroke=t_cpu / n_cpu * 1e3, gpu_ms_per_stroke=best / n_strokes * 1e3,
                speedup_per_stroke=(t_cpu / n_cpu) / (best / n_strokes), max_abs_err_prefix=err, cpu_pixels_prefix=cpu_px)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tex-strokes", type=int, default=2000)
    ap.add_argument("--tex-cpu-strokes", type=int, default=40)
    args = ap.parse_args()
    ocpu.build()
    cpu = ocpu.Cpu("ref" if ocpu.have_ref() else "port")
    ctx = api.Context(0, api.F32)
    for rows, cols in ((768, 1024), (1024, 1024)):
        print(json.dumps(gui_config(ctx, cpu, rows, cols)), flush=True)
    print(json.dumps(texture_config(ctx, cpu, args.tex_strokes, args.tex_cpu_strokes)), flush=True)


if __name__ == "__main__":
    main()
