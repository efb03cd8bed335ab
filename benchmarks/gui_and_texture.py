#!/usr/bin/env python
"""SURVEY.md §8d config 1 (painty_gui default stroke) and config 3 (8K canvas, texture-brush strokes), GPU vs the
CPU reference on the same inputs, with parity of the result checked in the same run.

  python benchmarks/gui_and_texture.py [--tex-strokes 100000] [--tex-cpu-strokes 24] [--out FILE]
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


def texture_config(ctx, cpu, n_strokes, n_cpu, rows=4320, cols=7680, seed=4321):
    """BASELINE configs[2] / SURVEY §8d config 3: 7680x4320, R0 = canvas_patterns/0.png (CanvasGpu.cxx:27-40), n_strokes
    sbr-shaped texture-brush strokes (config 2 x 10) of the 14-pigment palette, thickness texture per stroke chosen from the
    236 data/textures maps by the dictionary rule (TextureBrushDictionary.cxx:25-69, the random draw among the candidates
    seeded and recorded in the stroke list), thicknessScale 0.05, smudge off. CPU semantics: TextureBrush::paintStroke with
    that texture installed as the brush's thickness map."""
    pk, ps = assets.palette("lindemeier_measured")
    strokes = sbr_strokes(rows, cols, n_strokes, seed=seed, palette=(pk, ps))
    tex = assets.brush_textures()
    dic = api.TextureBrushDictionary([t[1] for t in tex], [t[2] for t in tex], [t[3].shape[0] for t in tex],
                                     [t[3].shape[1] for t in tex])
    R0 = assets.canvas_pattern(rows, cols)
    cv = api.Canvas(ctx, rows, cols)
    tb = api.TextureBrush(ctx)
    ids = [tb.addTexture(t[3]) for t in tex]  # the whole atlas lives in HBM (20.9 M texels, 168 MB f64)
    rng = np.random.default_rng(seed + 1)
    rec = np.zeros(len(strokes), dtype=api.TSTROKE_DTYPE)
    verts, first, picks = [], 0, []
    t0 = time.perf_counter()
    for i, s in enumerate(strokes):
        _, _, cand = dic.lookup(s["path"], 2.0 * s["radius"])  # TextureBrushGpu.cxx:238
        pick = int(cand[int(rng.integers(0, len(cand)))])
        picks.append(pick)
        rec[i] = (s["radius"], s["K"], s["S"], 0.05, first, len(s["path"]), ids[pick])  # thicknessScale 0.05 like sbr_config.json
        first += len(s["path"])
        verts.append(s["path"])
    lookup_s = time.perf_counter() - t0
    verts = np.concatenate(verts)
    stream = torch.cuda.ExternalStream(ctx.stream)
    best, e2e = 1e9, 1e9
    out = np.empty((rows, cols, 3))
    for _ in range(3):
        cv.setBackground(R0)  # clear + substrate
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0 = tb.counters()
        t0 = time.perf_counter()
        e0.record(stream)
        tb.stroke_batch(cv, rec, verts)
        e1.record(stream)
        cv.compose(out)  # host AoS f64 like Renderer::compose
        e2e = min(e2e, time.perf_counter() - t0)
        best = min(best, e0.elapsed_time(e1) * 1e-3)
        pixels = tb.counters() - p0
    checksum = float(out.sum())
    # CPU: a prefix of n_cpu strokes on a fresh canvas (single thread, like the reference); parity on the same prefix
    cvo = cpu.canvas(rows, cols)
    cvo.set_background(R0)
    cpu_brushes = {}
    t_cpu = 0.0
    state = 1.0e9  # the ONE brush's radius: setRadius only acts on a change >= 0.5 (TextureBrush.hxx:33-41)
    tb.setRadius(state)
    for s, pick in zip(strokes[:n_cpu], picks[:n_cpu]):
        if pick not in cpu_brushes:
            cpu_brushes[pick] = cpu.texture_brush(tex[pick][3])
        tbo = cpu_brushes[pick]
        if not abs(state - s["radius"]) < 0.5:
            state = s["radius"]
        tbo.set_radius(state + 1000.0)  # force the per-texture CPU brush to exactly the single brush's effective radius
        tbo.set_radius(state)
        tbo.dip(s["K"], s["S"])
        tbo.set_thickness_scale(0.05)
        t_cpu += tbo.paint_stroke(cvo, s["path"])
    cv.setBackground(R0)
    nv = int(rec["first_vertex"][n_cpu]) if n_cpu < len(rec) else len(verts)
    tb.stroke_batch(cv, rec[:n_cpu], verts[:nv])
    got, want = cv.compose(), cvo.compose()
    err = float(np.abs(got - want).max())
    wet_equal = bool(np.array_equal(cv.download("V")["V"] > 0, cvo.get()["V"] > 0))
    return dict(config="config 3: %dx%d, canvas-pattern substrate, %d texture strokes, 236-texture dictionary" % (cols, rows, n_strokes),
                gpu_s=best, gpu_e2e_s_incl_compose_and_d2h=e2e, stroke_pixels=int(pixels), gpu_stroke_px_per_s=pixels / best,
                deposit_GBps_at_56B_per_px=56 * pixels / best / 1e9, host_dictionary_lookup_s=lookup_s,
                textures_used=len(set(picks)), reflectance_checksum=checksum,
                cpu_kind=cpu.kind, cpu_strokes=n_cpu, cpu_s=t_cpu, cpu_ms_per_stroke=t_cpu / n_cpu * 1e3,
                gpu_ms_per_stroke=best / n_strokes * 1e3, speedup_per_stroke=(t_cpu / n_cpu) / (best / n_strokes),
                parity_prefix={"max_abs_err": err, "tolerance": 1e-4, "wet_px_equal": wet_equal, "strokes": n_cpu})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tex-strokes", type=int, default=100000)
    ap.add_argument("--tex-cpu-strokes", type=int, default=24)
    ap.add_argument("--out", default=None, help="also write the records to this JSON file (profiles/r02_config1_config3.json)")
    args = ap.parse_args()
    ocpu.build()
    cpu = ocpu.Cpu("ref" if ocpu.have_ref() else "port")
    ctx = api.Context(0, api.F32)
    records = []
    for rows, cols in ((768, 1024), (1024, 1024)):
        records.append(gui_config(ctx, cpu, rows, cols))
        print(json.dumps(records[-1]), flush=True)
    records.append(texture_config(ctx, cpu, args.tex_strokes, args.tex_cpu_strokes))
    print(json.dumps(records[-1]), flush=True)
    if args.out:
        json.dump(records, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
