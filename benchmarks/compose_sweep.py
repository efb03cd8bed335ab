#!/usr/bin/env python
"""SURVEY.md §8d config 5 — KM compose-only throughput sweep: canvases 4K .. 32K, 1..8 stacked paint layers,
FP32 vs the FP64 validation mode, every pixel wet (roofline variant). Prints one JSON line per point and a table.

  python benchmarks/compose_sweep.py [--max-gb 120] [--cpu]

Timing: CUDA events on the context stream around `reps` back-to-back launches (no launch gaps), inputs larger than
the 126 MB L2 (smallest case 4K: 431 MB touched per launch). Bytes per pixel: (28 L + 24) FP32, x2 FP64
(L layers of 7 planes read once, R0 read once, R written once; R stays in registers between layers).
32K x 32K x 8 layers (240 GB FP32) does not fit one B200: points that exceed --max-gb are skipped and noted.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from painty_b200 import api  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gb", type=float, default=120.0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cpu", action="store_true", help="also time the CPU reference on a 1024x768 sample")
    args = ap.parse_args()
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    sizes = [("4K", 2160, 3840), ("8K", 4320, 7680), ("16K", 16384, 16384), ("32K", 32768, 32768)]
    rows_out = []
    for prec, dt, es in ((api.F32, torch.float32, 4), (api.F64, torch.float64, 8)):
        ctx = api.Context(0, prec)
        stream = torch.cuda.ExternalStream(ctx.stream, device=0)
        for name, rows, cols in sizes:
            n = rows * cols
            for L in (1, 2, 4, 8):
                need = (7 * L + 6) * n * es / 1e9
                if need > args.max_gb:
                    rows_out.append(dict(canvas=name, layers=L, dtype="f64" if prec else "f32", skipped="needs %.0f GB" % need))
                    continue
                g = torch.Generator(device="cuda").manual_seed(42)
                layers = []
                for l in range(L):
                    pl = torch.empty((7, n), dtype=dt, device="cuda")
                    pl[0:3] = torch.exp(torch.rand((3, n), device="cuda", generator=g, dtype=dt) * (np.log(4.32) - np.log(1e-3)) + np.log(1e-3))
                    pl[3:6] = torch.exp(torch.rand((3, n), device="cuda", generator=g, dtype=dt) * (np.log(1.21) - np.log(1e-3)) + np.log(1e-3))
                    pl[6] = torch.rand(n, device="cuda", generator=g, dtype=dt) * 0.9 + 0.05
                    layers.append(pl)
                r0 = torch.rand((3, n), device="cuda", generator=g, dtype=dt) * 0.96 + 0.02
                out = torch.empty((3, n), dtype=dt, device="cuda")
                torch.cuda.synchronize()
                K = [[pl[i].data_ptr() for i in range(3)] for pl in layers]
                S = [[pl[3 + i].data_ptr() for i in range(3)] for pl in layers]
                V = [pl[6].data_ptr() for pl in layers]
                R0 = [r0[i].data_ptr() for i in range(3)]
                R = [out[i].data_ptr() for i in range(3)]

                def launch():
                    if L == 1:
                        ctx.km_compose_planes(n, K[0], S[0], V[0], R0, R)
                    else:
                        ctx.km_compose_stacked_planes(n, K, S, V, R0, R)

                for _ in range(3):
                    launch()
                ctx.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.reps):
                    launch()
                e1.record(stream)
                ctx.synchronize()
                ms = e0.elapsed_time(e1) / args.reps
                bpp = (7 * L + 6) * es
                gbs = bpp * n / ms / 1e6
                row = dict(canvas=name, rows=rows, cols=cols, layers=L, dtype="f64" if prec else "f32", us=ms * 1e3, gpx_s=n / ms / 1e6,
                           bytes_per_px=bpp, gb_s=gbs, frac_of_measured_peak=gbs / peak, frac_of_8tbs=gbs / 8000.0)
                rows_out.append(row)
                print(json.dumps(row), flush=True)
                del layers, r0, out
                torch.cuda.empty_cache()
        ctx.close()
    if args.cpu:
        from oracle import cpu as ocpu
        from tests.workloads import km_random_planes

        c = ocpu.Cpu("ref" if ocpu.have_ref() else "port")
        K, S, V, R0 = km_random_planes(768, 1024, seed=42)
        for th in (1, os.cpu_count() or 1):
            t, _ = c.compose_timed(K, S, V, R0, threads=th)
            print(json.dumps(dict(cpu=c.kind, threads=th, canvas="1024x768", mpx_s=768 * 1024 / t / 1e6)), flush=True)
    print("\n| canvas | layers | dtype | us | Gpx/s | GB/s | of measured peak |")
    print("|---|---|---|---|---|---|---|")
    for r in rows_out:
        if "skipped" in r:
            print("| %s | %d | %s | skipped: %s | | | |" % (r["canvas"], r["layers"], r["dtype"], r["skipped"]))
        else:
            print("| %s | %d | %s | %.1f | %.1f | %.0f | %.2f |" % (r["canvas"], r["layers"], r["dtype"], r["us"], r["gpx_s"], r["gb_s"], r["frac_of_measured_peak"]))


if __name__ == "__main__":
    main()
