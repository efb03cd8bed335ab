#!/usr/bin/env python
"""Parity at the bench's full size without a CPU run: the FP64 validation mode reproduces the CPU renderer's
K/S/V bit for bit (tests/test_imprint_gpu.py), so the FP32 product mode is compared against it on the whole
sbr-style 4K workload: max |R32 - R64| must stay within the 1e-4 budget of BASELINE.json.

  python benchmarks/parity_at_scale.py [--strokes 10000]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from painty_b200 import api  # noqa: E402


def render(prec, rec, cx, cy, th, radii):
    ctx = api.Context(0, prec)
    cv = api.Canvas(ctx, bench.ROWS, bench.COLS)
    br = api.FootprintBrush(ctx, radii[0])
    for r in radii:
        br.register_radius(r)
    br.stroke_batch(cv, rec, cx, cy, th)
    R = cv.compose()
    V = cv.download("V")["V"]
    del br, cv
    ctx.close()
    return R, V


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strokes", type=int, default=10000)
    args = ap.parse_args()
    _, rec, cx, cy, th, radii = bench.build_workload(args.strokes)
    R32, V32 = render(api.F32, rec, cx, cy, th, radii)
    R64, V64 = render(api.F64, rec, cx, cy, th, radii)
    err = np.abs(R32 - R64)
    out = dict(strokes=int(args.strokes), imprints=int(len(cx)), wet_pixels=int((V64 > 0).sum()), max_abs_err_R=float(np.nanmax(err)),
               pixels_over_1e_4=int((np.nanmax(err, axis=2) > 1e-4).sum()), p9999_err=float(np.nanquantile(err, 0.9999)),
               nan_mismatch=int((np.isnan(R32) != np.isnan(R64)).sum()), max_rel_err_V=float(np.max(np.abs(V32 - V64) / np.maximum(V64, 1e-3))),
               checksum_R64=float(np.nansum(R64)))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
