#!/usr/bin/env python
"""BASELINE configs[3] / SURVEY.md §8d config 4 — one 16384 x 16384 canvas, 1 M synthetic sbr-shaped footprint strokes
(config 2 x 100: the same four brush-size passes and radii, 26..151 px, spread over the 16K sheet), row bands over
1/2/4/8 GPUs. STRONG scaling: the same stroke list on every GPU count.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      benchmarks/config4_16k.py [--strokes 1000000] [--batch 100000] [--size 16384] [--out FILE] [--cpu-strokes 48]

The list is submitted in batches of --batch strokes like a painter that keeps producing strokes: the host half of batch
k+1 (dataflow graph, claim order, per-imprint constants: pb_fbrush_plan_stroke_batch) runs on a helper thread while the
GPUs execute batch k. Timed: all batches + compose with the band-gather epilogue (the image is assembled on rank 0 by
the compose kernels' peer stores), max over ranks. Stroke generation and stroke -> imprint expansion are input
preparation and untimed. Reported per rank: host planning, records<->planes conversion, kernel time (CUDA events) and
barrier waits, so that an unmodelled phase shows up. Checks: SHA-256 of the assembled FP32 reflectance image (must not
depend on N) and parity with the CPU reference on a bounded spatial sub-sample (strokes whose whole region lies in one
2048 x 2048 window, rendered alone by the CPU and by the GPU on the 16K canvas).
"""
import argparse
import hashlib
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from painty_b200 import api, assets  # noqa: E402
from painty_b200.dist import DistCanvas  # noqa: E402


def make_strokes(n_strokes, rows, cols, seed=1234):
    """Vectorised version of tests/workloads.sbr_strokes for large counts: 4 passes (brush sizes 80/60/30/20 image px at a
    3840-wide canvas), radius ~ U[0.35, 0.5] * size * 3.75 snapped to an OOB-free radius, 5..20 control points spaced
    0.25 * radius with a smooth heading random walk, 5 palette mixes, strokes of a pass grouped by colour index.
    Returns (radius[n], K[n,3], S[n,3], first_vertex[n], n_vertices[n], path_xy[sum,2])."""
    rng = np.random.default_rng(seed)
    pk, ps = assets.palette("lindemeier_measured")
    mixes = []
    for _ in range(5):
        w = rng.dirichlet(np.ones(len(pk)))
        mixes.append(((w[:, None] * pk).sum(0), (w[:, None] * ps).sum(0)))
    mixK, mixS = np.array([m[0] for m in mixes]), np.array([m[1] for m in mixes])
    safe = {r: assets.snap_to_safe_radius(float(r)) for r in range(1, 400)}
    per = [n_strokes // 4] * 4
    per[0] += n_strokes - sum(per)
    R, Kc, NV, paths = [], [], [], []
    for size, cnt in zip((80, 60, 30, 20), per):
        r = np.maximum(1.0, rng.uniform(0.35, 0.5, cnt) * size * 3.75)
        r = np.array([safe[int(round(v))] for v in r])
        npts = rng.integers(5, 21, cnt)
        ci = rng.integers(0, 5, cnt)
        order = np.argsort(ci, kind="stable")  # grouped by colour index (std::map order)
        r, npts, ci = r[order], npts[order], ci[order]
        p0 = np.stack([rng.uniform(0, cols, cnt), rng.uniform(0, rows, cnt)], axis=1)
        ang0 = rng.uniform(0, 2 * np.pi, cnt)
        dang = 0.15 * rng.normal(size=(cnt, 19))
        ang = ang0[:, None] + np.cumsum(dang, axis=1)
        step = (0.25 * r)[:, None]
        dx, dy = step * np.cos(ang), step * np.sin(ang)
        xs = np.concatenate([p0[:, :1], p0[:, :1] + np.cumsum(dx, axis=1)], axis=1)
        ys = np.concatenate([p0[:, 1:], p0[:, 1:] + np.cumsum(dy, axis=1)], axis=1)
        mask = np.arange(20)[None, :] < npts[:, None]
        paths.append(np.stack([xs[mask], ys[mask]], axis=1))
        R.append(r), Kc.append(ci), NV.append(npts)
    R, Kc, NV = np.concatenate(R), np.concatenate(Kc), np.concatenate(NV).astype(np.int32)
    first = np.concatenate([[0], np.cumsum(NV)[:-1]]).astype(np.int64)
    return R, mixK[Kc], mixS[Kc], first, NV, np.concatenate(paths)


def batch_records(R, K, S, first, nv, path, a, b):
    """Stroke records + imprints of strokes [a, b)."""
    v0, v1 = int(first[a]), int(first[b - 1] + nv[b - 1])
    cx, cy, th, fi, ni = api.expand_strokes(first[a:b] - v0, nv[a:b], path[v0:v1])
    rec = np.zeros(b - a, dtype=api.STROKE_DTYPE)
    rec["radius"], rec["K"], rec["S"], rec["first_imprint"], rec["n_imprints"] = R[a:b], K[a:b], S[a:b], fi, ni
    return rec, cx, cy, th


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strokes", type=int, default=1000000)
    ap.add_argument("--batch", type=int, default=100000)
    ap.add_argument("--size", type=int, default=16384)
    ap.add_argument("--cpu-strokes", type=int, default=48)
    ap.add_argument("--out", default=None)
    ap.add_argument("--stroke-pixels", type=int, default=0,
                    help="exact stroke-pixel count of this stroke list from an earlier run (it does not depend on N); 0 = count it "
                         "in an extra untimed pass (a pass over all footprint cells of all imprints)")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rows = cols = args.size
    t0 = time.perf_counter()
    R, K, S, first, nv, path = make_strokes(args.strokes, rows, cols)
    gen_s = time.perf_counter() - t0
    radii = sorted(set(float(r) for r in R))
    ctx = api.Context(local, api.F32)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    dc = DistCanvas(ctx, rows, cols, dist)
    br = api.FootprintBrush(ctx, radii[0])
    for r in radii:
        br.register_radius(r)
    dc.attach(br)
    dc.attach_image(root=0)
    edges = list(range(0, args.strokes, args.batch)) + [args.strokes]
    t0 = time.perf_counter()
    batches = [batch_records(R, K, S, first, nv, path, a, b) for a, b in zip(edges[:-1], edges[1:])]  # input preparation
    expand_s = time.perf_counter() - t0
    n_imprints = int(sum(len(b[1]) for b in batches))

    phases = dict(plan_wait_s=0.0, plan_s=0.0, convert_s=0.0, kernel_s=0.0, barrier_wait_s=0.0, compose_gather_s=0.0)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    pool = ThreadPoolExecutor(1)

    def plan(i):
        t = time.perf_counter()
        p = dc.plan(br, *batches[i])
        return p, time.perf_counter() - t

    def run_all(count_visited=False):
        """All batches + compose/gather. Returns wall seconds of this rank (max over ranks taken by the caller)."""
        dc.canvas.clear()
        br.updateSnapshot(dc.canvas)
        br.enable_visited_count(count_visited)
        ctx.synchronize()
        dist.barrier()
        for k in phases:
            phases[k] = 0.0
        kernel_events = []
        t_start = time.perf_counter()
        nxt = pool.submit(plan, 0)
        lib = api.lib()
        d = dc._desc[id(br)]
        for i in range(len(batches)):
            t = time.perf_counter()
            p, plan_s = nxt.result()
            phases["plan_wait_s"] += time.perf_counter() - t  # planning that was NOT hidden behind the previous batch
            phases["plan_s"] += plan_s
            t = time.perf_counter()
            api._chk(lib.pb_fbrush_dist_begin(br.h, dc.canvas.h))
            phases["convert_s"] += time.perf_counter() - t
            t = time.perf_counter()
            dist.barrier()
            phases["barrier_wait_s"] += time.perf_counter() - t
            e0, e1 = ev(), ev()
            e0.record(stream)
            br.run_batch_plan(dc.canvas, p, dist_desc=d)
            e1.record(stream)
            kernel_events.append((e0, e1))
            if i + 1 < len(batches):
                nxt = pool.submit(plan, i + 1)  # the brush state of batch i is set: plan the next batch now
            ctx.synchronize()
            t = time.perf_counter()
            dist.barrier()
            phases["barrier_wait_s"] += time.perf_counter() - t
            t = time.perf_counter()
            api._chk(lib.pb_fbrush_dist_end(br.h, dc.canvas.h))
            phases["convert_s"] += time.perf_counter() - t
        t = time.perf_counter()
        dc.compose_gather()
        dc.finish_gather()
        phases["compose_gather_s"] = time.perf_counter() - t
        wall = time.perf_counter() - t_start
        phases["kernel_s"] = sum(a.elapsed_time(b) for a, b in kernel_events) * 1e-3
        return wall

    # pass 1 (untimed): exact stroke-pixel count; pass 2: timed
    visited = args.stroke_pixels
    if visited <= 0:
        v0 = br.counters()[0]
        run_all(count_visited=True)
        tv = torch.tensor([br.counters()[0] - v0], dtype=torch.int64, device="cuda")
        if world > 1:
            dist.all_reduce(tv)
        visited = int(tv.item())
    wall = run_all(count_visited=False)
    tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    seconds = float(tw.item())
    everyone = [None] * world
    dist.all_gather_object(everyone, {k: round(v, 3) for k, v in phases.items()})

    record = None
    if rank == 0:
        # read the image back through the C ABI (host AoS f64) in one piece: 16K^2 x 3 x 8 B = 6.4 GB
        host = dc.download_image()
        r32 = host.astype(np.float32)
        sha = hashlib.sha256(r32.tobytes()).hexdigest()
        checksum = float(host.sum())
        record = dict(config="config 4: %dx%d, %d footprint strokes in batches of %d (%d imprints), strong scaling" % (
            cols, rows, args.strokes, args.batch, n_imprints), n_gpus=world, seconds=seconds, stroke_pixels=visited,
            stroke_pixels_per_s=visited / seconds, reflectance_sha256_f32=sha, reflectance_sum=checksum,
            phases_per_rank=everyone, input_generation_s=round(gen_s, 2), input_expansion_s=round(expand_s, 2))
        del host, r32

    # CPU parity on a spatial sub-sample: strokes whose whole region lies inside one 2048 x 2048 window, rendered alone
    if args.cpu_strokes > 0:
        win = 2048
        wx0, wy0 = (cols - win) // 2, (rows // max(world, 1)) - win // 2 if world > 1 else (rows - win) // 2  # straddles the first band boundary
        wy0 = max(0, min(rows - win, wy0))
        m = R * 2.5 + 4.0  # half footprint side (1.42 r) + snapshot ring (r) + margin
        x_lo, x_hi = np.minimum.reduceat(path[:, 0], first), np.maximum.reduceat(path[:, 0], first)
        y_lo, y_hi = np.minimum.reduceat(path[:, 1], first), np.maximum.reduceat(path[:, 1], first)
        inside = (x_lo - m >= wx0) & (x_hi + m < wx0 + win) & (y_lo - m >= wy0) & (y_hi + m < wy0 + win)
        sel = [int(v) for v in np.nonzero(inside)[0]]
        # a bounded, size-balanced sample: every k-th contained stroke up to the budget
        sel = sel[::max(1, len(sel) // args.cpu_strokes)][:args.cpu_strokes]
        sub = np.array(sel, dtype=np.int64)
        v_first = np.concatenate([[0], np.cumsum(nv[sub])[:-1]]).astype(np.int64)
        sub_path = np.concatenate([path[first[s]:first[s] + nv[s]] for s in sel])
        cx, cy, th, fi, ni = api.expand_strokes(v_first, nv[sub], sub_path)
        rec = np.zeros(len(sel), dtype=api.STROKE_DTYPE)
        rec["radius"], rec["K"], rec["S"], rec["first_imprint"], rec["n_imprints"] = R[sub], K[sub], S[sub], fi, ni
        dc.canvas.clear()
        br.updateSnapshot(dc.canvas)
        dc.stroke_batch(br, rec, cx, cy, th)
        dc.compose_gather()
        dc.finish_gather()
        if rank == 0:
            from oracle import cpu as ocpu  # the checker (CPU reference), rank 0 only

            ocpu.build()
            cpu = ocpu.Cpu("ref" if ocpu.have_ref() else "port")
            cvo = cpu.canvas(win, win)
            bro = cpu.footprint_brush(float(R[sub[0]]))
            t_cpu = 0.0
            for j, s in enumerate(sel):
                bro.dip(K[s], S[s])
                bro.set_radius(float(R[s]))
                a, m = int(fi[j]), int(ni[j])
                t_cpu += bro.imprint_batch(cvo, cx[a:a + m] - wx0, cy[a:a + m] - wy0, th[a:a + m])
            want = cvo.compose()
            got = dc.download_image()[wy0:wy0 + win, wx0:wx0 + win]
            record["cpu_parity"] = dict(window=[wx0, wy0, win, win], strokes=len(sel), imprints=int(len(cx)), cpu_kind=cpu.kind,
                                        cpu_seconds=round(t_cpu, 2), max_abs_err=float(np.abs(got - want).max()), tolerance=1e-4,
                                        painted_px=int((want != 1.0).any(axis=2).sum()),
                                        what="strokes whose region lies inside the window (which straddles a band boundary at N > 1), "
                                             "rendered alone by the CPU reference on a %dx%d canvas and by the GPUs on the 16K canvas" % (win, win))
    if rank == 0:
        print(json.dumps(record), flush=True)
        if args.out:
            prev = json.load(open(args.out)) if os.path.exists(args.out) else []
            prev.append(record)
            json.dump(prev, open(args.out, "w"), indent=1)
    pool.shutdown()
    dc.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
