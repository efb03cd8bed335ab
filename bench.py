#!/usr/bin/env python
"""bench.py — stroke-pixels/s (imprint + Kubelka-Munk compose) at a 4K canvas, with the KM compose
kernel's HBM roofline fraction and the reference CPU path timed beside it.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--strokes S]

Workload (BASELINE.json configs[1]): 3840x2160 canvas, S (default 10000)
synthetic sbr_painter-shaped footprint-brush strokes (4 brush-size passes, radii snapped to OOB-free
values, 5..20 control points, 5 palette mixes of the 14-pigment measured palette thinned with medium),
each stroke = dip -> setRadius -> paintStroke, then Renderer::compose of the whole canvas.
A step = clear canvas -> all strokes -> compose. stroke-pixel = a footprint cell passing both bounds checks of
FootprintBrush::imprint (the reference's own `counter`, FootprintBrush.hxx:119), counted exactly on the device
in an untimed pass.

  value    : stroke-pixels/s, inputs (stroke list, footprints) already resident on the device side of the
             C ABI call; timed with CUDA events on the context's stream (imprint kernel + compose kernel).
  e2e      : same metric through the C ABI with HOST buffers: stroke list H2D + kernels + reflectance D2H
             (AoS f64 like Renderer::compose returns) inside the timed region.
  roofline : the KM compose kernel (the path's HBM-bound kernel): 52 B/px algorithmic / event time.
N > 1: one process per GPU (torchrun), weak scaling of ONE canvas: the canvas grows to (N x 2160) x 3840 with N x S
strokes of the same size distribution and is cut into N row bands, one per GPU. A stroke is executed by the GPU
whose band holds its first imprint; where it leaves the band the kernel reads / writes the neighbour's HBM through
NVLink peer mappings and strokes wait on completion flags of conflicting earlier strokes on any GPU
(painty_b200/dist.py, bit-exact vs one GPU: tests/test_dist_gpu.py). The reflectance bands are assembled with one NCCL
all_gather inside the timed step.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROWS, COLS = 2160, 3840
METRIC = "stroke-pixels/sec (imprint+KM compose) at 4K canvas; % of HBM roofline"
COMPOSE_BYTES_PER_PX = 52  # 7 layer planes + 3 R0 read, 3 R written, FP32


def build_workload(n_strokes, rows=ROWS, cols=COLS, seed=1234):
    from painty_b200 import api, assets
    from tests.workloads import sbr_strokes

    pk, ps = assets.palette("lindemeier_measured")
    tk, ts = assets.palette("thinning_medium")
    strokes = sbr_strokes(rows, cols, n_strokes, seed=seed, safe_radius=assets.snap_to_safe_radius, palette=(pk, ps))
    rec = np.zeros(len(strokes), dtype=api.STROKE_DTYPE)
    xs, ys, ts_ = [], [], []
    first = 0
    for i, s in enumerate(strokes):
        K, S = api.mixed(s["K"], s["S"], 1.0, tk[0], ts[0], 0.0)  # mixed(p,1,thinner,0) like the sbr painter
        cx, cy, th = api.expand_stroke(s["path"], mode=0)
        rec[i] = (s["radius"], K, S, first, len(cx))
        first += len(cx)
        xs.append(cx), ys.append(cy), ts_.append(th)
    radii = sorted(set(float(s["radius"]) for s in strokes))
    return rec, np.concatenate(xs), np.concatenate(ys), np.concatenate(ts_), radii


class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_sample(rec, cx, cy, th, rows, cols, budget_cells=1.5e8, threads=None):
    """Time the CPU reference path (oracle/_ref when present, else the oracle port) on a bounded sample of the
    same workload: strokes taken round-robin over the brush-size passes until ~budget visited cells, rendered
    in submission order on a fresh canvas, plus one threaded ComputeReflectance pass over a canvas slice."""
    from oracle import cpu as ocpu
    from painty_b200 import assets

    ocpu.build()
    kind = "reference" if ocpu.have_ref() else "port"
    c = ocpu.Cpu("ref" if kind == "reference" else "port")
    threads = threads or os.cpu_count() or 1
    n = len(rec)
    order = []
    per_pass = max(1, n // 4)
    k = 0
    est = 0.0
    while est < budget_cells and k < per_pass:
        for p in range(4):
            i = min(p * per_pass + k, n - 1)
            side = assets.footprint_geometry(float(rec["radius"][i]))[3]
            est += float(rec["n_imprints"][i]) * side * side * 0.8
            order.append(i)
            if est >= budget_cells:
                break
        k += 1
    order = sorted(set(order))
    cv = c.canvas(rows, cols)
    br = c.footprint_brush(float(rec["radius"][order[0]]))
    counter = ocpu.Cpu("port")  # visited-cell count comes from the port's counters (the reference discards its own)
    t_imp = 0.0
    for i in order:
        a, m = int(rec["first_imprint"][i]), int(rec["n_imprints"][i])
        br.dip(rec["K"][i], rec["S"][i])
        br.set_radius(float(rec["radius"][i]))
        t_imp += br.imprint_batch(cv, cx[a:a + m], cy[a:a + m], th[a:a + m])
    # visited cells of the sample, exact, from the device-independent port on a tiny canvas is not
    # possible (bounds depend on the canvas) -> count analytically with the port on the same canvas
    cvp = counter.canvas(rows, cols) if kind == "reference" else None
    if kind == "reference":
        brp = counter.footprint_brush(float(rec["radius"][order[0]]))
        for i in order:
            a, m = int(rec["first_imprint"][i]), int(rec["n_imprints"][i])
            brp.dip(rec["K"][i], rec["S"][i])
            brp.set_radius(float(rec["radius"][i]))
            brp.imprint_batch(cvp, cx[a:a + m], cy[a:a + m], th[a:a + m])
        visited = brp.counters()[0]
    else:
        visited = br.counters()[0]
    # compose: threaded row-split of the reference's per-pixel function over a 1/8 slice of the canvas
    st = cv.get()
    sl = slice(0, max(1, rows // 8))
    t_cmp_slice, _ = c.compose_timed(st["K"][sl], st["S"][sl], st["V"][sl], st["R0"][sl], threads=threads)
    t_cmp_full = t_cmp_slice * rows / (sl.stop - sl.start)
    return dict(kind=kind, visited=int(visited), t_imprint=t_imp, t_compose_full=t_cmp_full, n_sample=len(order),
                threads=threads)


def run_reference(args, rank, world):
    if rank != 0:
        return
    rec, cx, cy, th, _ = build_workload(args.strokes)
    vals = []
    info = None
    budget = 1.2e8
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        info = cpu_sample(rec, cx, cy, th, ROWS, COLS, budget_cells=budget)
        # compose prorated to the sample's share of the canvas work is negligible next to imprint; we charge
        # the full-canvas compose scaled by sample/total strokes
        t = info["t_imprint"] + info["t_compose_full"] * info["n_sample"] / len(rec)
        if it >= args.warmup:
            vals.append(info["visited"] / t)
        if time.time() - t0 > 40:
            budget *= 0.5
    v = float(np.mean(vals))
    line = {"metric": METRIC, "value": v, "unit": "stroke-pixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": "sbr-style 3840x2160, %d footprint strokes + KM compose (bounded CPU sample)" % args.strokes},
            "cpu_baseline": {"value": v, "unit": "stroke-pixels/s", "cores": 1, "kind": info["kind"],
                             "sample": "%d of %d strokes (round-robin over the 4 brush-size passes, %d visited cells), imprint single-threaded "
                                       "as in the reference; compose prorated from a %d-thread row-split" % (
                                           info["n_sample"], len(rec), info["visited"], info["threads"])},
            "e2e": {"value": v, "unit": "stroke-pixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--strokes", type=int, default=10000)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    from painty_b200 import api, build

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: painty_b200 has no CPU fallback")
    build.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    rows_total = ROWS * world
    rec, cx, cy, th, radii = build_workload(args.strokes * world, rows=rows_total)  # same list on every rank
    ctx = api.Context(local, api.F32)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local)
    dc = None
    if world > 1:
        from painty_b200.dist import DistCanvas

        dc = DistCanvas(ctx, rows_total, COLS, dist)
        cv = dc.canvas
    else:
        cv = api.Canvas(ctx, ROWS, COLS)
    br = api.FootprintBrush(ctx, radii[0])
    for r in radii:
        br.register_radius(r)
    if dc is not None:
        dc.attach(br)
    n_px = cv.store_rows * COLS
    d_R = torch.empty((3, n_px), dtype=torch.float32, device="cuda")
    h_R = torch.empty((cv.store_rows, COLS, 3), dtype=torch.float64).pin_memory()
    h_R_np = h_R.numpy()
    gathered = torch.empty((world, 3, n_px), dtype=torch.float32, device="cuda") if world > 1 else None

    def strokes_all():
        if dc is not None:
            dc.stroke_batch(br, rec, cx, cy, th)
        else:
            br.stroke_batch(cv, rec, cx, cy, th)

    # untimed: exact stroke-pixel count of the workload (reference's `counter`)
    br.enable_visited_count(True)
    cv.clear()
    br.updateSnapshot(cv)
    strokes_all()
    ctx.synchronize()
    visited, active = br.counters()
    br.enable_visited_count(False)

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step_device(timers=None):
        cv.clear()
        br.updateSnapshot(cv)  # FootprintBrush::updateSnapshot(canvas): brush state == a freshly constructed brush
        if timers is not None:
            timers[0].record(stream)
        strokes_all()
        if timers is not None:
            timers[1].record(stream)
        cv.compose_device(d_R.data_ptr(), n_px)
        if timers is not None:
            timers[2].record(stream)
        if world > 1:
            # NCCL work is ordered after the compose on the context's stream, and the stream waits for it
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(gathered.view(-1), d_R.view(-1))

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = ctx.launches
    t_imp = t_cmp = 0.0
    timers = [[ev(), ev(), ev()] for _ in range(args.steps)]
    e0, e1 = ev(), ev()
    barrier()
    e0.record(stream)
    for k in range(args.steps):
        step_device(timers[k])
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = ctx.launches - launches0
    for t in timers:
        t_imp += t[0].elapsed_time(t[1])
        t_cmp += t[1].elapsed_time(t[2])
    ms_step = total_ms / args.steps
    if world > 1:
        tt = torch.tensor([ms_step], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_step = float(tt.item())
    visited_all = visited
    if world > 1:
        tv = torch.tensor([visited], dtype=torch.int64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.SUM)
        visited_all = int(tv.item())
    value = visited_all / (ms_step * 1e-3)

    # e2e: C ABI with host buffers (stroke list H2D, kernels, reflectance D2H as AoS f64)
    def step_e2e():
        cv.clear()
        br.updateSnapshot(cv)
        strokes_all()
        cv.compose(h_R_np)

    # one warm-up pass (first use of the host-buffer path allocates staging memory), one timed pass
    e2e_steps = 1
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        tt = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt.item())
    h2d = rec.nbytes + cx.nbytes * 4  # stroke records + (cx, cy, cos, sin) per imprint
    d2h = n_px * 3 * 8

    # roofline of the KM compose kernel, measured live with events around each launch
    cmp_ms = t_cmp / args.steps
    roofline_how = "events around the compose launch of every timed step"
    if world > 1:
        # at N > 1 the in-step launch follows a host-side process-group barrier (idle GPU + launch latency inside the
        # event pair); the kernel itself is timed with back-to-back launches right after the timed steps
        ce0, ce1 = ev(), ev()
        cv.compose_device(d_R.data_ptr(), n_px)
        ce0.record(stream)
        for _ in range(5):
            cv.compose_device(d_R.data_ptr(), n_px)
        ce1.record(stream)
        ctx.synchronize()
        cmp_ms = ce0.elapsed_time(ce1) / 5
        roofline_how = "5 back-to-back launches after the timed steps (the in-step launch follows a host barrier)"
    achieved = COMPOSE_BYTES_PER_PX * n_px / (cmp_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback"
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured"
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "stroke-pixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "sbr-style 3840x%d, %d footprint strokes (%d imprints) + KM compose" % (rows_total, len(rec), len(cx)),
                   "parallelism": "single GPU" if world == 1 else
                   "one %dx%d canvas in %d row bands (one per GPU), strokes cross bands through NVLink peer memory, NCCL all_gather of "
                   "reflectance" % (rows_total, COLS, world),
                   "stroke_pixels_per_step": int(visited_all), "stroke_pixels_per_step_this_rank": int(visited),
                   "active_stroke_pixels_per_step_this_rank": int(active),
                   "l2": "canvas working set 8.3 Mpx x 14 planes x 4 B = 464 MB > 126 MB L2; canvas cleared every step",
                   "imprint_ms": t_imp / args.steps, "compose_ms": cmp_ms},
        "clocks": clocks,
        "e2e": {"value": visited_all / (e2e_ms * 1e-3), "unit": "stroke-pixels/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "km_compose_kernel<float>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "measured": roofline_how,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one 4K launch, ncu --set full (profiles/r01_compose_f32_raw.csv)
                     "traffic": 401643264 if world == 1 else None, "algorithmic_bytes_per_launch": COMPOSE_BYTES_PER_PX * n_px,
                     "frac_of_8TBs_nominal": achieved / 8000.0},
    }
    # the step's dominant kernel is the imprint chain: latency bound (dependent imprints, L2-resident working set), so no
    # roofline claim — reported with the upper-bound byte model of SURVEY.md §8d (88 B per active stroke-pixel)
    imp_ms = t_imp / args.steps
    line["imprint"] = {"kernel": "imprint_kernel<float>", "bound": "latency (dependency chain of imprints, see DESIGN.md §5)",
                       "ms_per_step": imp_ms, "ms_each_step": [round(t[0].elapsed_time(t[1]), 1) for t in timers],
                       "imprints_per_s": world * len(cx) / max(world, 1) / (imp_ms * 1e-3),
                       "active_stroke_pixels_per_s": active / (imp_ms * 1e-3),
                       "model_bytes_per_active_px": 88, "model_GBps": 88 * active / (imp_ms * 1e-3) / 1e9}
    try:  # host-side share of a batch: dataflow planning, per-imprint constants, the planner's model of the step
        line["host"] = br.batch_stats()
    except Exception as exc:  # diagnostics only
        line["host"] = {"unavailable": str(exc)}
    if rank == 0 and not args.no_cpu and world == 1:
        info = cpu_sample(rec, cx, cy, th, ROWS, COLS)
        t = info["t_imprint"] + info["t_compose_full"] * info["n_sample"] / len(rec)
        line["cpu_baseline"] = {"value": info["visited"] / t, "unit": "stroke-pixels/s", "cores": 1, "kind": info["kind"],
                                "sample": "%d of %d strokes (round-robin over the 4 brush-size passes, %d visited cells, %.1f s imprint "
                                          "single-threaded as in the reference); full-canvas compose %.2f s on %d threads, prorated" % (
                                              info["n_sample"], len(rec), info["visited"], info["t_imprint"], info["t_compose_full"],
                                              info["threads"])}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
