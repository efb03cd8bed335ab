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
N > 1: one process per GPU (torchrun), weak scaling of ONE canvas: the canvas grows to (N x 2160) x 3840 and is cut into N
row bands, one per GPU; the stroke list is the single-GPU list repeated per band (copy k shifted down by k x 2160 rows), so
every GPU receives exactly the single-GPU work and the copies interact across the band boundaries. A stroke is executed by the GPU
whose band holds its first imprint; where it leaves the band the kernel reads / writes the neighbour's HBM through
NVLink peer mappings and strokes wait on completion flags of conflicting earlier strokes on any GPU
(painty_b200/dist.py, bit-exact vs one GPU: tests/test_dist_gpu.py). The reflectance image is assembled inside the timed
step by the compose kernel itself: every rank stores its band's rows straight into every rank's image over NVLink
(pb_canvas_compose_gather) — no separate collective.
"""
import argparse
import json
import math
import os

os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")  # persistent kernels on several streams / GPUs: no load-time syncs
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


def build_strokes(n_strokes, rows=ROWS, cols=COLS, seed=1234):
    """The synthetic stroke list (numpy only — the reference arm must not load the product library): list of
    dict(radius, K, S, path[n,2]); paint = palette mix thinned like the sbr painter's mixed(p, 1, thinner, 0)
    (PaintMixer.cxx:539-545: ((1*K1) + (0*K2)) * (1 / (1 + 0)))."""
    from painty_b200 import assets
    from tests.workloads import sbr_strokes

    pk, ps = assets.palette("lindemeier_measured")
    tk, ts = assets.palette("thinning_medium")
    strokes = sbr_strokes(rows, cols, n_strokes, seed=seed, safe_radius=assets.snap_to_safe_radius, palette=(pk, ps))
    inv = 1.0 / (1.0 + 0.0)
    for s in strokes:
        s["K"] = ((1.0 * s["K"]) + (0.0 * tk[0])) * inv
        s["S"] = ((1.0 * s["S"]) + (0.0 * ts[0])) * inv
    return strokes


def imprint_counts(strokes):
    """Imprints per stroke of FootprintBrush::paintStroke (:251-267): sum over segments of int(|p1 - p0|)."""
    out = np.zeros(len(strokes), dtype=np.int64)
    for i, s in enumerate(strokes):
        d = np.diff(s["path"], axis=0)
        out[i] = int(np.sum(np.floor(np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]))))
    return out


def build_workload(n_strokes, rows=ROWS, cols=COLS, seed=1234, tiles=1):
    """Stroke records + imprint arrays for the product arm (expansion = pb_expand_stroke, host f64).
    tiles > 1 (weak scaling over GPUs): the (rows x cols) stroke list is repeated `tiles` times, tile k shifted down by
    k * rows — every band of the taller canvas receives exactly the single-GPU workload, and the copies meet at the band
    boundaries (strokes that leave the single-GPU canvas at its top / bottom border reach into the neighbouring copy).
    Submission order: original order, the tiles of one stroke next to each other."""
    from painty_b200 import api

    base = build_strokes(n_strokes, rows, cols, seed)
    if tiles > 1:
        strokes = []
        for s in base:
            for k in range(tiles):
                t = dict(s)
                t["path"] = s["path"] + np.array([0.0, float(k * rows)])
                strokes.append(t)
    else:
        strokes = base
    rec = np.zeros(len(strokes), dtype=api.STROKE_DTYPE)
    xs, ys, ts_ = [], [], []
    first = 0
    for i, s in enumerate(strokes):
        cx, cy, th = api.expand_stroke(s["path"], mode=0)
        rec[i] = (s["radius"], s["K"], s["S"], first, len(cx))
        first += len(cx)
        xs.append(cx), ys.append(cy), ts_.append(th)
    radii = sorted(set(float(s["radius"]) for s in strokes))
    return strokes, rec, np.concatenate(xs), np.concatenate(ys), np.concatenate(ts_), radii


def expand_with_oracle(cpu, path):
    """FootprintBrush::paintStroke's expansion (:251-267) evaluated with the CPU checker's Catmull-Rom — the reference
    arm's own stroke -> imprint step (no product code)."""
    path = np.asarray(path, dtype=np.float64)
    cx, cy, th = [], [], []
    n = len(path)
    for i in range(n - 1):
        p_pre, p0, p1, p_next = path[max(i - 1, 0)], path[i], path[i + 1], path[min(i + 2, n - 1)]
        dist = math.sqrt((p1[0] - p0[0]) * (p1[0] - p0[0]) + (p1[1] - p0[1]) * (p1[1] - p0[1]))
        for pd in range(1, int(dist) + 1):
            t = pd / dist
            d = cpu.catmull_rom(p_pre, p0, p1, p_next, t, True)
            q = cpu.catmull_rom(p_pre, p0, p1, p_next, t)
            cx.append(q[0]), cy.append(q[1]), th.append(math.atan2(d[1], d[0]))
    return np.array(cx), np.array(cy), np.array(th)


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


def cpu_sample(strokes, rows, cols, budget_cells=1.5e8, threads=None, want_image=False):
    """Time the CPU reference path (oracle/_ref when present, else the oracle port) on a bounded sample of the
    same workload: strokes taken round-robin over the brush-size passes until ~budget visited cells, expanded with the
    checker's own Catmull-Rom and rendered in submission order on a fresh canvas, plus one threaded ComputeReflectance
    pass over a canvas slice. No product code runs here."""
    from oracle import cpu as ocpu
    from painty_b200 import assets

    ocpu.build()
    kind = "reference" if ocpu.have_ref() else "port"
    c = ocpu.Cpu("ref" if kind == "reference" else "port")
    threads = threads or os.cpu_count() or 1
    n = len(strokes)
    counts = imprint_counts(strokes)
    order = []
    per_pass = max(1, n // 4)
    k = 0
    est = 0.0
    while est < budget_cells and k < per_pass:
        for p in range(4):
            i = min(p * per_pass + k, n - 1)
            side = assets.footprint_geometry(float(strokes[i]["radius"]))[3]
            est += float(counts[i]) * side * side * 0.8
            order.append(i)
            if est >= budget_cells:
                break
        k += 1
    order = sorted(set(order))
    expanded = {i: expand_with_oracle(c, strokes[i]["path"]) for i in order}
    cv = c.canvas(rows, cols)
    br = c.footprint_brush(float(strokes[order[0]]["radius"]))
    t_imp = 0.0
    for i in order:
        s = strokes[i]
        br.dip(s["K"], s["S"])
        br.set_radius(float(s["radius"]))
        t_imp += br.imprint_batch(cv, *expanded[i])
    # visited cells of the sample (the reference discards its own `counter`): replay on the port, which counts them
    if kind == "reference":
        counter = ocpu.Cpu("port")
        cvp = counter.canvas(rows, cols)
        brp = counter.footprint_brush(float(strokes[order[0]]["radius"]))
        for i in order:
            s = strokes[i]
            brp.dip(s["K"], s["S"])
            brp.set_radius(float(s["radius"]))
            brp.imprint_batch(cvp, *expanded[i])
        visited = brp.counters()[0]
    else:
        visited = br.counters()[0]
    # compose: threaded row-split of the reference's per-pixel function over a 1/8 slice of the canvas
    st = cv.get()
    sl = slice(0, max(1, rows // 8))
    t_cmp_slice, _ = c.compose_timed(st["K"][sl], st["S"][sl], st["V"][sl], st["R0"][sl], threads=threads)
    t_cmp_full = t_cmp_slice * rows / (sl.stop - sl.start)
    out = dict(kind=kind, visited=int(visited), t_imprint=t_imp, t_compose_full=t_cmp_full, n_sample=len(order), threads=threads,
               order=order, expanded=expanded)
    if want_image:  # for the parity field: the reference's reflectance image and wet-pixel set of the sample
        out["R"] = cv.compose()
        out["wet"] = st["V"] > 0
    return out


def run_reference(args, rank, world):
    if rank != 0:
        return
    strokes = build_strokes(args.strokes)
    vals, times = [], []
    info = None
    budget = 1.2e8
    for it in range(args.warmup + args.steps):
        t0 = time.time()
        info = cpu_sample(strokes, ROWS, COLS, budget_cells=budget)
        # compose prorated to the sample's share of the canvas work is negligible next to imprint; we charge
        # the full-canvas compose scaled by sample/total strokes
        t = info["t_imprint"] + info["t_compose_full"] * info["n_sample"] / len(strokes)
        if it >= args.warmup:
            vals.append(info["visited"] / t)
            times.append(t * 1e3)
        if time.time() - t0 > 40:
            budget *= 0.5
    v = float(np.mean(vals))
    line = {"metric": METRIC, "value": v, "unit": "stroke-pixels/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            # a step of this arm is ONE bounded sample of the workload (see cpu_baseline.sample), not the whole stroke list
            "ms_per_step": float(np.mean(times)), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "sbr-style 3840x2160, %d footprint strokes + KM compose (bounded CPU sample)" % args.strokes},
            "cpu_baseline": {"value": v, "unit": "stroke-pixels/s", "cores": 1, "kind": info["kind"],
                             "sample": "%d of %d strokes (round-robin over the 4 brush-size passes, %d visited cells), stroke expansion and "
                                       "imprint single-threaded as in the reference; compose prorated from a %d-thread row-split" % (
                                           info["n_sample"], len(strokes), info["visited"], info["threads"])},
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
    strokes, rec, cx, cy, th, radii = build_workload(args.strokes, rows=ROWS, tiles=world)  # same list on every rank
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
    # host result of the e2e leg: the whole reflectance image (AoS f64 like Renderer::compose returns); at N > 1 it is
    # assembled in rank 0's HBM by the compose kernels' peer stores and read back by rank 0 alone
    h_rows = rows_total if rank == 0 else 1
    h_R = torch.empty((h_rows, COLS, 3), dtype=torch.float64).pin_memory()
    h_R_np = h_R.numpy()
    if dc is not None:
        dc.attach_image(root=None)  # every rank receives the assembled image (all-gather semantics)

    # N > 1: the host half of a batch (dataflow planning of the global stroke list, 0.1-0.4 s) runs on a helper thread while
    # the GPUs execute the previous batch — every step still plans its own batch, one step ahead (pb_fbrush_plan_stroke_batch)
    from concurrent.futures import ThreadPoolExecutor

    planner = ThreadPoolExecutor(1) if dc is not None else None
    pending = [None]

    def plan_next():
        pending[0] = planner.submit(dc.plan, br, rec, cx, cy, th)

    def strokes_all(r=None, x=None, y=None, t=None):
        if dc is not None and r is None:
            if pending[0] is None:
                plan_next()
            plan = pending[0].result()
            dc.stroke_batch(br, plan=plan, after_launch=plan_next)
        elif dc is not None:
            dc.stroke_batch(br, r, x, y, t)
        else:
            br.stroke_batch(cv, rec if r is None else r, cx if x is None else x, cy if y is None else y, th if t is None else t)

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
        if world > 1:
            dc.compose_gather()  # compose + band gather in one kernel: rows go straight into every rank's image over NVLink
        else:
            cv.compose_device(d_R.data_ptr(), n_px)
        if timers is not None:
            timers[2].record(stream)

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
        if world > 1:
            dc.compose_gather()
            dc.finish_gather()  # context sync + process-group barrier: every band has arrived
            if rank == 0:
                dc.download_image(h_R_np)
        else:
            cv.compose(h_R_np)

    # one warm-up pass (first use of the host-buffer path allocates staging memory), then timed passes
    e2e_steps = max(3, args.steps)
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
    h2d = rec.nbytes + cx.nbytes * 3  # stroke records + (cx, cy, theta) per imprint (the library derives the rest)
    d2h = n_px * world * 3 * 8  # the assembled image, read back by rank 0

    # Roofline of the KM compose kernel, measured live. "in_step": events around the compose launch of every timed step
    # (behind the imprint kernel, cold L2; at N > 1 the launch also follows a host-side process-group barrier, so the
    # event pair includes an idle gap). "back_to_back": 5 launches in a row right after the timed steps. The line's
    # roofline uses the in-step time at N = 1 and the back-to-back time at N > 1 and says which.
    cmp_in_step_ms = t_cmp / args.steps
    ce0, ce1 = ev(), ev()
    cv.compose_device(d_R.data_ptr(), n_px)
    ce0.record(stream)
    for _ in range(5):
        cv.compose_device(d_R.data_ptr(), n_px)
    ce1.record(stream)
    ctx.synchronize()
    cmp_b2b_ms = ce0.elapsed_time(ce1) / 5
    cmp_ms = cmp_in_step_ms if world == 1 else cmp_b2b_ms
    roofline_how = "events around the compose launch of every timed step" if world == 1 else \
        "5 back-to-back launches of the plain compose kernel after the timed steps (the in-step launch is the compose + gather " \
        "kernel, NVLink bound: see gather)"
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
        "config": {"workload": "sbr-style 3840x%d, %d footprint strokes (%d imprints) + KM compose%s" % (
            rows_total, len(rec), len(cx), "" if world == 1 else "; the 3840x2160 / %d-stroke list repeated in each of the %d bands" % (args.strokes, world)),
                   "parallelism": "single GPU" if world == 1 else
                   "one %dx%d canvas in %d row bands (one per GPU), strokes cross bands through NVLink peer memory, reflectance image "
                   "assembled on every rank by the compose kernel's peer stores; host planning of a batch overlaps the previous batch's "
                   "execution (plan thread)" % (rows_total, COLS, world),
                   "stroke_pixels_per_step": int(visited_all), "stroke_pixels_per_step_this_rank": int(visited),
                   "active_stroke_pixels_per_step_this_rank": int(active),
                   "l2": "canvas working set 8.3 Mpx x (2 x 32 B records + 1 B) = 539 MB > 126 MB L2; canvas cleared every step",
                   "imprint_ms": t_imp / args.steps, "compose_ms": cmp_ms},
        "clocks": clocks,
        "e2e": {"value": visited_all / (e2e_ms * 1e-3), "unit": "stroke-pixels/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "passes": e2e_steps},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "km_compose_kernel<float>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "measured": roofline_how,
                     "compose_ms_in_step": cmp_in_step_ms, "compose_ms_back_to_back": cmp_b2b_ms,
                     # NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one 4K launch from the committed
                     # ncu --set full capture (profiles/r01_compose_f32_raw.csv); null where the launch has another size
                     "traffic": 404830464 if world == 1 else None, "traffic_source": "ncu constant (profiles/r02_compose_f32_raw.csv)",
                     "algorithmic_bytes_per_launch": COMPOSE_BYTES_PER_PX * n_px, "frac_of_8TBs_nominal": achieved / 8000.0},
    }
    # the step's dominant kernel is the imprint chain: a dependency chain of imprints whose working set lives in L2, bound by
    # the SM's L1 wavefront rate and by barrier latency, not by HBM — reported with its own figures, no roofline claim
    imp_ms = t_imp / args.steps
    line["imprint"] = {"kernel": "imprint_kernel<float>", "bound": "latency (dependency chain of imprints, see DESIGN.md)",
                       "ms_per_step": imp_ms, "ms_each_step": [round(t[0].elapsed_time(t[1]), 1) for t in timers],
                       "imprints_per_s_this_rank": len(cx) / max(world, 1) / (imp_ms * 1e-3),
                       "active_stroke_pixels_per_s": active / (imp_ms * 1e-3),
                       "record_bytes_per_active_px": 101, "record_GBps": 101 * active / (imp_ms * 1e-3) / 1e9}
    if world > 1:  # compose with the band-gather epilogue: (40 + 12 N) B per pixel, 12 (N - 1) of them over NVLink
        line["gather"] = {"kernel": "km_compose_gather_kernel<float>", "ms_in_step": cmp_in_step_ms,
                          "nvlink_bytes_out_per_rank": 12 * (world - 1) * n_px,
                          "nvlink_GBps_out_per_rank": 12 * (world - 1) * n_px / (cmp_in_step_ms * 1e-3) / 1e9,
                          "what": "every rank stores its band's reflectance rows into every rank's image (all-gather by peer stores)"}
    try:  # host-side share of a batch: dataflow planning, per-imprint constants, the planner's model of the step
        line["host"] = br.batch_stats()
    except Exception as exc:  # diagnostics only
        line["host"] = {"unavailable": str(exc)}

    if world > 1:
        # Proof that the band-sharded render is the single-GPU render: a bounded sub-list (every k-th stroke, ~3000 strokes
        # over all bands) is rendered once across the N GPUs and once on rank 0 alone; K/S/V must agree bit for bit.
        k = max(1, len(rec) // 3000)
        sub = np.arange(0, len(rec), k)
        srec = rec[sub].copy()
        sx, sy, st_, first = [], [], [], 0
        for j, i in enumerate(sub):
            a, m = int(rec["first_imprint"][i]), int(rec["n_imprints"][i])
            sx.append(cx[a:a + m]), sy.append(cy[a:a + m]), st_.append(th[a:a + m])
            srec["first_imprint"][j] = first
            first += m
        sx, sy, st_ = np.concatenate(sx), np.concatenate(sy), np.concatenate(st_)
        cv.clear()
        br.updateSnapshot(cv)
        strokes_all(srec, sx, sy, st_)
        ctx.synchronize()
        band = cv.download("KSV")
        mine = np.concatenate([band["K"].reshape(-1), band["S"].reshape(-1), band["V"].reshape(-1)])
        everyone = [None] * world if rank == 0 else None
        dist.gather_object(mine, everyone, dst=0)
        if rank == 0:
            import hashlib

            full = api.Canvas(ctx, rows_total, COLS)
            br1 = api.FootprintBrush(ctx, radii[0])
            for r in radii:
                br1.register_radius(r)
            br1.stroke_batch(full, srec, sx, sy, st_)
            ref = full.download("KSV")
            exact = True
            h = hashlib.sha256()
            for r_ in range(world):
                r0, r1 = r_ * ROWS, (r_ + 1) * ROWS
                want = np.concatenate([ref["K"][r0:r1].reshape(-1), ref["S"][r0:r1].reshape(-1), ref["V"][r0:r1].reshape(-1)])
                exact = exact and np.array_equal(want, everyone[r_])
                h.update(everyone[r_].tobytes())
            line["parity_vs_1gpu"] = {"bit_exact": bool(exact), "strokes": int(len(srec)), "imprints": int(len(sx)),
                                      "checksum_sha256_KSV": h.hexdigest()[:16],
                                      "what": "every %d-th stroke of the workload rendered on %d GPUs vs on one GPU (rank 0)" % (k, world)}
            del br1, full
        if rank == 0 and not args.no_cpu:  # the other ranks wait in the barrier below
            # the same CPU leg as at N = 1, on ONE band's stroke list (every band holds a copy of it): the per-GPU workload on a host core
            try:  # never leave the other ranks in the barrier because of the reporting leg
                tile_strokes = build_workload(args.strokes, rows=ROWS, tiles=1)[0]
                info = cpu_sample(tile_strokes, ROWS, COLS)
                t = info["t_imprint"] + info["t_compose_full"] * info["n_sample"] / len(tile_strokes)
                line["cpu_baseline"] = {"value": info["visited"] / t, "unit": "stroke-pixels/s", "cores": 1, "kind": info["kind"],
                                        "sample": "%d of the %d strokes of one band (%d visited cells, %.1f s imprint single-threaded as "
                                                  "in the reference); one band's compose %.2f s on %d threads, prorated" % (
                                                      info["n_sample"], len(tile_strokes), info["visited"], info["t_imprint"],
                                                      info["t_compose_full"], info["threads"])}
            except Exception as exc:
                line["cpu_baseline"] = {"unavailable": str(exc)}
        dist.barrier()

    if rank == 0 and not args.no_cpu and world == 1:
        info = cpu_sample(strokes, ROWS, COLS, want_image=True)
        t = info["t_imprint"] + info["t_compose_full"] * info["n_sample"] / len(rec)
        line["cpu_baseline"] = {"value": info["visited"] / t, "unit": "stroke-pixels/s", "cores": 1, "kind": info["kind"],
                                "sample": "%d of %d strokes (round-robin over the 4 brush-size passes, %d visited cells, %.1f s imprint "
                                          "single-threaded as in the reference); full-canvas compose %.2f s on %d threads, prorated" % (
                                              info["n_sample"], len(rec), info["visited"], info["t_imprint"], info["t_compose_full"],
                                              info["threads"])}
        # parity of the product path on the very strokes the CPU just rendered: same sample, fresh canvas, FP32 mode
        order = info["order"]
        prec = np.zeros(len(order), dtype=api.STROKE_DTYPE)
        px_, py_, pt_, first = [], [], [], 0
        for j, i in enumerate(order):
            a, m = int(rec["first_imprint"][i]), int(rec["n_imprints"][i])
            px_.append(cx[a:a + m]), py_.append(cy[a:a + m]), pt_.append(th[a:a + m])
            prec[j] = (rec["radius"][i], rec["K"][i], rec["S"][i], first, m)
            first += m
            ex = info["expanded"][i]  # the oracle's own expansion must equal the library's, bit for bit
            assert np.array_equal(ex[0], px_[-1]) and np.array_equal(ex[1], py_[-1]) and np.array_equal(ex[2], pt_[-1])
        cv.clear()
        br.updateSnapshot(cv)  # the CPU sample starts with a fresh brush: its first imprint copies the whole canvas (:281-284)
        br.stroke_batch(cv, prec, np.concatenate(px_), np.concatenate(py_), np.concatenate(pt_))
        got = cv.compose()
        wet = cv.download("V")["V"] > 0
        line["parity"] = {"max_abs_err": float(np.abs(got - info["R"]).max()), "tolerance": 1e-4, "strokes": len(order),
                          "wet_px": int(wet.sum()), "wet_px_equal": bool(np.array_equal(wet, info["wet"])),
                          "against": "oracle/_ref (the reference's headers)" if info["kind"] == "reference" else "oracle port",
                          "what": "reflectance of the CPU sample's strokes rendered by the product in FP32 mode on a fresh 4K canvas"}
    if rank == 0:
        print(json.dumps(line))
    if planner is not None:
        if pending[0] is not None:
            pending[0].result()
        planner.shutdown()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
