"""Per-imprint latency inside the multi-GPU kernel: one stroke that stays in rank 0's band, and one that runs along the band
boundary (half of its footprint in rank 1's band: one staging window per dataflow segment).
torchrun --nproc-per-node 2 scratch/dist_micro.py [R1,R2,..] [N_IMPRINTS]   (2 GPUs)"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
from painty_b200 import api, assets
from painty_b200.dist import DistCanvas
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
radii = [float(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "30,112,151".split(","))]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400
rows, cols = 2160 * world, 3840
ctx = api.Context(local, api.F32)
stream = torch.cuda.ExternalStream(ctx.stream, device=local)
dc = DistCanvas(ctx, rows, cols, dist)
br = api.FootprintBrush(ctx, assets.snap_to_safe_radius(radii[0]))
for r in radii:
    br.register_radius(assets.snap_to_safe_radius(r))
dc.attach(br)
lib = api.lib()
if os.environ.get("PB_TRACE"):
    api._chk(lib.pb_fbrush_enable_trace(br.h, 1))
for r in radii:
    r = assets.snap_to_safe_radius(r)
    for name, y0 in (("inside band 0", 1000.0), ("along the boundary", 2150.0)):
        cx = np.linspace(600, 600 + n, n); cy = np.full(n, y0) + np.linspace(0, 3.0, n); th = np.full(n, 0.79)
        rec = np.zeros(1, dtype=api.STROKE_DTYPE)
        rec[0] = (r, [.3, .2, .1], [.2, .4, .3], 0, n)
        best = 1e30
        for rep in range(4):
            dc.canvas.clear(); br.updateSnapshot(dc.canvas)
            d = dc._desc[id(br)]
            api._chk(lib.pb_fbrush_dist_begin(br.h, dc.canvas.h))
            plan = dc.plan(br, rec, cx, cy, th)
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); br.run_batch_plan(dc.canvas, plan, dist_desc=d); e1.record(stream)
            ctx.synchronize(); dist.barrier()
            api._chk(lib.pb_fbrush_dist_end(br.h, dc.canvas.h))
            if rep and rank == 0: best = min(best, e0.elapsed_time(e1))
        if rank == 0:
            print("r=%5.1f %-20s %7.2f ms  %6.2f us/imprint%s" % (r, name, best, best * 1e3 / n, ""), flush=True)
        if os.environ.get("PB_TRACE") and rank == 0:  # cycle stamps of the last repetition's stroke (thread 0)
            import ctypes as C
            out = np.zeros((256, 2, 8), dtype=np.uint64)
            api._chk(lib.pb_fbrush_read_trace(br.h, out.ctypes.data_as(C.c_void_p)))
            t = out[:min(n, 256), 0, :5].astype(np.int64)
            per = np.diff(t[:, 0])
            seg = [i for i in range(1, len(per)) if (i + 1) % 64 == 0]
            print("   cycles per imprint: median %d; imprints that start a segment (incl. staging): %s; first interval (ring) there: %s" % (
                np.median(per), [int(per[i]) for i in seg], [int(t[i + 1, 1] - t[i + 1, 0]) for i in seg]), flush=True)
dc.close()
dist.destroy_process_group()
