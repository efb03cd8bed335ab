"""torchrun probe: weak-scaling bench variants to see what straddlers cost. Usage: torchrun ... scratch/dist_probe.py <strokes_per_gpu> <mode>
mode: all | confined (drop strokes whose region leaves their band)"""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
import bench
from painty_b200 import api, bands
from painty_b200.dist import DistCanvas
n_per, mode = int(sys.argv[1]), sys.argv[2]
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rows = bench.ROWS * world
_, rec, cx, cy, th, radii = bench.build_workload(n_per * world, rows=rows)
if mode == "confined":
    keep = []
    rpb = (rows + world - 1) // world
    for i in range(len(rec)):
        a, m = int(rec["first_imprint"][i]), int(rec["n_imprints"][i])
        if m == 0: continue
        lo, hi = bands.footprint_stroke_rows(float(rec["radius"][i]), cy[a:a+m])
        if max(lo,0) // rpb == min(hi, rows-1) // rpb: keep.append(i)
    rec = rec[keep]
ctx = api.Context(local, api.F32)
dc = DistCanvas(ctx, rows, bench.COLS, dist)
br = api.FootprintBrush(ctx, radii[0])
for r in radii: br.register_radius(r)
dc.attach(br)
ts = []
for it in range(3):
    dc.canvas.clear(); br.updateSnapshot(dc.canvas); ctx.synchronize(); dist.barrier()
    t0 = time.perf_counter(); dc.stroke_batch(br, rec, cx, cy, th); ts.append(time.perf_counter() - t0)
if rank == 0: print(f"world={world} mode={mode} strokes={len(rec)} best={min(ts):.3f}s")
dc.close(); dist.destroy_process_group()
