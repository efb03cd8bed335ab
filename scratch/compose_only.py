"""One fully-wet 4K compose through the C ABI (device planes), for ncu captures."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
from painty_b200 import api
rows, cols = 2160, 3840
n = rows * cols
prec = api.F64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else api.F32
dt = torch.float64 if prec == api.F64 else torch.float32
ctx = api.Context(0, prec)
g = torch.Generator(device="cuda").manual_seed(42)
pl = torch.empty((10, n), dtype=dt, device="cuda")
pl[0:3] = torch.exp(torch.rand((3, n), device="cuda", generator=g, dtype=dt) * (np.log(4.32) - np.log(1e-3)) + np.log(1e-3))
pl[3:6] = torch.exp(torch.rand((3, n), device="cuda", generator=g, dtype=dt) * (np.log(1.21) - np.log(1e-3)) + np.log(1e-3))
pl[6] = torch.rand(n, device="cuda", generator=g, dtype=dt) * 0.9 + 0.05
pl[7:10] = torch.rand((3, n), device="cuda", generator=g, dtype=dt) * 0.96 + 0.02
out = torch.empty((3, n), dtype=dt, device="cuda")
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
p = [pl[i].data_ptr() for i in range(10)]
stream = torch.cuda.ExternalStream(ctx.stream, device=0)
for it in range(6):
    flush.fill_(float(it)); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    ctx.km_compose_planes(n, p[0:3], p[3:6], p[6], p[7:10], [out[i].data_ptr() for i in range(3)])
    e1.record(stream); ctx.synchronize()
    ms = e0.elapsed_time(e1)
    bpp = 104 if prec == api.F64 else 52
    print(f"compose {rows}x{cols} {'f64' if prec else 'f32'}: {ms*1e3:.1f} us  {bpp*n/ms/1e6:.0f} GB/s  {n/ms/1e6:.1f} Gpx/s")
