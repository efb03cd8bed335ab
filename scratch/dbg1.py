import sys, numpy as np, torch
sys.path.insert(0,'.')
from painty_b200 import api
from oracle.cpu import Cpu
from tests.workloads import km_random_planes
port = Cpu("port")
ctx = api.Context(0, api.F32)
rows, cols = 33, 47; n = rows*cols
K,S,V,R0 = km_random_planes(rows, cols, seed=100)
want = port.compose_onto(K,S,V,R0)
t = torch.tensor(np.concatenate([K.reshape(n,3).T, S.reshape(n,3).T, V.reshape(1,n)]), dtype=torch.float32, device="cuda")
r0 = torch.tensor(R0.reshape(n,3).T.copy(), dtype=torch.float32, device="cuda")
for name in ("plain","stacked"):
    out = torch.full_like(r0, -7.0)
    torch.cuda.synchronize()
    if name=="plain":
        ctx.km_compose_planes(n, [t[i].data_ptr() for i in range(3)], [t[3+i].data_ptr() for i in range(3)], t[6].data_ptr(), [r0[i].data_ptr() for i in range(3)], [out[i].data_ptr() for i in range(3)])
    else:
        ctx.km_compose_stacked_planes(n, [[t[i].data_ptr() for i in range(3)]], [[t[3+i].data_ptr() for i in range(3)]], [t[6].data_ptr()], [r0[i].data_ptr() for i in range(3)], [out[i].data_ptr() for i in range(3)])
    ctx.synchronize()
    got = out.cpu().numpy().T.reshape(rows,cols,3).astype(np.float64)
    err = np.abs(got-want)
    print(name, err.max(), np.argwhere(err.reshape(-1,3).max(1)>1e-4)[:10].ravel(), got.reshape(-1,3)[:3], want.reshape(-1,3)[:3])
