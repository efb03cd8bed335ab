import sys, heapq, numpy as np
sys.path.insert(0,'.')
import bench
from painty_b200 import assets
n = int(sys.argv[1]) if len(sys.argv)>1 else 10000
world = int(sys.argv[2]) if len(sys.argv)>2 else 1
rows, cols, tile = 2160*world, 3840, 64
_, rec, cx, cy, th, radii = bench.build_workload(n, rows=rows)
tx, ty = (cols+tile-1)//tile, (rows+tile-1)//tile
last = -np.ones((ty,tx), dtype=np.int64); ring = [[[] for _ in range(tx)] for _ in range(ty)]
def tmodel(r): return np.interp(r, [11,30,64,112,151,200], [5.3,6.0,10,16,27,45])*1e-6
preds=[]; dur=np.zeros(n)
for s in range(n):
    a,m = int(rec["first_imprint"][s]), int(rec["n_imprints"][s]); r=float(rec["radius"][s])
    side = assets.footprint_geometry(r)[3]; wr=(side-1)//2
    dur[s] = m*tmodel(r)
    if m==0: preds.append([]); continue
    def reg(mg):
        mm = wr+mg+2
        return (max(0,int(np.floor(cx[a:a+m].min()-mm)))//tile, max(0,int(np.floor(cy[a:a+m].min()-mm)))//tile,
                min(cols-1,int(np.ceil(cx[a:a+m].max()+mm)))//tile, min(rows-1,int(np.ceil(cy[a:a+m].max()+mm)))//tile)
    bx = reg(0); al = reg(r); P=set()
    if al[2]>=al[0] and al[3]>=al[1]:
        P.update(int(v) for v in np.unique(last[al[1]:al[3]+1, al[0]:al[2]+1]) if v>=0)
    if bx[2]>=bx[0] and bx[3]>=bx[1]:
        for y in range(bx[1],bx[3]+1):
            for x in range(bx[0],bx[2]+1): P.update(ring[y][x])
    for y in range(al[1],al[3]+1):
        for x in range(al[0],al[2]+1): ring[y][x].append(s)
    for y in range(bx[1],bx[3]+1):
        for x in range(bx[0],bx[2]+1): ring[y][x]=[]; last[y,x]=s
    P.discard(s); preds.append(sorted(P))
# critical path
fin=np.zeros(n)
for s in range(n): fin[s]=(max(fin[p] for p in preds[s]) if preds[s] else 0)+dur[s]
print("serial sum %.2fs  critical path %.2fs  strokes %d avg preds %.1f"%(dur.sum(), fin.max(), n, np.mean([len(p) for p in preds])))
def sim(slots):
    free=[0.0]*slots; heapq.heapify(free); fin=np.zeros(n)
    for s in range(n):
        t0=heapq.heappop(free); start=max([t0]+[fin[p] for p in preds[s]]); fin[s]=start+dur[s]; heapq.heappush(free, fin[s])
    return fin.max()
for sl in (2,4,9,18,36,74,148): print("slots",sl,"makespan %.2fs"%sim(sl))
