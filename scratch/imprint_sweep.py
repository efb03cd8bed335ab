"""Launch-shape sweep of the imprint chain: single-stroke latency per imprint for cluster size x block size per footprint
class (env knobs PB_IMPRINT_CLUSTER16/17, PB_IMPRINT_BLOCK16/17, PB_RING_THREADS). Prints latency and the SM-time per
imprint (latency x CTAs), the figure that matters when a batch has more independent strokes than cluster slots. (GPU)"""
import os, subprocess, sys
N = sys.argv[1] if len(sys.argv) > 1 else "400"
def run(radii, env):
    e = dict(os.environ); e.update({k: str(v) for k, v in env.items()})
    r = subprocess.run([sys.executable, "scratch/imprint_micro.py", radii, N], env=e, capture_output=True, text=True, timeout=600)
    return [l for l in r.stdout.splitlines() if l.startswith("r=")] or [r.stderr[-300:]]
print("== class 16 (<= 4096 active cells): r = 30, 45, 56")
for cl, bl in ((16, 128), (8, 128), (8, 256), (4, 256), (4, 512), (2, 512), (1, 512)):
    for line in run("30,45,56", {"PB_IMPRINT_CLUSTER16": cl, "PB_IMPRINT_BLOCK16": bl}):
        print("cluster %2d x %3d: %s" % (cl, bl, line), flush=True)
print("== class 17: r = 64, 80 (<= 8192 cells), 112, 129, 151")
for cl, bl, radii in ((16, 256, "64,80"), (8, 256, "64,80"), (8, 512, "64,80"), (4, 512, "64,80"), (16, 512, "64,80,112,129,151"),
                      (8, 512, "112,129,151")):
    for line in run(radii, {"PB_IMPRINT_CLUSTER17": cl, "PB_IMPRINT_BLOCK17": bl}):
        print("cluster %2d x %3d: %s" % (cl, bl, line), flush=True)
print("== ring threads (default 64): r = 30, 112, 151")
for rt in (32, 64, 128, 256):
    for line in run("30,112,151", {"PB_RING_THREADS": rt}):
        print("ring threads %3d: %s" % (rt, line), flush=True)
