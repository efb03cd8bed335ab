"""Summarise an ncu report's source page: per CUDA source line, stall samples and instructions executed."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
si, ie = h.index("# Samples"), h.index("Instructions Executed")
lines = {}
cur = None
for r in rows[hi + 1:]:
    if len(r) <= ie: continue
    if r[0] != "":
        cur = (r[0], r[1]); lines.setdefault(cur, [0, 0]); continue
    if cur is None: continue
    try:
        lines[cur][0] += int(r[si]); lines[cur][1] += int(r[ie])
    except ValueError:
        pass
tot_s = sum(v[0] for v in lines.values()); tot_i = sum(v[1] for v in lines.values())
print("total samples", tot_s, "total warp-instructions", tot_i)
for (ln, src), (s, i) in sorted(lines.items(), key=lambda kv: -kv[1][int(sys.argv[3]) if len(sys.argv) > 3 else 0])[:top]:
    print(f"{100*s/max(tot_s,1):5.1f}% smp {100*i/max(tot_i,1):5.1f}% ins  L{ln:>4} {src.strip()[:130]}")
