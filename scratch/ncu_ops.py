"""Per source line, which SASS opcodes were executed (ncu report with source info)."""
import csv, subprocess, collections, sys
rep = sys.argv[1]; keys = sys.argv[2].split(",") if len(sys.argv) > 2 else ["IMAD", "ISETP", "LDC"]
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","sass,cuda","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hi=next(i for i,r in enumerate(rows) if r and r[0]=="Line No")
h=rows[hi]; si=h.index("# Samples"); ie=h.index("Instructions Executed")
cur=None; per=collections.defaultdict(collections.Counter); smp=collections.Counter(); lines={}
for r in rows[hi+1:]:
    if len(r)<=ie: continue
    if r[0]!="":
        try: cur=int(r[0])
        except ValueError: continue
        lines[cur]=r[1].strip(); continue
    try: n=int(r[ie]); s=int(r[si])
    except ValueError: continue
    ins=r[3].strip().split()
    if not ins: continue
    op=ins[1] if ins[0].startswith("@") else ins[0]
    per[cur][op.split(".")[0]]+=n; smp[cur]+=s
tot=sum(sum(c.values()) for c in per.values())
print("total warp instructions", tot)
print("== all opcodes by line")
for ln,c in sorted(per.items(), key=lambda kv:-sum(kv[1].values()))[:45]:
    print("   %5.2f%% ins %5.2f%% smp L%4d %-90s %s"%(100*sum(c.values())/tot, 100*smp[ln]/max(sum(smp.values()),1), ln, lines[ln][:90], dict(c.most_common(4))))
for key in keys:
    print("==",key, "%.1f%% of all"%(100*sum(c[key] for c in per.values())/tot))
    for ln,c in sorted(per.items(), key=lambda kv:-kv[1][key])[:8]:
        print("   %5.2f%%  L%4d %s"%(100*c[key]/tot, ln, lines[ln][:110]))
