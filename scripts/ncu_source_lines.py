"""Stall samples, executed instructions and shared-memory wavefronts of one .ncu-rep by source line:
    python scripts/ncu_source_lines.py file.ncu-rep   (needs --import-source on at capture time)"""
import csv, subprocess, sys
out = subprocess.run(["ncu","-i",sys.argv[1],"--page","source","--csv","--print-source","cuda,sass"]+sys.argv[2:],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
# find header
hi = next(i for i,r in enumerate(rows) if r and r[0]=="Line No")
hdr = rows[hi]
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed"); wi = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
lines=[]
for r in rows[hi+1:]:
    if len(r) <= si: continue
    if r[0] != "":   # source line aggregate row
        try: lines.append((int(r[si]), int(r[ii]), (int(r[wi] or 0) if wi is not None else 0), r[0], r[1][:110]))
        except ValueError: pass
tot = sum(l[0] for l in lines); toti = sum(l[1] for l in lines); totw = sum(l[2] for l in lines)
print("total samples", tot, "instr", toti, "smem wavefronts", totw)
for s,i,w,ln,src in sorted(lines, reverse=True)[:int(45)]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% ins {100*w/max(totw,1):5.1f}% wav  L{ln}: {src}")
