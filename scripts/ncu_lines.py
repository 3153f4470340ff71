"""Aggregate the warp-stall samples / executed instructions of an .ncu-rep per CUDA source line.
usage: python scripts/ncu_lines.py rep.ncu-rep [kernel-regex] [top-n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kre:
    cmd += ["-k", f"regex:{kre}"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
agg = collections.defaultdict(lambda: [0, 0, ""])
fname = ""
hdr = None
seen_kernel = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Kernel Name":
        seen_kernel += 1
        if seen_kernel > 1: break
        print("kernel:", r[1][:100]); continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r; iS = r.index("Warp Stall Sampling (All Samples)"); iE = r.index("Instructions Executed"); continue
    if hdr is None or len(r) <= iE: continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    key = (fname, ln)
    if r[2] == "":                      # the CUDA source line itself
        agg[key][2] = r[1].strip()
    else:
        num = lambda x: int(x) if x.strip().isdigit() else 0
        agg[key][0] += num(r[iS]); agg[key][1] += num(r[iE])
tot = sum(v[0] for v in agg.values()); tote = sum(v[1] for v in agg.values())
print(f"samples {tot}  executed warp-instrs {tote}")
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100 * v[0] / max(tot, 1):5.1f}% smp {100 * v[1] / max(tote, 1):5.1f}% ins  {f}:{ln:<4d} {v[2][:110]}")
