"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total and share."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
seq = []
for r in rows:
    if r[0] == "ID":
        hdr = r; continue
    if hdr is None: continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum": continue
    if int(d["ID"]) < skip: continue
    name = re.sub(r"\(.*", "", d["Kernel Name"])
    name = re.sub(r"^void ", "", name)[:90]
    v = float(d["Metric Value"].replace(",", ""))
    u = d["Metric Unit"]
    v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
    agg[name][0] += 1; agg[name][1] += v
    seq.append((int(d["ID"]), name, v, d["Grid Size"], d["Block Size"]))
tot = sum(v[1] for v in agg.values())
print(f"total {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v[1]:10.1f} us {v[0]:5d}  {100 * v[1] / tot:5.1f}%  avg {v[1] / v[0]:8.1f}  {k}")
if len(sys.argv) > 3:
    for s in seq: print(s)
