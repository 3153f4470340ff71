"""Write profiles/<tag>_chain_traffic.json from an `ncu --set full` capture of the chain kernels: DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum) of the forward and the backward kernel.  bench.py reads the newest such file for
`roofline.traffic` and names it in the JSON line.   usage: python scripts/ncu_traffic.py rep.ncu-rep profiles/r02_chain_traffic.json"""
import csv, io, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
res = {"source": rep.split("/")[-1]}
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    name = d.get("Kernel Name", "")
    key = "bwd" if "bwd" in name else ("fwd" if "chain" in name else ("rank" if "lr_rank" in name else ("wgrad" if "umma_gemm_kernel<256, 1, 1, 48" in name else None)))
    if key is None or key in res:
        continue
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(d[m].replace(",", "")) * UNIT.get(u[m], 1.0)
    res[key] = tot
    res[key + "_kernel"] = name.split("(")[0]
    res[key + "_us_under_ncu"] = float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}.get(u["gpu__time_duration.sum"], 1.0)
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
