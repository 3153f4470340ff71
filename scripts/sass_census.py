"""Per-kernel census of the Blackwell instructions in the shipped library (tcgen05.mma = UTCHMMA, tcgen05.ld/st = LDTM/STTM,
TMA = UTMALDG/UTMASTG/UTMAPF, bulk copies = UBLKCP, legacy tensor-core = HMMA, packed fp32 = FADD2/FMUL2/FFMA2).
usage: python scripts/sass_census.py [lib.so] > profiles/r02_sass_census.txt"""
import collections, re, subprocess, sys, os
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "iisan_b200", "lib", "libiisan_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "FFMA2", "FADD2", "FMUL2", "MUFU.EX2", "RED", "ATOM"]
cur = None
tab = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        tab[cur] = collections.Counter(); tab[cur]["instructions"] = 0
        continue
    if cur is None or not re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        continue
    tab[cur]["instructions"] += 1
    for k in KEYS:
        if re.search(r"\b" + re.escape(k), line):
            tab[cur][k] += 1
print(f"{'kernel':70s} {'instr':>6s} " + " ".join(f"{k:>8s}" for k in KEYS))
for name, c in tab.items():
    if not any(c[k] for k in KEYS[:8]) and c["HMMA"] == 0:
        continue
    print(f"{name[:70]:70s} {c['instructions']:6d} " + " ".join(f"{c[k]:8d}" for k in KEYS))
