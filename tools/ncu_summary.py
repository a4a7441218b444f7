"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python tools/ncu_summary.py in.csv out.csv"""
import csv, re, sys
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = {}
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = name.split("(")[0][:90]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
with open(sys.argv[2], "w") as f:
    f.write("kernel,launches,total_ms,share\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{v[0]},{v[1]:.3f},{v[1]/tot:.3f}\n")
print(open(sys.argv[2]).read()[:3000])
