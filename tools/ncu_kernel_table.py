"""Per-kernel table out of an `ncu --set full` report:  python tools/ncu_kernel_table.py report.ncu-rep [out.txt]
Keeps the LAST launch of every kernel name (the measured one of tools/gram_kernels.py ncu) and prints duration, DRAM
bytes, DRAM / SM throughput in % of peak, pipe utilisations (FMA, FP64, XU = MUFU, ALU, LSU), registers, occupancy."""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_cyc%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_cyc%"),
    ("sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active", "xu_cyc%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
]

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
kn = col["Kernel Name"]
last = {}
for r in body:
    name = re.sub(r"\(anonymous namespace\)::|^void ", "", r[kn]).split("(")[0]
    last[name] = r


def fmt(metric, r):
    if metric not in col:
        return "-"
    v, u = r[col[metric]], units[col[metric]]
    try:
        f = float(v.replace(",", ""))
    except ValueError:
        return v or "-"
    if metric.startswith("dram__bytes"):
        f *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return f"{f / 1e6:.1f}MB"
    if metric.startswith("gpu__time"):
        f *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3}.get(u, 1)
        return f"{f:.1f}us"
    return f"{f:.1f}"


out = ["kernel | " + " | ".join(s for _, s in WANT)]
for name, r in last.items():
    out.append(name[:70] + " | " + " | ".join(fmt(m, r) for m, _ in WANT))
text = "\n".join(out) + "\n"
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(text)
print(text)
