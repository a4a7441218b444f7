"""Opcode histogram per kernel of libdcgp.so (cuobjdump -sass): the Blackwell-native evidence of B200_PROFILING.md
(UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = TMA, DMMA = FP64 tensor pipe, HMMA.*TF32 = legacy
mma.sync TF32, MUFU = special-function unit).   python tools/sass_histogram.py [lib] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "deconditional_downscaling_b200", "lib", "libdcgp.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "DMMA", "HMMA", "MUFU.EX2", "MUFU.RSQ",
        "MUFU.RCP", "DFMA", "FFMA", "SHFL", "LDGSTS", "ATOMG", "REDG", "RED.", "BAR.SYNC", "LDG", "STG", "LDS", "STS"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"^void ", "", name)
        cur = name.split("(")[0]
        per.setdefault(cur, collections.Counter())
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                per[cur][k] += 1
print(f"# {os.path.relpath(lib, ROOT)}: {len(per)} kernels; columns = instruction counts in the SASS (static)")
tot = collections.Counter()
for name, c in sorted(per.items(), key=lambda kv: kv[0]):
    tot.update(c)
    interesting = {k: v for k, v in c.items() if k != "_total" and v and k in ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "DMMA", "HMMA", "MUFU.EX2", "MUFU.RSQ", "DFMA", "SHFL", "LDGSTS", "REDG", "ATOMG", "RED.")}
    print(f"{name[:110]:110s} total {c['_total']:6d}  " + "  ".join(f"{k} {v}" for k, v in interesting.items()))
print("# whole library: " + "  ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))
