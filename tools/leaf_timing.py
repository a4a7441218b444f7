"""Cycle breakdown of the serial chain of the FP32 look-ahead factorisation (potrf_step_kernel: load / P / syrk / potrf /
trtri + stores; block_potrf: pivots / panel / trailing), from the clock64() instrumentation behind -DDCGP_LEAF_TIMING.

    python tools/leaf_timing.py build      # build container (no GPU): lib/timing/libdcgp.so, travels with gpurun
    python tools/leaf_timing.py run [n]    # GPU box: one FP32 potrf of n (default 8192), prints the median link
"""
import os
import re
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "deconditional_downscaling_b200")
OUT = os.path.join(PKG, "lib", "timing")
LIB = os.path.join(OUT, "libdcgp.so")


def build():
    sys.path.insert(0, ROOT)
    from deconditional_downscaling_b200 import build as b
    os.makedirs(OUT, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs, procs = [], []
    for src in b._sources():
        obj = os.path.join(OUT, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen([nvcc, *b.NVCC_FLAGS, "-DDCGP_LEAF_TIMING", "-c", src, "-o", obj]))
    if any(p.wait() for p in procs):
        raise SystemExit("nvcc failed")
    subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"], check=True)
    print(LIB)


def run(n):
    env = dict(os.environ, DCGP_LIB=LIB)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "one_potrf.py"), "f32", str(n), "16"], env=env,
                         capture_output=True, text=True).stdout
    steps = [list(map(int, m.groups())) for m in re.finditer(r"step kernel: load (\d+)  P (\d+)  syrk (\d+)  potrf (\d+)  rest (\d+)", out)]
    leaf = [list(map(int, m.groups())) for m in re.finditer(r"potrf steps: \(first diag\) (\d+)  panel (\d+)  trailing\+next diag (\d+)", out)]
    tri = [list(map(int, m.groups())) for m in re.finditer(r"trtri: level0 (\d+)  A (\d+)  B (\d+)  C (\d+)", out)]
    tail = [list(map(int, m.groups())) for m in re.finditer(r"tail: storeL (\d+)  trtri (\d+)  storeinv (\d+)", out)]
    for name, rows, cols in (("potrf_step_kernel", steps, ["load", "P", "syrk", "potrf", "trtri+stores"]),
                             ("block_potrf", leaf, ["first diag", "panels", "trailing + next diag"]),
                             ("block_trtri", tri, ["level 0", "steps A", "steps B", "steps C"]),
                             ("step tail", tail, ["store L", "trtri", "store inverse"])):
        if rows:
            med = [statistics.median(c) for c in zip(*rows)]
            print(name, {k: f"{v:.0f} cyc = {v / 1.965e3:.1f} us" for k, v in zip(cols, med)}, f"({len(rows)} samples)")
    print([l for l in out.splitlines() if l.startswith("potrf ")])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build()
    else:
        run(int(sys.argv[2]) if len(sys.argv) > 2 else 8192)
