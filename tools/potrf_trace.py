"""Device timeline of the two look-ahead lanes of dcgp_potrf (events behind -DDCGP_TRACE in csrc/chol.cu).

    python tools/potrf_trace.py build            # build container: lib/trace/libdcgp.so (travels with gpurun)
    python tools/potrf_trace.py run f32 8192     # GPU box: prints the TRACE records of one factorisation
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "deconditional_downscaling_b200", "lib", "trace")
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
        procs.append(subprocess.Popen([nvcc, *b.NVCC_FLAGS, "-DDCGP_TRACE", "-c", src, "-o", obj]))
    if any(p.wait() for p in procs):
        raise SystemExit("nvcc failed")
    subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"], check=True)
    print(LIB)


if __name__ == "__main__":
    if sys.argv[1] == "build":
        build()
    else:
        env = dict(os.environ, DCGP_LIB=LIB)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "one_potrf.py"), sys.argv[2], sys.argv[3], "16"],
                             env=env, capture_output=True, text=True)
        print(out.stdout[-400000:])
        print(out.stderr[-2000:])
