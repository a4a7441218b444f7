"""Gram / bag kernels at the sizes the benchmarks issue them, one JSON line each with the algorithmic bytes and kernel
evaluations per second:   python tools/gram_kernels.py [ncu]
With the `ncu` argument every kernel runs exactly twice (one warm-up, one measured launch) so that
    ncu --set full --clock-control none -k regex:'gram_|bag' python tools/gram_kernels.py ncu
captures a short list; tools/ncu_kernel_table.py turns the report into the table under profiles/."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops  # noqa: E402
from deconditional_downscaling_b200.exact import rbf_spec  # noqa: E402

NCU = len(sys.argv) > 1 and sys.argv[1] == "ncu"
dev = torch.device("cuda", 0)
torch.manual_seed(0)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    if NCU:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) * 1e-3)
    return best


def line(kernel, dt, shape, sec, nbytes, evals):
    print(json.dumps({"kernel": kernel, "dtype": str(dt).split(".")[-1], "shape": shape, "ms": round(sec * 1e3, 4),
                      "algorithmic_GBps": round(nbytes / sec / 1e9, 1), "Gevals_per_s": round(evals / sec / 1e9, 1)}), flush=True)


for dt, N, d in [(torch.float32, 8192, 5), (torch.float64, 16384, 3)]:
    es = 4 if dt == torch.float32 else 8
    x = torch.randn(N, d, dtype=dt, device=dev)
    spec = rbf_spec(d, dt, dev)
    K = torch.empty(N, N, dtype=dt, device=dev)
    t = timed(lambda: ops.gram(spec, x, None, diag_const=0.5, out=K))
    line("gram_fwd", dt, f"{N}x{N} d={d}", t, N * N * es, N * N)       # one write of K
    G = torch.randn(N, N, dtype=dt, device=dev)
    t = timed(lambda: ops.gram_bwd(spec, x, None, G))
    line("gram_bwd(hypers)", dt, f"{N}x{N} d={d}", t, N * N * es, N * N)   # one read of G
    t = timed(lambda: ops.gram_bwd(spec, x, None, G, need_dx1=True))
    line("gram_bwd(hypers+dx)", dt, f"{N}x{N} d={d}", t, N * N * es, N * N)
    del K, G

# bag kernels of the VBagg step (BASELINE config 3: 1024 bags of 64 pixels, M = 1024 inducing points, d = 5)
for dt in (torch.float32, torch.float64):
    es = 4 if dt == torch.float32 else 8
    nb, per, M, d = 1024, 64, 1024, 5
    x = torch.randn(nb * per, d, dtype=dt, device=dev)
    z = torch.randn(M, d, dtype=dt, device=dev)
    off = torch.arange(0, nb * per + 1, per, dtype=torch.int64, device=dev)
    spec = rbf_spec(d, dt, dev)
    t = timed(lambda: ops.gram_bagsum(spec, z, x, off))
    line("gram_bagsum", dt, f"{nb} bags x {per}, M={M}", t, (nb * per * d + M * d + nb * M) * es, nb * per * M)
    Gs = torch.randn(nb, M, dtype=dt, device=dev)
    t = timed(lambda: ops.gram_bagsum_bwd(spec, z, x, off, Gs))
    line("gram_bagsum_bwd", dt, f"{nb} bags x {per}, M={M}", t, (nb * per * d + 2 * M * d + nb * M) * es, nb * per * M)
    t = timed(lambda: ops.gram_bag_blocksum(spec, x, off))
    line("gram_bag_blocksum", dt, f"{nb} bags x {per}", t, (nb * per * d + nb) * es, nb * per * per)
    g = torch.randn(nb, dtype=dt, device=dev)
    t = timed(lambda: ops.gram_bag_blocksum_bwd(spec, x, off, g))
    line("gram_bag_blocksum_bwd", dt, f"{nb} bags x {per}", t, (nb * per * d + nb) * es, nb * per * per)
