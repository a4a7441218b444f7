"""Kernel micro-benchmarks (run on the GPU box): every libdcgp hot kernel timed with
CUDA events after warm-up, next to the cuBLAS/cuSOLVER call torch makes for the same op
(the measured FP64 / TF32 peaks used as roofline denominators).  Writes
gpurun_out/kbench.json.  Usage: python tools/kbench.py [--quick]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deconditional_downscaling_b200 import ops  # noqa: E402

DEV = "cuda:0"
QUICK = "--quick" in sys.argv


def timeit(fn, warmup=2, iters=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e-3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    res = {"device": torch.cuda.get_device_name(0), "results": []}

    def rec(name, **kw):
        kw["name"] = name
        res["results"].append(kw)
        print(json.dumps(kw), flush=True)

    # ---------------- GEMM ----------------
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "tf32")):
        n = 4096 if QUICK else 8192
        A = torch.randn(n, n, dtype=dtype, device=DEV)
        B = torch.randn(n, n, dtype=dtype, device=DEV)
        C = torch.empty(n, n, dtype=dtype, device=DEV)
        flops = 2.0 * n ** 3
        torch.backends.cuda.matmul.allow_tf32 = True
        med, best = timeit(lambda: torch.matmul(A, B.t(), out=C))
        rec(f"cublas_gemm_nt_{tag}", n=n, ms=med * 1e3, tflops=flops / med / 1e12, tflops_best=flops / best / 1e12)
        for ta, tb, lay in ((False, True, "nt"), (False, False, "nn"), (True, False, "tn")):
            med, best = timeit(lambda: ops.gemm(A, B, transa=ta, transb=tb, C=C))
            rec(f"dcgp_gemm_{lay}_{tag}", n=n, ms=med * 1e3, tflops=flops / med / 1e12, tflops_best=flops / best / 1e12)
        # SYRK-shaped trailing update (K = 128, lower tiles only)
        P = torch.randn(n, 128, dtype=dtype, device=DEV)
        med, best = timeit(lambda: ops.gemm(P, P, transb=True, alpha=-1.0, beta=1.0, C=C, lower=True))
        rec(f"dcgp_syrk_k128_{tag}", n=n, ms=med * 1e3, tflops=(n * n * 128.0) / med / 1e12)
        del A, B, C, P

    # ---------------- Gram ----------------
    from oracle_free_spec import rbf_spec
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "f32")):
        for n in ((8192,) if QUICK else (16384, 32768)):
            x = torch.randn(n, 3, dtype=dtype, device=DEV)
            K = torch.empty(n, n, dtype=dtype, device=DEV)
            spec = rbf_spec(3, dtype)
            med, best = timeit(lambda: ops.gram(spec, x, None, diag_const=1.0, out=K))
            nbytes = n * n * K.element_size()
            rec(f"dcgp_gram_rbf_{tag}", n=n, ms=med * 1e3, gbs=nbytes / med / 1e9, gbs_best=nbytes / best / 1e9)
            if n <= 16384:
                G = torch.randn(n, n, dtype=dtype, device=DEV)
                med, best = timeit(lambda: ops.gram_bwd(spec, x, None, G, need_dx1=False))
                rec(f"dcgp_gram_bwd_rbf_{tag}", n=n, ms=med * 1e3, gbs=nbytes / med / 1e9)
                del G
            del K

    # ---------------- Cholesky / solves ----------------
    for dtype, tag in ((torch.float64, "f64"), (torch.float32, "tf32")):
        for n in ((4096,) if QUICK else (8192, 16384, 32768)):
            if dtype == torch.float32 and n > 16384:
                continue
            x = torch.randn(n, 3, dtype=dtype, device=DEV)
            spec = rbf_spec(3, dtype)
            A0 = ops.gram(spec, x, None, diag_const=0.01 * n)
            A = torch.empty_like(A0)

            def run():
                A.copy_(A0)
                return ops.potrf_(A)

            t_copy, _ = timeit(lambda: A.copy_(A0))
            med, best = timeit(run, warmup=1, iters=3)
            med -= t_copy
            info = run().item()
            rec(f"dcgp_potrf_{tag}", n=n, ms=med * 1e3, tflops=n ** 3 / 3.0 / med / 1e12, info=info)
            if n <= 16384:
                t0, _ = timeit(lambda: torch.linalg.cholesky(A0), warmup=1, iters=3)
                rec(f"cusolver_potrf_{tag}", n=n, ms=t0 * 1e3, tflops=n ** 3 / 3.0 / t0 / 1e12)
            m = n // 100
            B0 = torch.randn(n, m, dtype=dtype, device=DEV)
            B = torch.empty_like(B0)

            def solve():
                B.copy_(B0)
                ops.potrs_(A, B)

            med, best = timeit(solve, warmup=1, iters=3)
            rec(f"dcgp_potrs_{tag}", n=n, m=m, ms=med * 1e3, tflops=4.0 * n * n * m / med / 1e12)
            if n <= 16384:
                resid = (A0.double() @ B.double() - B0.double()).norm() / B0.double().norm()
                rec(f"potrs_residual_{tag}", n=n, resid=resid.item())
            del A0, A, B0, B

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kbench.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
