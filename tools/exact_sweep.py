"""BASELINE config 4: ExactCMP Gram + Cholesky + solve composite over N = 4k .. 64k in FP64 and FP32
(3xTF32 factorisation / solves, single-pass TF32 products), one GPU.  python tools/exact_sweep.py [out.json] [Nmax]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import exact  # noqa: E402


def one(N, dt, device):
    g = torch.Generator().manual_seed(7)
    n = N // 100
    x = torch.randn(N, 3, generator=g, dtype=dt)
    x = (x - x.mean(0)) / x.std(0)
    ybag = torch.randn(n, 3, generator=g, dtype=dt)
    ext = ybag[torch.arange(N) * n // N] + 0.3 * torch.randn(N, 3, generator=g, dtype=dt)   # r = 3 continuous variant
    z = torch.randn(n, generator=g, dtype=dt)
    x, ext, ybag, z = (t.to(device) for t in (x, ext, ybag, z))
    res = exact.exact_cmp_phases(x, ext, ybag, z, x, 0.01)          # warm-up at full size: allocator, kernel attributes
    del res
    torch.cuda.synchronize()
    timer = exact.PhaseTimer(True)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    res = exact.exact_cmp_phases(x, ext, ybag, z, x, 0.01, timer=timer)
    e.record()
    torch.cuda.synchronize()
    total = s.elapsed_time(e) * 1e-3
    ph = timer.seconds()
    fl = exact.exact_cmp_flops(N, n, 3, N)
    size = 8 if dt == torch.float64 else 4
    out = {"N": N, "n_bags": n, "dtype": "f64" if dt == torch.float64 else "f32/3xTF32", "seconds": total,
           "composite_tflops": sum(fl.values()) / total / 1e12,
           "phases_ms": {k: round(1e3 * v, 3) for k, v in ph.items()},
           "potrf_tflops": fl["potrf"] / ph["potrf"] / 1e12, "potrs_tflops": fl["potrs"] / ph["potrs"] / 1e12,
           "KW_tflops": fl["KW"] / ph["KW"] / 1e12, "gram_K_gbs": N * N * size / ph["gram_K"] / 1e9,
           "potrf_info": int(res["info"].item()), "mll": float(res["mll"].item()),
           "post_var_min": float(res["post_var"].min().item())}
    del res
    torch.cuda.empty_cache()
    return out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else None
    nmax = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
    device = torch.device("cuda", 0)
    # kernels / look-ahead lanes once
    one(4096, torch.float32, device)
    one(4096, torch.float64, device)
    rows = []
    for dt in (torch.float64, torch.float32):
        for N in (4096, 8192, 16384, 32768, 65536):
            if N > nmax:
                continue
            r = one(N, dt, device)
            rows.append(r)
            print(json.dumps(r), flush=True)
            if out_path:
                with open(out_path, "w") as f:
                    json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
