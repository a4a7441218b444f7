"""Latency (single launch between events) and throughput (20 back-to-back) of dcgp_gemm for chain shapes."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops, _lib
lib = _lib.load()
def run(m, n, k, dtype, flags, beta=1.0):
    A = torch.randn(m, k, dtype=dtype, device="cuda"); B = torch.randn(k, n, dtype=dtype, device="cuda")
    C = torch.zeros(m, n, dtype=dtype, device="cuda")
    dt = 0 if dtype == torch.float64 else 1
    call = lambda: lib.dcgp_gemm(0, 0, m, n, k, -1.0, A.data_ptr(), k, B.data_ptr(), n, beta, C.data_ptr(), n, dt, flags, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): call()
    lat = 1e9
    for _ in range(5):
        torch.cuda.synchronize()
        s, e = (torch.cuda.Event(enable_timing=True) for _ in range(2))
        s.record(); call(); e.record(); torch.cuda.synchronize(); lat = min(lat, s.elapsed_time(e))
    s.record()
    for _ in range(20): call()
    e.record(); torch.cuda.synchronize()
    return lat * 1e3, s.elapsed_time(e) / 20 * 1e3
X3, NOTC = _lib.GEMM_TF32X3, 8
for (m, n, k) in [(512, 1024, 512), (128, 128, 128), (1024, 128, 128), (8192, 128, 128), (7168, 1024, 512), (3584, 1024, 512), (4096, 1024, 4096), (1024, 1024, 1024), (8192, 8192, 128), (4096, 4096, 128)]:
    for name, dtype, fl in [("f32 x3 mma", torch.float32, X3 | NOTC), ("f32 x3 auto", torch.float32, X3), ("f32 fast auto", torch.float32, 0), ("f64", torch.float64, 0)]:
        lat, thr = run(m, n, k, dtype, fl)
        print(f"{m}x{n}x{k} {name:14s} latency {lat:7.1f} us  back-to-back {thr:7.1f} us  ({2*m*n*k/thr/1e6:.1f} TF/s)", flush=True)
