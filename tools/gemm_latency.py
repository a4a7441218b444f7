"""Per-launch cost of the tcgen05 GEMM on small problems issued back to back on one stream (what the bulk lane of the
look-ahead factorisation pays per panel): python tools/gemm_latency.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
def timed(fn, reps=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3
g = torch.cuda.CUDAGraph()
for (m, n, k, x3) in ((256, 256, 128, True), (1024, 128, 128, True), (1024, 1024, 128, True), (4096, 4096, 128, True), (8192, 128, 128, True),
                      (1024, 1024, 128, False), (512, 1024, 512, True)):
    A = torch.randn(m, k, device="cuda"); B = torch.randn(n, k, device="cuda"); C = torch.randn(m, n, device="cuda")
    Alo, Blo = ops.split_lo(A), ops.split_lo(B)
    fn = (lambda: ops.gemm_x3(A, B, transb=True, alpha=-1.0, beta=1.0, C=C, Alo=Alo, Blo=Blo)) if x3 else \
         (lambda: ops.gemm(A, B, transb=True, alpha=-1.0, beta=1.0, C=C))
    t_eager = timed(fn)
    gr = torch.cuda.CUDAGraph()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=st):
            for _ in range(50): fn()
    torch.cuda.synchronize()
    t_graph = timed(lambda: gr.replay(), reps=10) / 50
    print(f"{m}x{n}x{k} {'3xTF32' if x3 else 'TF32'} beta=1: {t_eager:.1f} us per call issued eagerly, {t_graph:.1f} us per call inside a graph (50 dependent launches)", flush=True)
