"""One fused Gram-times-matrix product (csrc/gram_mm.cu) for ncu / timing: python tools/one_gram_mm.py [N] [m] [d]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = int(sys.argv[2]) if len(sys.argv) > 2 else 328
d = int(sys.argv[3]) if len(sys.argv) > 3 else 3
x = torch.randn(N, d, device="cuda")
W = torch.randn(N, m, device="cuda")
spec = rbf_spec(d, torch.float32, "cuda")
for fused in (True, False):
    ops.FUSED_GRAM_MATMUL = fused
    for _ in range(2):
        out = ops.gram_matmul(spec, x, None, W)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); out = ops.gram_matmul(spec, x, None, W); e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e)
    print(f"{'fused tcgen05' if fused else 'panels + GEMM'}: N={N} m={m} d={d}: {ms:.3f} ms = {2.0*N*N*m/ms/1e9:.1f} TF/s, {N*N/ms/1e6:.0f} G kernel evaluations/s x {-(-m//256) if fused else 1}")
