"""Low-rank Gram adjoint (G = -X A^T never stored) against product + dense adjoint at the ELBO-step shape."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
dev = "cuda"
N, n = 8192, 1024
for kinds, dims in ((["matern32", "rbf"], [[0, 1], [2]]), (["rbf"], [[0, 1, 2, 3, 4]])):
    d = max(max(dm) for dm in dims) + 1
    x = torch.randn(N, d, device=dev)
    spec = ops.FlatSpec(kinds, dims, [torch.ones(len(dm), device=dev) for dm in dims], [torch.ones(1, device=dev) for _ in dims])
    P, Q = ops.round_tf32(torch.randn(N, n, device=dev)), ops.round_tf32(torch.randn(N, n, device=dev))
    def fused():
        return ops.gram_bwd_lowrank(spec, x, None, P, Q, alpha=-1.0, fused=True)
    def dense():
        G = ops.gemm(P, Q, transb=True, alpha=-1.0)
        return ops.gram_bwd(spec, x, None, G)
    for name, f in (("fused", fused), ("product + dense adjoint", dense)):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); out = f(); e.record(); torch.cuda.synchronize()
            best = min(best, s.elapsed_time(e))
        print(kinds, name, f"{best * 1e3:.1f} us", out[0][0, :2].tolist())
