"""Two forward + backward passes of dcgp_cmp_elbo at the bench shape, FP32 (ncu launch lists): python tools/cmp_composite_once.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
N, n, M, d, r = 8192, 1024, 1024, 5, 3
t = lambda v: v.to(device=dev, dtype=torch.float32)
k = ops.FlatSpec(["rbf"], [list(range(d))], [t(torch.full((d,), 0.6931))], [t(torch.full((1,), 0.6931))])
l = ops.FlatSpec(["matern32", "rbf"], [[0, 1], [2]], [t(torch.full((2,), 0.6931)), t(torch.full((1,), 0.6931))],
                 [t(torch.full((1,), 0.6931)), t(torch.full((1,), 0.6931))])
Z, x, ext, y, z = (t(torch.randn(M, d, generator=g)), t(torch.randn(N, d, generator=g)), t(torch.randn(N, r, generator=g)),
                   t(torch.randn(n, r, generator=g)), t(torch.randn(n, generator=g)))
m, Ls = t(torch.zeros(M)), t(torch.eye(M))
noise, ind = t(torch.full((1,), 0.6932)), t(torch.full((1,), 0.6931))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    ops.cmp_elbo(k, l, Z, x, ext, y, z, m, Ls, noise, ind, 1e-4, 10240.0)
torch.cuda.synchronize()
print("ok")
