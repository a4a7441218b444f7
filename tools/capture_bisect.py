import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
dev = "cuda:0"
torch.manual_seed(0)
def spd(n, dt):
    x = torch.randn(n, 3, dtype=dt, device=dev)
    return ops.gram(rbf_spec(3, dt, dev), x, None, diag_const=0.5)
cases = {}
for n, dt in ((640, torch.float32), (48, torch.float64), (40, torch.float32), (300, torch.float32), (1100, torch.float32), (2300, torch.float32), (1024, torch.float64)):
    A = spd(n, dt)
    L = A.clone(); ops.potrf_(L); L = L.tril()
    B = torch.randn(n, 40, dtype=dt, device=dev)
    cases[f"potrf {n} {dt}"] = lambda A=A: ops.potrf_(A.clone())
    cases[f"potrs {n} {dt}"] = lambda L=L, B=B: ops.potrs_(L, B.clone())
    cases[f"trsmT {n} {dt}"] = lambda L=L, B=B: ops.trsm_(L, B.clone(), trans=True)
    cases[f"trtri {n} {dt}"] = lambda L=L: ops.trtri(L)
x = torch.randn(640, 3, device=dev); G = torch.randn(640, 640, device=dev)
cases["gram_bwd"] = lambda: ops.gram_bwd(rbf_spec(3, torch.float32, dev), x, None, G)
P = torch.randn(48, 640, device=dev)
cases["gemm_x3 small"] = lambda: ops.gemm_x3(P, P, transb=True)
cases["gemm splitk"] = lambda: ops.gemm(P, P, transb=True)
s = torch.cuda.Stream()
for name, fn in cases.items():
    for _ in range(2): fn()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g, stream=s):
            fn()
        g.replay(); torch.cuda.synchronize()
        print("ok  ", name)
    except Exception as e:
        print("FAIL", name, str(e).splitlines()[0][:100])
        torch.cuda.synchronize()
