"""Single dcgp_gram_bwd driver for ncu: python tools/one_gram_bwd.py [n] [d]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
d = int(sys.argv[2]) if len(sys.argv) > 2 else 5
x = torch.randn(n, d, device="cuda")
G = torch.randn(n, n, device="cuda")
spec = rbf_spec(d, torch.float32, "cuda")
for _ in range(3):
    ops.gram_bwd(spec, x, None, G)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); ops.gram_bwd(spec, x, None, G); e.record(); torch.cuda.synchronize()
print("gram_bwd ms", s.elapsed_time(e))
