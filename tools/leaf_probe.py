import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import ops
from deconditional_downscaling_b200.exact import rbf_spec
for dtype in (torch.float32, torch.float64):
    for n in (128, 256, 512, 1024, 2048):
        x = torch.randn(n, 3, dtype=dtype, device="cuda")
        A0 = ops.gram(rbf_spec(3, dtype, "cuda"), x, None, diag_const=0.5)
        A = A0.clone()
        reps = 50
        for _ in range(5):
            A.copy_(A0); ops.potrf_(A)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            A.copy_(A0); ops.potrf_(A)
        e.record(); torch.cuda.synchronize()
        t = s.elapsed_time(e) / reps
        s.record()
        for _ in range(reps):
            A.copy_(A0)
        e.record(); torch.cuda.synchronize()
        print(f"{dtype} n={n}: potrf {1e3*(t - s.elapsed_time(e)/reps):.1f} us")
