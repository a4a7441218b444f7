import torch
C = torch.randn(8192, 8192, device="cuda")
L = torch.randn(8192, 8192, device="cuda").tril()
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
ms = t(lambda: C.mul_(1.0001))
print(f"in-place scale 8192^2 f32: {ms*1e3:.1f} us = {2*C.numel()*4/ms/1e9:.2f} TB/s")
D = torch.empty_like(C)
ms = t(lambda: D.copy_(C))
print(f"copy 8192^2 f32: {ms*1e3:.1f} us = {2*C.numel()*4/ms/1e9:.2f} TB/s")
