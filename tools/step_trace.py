"""One bench step with every dcgp op call listed (shape, dtype, ms): python tools/step_trace.py
Also wraps gram / gram_bwd / bag ops so the whole step is covered; prints the uncovered remainder."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deconditional_downscaling_b200 import ops

wl = bench.WL
dev = torch.device("cuda", 0)
trainer = bench.build_trainer(wl, dev)
shard = bench.make_shard(wl, 0)
host = bench.minibatches(wl, shard, 5, seed=99)
devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
rec, on = [], [False]
names = [n for n in ("gemm", "potrf_", "trsm_", "potrs_", "gram", "gram_bwd", "gram_bagsum", "gram_bagsum_bwd",
                     "gram_bag_blocksum", "gram_bag_blocksum_bwd", "bag_sum", "whitened_kl", "logdet_chol", "gemm_x3", "trtri", "split_lo", "solve_plan") if hasattr(ops, n)]
for name in names:
    orig = getattr(ops, name)
    def wrapped(*a, _o=orig, _n=name, **k):
        if not on[0]:
            return _o(*a, **k)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = _o(*a, **k); e.record()
        shp = " ".join(f"{tuple(t.shape)}:{str(t.dtype)[6:]}" for t in a if torch.is_tensor(t))
        kw = " ".join(f"{kk}={vv}" for kk, vv in k.items() if not torch.is_tensor(vv) and vv not in (None, False))
        rec.append((_n, shp + " " + kw, s, e))
        return out
    setattr(ops, name, wrapped)

def step(i):
    b = dict(devb[i]); return trainer.step(b.pop("x"), b.pop("z"), **b)
for i in range(3): step(i)
torch.cuda.synchronize()
on[0] = True
s0, e0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record(); step(3); e0.record(); torch.cuda.synchronize()
tot = 0.0
for n, d, s, e in rec:
    t = s.elapsed_time(e); tot += t
    print(f"{t:8.3f} ms  {n:22s} {d}")
print(f"step {s0.elapsed_time(e0):.3f} ms, covered {tot:.3f} ms, calls {len(rec)}")
