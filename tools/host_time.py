"""Host issue time vs device time of the bench step: python tools/host_time.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = bench.WL
dev = torch.device("cuda", 0)
trainer = bench.build_trainer(wl, dev)
shard = bench.make_shard(wl, 0)
host = bench.minibatches(wl, shard, 14, seed=99)
devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
def step(i):
    b = dict(devb[i]); return trainer.step(b.pop("x"), b.pop("z"), **b)
for i in range(4): step(i)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); s.record()
for i in range(4, 14): step(i)
e.record(); t1 = time.perf_counter()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host issue {1e2*(t1-t0):.2f} ms/step, device {s.elapsed_time(e)/10:.2f} ms/step, wall {1e2*(t2-t0):.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(4, 8): step(i)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
