"""Per-step wall times of the end-to-end arm (pinned host -> device, step, loss read-back)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
wl = bench.WL
dev = torch.device("cuda", 0)
trainer = bench.build_trainer(wl, dev)
shard = bench.make_shard(wl, 0)
host = bench.minibatches(wl, shard, 16, seed=99)
host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
for i in range(3): trainer.step_from_host(host[i], dev)
torch.cuda.synchronize()
ts = []
for i in range(3, 16):
    t0 = time.perf_counter(); trainer.step_from_host(host[i], dev); ts.append((time.perf_counter() - t0) * 1e3)
print("e2e step ms:", " ".join(f"{t:.1f}" for t in ts))
devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
ts = []
for i in range(3, 16):
    b = dict(devb[i]); t0 = time.perf_counter(); l = trainer.step(b.pop("x"), b.pop("z"), **b); l.item(); ts.append((time.perf_counter() - t0) * 1e3)
print("dev step+item ms:", " ".join(f"{t:.1f}" for t in ts))
