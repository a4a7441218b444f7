"""Prototype: capture one ELBO step (loss + backward + all-reduce-free + Adam) in a CUDA graph and replay it."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from deconditional_downscaling_b200 import functional as F
wl = bench.WL
dev = torch.device("cuda", 0)
trainer = bench.build_trainer(wl, dev)
trainer.opt = torch.optim.Adam(trainer.params, lr=wl["lr"], capturable=True)
shard = bench.make_shard(wl, 0)
host = bench.minibatches(wl, shard, 16, seed=99)
devb = [{k: v.to(dev) for k, v in b.items()} for b in host]
static = {k: v.clone() for k, v in devb[0].items()}
def body():
    trainer.grads.zero()
    with F.deferred_pd_checks() as pending:
        b = dict(static)
        loss = trainer.loss(b.pop("x"), b.pop("z"), **b)
        loss.backward()
    trainer.grads.allreduce_mean()
    trainer.opt.step()
    return loss.detach(), pending.infos
# warm-up on a side stream (torch recipe)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        loss, infos = body()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    gl, ginfos = body()
torch.cuda.synchronize()
print("captured; infos", len(ginfos))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
losses = []
t0 = time.perf_counter(); ev0.record()
for i in range(1, 11):
    for k in static: static[k].copy_(devb[i][k])
    g.replay()
    losses.append(gl.clone())
ev1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"graph replay: device {ev0.elapsed_time(ev1)/10:.3f} ms/step, host {(t1-t0)*100:.3f} ms/step")
print("losses", [round(float(l), 5) for l in losses][:5], "infos", torch.cat([i.reshape(1) for i in ginfos]).tolist())
