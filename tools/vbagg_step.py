"""VBagg ELBO step at BASELINE config 5 scale: 1024 bags x ~102 individuals per step, M = 1024, d = 5 (FP32)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M, ops
from deconditional_downscaling_b200.training import ElboTrainer
dev = torch.device("cuda", 0)
torch.manual_seed(0)
nb, per, Mz, d = 1024, 102, 1024, 5
sizes = [per] * nb
N = nb * per
km = K.ScaleKernel(K.RBFKernel(ard_num_dims=d))
model = M.VariationalGP(inducing_points=torch.randn(Mz, d), mean_module=Mn.ZeroMean(), covar_module=km)
lik = L.VBaggGaussianLikelihood()
model, lik = model.to(dev), lik.to(dev)
tr = ElboTrainer(model, lik, num_data=10240, beta=1.0, lr=0.01)
batches = [(torch.randn(N, d, device=dev), torch.randn(nb, device=dev)) for _ in range(6)]
rec, on = [], [False]
for name in ("gemm", "potrf_", "trsm_", "potrs_", "gram", "gram_bwd", "gram_bagsum", "gram_bagsum_bwd", "gram_bag_blocksum",
             "gram_bag_blocksum_bwd", "bag_sum", "whitened_kl", "gemm_x3", "trtri", "split_lo", "vbagg_ell"):
    orig = getattr(ops, name)
    def wrapped(*a, _o=orig, _n=name, **k):
        if not on[0]: return _o(*a, **k)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); out = _o(*a, **k); e.record()
        rec.append((_n, " ".join(str(tuple(t.shape)) for t in a if torch.is_tensor(t)), s, e)); return out
    setattr(ops, name, wrapped)
for i in range(3): tr.step(batches[i][0], batches[i][1], bags_sizes=sizes)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for i in range(3, 6): tr.step(batches[i][0], batches[i][1], bags_sizes=sizes)
e.record(); torch.cuda.synchronize()
print(f"vbagg step {s.elapsed_time(e)/3:.3f} ms  ({3e3/s.elapsed_time(e):.1f} steps/s)")
on[0] = True
s.record(); tr.step(batches[0][0], batches[0][1], bags_sizes=sizes); e.record(); torch.cuda.synchronize()
tot = 0
for n, d_, a, b in rec:
    t = a.elapsed_time(b); tot += t
    print(f"{t:8.3f} ms {n:22s} {d_}")
print(f"traced step {s.elapsed_time(e):.3f} ms covered {tot:.3f}")
