import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M, functional as F
from deconditional_downscaling_b200.training import ElboTrainer
dev = torch.device("cuda", 0)
torch.manual_seed(0)
nb, per, Mz, d = 1024, 102, 1024, 5
sizes = [per] * nb
model = M.VariationalGP(inducing_points=torch.randn(Mz, d), mean_module=Mn.ZeroMean(), covar_module=K.ScaleKernel(K.RBFKernel(ard_num_dims=d)))
lik = L.VBaggGaussianLikelihood()
model, lik = model.to(dev), lik.to(dev)
tr = ElboTrainer(model, lik, num_data=10240, beta=1.0, lr=0.01)
batches = [(torch.randn(nb * per, d, device=dev), torch.randn(nb, device=dev)) for _ in range(24)]
for i in range(4): tr.step(batches[i][0], batches[i][1], bags_sizes=sizes)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); s.record()
for i in range(4, 24): tr.step(batches[i][0], batches[i][1], bags_sizes=sizes)
e.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"vbagg: wall/step {(t1-t0)*50:.3f} ms  device/step {s.elapsed_time(e)/20:.3f} ms")
# host-only cost: time the issue part with the check disabled
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(4, 8): tr.step(batches[i][0], batches[i][1], bags_sizes=sizes)
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("cumulative").print_stats(8)
