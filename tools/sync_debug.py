import os, sys, torch, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M
from deconditional_downscaling_b200.training import ElboTrainer
DEV = "cuda:0"
kind = sys.argv[1] if len(sys.argv) > 1 else "vcmp"
torch.manual_seed(5)
Mz, d, r, Nb, nb = 48, 3, 2, 640, 40
k = K.ScaleKernel(K.RBFKernel(ard_num_dims=d))
if kind == "vcmp":
    model = M.VariationalCMP(inducing_points=torch.randn(Mz, d), individuals_mean=Mn.ZeroMean(), individuals_kernel=k,
                             bag_kernel=K.ScaleKernel(K.RBFKernel(ard_num_dims=r)), lbda=1e-3, use_individuals_noise=True)
    lik = L.CMPLikelihood(use_individuals_noise=True)
    kw = dict(bags_values=torch.randn(nb, r, device=DEV), extended_bags_values=torch.randn(Nb, r, device=DEV))
else:
    model = M.VariationalGP(inducing_points=torch.randn(Mz, d), mean_module=Mn.ZeroMean(), covar_module=k)
    lik = L.VBaggGaussianLikelihood()
    kw = dict(bags_sizes=[Nb // nb] * nb)
model, lik = model.to(DEV), lik.to(DEV)
tr = ElboTrainer(model, lik, num_data=nb, beta=1.0, lr=0.01)
x, z = torch.randn(Nb, d, device=DEV), torch.randn(nb, device=DEV)
for _ in range(2): tr.step(x, z, **kw)
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("error")
try:
    tr._backward_pass(x, z, kw)
    print("no synchronising call in the pass")
except Exception:
    traceback.print_exc()
