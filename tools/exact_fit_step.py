"""ExactCMP fit step + predict at the reference's swiss-roll size (N = 5000, d = 3, r = 1, n = 50)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, mlls as E, models as M
dev = "cuda:0"
for dt in (torch.float64, torch.float32):
    torch.manual_seed(42)
    N, n = 5000, 50
    x = torch.randn(N, 3, dtype=dt)
    yb = torch.randn(n, 1, dtype=dt)
    ext = yb[torch.arange(N) * n // N] + 0.05 * torch.randn(N, 1, dtype=dt)
    z = torch.randn(n, dtype=dt)
    model = M.ExactCMP(train_individuals=x, extended_train_bags=ext, train_bags=yb, train_aggregate_targets=z,
                       individuals_mean=Mn.ZeroMean(), individuals_kernel=K.ScaleKernel(K.RBFKernel(ard_num_dims=3)),
                       bag_kernel=K.ScaleKernel(K.RBFKernel(ard_num_dims=1)), lbda=0.01, likelihood=L.GaussianLikelihood(),
                       use_individuals_noise=False).to(dev).to(dt)
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    mll = E.ExactMarginalLogLikelihood(model.likelihood, model)
    def step():
        opt.zero_grad()
        loss = -mll(model(model.train_bags), model.train_aggregate_targets)
        loss.backward(); opt.step(); model.update_cmp_estimate_parameters()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): l = step()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    with torch.no_grad():
        xd = x.to(dev); model.eval(); model.predict(xd).mean; torch.cuda.synchronize(); t2 = time.perf_counter()
        p = model.predict(xd); m_, v_ = p.mean, p.variance; torch.cuda.synchronize(); t3 = time.perf_counter()
    print(f"{dt}: fit step {1e2*(t1-t0):.2f} ms, predict(5000) mean+variance {1e3*(t3-t2):.2f} ms, loss {l.item():.5f}")
