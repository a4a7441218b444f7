"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/src) over the restated gpytorch (oracle/gpytorch_min).
Run in the build container only:  python tests/golden/make_golden.py

All cases run in float64 (torch default dtype switched) so the GPU FP64 path
can be held to 1e-6 relative; the TF32 path is held to 1e-3 against the same
vectors.  Inputs, every parameter value and every output/gradient are stored,
so the tests need neither the reference nor this script at run time.
"""
import os
import sys
import math
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference_src  # noqa: E402

load_reference_src()
import gpytorch  # noqa: E402  (the shim)
from src.models import BaggedGP, ExactCMP, VariationalCMP, VariationalGP  # noqa: E402
from src.likelihoods import CMPLikelihood, VBaggGaussianLikelihood  # noqa: E402
from src.mlls import BagVariationalELBO  # noqa: E402
from src.kernels import RFFKernel  # noqa: E402

torch.set_default_dtype(torch.float64)


def npy(t):
    return t.detach().cpu().numpy().copy()


def save(name, **arrays):
    out = {k: (npy(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()}
    np.savez(os.path.join(HERE, name + ".npz"), **out)
    print(f"wrote {name}.npz: {sorted(out)}")


def scale_rbf(d, ls, os_, active_dims=None):
    k = gpytorch.kernels.RBFKernel(ard_num_dims=d, active_dims=active_dims)
    k.initialize(raw_lengthscale=torch.as_tensor(ls))
    k = gpytorch.kernels.ScaleKernel(k)
    k.initialize(raw_outputscale=torch.tensor(os_))
    return k


def scale_matern(d, ls, os_, active_dims=None):
    k = gpytorch.kernels.MaternKernel(nu=1.5, ard_num_dims=d, active_dims=active_dims)
    k.initialize(raw_lengthscale=torch.as_tensor(ls))
    k = gpytorch.kernels.ScaleKernel(k)
    k.initialize(raw_outputscale=torch.tensor(os_))
    return k


def grads_of(named):
    return {"grad__" + n.replace(".", "__"): (p.grad if p.grad is not None else torch.zeros_like(p)) for n, p in named}


def params_of(named):
    return {"param__" + n.replace(".", "__"): p for n, p in named}


# ---------------------------------------------------------------- gram
def golden_gram():
    g = torch.Generator().manual_seed(1)
    x1 = torch.randn(37, 3, generator=g)
    x2 = torch.randn(23, 3, generator=g)
    k_rbf = scale_rbf(3, [0.3, -0.2, 0.9], 0.4)
    k_mat = scale_matern(2, [0.1, 0.7], -0.3, active_dims=[0, 1])
    k_add = scale_matern(2, [0.1, 0.7], -0.3, active_dims=[0, 1]) + scale_rbf(1, [0.5], 0.2, active_dims=[2])
    with torch.no_grad():
        save("gram",
             x1=x1, x2=x2,
             rbf_raw_lengthscale=torch.tensor([0.3, -0.2, 0.9]), rbf_raw_outputscale=torch.tensor(0.4),
             mat_raw_lengthscale=torch.tensor([0.1, 0.7]), mat_raw_outputscale=torch.tensor(-0.3),
             add_rbf_raw_lengthscale=torch.tensor([0.5]), add_rbf_raw_outputscale=torch.tensor(0.2),
             rbf_x1x2=k_rbf(x1, x2).evaluate(), rbf_x1x1=k_rbf(x1).evaluate(),
             mat_x1x2=k_mat(x1, x2).evaluate(), mat_x1x1=k_mat(x1).evaluate(),
             add_x1x2=k_add(x1, x2).evaluate(), add_x1x1=k_add(x1).evaluate())


# ---------------------------------------------------------------- exact CMP
def toy_bags(seed, bag_sizes, d, r):
    g = torch.Generator().manual_seed(seed)
    N, n = sum(bag_sizes), len(bag_sizes)
    x = torch.randn(N, d, generator=g)
    y = torch.randn(n, r, generator=g).squeeze(-1) if r == 1 else torch.randn(n, r, generator=g)
    ext = torch.cat([y[i:i + 1].expand(b, *y.shape[1:]) for i, b in enumerate(bag_sizes)])
    if r > 1:  # continuous extended bag values (not constant within bag) for the r>1 case
        ext = ext + 0.3 * torch.randn(N, r, generator=g)
    z = torch.randn(n, generator=g)
    return x, ext, y, z


def golden_exact_cmp(noise):
    bag_sizes = [9, 14, 5, 11, 13, 8]
    x, ext, y, z = toy_bags(2, bag_sizes, d=3, r=1)
    k = scale_rbf(3, [0.6, 0.2, -0.1], 0.3)
    l = scale_rbf(1, [0.4], -0.2)
    lik = gpytorch.likelihoods.GaussianLikelihood()
    lik.initialize(**{"noise_covar.raw_noise": torch.tensor([-1.0])})
    model = ExactCMP(train_individuals=x, extended_train_bags=ext, train_bags=y, train_aggregate_targets=z,
                     individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=k, bag_kernel=l,
                     lbda=0.01, likelihood=lik, use_individuals_noise=noise)
    if noise:
        model.noise_kernel.initialize(raw_outputscale=torch.tensor(-2.0))
    model.train()
    lik.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    named = list(model.named_parameters())
    out = {"x": x, "ext": ext, "y": y, "z": z, "lbda": 0.01, "bag_sizes": np.array(bag_sizes)}
    out.update(params_of(named))
    # epoch 1: R was built under no_grad in __init__ (src/models/cmp.py:31)
    prior = model(model.train_bags)
    loss = -mll(prior, model.train_aggregate_targets)
    loss.backward()
    out.update(e1_mean=prior.mean, e1_covar=prior.covariance_matrix, e1_loss=loss)
    out.update({"e1_" + a: b for a, b in grads_of(named).items()})
    # epoch 2 semantics: refresh with grad enabled (src/models/cmp.py:70-85)
    for _, p in named:
        p.grad = None
    model.update_cmp_estimate_parameters()
    prior = model(model.train_bags)
    loss = -mll(prior, model.train_aggregate_targets)
    loss.backward()
    out.update(e2_mean=prior.mean, e2_covar=prior.covariance_matrix, e2_loss=loss)
    out.update({"e2_" + a: b for a, b in grads_of(named).items()})
    # posterior on individuals (src/models/exact_cmp.py:126-155)
    g = torch.Generator().manual_seed(3)
    xs = torch.randn(17, 3, generator=g)
    with torch.no_grad():
        post = model.predict(xs)
        out.update(xs=xs, post_mean=post.mean, post_covar=post.covariance_matrix, post_var=post.variance,
                   post_logprob=post.log_prob(torch.linspace(-1, 1, 17)))
    save("exact_cmp_noise" if noise else "exact_cmp", **out)


# ---------------------------------------------------------------- variational CMP
def fix_qu(model, M, seed):
    g = torch.Generator().manual_seed(seed)
    vd = model.variational_strategy._variational_distribution
    vd.variational_mean.data.copy_(0.3 * torch.randn(M, generator=g))
    Ls = torch.eye(M) + 0.2 * torch.randn(M, M, generator=g).tril()
    # leave garbage above the diagonal: the lower mask must hide it
    vd.chol_variational_covar.data.copy_(Ls + torch.randn(M, M, generator=g).triu(1))
    model.variational_strategy.variational_params_initialized.fill_(1)


def golden_variational_cmp(noise, additive):
    bag_sizes = [7, 12, 4, 9, 10, 6, 8]
    M = 9
    if additive:
        d, r = 5, 3
        x, ext, y, z = toy_bags(4, bag_sizes, d=d, r=r)
        k = scale_matern(2, [0.3, 0.5], 0.1, active_dims=[0, 1]) + scale_rbf(3, [0.2, 0.4, 0.6], -0.4, active_dims=[2, 3, 4])
        l = scale_matern(2, [0.1, 0.7], -0.3, active_dims=[0, 1]) + scale_rbf(1, [0.5], 0.2, active_dims=[2])
        lbda = 1e-3
    else:
        d, r = 3, 1
        x, ext, y, z = toy_bags(5, bag_sizes, d=d, r=r)
        k = scale_rbf(3, [0.6, 0.2, -0.1], 0.3)
        l = scale_rbf(1, [0.4], -0.2)
        lbda = 1e-2
    # D2 bags are other bags than the D1 ones (unmatched setting): take 5 fresh bag values
    g = torch.Generator().manual_seed(6)
    y2 = torch.randn(5, r, generator=g).squeeze(-1) if r == 1 else torch.randn(5, r, generator=g)
    z2 = torch.randn(5, generator=g)
    Z = torch.randn(M, d, generator=g)
    model = VariationalCMP(inducing_points=Z, individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=k,
                           bag_kernel=l, lbda=lbda, use_individuals_noise=noise)
    if noise:
        model.noise_kernel.initialize(raw_outputscale=torch.tensor(-1.5))
    fix_qu(model, M, 7)
    lik = CMPLikelihood(use_individuals_noise=noise)
    lik.initialize(**{"noise_covar.raw_noise": torch.tensor([-0.5])})
    model.train()
    lik.train()
    elbo = BagVariationalELBO(lik, model, num_data=len(z2), beta=0.7)
    named = [("model." + n, p) for n, p in model.named_parameters()] + [("lik." + n, p) for n, p in lik.named_parameters()]
    q = model(x)
    kw = model.get_elbo_computation_parameters(bags_values=y2, extended_bags_values=ext)
    ell = lik.expected_log_prob(z2, q, **kw)
    kl = model.variational_strategy.kl_divergence()
    loss = -elbo(variational_dist_f=q, target=z2, **kw)
    loss.backward()
    out = {"x": x, "ext": ext, "y": y2, "z": z2, "lbda": lbda, "beta": 0.7, "M": M,
           "q_mean": q.mean, "q_covar": q.covariance_matrix, "ell": ell, "kl": kl, "loss": loss}
    out.update(params_of(named))
    out.update(grads_of(named))
    name = "variational_cmp" + ("_noise" if noise else "") + ("_additive" if additive else "")
    save(name, **out)


def golden_variational_cmp_rff():
    """downscaling-style individuals kernel: Scale(RFF-Matern1.5 dims[0,1]) + Scale(RFF-RBF dims[2,3,4]),
    experiments/downscaling/models/variational_cmp.py:36-55 with D=16 instead of 1000."""
    bag_sizes = [7, 12, 4, 9, 10, 6, 8]
    M, d, r, D = 9, 5, 3, 16
    x, ext, y, z = toy_bags(8, bag_sizes, d=d, r=r)
    torch.manual_seed(9)
    k1 = RFFKernel(nu=1.5, num_samples=D, ard_num_dims=2, active_dims=[0, 1])
    k1.initialize(raw_lengthscale=torch.tensor([0.3, 0.5]))
    k2 = gpytorch.kernels.RFFKernel(num_samples=D, ard_num_dims=3, active_dims=[2, 3, 4])
    k2.initialize(raw_lengthscale=torch.tensor([0.2, 0.4, 0.6]))
    s1, s2 = gpytorch.kernels.ScaleKernel(k1), gpytorch.kernels.ScaleKernel(k2)
    s1.initialize(raw_outputscale=torch.tensor(0.1))
    s2.initialize(raw_outputscale=torch.tensor(-0.4))
    k = s1 + s2
    l = scale_matern(2, [0.1, 0.7], -0.3, active_dims=[0, 1]) + scale_rbf(1, [0.5], 0.2, active_dims=[2])
    g = torch.Generator().manual_seed(10)
    y2, z2, Z = torch.randn(5, r, generator=g), torch.randn(5, generator=g), torch.randn(M, d, generator=g)
    model = VariationalCMP(inducing_points=Z, individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=k,
                           bag_kernel=l, lbda=1e-3, use_individuals_noise=True)
    model.noise_kernel.initialize(raw_outputscale=torch.tensor(-1.5))
    fix_qu(model, M, 11)
    lik = CMPLikelihood(use_individuals_noise=True)
    lik.initialize(**{"noise_covar.raw_noise": torch.tensor([-0.5])})
    model.train()
    lik.train()
    elbo = BagVariationalELBO(lik, model, num_data=len(z2), beta=1.0)
    named = [("model." + n, p) for n, p in model.named_parameters()] + [("lik." + n, p) for n, p in lik.named_parameters()]
    q = model(x)
    kw = model.get_elbo_computation_parameters(bags_values=y2, extended_bags_values=ext)
    ell = lik.expected_log_prob(z2, q, **kw)
    kl = model.variational_strategy.kl_divergence()
    loss = -elbo(variational_dist_f=q, target=z2, **kw)
    loss.backward()
    out = {"x": x, "ext": ext, "y": y2, "z": z2, "lbda": 1e-3, "beta": 1.0, "M": M, "D": D,
           "rand_weights": k1.rand_weights, "randn_weights": k2.randn_weights,
           "q_mean": q.mean, "q_covar": q.covariance_matrix, "ell": ell, "kl": kl, "loss": loss}
    out.update(params_of(named))
    out.update(grads_of(named))
    save("variational_cmp_rff", **out)


# ---------------------------------------------------------------- VBagg
def golden_vbagg():
    bag_sizes = [5, 9, 1, 12, 13]
    M, d = 8, 3
    x, _, _, z = toy_bags(12, bag_sizes, d=d, r=1)
    g = torch.Generator().manual_seed(13)
    Z = torch.randn(M, d, generator=g)
    k = scale_rbf(3, [0.6, 0.2, -0.1], 0.3)
    model = VariationalGP(inducing_points=Z, mean_module=gpytorch.means.ZeroMean(), covar_module=k)
    fix_qu(model, M, 14)
    lik = VBaggGaussianLikelihood()
    lik.initialize(**{"noise_covar.raw_noise": torch.tensor([-0.5])})
    model.train()
    lik.train()
    elbo = BagVariationalELBO(lik, model, num_data=len(z), beta=1.0)
    named = [("model." + n, p) for n, p in model.named_parameters()] + [("lik." + n, p) for n, p in lik.named_parameters()]
    q = model(x)
    ell = lik.expected_log_prob(z, q, bags_sizes=bag_sizes)
    kl = model.variational_strategy.kl_divergence()
    loss = -elbo(variational_dist_f=q, target=z, bags_sizes=bag_sizes)
    loss.backward()
    out = {"x": x, "z": z, "bag_sizes": np.array(bag_sizes), "beta": 1.0, "M": M,
           "q_mean": q.mean, "q_covar": q.covariance_matrix, "ell": ell, "kl": kl, "loss": loss}
    out.update(params_of(named))
    out.update(grads_of(named))
    # eval-mode posterior on fresh points (predict path of the trainers)
    model.eval()
    xs = torch.randn(11, d, generator=g)
    with torch.no_grad():
        qs = model(xs)
        out.update(xs=xs, eval_mean=qs.mean, eval_covar=qs.covariance_matrix)
    save("vbagg", **out)


def golden_bagged_gp():
    """BaggedGP baseline (src/models/bagged_gp.py, prediction_strategies/bagged_gp_prediction_strategy.py) driven as
    experiments/swiss_roll/models/bagged_gp.py:52-79 drives it: aggregate prior, exact MLL with the bag-size scaled
    noise of VBaggGaussianLikelihood.marginal, every gradient, posterior on the training individuals."""
    bag_sizes = [5, 9, 1, 12, 13, 4]
    d = 3
    x, _, _, z = toy_bags(21, bag_sizes, d=d, r=1)
    k = scale_rbf(3, [0.4, -0.2, 0.1], 0.25)
    lik = VBaggGaussianLikelihood()
    lik.initialize(**{"noise_covar.raw_noise": torch.tensor([-0.3])})
    model = BaggedGP(train_individuals=x, bags_sizes=bag_sizes, train_aggregate_targets=z,
                     individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=k, likelihood=lik)
    model.train()
    lik.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(lik, model)
    named = [("model." + n, p) for n, p in model.named_parameters()]
    output = model(model.train_individuals, model.bags_sizes)
    loss = -mll(output, model.train_aggregate_targets, model.bags_sizes)
    loss.backward()
    out = {"x": x, "z": z, "bag_sizes": np.array(bag_sizes), "prior_mean": output.mean,
           "prior_var": output.covariance_matrix.diagonal(), "loss": loss}
    out.update(params_of(named))
    out.update(grads_of(named))
    model.eval()
    lik.eval()
    with torch.no_grad():
        post = model.predict(x, bag_sizes)
        out.update(post_mean=post.mean, post_covar=post.covariance_matrix)
    save("bagged_gp", **out)


if __name__ == "__main__":
    golden_gram()
    golden_exact_cmp(noise=False)
    golden_exact_cmp(noise=True)
    golden_variational_cmp(noise=False, additive=False)
    golden_variational_cmp(noise=True, additive=False)
    golden_variational_cmp(noise=True, additive=True)
    golden_variational_cmp(noise=False, additive=True)
    golden_variational_cmp_rff()
    golden_vbagg()
    golden_bagged_gp()
