"""The reference's swiss-roll ExactCMP builder / trainer / evaluation code paths, written as the runner writes
them (`import gpytorch`, `from src.models import ExactCMP`: experiments/swiss_roll/models/exact_cmp.py:10-160,
core/metrics.py:6-32) and executed on libdcgp through the compat namespaces, against the golden vectors the
unmodified reference produced (tests/golden/exact_cmp*.npz)."""
import os
import sys

import numpy as np
import pytest
import torch

import golden_util as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture()
def runner_namespaces():
    from deconditional_downscaling_b200 import compat
    saved = {k: v for k, v in sys.modules.items() if k == "gpytorch" or k.startswith("gpytorch.") or k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    compat.install()
    yield
    compat.uninstall()
    sys.modules.update(saved)


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def test_swiss_roll_exact_cmp_runner_flow(runner_namespaces, tmp_path):
    import gpytorch
    from src.models import ExactCMP
    from src.utils import Registry
    MODELS, TRAINERS = Registry(), Registry()
    g = G.load("exact_cmp_noise")

    @MODELS.register("exact_cmp")
    def build_swiss_roll_exact_cmp(individuals, extended_bags_values, bags_values, aggregate_targets, lbda, use_individuals_noise, **kwargs):
        inv_softplus = lambda x, n: torch.log(torch.exp(x * torch.ones(n)) - 1)  # noqa: E731
        individuals_mean = gpytorch.means.ZeroMean()
        base_individuals_kernel = gpytorch.kernels.RBFKernel(ard_num_dims=3)
        base_individuals_kernel.initialize(raw_lengthscale=inv_softplus(x=1, n=3))
        individuals_kernel = gpytorch.kernels.ScaleKernel(base_individuals_kernel)
        base_bag_kernel = gpytorch.kernels.RBFKernel(ard_num_dims=1)
        base_bag_kernel.initialize(raw_lengthscale=inv_softplus(x=1, n=1))
        bag_kernel = gpytorch.kernels.ScaleKernel(base_bag_kernel)
        return ExactCMP(individuals_mean=individuals_mean, individuals_kernel=individuals_kernel, bag_kernel=bag_kernel,
                        train_individuals=individuals, train_bags=bags_values, train_aggregate_targets=aggregate_targets,
                        extended_train_bags=extended_bags_values, lbda=lbda, use_individuals_noise=use_individuals_noise,
                        likelihood=gpytorch.likelihoods.GaussianLikelihood())

    def predict(model, individuals, **kwargs):
        with torch.no_grad():
            return model.predict(individuals)

    def compute_chunked_nll(model, individuals, targets, chunk_size):
        try:
            nlls = []
            with gpytorch.settings.fast_computations(log_prob=False, covar_root_decomposition=False):
                for X, t in zip(individuals.split(chunk_size), targets.split(chunk_size)):
                    sub = predict(model=model, individuals=X)
                    nlls.append((-sub.log_prob(t).div(chunk_size)).item())
            return float(np.mean(nlls))
        except gpytorch.utils.errors.NotPSDError:
            return np.nan

    @TRAINERS.register("exact_cmp")
    def train(model, lr, n_epochs, dump_dir, **kwargs):
        device = torch.device(DEV)
        model = model.to(device)
        model.train()
        model.likelihood.train()
        optimizer = torch.optim.Adam(params=model.parameters(), lr=lr)
        mll = gpytorch.mlls.ExactMarginalLogLikelihood(model.likelihood, model)
        losses = []
        for epoch in range(n_epochs):
            optimizer.zero_grad()
            output = model(model.train_bags)
            loss = -mll(output, model.train_aggregate_targets)
            loss.backward()
            losses.append(loss.item())
            if kwargs.get("step", True):
                optimizer.step()
            model.update_cmp_estimate_parameters()
        state = {"epoch": n_epochs, "state_dict": model.state_dict(), "optimizer": optimizer.state_dict()}
        torch.save(state, os.path.join(dump_dir, "state.pt"))
        return model, losses

    dt = torch.float64
    model = MODELS["exact_cmp"](individuals=g["x"].to(dt), extended_bags_values=g["ext"].to(dt), bags_values=g["y"].to(dt),
                                aggregate_targets=g["z"].to(dt), lbda=g["lbda"], use_individuals_noise=True).double()
    # same hyper-parameters as the golden run (its raw values are stored under the reference's parameter names)
    sd = model.state_dict()
    for k in g:
        if k.startswith("param__"):
            name = k[len("param__"):].replace("__", ".")
            sd[name] = g[k].to(dt).reshape(sd[name].shape)
    model.load_state_dict(sd)
    model, losses = TRAINERS["exact_cmp"](model, lr=0.01, n_epochs=2, dump_dir=str(tmp_path), step=False)
    assert rel(losses[0], g["e1_loss"]) < 1e-6 and rel(losses[1], g["e2_loss"]) < 1e-6
    # evaluation side (core/metrics.py): posterior summaries and the chunked NLL
    xs = g["xs"].to(device=DEV, dtype=dt)
    post = predict(model, xs)
    assert rel(post.mean, g["post_mean"]) < 1e-6
    lower, upper = post.confidence_region()
    assert torch.all(upper > lower)
    t = torch.linspace(-1, 1, xs.size(0), dtype=dt, device=DEV)
    nll = compute_chunked_nll(model, xs, t, chunk_size=xs.size(0))
    assert rel(-nll * xs.size(0), g["post_logprob"]) < 1e-5
    model._clear_cache()
    # the checkpoint written by the trainer is in the reference's wire format and loads back
    state = torch.load(os.path.join(str(tmp_path), "state.pt"), weights_only=True)
    assert set(state) == {"epoch", "state_dict", "optimizer"}
    assert "likelihood.noise_covar.raw_noise_constraint.lower_bound" in state["state_dict"]
    assert "mean_module.bag_kernel.base_kernel.raw_lengthscale" in state["state_dict"]
    model.load_state_dict(state["state_dict"])


def test_bagged_gp_baseline_matches_dense_formulas(runner_namespaces):
    """src.models.BaggedGP (bagged_gp.py:81-180): aggregate prior and individuals posterior against the dense
    per-bag block formulas evaluated with torch on the same kernel matrix."""
    import gpytorch
    from src.likelihoods import VBaggGaussianLikelihood
    from src.models import BaggedGP
    torch.manual_seed(1)
    dt = torch.float64
    sizes = [7, 12, 5, 9]
    x = torch.randn(sum(sizes), 3, dtype=dt, device=DEV)
    z = torch.randn(len(sizes), dtype=dt, device=DEV)
    kern = gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=3))
    model = BaggedGP(train_individuals=x, bags_sizes=sizes, train_aggregate_targets=z, individuals_mean=gpytorch.means.ZeroMean(),
                     individuals_kernel=kern, likelihood=VBaggGaussianLikelihood()).to(DEV).double()
    K = kern(x).evaluate()
    off = np.cumsum([0] + sizes)
    blocks = [K[i:j, i:j] for i, j in zip(off[:-1], off[1:])]
    prior = model(x, torch.tensor(sizes))
    var_ref = torch.stack([b.sum() / b.numel() for b in blocks])
    assert rel(prior.variance, var_ref.cpu()) < 1e-12
    assert rel(prior.covariance_matrix, torch.diag(var_ref).cpu()) < 1e-12
    with torch.no_grad():
        post = model.predict(x, sizes)
    noise = model.likelihood.noise.double()
    n = torch.tensor(sizes, dtype=dt, device=DEV)
    C = torch.zeros(len(x), len(sizes), dtype=dt, device=DEV)
    for a, (i, j) in enumerate(zip(off[:-1], off[1:])):
        C[i:j, a] = K[i:j, i:j].sum(1) / (j - i)
    tot = var_ref + noise / n
    mean_ref = C @ (z / tot)
    cov_ref = K - (C / tot) @ C.t()
    assert rel(post.mean, mean_ref.cpu()) < 1e-10
    assert rel(post.covariance_matrix, cov_ref.cpu()) < 1e-10
