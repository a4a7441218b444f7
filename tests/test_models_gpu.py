"""GPU parity tests of the reference-facing Python layer (deconditional_downscaling_b200:
kernels / means / likelihoods / mlls / variational / models) against the golden vectors that
the reference's own src/ code produced (tests/golden/*.npz), and against the oracle on other
seeds and sizes.  Every parameter is addressed by the reference's state-dict name, so these
tests also pin the parameter naming.

FP64: 1e-6 relative (north_star).  FP32 (TF32 tensor cores + FP64 island): 1e-3 relative on
values of well-conditioned cases.
"""
import math

import numpy as np
import pytest
import torch

import golden_util as G

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def pkg():
    import deconditional_downscaling_b200 as p
    from deconditional_downscaling_b200 import kernels, likelihoods, means, mlls, models
    return p, kernels, likelihoods, means, mlls, models


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu().reshape(-1)
    b = torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def load_params(module, g, prefix, dtype):
    """Copies golden parameters into `module` by reference state-dict name; returns the names."""
    names = []
    with torch.no_grad():
        for name, p in module.named_parameters():
            key = "param__" + prefix + name.replace(".", "__")
            assert key in g, f"parameter {name} has no counterpart in the reference ({key})"
            p.copy_(g[key].reshape(p.shape).to(p))
            names.append(name)
    ref_keys = {k for k in g if k.startswith("param__" + prefix)}
    assert len(ref_keys) == len(names), f"reference has parameters we lack: {sorted(ref_keys)} vs {names}"
    return names


def check_grads(module, g, prefix, tol, gprefix="grad__"):
    for name, p in module.named_parameters():
        key = gprefix + prefix + name.replace(".", "__")
        ref = g[key]
        got = p.grad if p.grad is not None else torch.zeros_like(p)
        if ref.abs().max() < 1e-14:
            assert got.abs().max().item() < 1e-10, name
        else:
            assert rel(got, ref) < tol, f"{name}: {rel(got, ref)}"


def rbf_scale(K, d, active_dims=None):
    return K.ScaleKernel(K.RBFKernel(ard_num_dims=d, active_dims=active_dims))


def mat_scale(K, d, active_dims=None):
    return K.ScaleKernel(K.MaternKernel(nu=1.5, ard_num_dims=d, active_dims=active_dims))


# ----------------------------------------------------------------------------- ExactCMP
@pytest.mark.parametrize("name", ["exact_cmp", "exact_cmp_noise"])
def test_exact_cmp_fit_and_posterior_match_reference(name):
    _, K, L, Mn, E, M = pkg()
    g = G.load(name)
    noise = name.endswith("noise")
    dt = torch.float64
    to = lambda t: t.to(device=DEV, dtype=dt)
    model = M.ExactCMP(train_individuals=to(g["x"]), extended_train_bags=to(g["ext"]), train_bags=to(g["y"]),
                       train_aggregate_targets=to(g["z"]), individuals_mean=Mn.ZeroMean(),
                       individuals_kernel=rbf_scale(K, 3), bag_kernel=rbf_scale(K, 1), lbda=g["lbda"],
                       likelihood=L.GaussianLikelihood(), use_individuals_noise=noise).to(DEV).double()
    load_params(model, g, "", dt)
    model.train()
    mll = E.ExactMarginalLogLikelihood(model.likelihood, model)
    # epoch 1: Lambda^{-1} carries no gradient (src/models/cmp.py:31)
    out = model(model.train_bags)
    loss = -mll(out, model.train_aggregate_targets)
    loss.backward()
    assert rel(out.mean, g["e1_mean"]) < 1e-6 or g["e1_mean"].abs().max() < 1e-12
    assert rel(out.covariance_matrix, g["e1_covar"]) < 1e-6
    assert rel(loss, g["e1_loss"]) < 1e-6
    check_grads(model, g, "", 1e-6, "e1_grad__")
    # epoch 2: refreshed with gradient (src/models/cmp.py:70-85)
    model.zero_grad()
    model.update_cmp_estimate_parameters()
    out = model(model.train_bags)
    loss = -mll(out, model.train_aggregate_targets)
    loss.backward()
    assert rel(out.covariance_matrix, g["e2_covar"]) < 1e-6
    assert rel(loss, g["e2_loss"]) < 1e-6
    check_grads(model, g, "", 1e-6, "e2_grad__")
    # posterior on individuals
    model.zero_grad()
    model.update_cmp_estimate_parameters()
    with torch.no_grad():
        post = model.predict(to(g["xs"]))
        assert rel(post.mean, g["post_mean"]) < 1e-6
        assert rel(post.variance, g["post_var"]) < 1e-6
        assert rel(post.covariance_matrix, g["post_covar"]) < 1e-6
        lo, hi = post.confidence_region()
        assert torch.all(hi >= lo)
        lp = post.log_prob(torch.linspace(-1, 1, 17, dtype=dt, device=DEV))
        assert rel(lp, g["post_logprob"]) < 1e-5   # log-det of a near-singular 17x17 posterior covariance
    model._clear_cache()


def test_exact_cmp_float32_tf32_path():
    _, K, L, Mn, E, M = pkg()
    g = G.load("exact_cmp")
    to = lambda t: t.to(device=DEV, dtype=torch.float32)
    model = M.ExactCMP(train_individuals=to(g["x"]), extended_train_bags=to(g["ext"]), train_bags=to(g["y"]),
                       train_aggregate_targets=to(g["z"]), individuals_mean=Mn.ZeroMean(),
                       individuals_kernel=rbf_scale(K, 3), bag_kernel=rbf_scale(K, 1), lbda=g["lbda"],
                       likelihood=L.GaussianLikelihood(), use_individuals_noise=False).to(DEV)
    load_params(model, g, "", torch.float32)
    model.train()
    mll = E.ExactMarginalLogLikelihood(model.likelihood, model)
    out = model(model.train_bags)
    loss = -mll(out, model.train_aggregate_targets)
    loss.backward()
    assert rel(out.covariance_matrix, g["e1_covar"]) < 1e-3
    assert rel(loss, g["e1_loss"]) < 1e-3
    with torch.no_grad():
        post = model.predict(to(g["xs"]))
    assert rel(post.mean, g["post_mean"]) < 1e-3


# ----------------------------------------------------------------------------- VariationalCMP
def build_vcmp(name, g, dtype):
    _, K, L, Mn, E, M = pkg()
    noise = "noise" in name or "rff" in name
    if "rff" in name:
        D = int(g["D"])
        k1 = K.RFFKernel(nu=1.5, num_samples=D, ard_num_dims=2, active_dims=[0, 1])
        k1._init_weights(rand_weights=g["rand_weights"].float())
        k2 = K.GPyTorchRFFKernel(num_samples=D, ard_num_dims=3, active_dims=[2, 3, 4])
        k2._init_weights(rand_weights=g["randn_weights"].float())
        k = K.ScaleKernel(k1) + K.ScaleKernel(k2)
    elif "additive" in name:
        k = mat_scale(K, 2, [0, 1]) + rbf_scale(K, 3, [2, 3, 4])
    else:
        k = rbf_scale(K, 3)
    if "additive" in name or "rff" in name:
        l = mat_scale(K, 2, [0, 1]) + rbf_scale(K, 1, [2])
    else:
        l = rbf_scale(K, 1)
    M_ = int(g["M"])
    d = g["x"].shape[1]
    model = M.VariationalCMP(inducing_points=torch.zeros(M_, d), individuals_mean=Mn.ZeroMean(), individuals_kernel=k,
                             bag_kernel=l, lbda=g["lbda"], use_individuals_noise=noise)
    lik = L.CMPLikelihood(use_individuals_noise=noise)
    model, lik = model.to(DEV).to(dtype), lik.to(DEV).to(dtype)
    if "rff" in name:  # buffers follow the module dtype
        k.kernels[0].base_kernel.rand_weights.copy_(g["rand_weights"].to(DEV).to(dtype))
        k.kernels[1].base_kernel.randn_weights.copy_(g["randn_weights"].to(DEV).to(dtype))
    load_params(model, g, "model__", dtype)
    load_params(lik, g, "lik__", dtype)
    model.variational_strategy.variational_params_initialized.fill_(1)
    return model, lik, E


@pytest.mark.parametrize("name", ["variational_cmp", "variational_cmp_noise", "variational_cmp_noise_additive",
                                  "variational_cmp_additive", "variational_cmp_rff"])
def test_variational_cmp_elbo_matches_reference(name):
    g = G.load(name)
    dt = torch.float64
    to = lambda t: t.to(device=DEV, dtype=dt)
    model, lik, E = build_vcmp(name, g, dt)
    model.train(); lik.train()
    elbo = E.BagVariationalELBO(lik, model, num_data=g["z"].numel(), beta=g["beta"])
    q = model(to(g["x"]))
    assert rel(q.mean, g["q_mean"]) < 1e-6
    assert rel(q.variance, g["q_covar"].diagonal()) < 1e-6
    assert rel(q.covariance_matrix, g["q_covar"]) < 1e-6
    kw = model.get_elbo_computation_parameters(bags_values=to(g["y"]), extended_bags_values=to(g["ext"]))
    ell = lik.expected_log_prob(to(g["z"]), q, **kw)
    kl = model.variational_strategy.kl_divergence()
    assert rel(ell, g["ell"]) < 1e-6
    assert rel(kl, g["kl"]) < 1e-9
    loss = -elbo(variational_dist_f=q, target=to(g["z"]), **kw)
    assert rel(loss, g["loss"]) < 1e-6
    loss.backward()
    check_grads(model, g, "model__", 1e-6)
    check_grads(lik, g, "lik__", 1e-6)
    ll, kld, _ = E.BagVariationalELBO(lik, model, num_data=g["z"].numel(), beta=g["beta"], combine_terms=False)(
        model(to(g["x"])), to(g["z"]), **model.get_elbo_computation_parameters(to(g["y"]), to(g["ext"])))
    assert rel(ll - kld, -g["loss"]) < 1e-6


@pytest.mark.parametrize("name", ["variational_cmp_noise", "variational_cmp"])
def test_variational_cmp_float32_tf32_path(name):
    g = G.load(name)
    to = lambda t: t.to(device=DEV, dtype=torch.float32)
    model, lik, E = build_vcmp(name, g, torch.float32)
    model.train(); lik.train()
    elbo = E.BagVariationalELBO(lik, model, num_data=g["z"].numel(), beta=g["beta"])
    q = model(to(g["x"]))
    kw = model.get_elbo_computation_parameters(bags_values=to(g["y"]), extended_bags_values=to(g["ext"]))
    loss = -elbo(variational_dist_f=q, target=to(g["z"]), **kw)
    assert rel(q.mean, g["q_mean"]) < 1e-3
    assert rel(loss, g["loss"]) < 1e-3
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)


# ----------------------------------------------------------------------------- VBagg
def test_vbagg_elbo_matches_reference():
    _, K, L, Mn, E, M = pkg()
    g = G.load("vbagg")
    dt = torch.float64
    to = lambda t: t.to(device=DEV, dtype=dt)
    model = M.VariationalGP(inducing_points=torch.zeros(int(g["M"]), 3), mean_module=Mn.ZeroMean(), covar_module=rbf_scale(K, 3))
    lik = L.VBaggGaussianLikelihood()
    model, lik = model.to(DEV).double(), lik.to(DEV).double()
    load_params(model, g, "model__", dt)
    load_params(lik, g, "lik__", dt)
    model.variational_strategy.variational_params_initialized.fill_(1)
    model.train(); lik.train()
    sizes = [int(s) for s in g["bag_sizes"]]
    elbo = E.BagVariationalELBO(lik, model, num_data=g["z"].numel(), beta=g["beta"])
    q = model(to(g["x"]))
    ell = lik.expected_log_prob(to(g["z"]), q, bags_sizes=sizes)
    assert rel(ell, g["ell"]) < 1e-6
    loss = -elbo(variational_dist_f=q, target=to(g["z"]), bags_sizes=sizes)
    assert rel(loss, g["loss"]) < 1e-6
    loss.backward()
    check_grads(model, g, "model__", 1e-6)
    check_grads(lik, g, "lik__", 1e-6)
    # dense-covariance fallback of the likelihood gives the same number
    from deconditional_downscaling_b200.distributions import MultivariateNormal
    with torch.no_grad():
        q2 = model(to(g["x"]))
        dense = MultivariateNormal(q2.mean, q2.covariance_matrix)
        assert rel(lik.expected_log_prob(to(g["z"]), dense, bags_sizes=sizes), g["ell"]) < 1e-6
    # eval-mode posterior on fresh points (predict path of the trainers)
    model.eval()
    with torch.no_grad():
        qs = model(to(g["xs"]))
        assert rel(qs.mean, g["eval_mean"]) < 1e-6
        assert rel(qs.covariance_matrix, g["eval_covar"]) < 1e-6
        assert rel((2.0 * qs.lazy_covariance_matrix).cpu().evaluate(), 2.0 * g["eval_covar"]) < 1e-6


def test_first_call_initialises_variational_distribution():
    """first call sets variational_mean ~ 1e-3 * randn, chol = I."""
    _, K, L, Mn, E, M = pkg()
    model = M.VariationalGP(inducing_points=torch.randn(16, 3), mean_module=Mn.ZeroMean(), covar_module=rbf_scale(K, 3)).to(DEV)
    vd = model.variational_strategy._variational_distribution
    assert vd.variational_mean.abs().max().item() == 0
    model(torch.randn(20, 3, device=DEV))
    assert 0 < vd.variational_mean.abs().max().item() < 0.02
    assert torch.equal(vd.chol_variational_covar.detach().cpu(), torch.eye(16))
    assert model.variational_strategy.variational_params_initialized.item() == 1


def test_not_psd_raises_reference_compatible_error():
    from deconditional_downscaling_b200 import functional as F
    A = -torch.eye(8, dtype=torch.float64, device=DEV)
    with pytest.raises(F.NotPSDError):
        F.psd_safe_cholesky(A)


def test_deferred_pd_checks_collect_and_raise():
    """Factorisations inside a deferred region do not read their failure word back; check() raises for the
    first failed one, and ElboTrainer-style callers can then repeat the work with immediate checks."""
    from deconditional_downscaling_b200 import functional as F
    good = torch.eye(200, dtype=torch.float64, device=DEV) * 2.0
    bad = good.clone()
    bad[150, 150] = -1.0
    with F.deferred_pd_checks() as pending:
        F.cholesky(good)
        F.cholesky(bad)          # no exception here
        F.cholesky(good)
    with pytest.raises(F.NotPSDError, match="#1"):
        pending.check()
    with pytest.raises(F.NotPSDError):
        F.cholesky(bad)          # immediate mode unchanged
    with F.deferred_pd_checks() as pending:
        F.cholesky(good)
    pending.check()


def _small_vcmp_trainer(lbda, seed=0):
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M
    from deconditional_downscaling_b200.training import ElboTrainer
    torch.manual_seed(seed)
    Mz, d, r, Nb, nb = 48, 3, 2, 600, 40
    k = K.ScaleKernel(K.RBFKernel(ard_num_dims=d))
    l = K.ScaleKernel(K.RBFKernel(ard_num_dims=r))
    model = M.VariationalCMP(inducing_points=torch.randn(Mz, d), individuals_mean=Mn.ZeroMean(), individuals_kernel=k,
                             bag_kernel=l, lbda=lbda, use_individuals_noise=True)
    lik = L.CMPLikelihood(use_individuals_noise=True)
    model, lik = model.to(DEV), lik.to(DEV)
    tr = ElboTrainer(model, lik, num_data=nb, beta=1.0, lr=0.01)
    batch = dict(x=torch.randn(Nb, d, device=DEV), z=torch.randn(nb, device=DEV), bags_values=torch.randn(nb, r, device=DEV),
                 extended_bags_values=torch.randn(Nb, r, device=DEV))
    return tr, batch


def test_trainer_step_failure_raises_and_leaves_parameters_untouched():
    """A step whose Lambda is not positive definite: the optimistic (deferred-check) pass fails, the eager pass
    retries with jitter and raises NotPSDError like the reference's psd_safe_cholesky; no update is applied."""
    from deconditional_downscaling_b200 import functional as F
    tr, b = _small_vcmp_trainer(lbda=-5.0)      # l(y~, y~) - 5 N I is negative definite: no jitter can repair it
    before = [p.detach().clone() for p in tr.params]
    with pytest.raises(F.NotPSDError):
        tr.step(b["x"], b["z"], bags_values=b["bags_values"], extended_bags_values=b["extended_bags_values"])
    assert all(torch.equal(p.detach(), q) for p, q in zip(tr.params, before))
    assert F._DEFERRED is None                   # the deferred region was left cleanly


def test_side_stream_prefetch_gives_identical_step(monkeypatch):
    """The optional side-stream K_ZZ factorisation (DCGP_SIDE_STREAM=1) only reorders work: loss and updated
    parameters match the single-stream step."""
    from deconditional_downscaling_b200 import variational
    out = []
    for side in (False, True):
        monkeypatch.setattr(variational, "SIDE_STREAM", side)
        tr, b = _small_vcmp_trainer(lbda=1e-3, seed=3)
        losses = [tr.step(b["x"], b["z"], bags_values=b["bags_values"], extended_bags_values=b["extended_bags_values"]).item()
                  for _ in range(3)]
        torch.cuda.synchronize()
        out.append((losses, [p.detach().clone() for p in tr.params]))
    (l0, p0), (l1, p1) = out
    assert all(abs(a - c) <= 1e-5 * max(1.0, abs(a)) for a, c in zip(l0, l1))
    assert all(torch.allclose(a, c, rtol=1e-4, atol=1e-5) for a, c in zip(p0, p1))


@pytest.mark.parametrize("kind", ["vcmp", "vbagg"])
def test_cuda_graph_trainer_matches_eager(kind):
    """ElboTrainer(cuda_graph=True): forward + backward captured once per minibatch shape and replayed; losses and
    parameters follow the eager trainer step by step (fresh minibatch values every step)."""
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M
    from deconditional_downscaling_b200.training import ElboTrainer

    def make(graph):
        torch.manual_seed(5)
        Mz, d, r, Nb, nb = 48, 3, 2, 640, 40
        k = K.ScaleKernel(K.RBFKernel(ard_num_dims=d))
        if kind == "vcmp":
            model = M.VariationalCMP(inducing_points=torch.randn(Mz, d), individuals_mean=Mn.ZeroMean(), individuals_kernel=k,
                                     bag_kernel=K.ScaleKernel(K.RBFKernel(ard_num_dims=r)), lbda=1e-3, use_individuals_noise=True)
            lik = L.CMPLikelihood(use_individuals_noise=True)
        else:
            model = M.VariationalGP(inducing_points=torch.randn(Mz, d), mean_module=Mn.ZeroMean(), covar_module=k)
            lik = L.VBaggGaussianLikelihood()
        model, lik = model.to(DEV), lik.to(DEV)
        tr = ElboTrainer(model, lik, num_data=nb, beta=1.0, lr=0.01, cuda_graph=graph)
        g = torch.Generator().manual_seed(9)
        batches = []
        for _ in range(7):
            b = dict(x=torch.randn(Nb, d, generator=g).to(DEV), z=torch.randn(nb, generator=g).to(DEV))
            if kind == "vcmp":
                b.update(bags_values=torch.randn(nb, r, generator=g).to(DEV), extended_bags_values=torch.randn(Nb, r, generator=g).to(DEV))
            else:
                b.update(bags_sizes=[Nb // nb] * nb)
            batches.append(b)
        losses = []
        for b in batches:
            b = dict(b)
            losses.append(tr.step(b.pop("x"), b.pop("z"), **b).item())
        return tr, losses

    te, le = make(False)
    tg, lg = make(True)
    assert len(tg._graphs) == 1 and tg.launches_replayed > 0          # steps 3.. were replays
    assert all(abs(a - b) <= 2e-4 * max(1.0, abs(a)) for a, b in zip(le, lg)), (le, lg)
    for p, q in zip(te.params, tg.params):
        assert torch.allclose(p, q, rtol=2e-3, atol=2e-4)


# ----------------------------------------------------------------------------- BaggedGP baseline
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-6), (torch.float32, 1e-3)])
def test_bagged_gp_matches_reference_golden(dtype, tol):
    """src/models/bagged_gp.py + prediction_strategies/bagged_gp_prediction_strategy.py driven as the swiss-roll trainer
    drives them (experiments/swiss_roll/models/bagged_gp.py:52-79): aggregate prior, exact MLL, every gradient and the
    posterior on the training individuals against vectors produced by the reference's own code
    (tests/golden/make_golden.py::golden_bagged_gp)."""
    _, K, L, Mn, E, M = pkg()
    g = G.load("bagged_gp")
    to = lambda t: t.to(device=DEV, dtype=dtype)
    sizes = [int(s) for s in g["bag_sizes"]]
    model = M.BaggedGP(train_individuals=to(g["x"]), bags_sizes=sizes, train_aggregate_targets=to(g["z"]),
                       individuals_mean=Mn.ZeroMean(), individuals_kernel=rbf_scale(K, 3),
                       likelihood=L.VBaggGaussianLikelihood()).to(DEV).to(dtype)
    load_params(model, g, "model__", dtype)
    model.train(); model.likelihood.train()
    mll = E.ExactMarginalLogLikelihood(model.likelihood, model)
    out = model(model.train_individuals, model.bags_sizes)
    assert rel(out.mean, g["prior_mean"]) < tol or g["prior_mean"].abs().max() == 0
    assert rel(out.covariance_matrix.diagonal(), g["prior_var"]) < tol
    loss = -mll(out, model.train_aggregate_targets, model.bags_sizes)
    assert rel(loss, g["loss"]) < tol
    loss.backward()
    check_grads(model, g, "model__", tol * 5 if dtype == torch.float32 else tol)
    model.eval(); model.likelihood.eval()
    with torch.no_grad():
        post = model.predict(to(g["x"]), sizes)
    assert rel(post.mean, g["post_mean"]) < tol
    assert rel(post.covariance_matrix, g["post_covar"]) < tol
