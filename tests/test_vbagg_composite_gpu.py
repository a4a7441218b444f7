"""dcgp_vbagg_elbo_fwd / dcgp_vbagg_elbo_bwd (the whole VBagg minibatch ELBO and all its gradients behind two C calls)
against the CPU oracle: q(f) of src/variational/variational_strategy.py:14-68, the expected log-likelihood of
src/likelihoods/vbagg_gaussian_likelihood.py:78-111 (literal per-bag block extraction) and the combination rule of
src/mlls/bag_variational_elbo.py:47-69, differentiated by torch autograd in FP64 on the CPU."""
import math

import pytest
import torch
from torch.nn.functional import softplus

from oracle import restate as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _problem(kind, M, n_bags, max_bag, d, seed, mean_const):
    g = torch.Generator().manual_seed(seed)
    sizes = torch.randint(1, max_bag + 1, (n_bags,), generator=g).tolist()
    N = sum(sizes)
    p = {
        "Z": torch.randn(M, d, generator=g, dtype=torch.float64),
        "var_mean": 0.3 * torch.randn(M, generator=g, dtype=torch.float64),
        "chol_var": (torch.eye(M, dtype=torch.float64) + 0.05 * torch.randn(M, M, generator=g, dtype=torch.float64)),
        "raw_noise": torch.tensor([0.3], dtype=torch.float64),
        "raw_ls0": 0.4 * torch.randn(d if kind == "rbf" else 2, generator=g, dtype=torch.float64) + 0.5,
        "raw_os0": torch.tensor(0.2, dtype=torch.float64),
    }
    if kind == "additive":
        p["raw_ls1"] = 0.4 * torch.randn(d - 2, generator=g, dtype=torch.float64) + 0.5
        p["raw_os1"] = torch.tensor(-0.3, dtype=torch.float64)
    if mean_const:
        p["mean_const"] = torch.tensor([0.7], dtype=torch.float64)
    x = torch.randn(N, d, generator=g, dtype=torch.float64)
    z = torch.randn(n_bags, generator=g, dtype=torch.float64)
    return p, x, z, sizes


def _oracle_spec(p, kind, d):
    if kind == "rbf":
        return R.KernelSpec([R.Component("rbf", p["raw_ls0"], p["raw_os0"], None)])
    return R.KernelSpec([R.Component("matern32", p["raw_ls0"], p["raw_os0"], [0, 1]),
                         R.Component("rbf", p["raw_ls1"], p["raw_os1"], list(range(2, d)))])


def _oracle(p, kind, x, z, sizes, num_data, beta):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    d = x.shape[1]
    k = _oracle_spec(q, kind, d)
    mu_x = q["mean_const"].expand(x.shape[0]) if "mean_const" in q else None
    mu_q, Sigma_q = R.svgp_q(k, q["Z"], x, q["var_mean"], q["chol_var"], mu_x=mu_x)
    ell = R.vbagg_ell(z, mu_q, Sigma_q, sizes, q["raw_noise"])
    elbo = R.bag_elbo(ell, R.whitened_kl(q["var_mean"], q["chol_var"]), z.numel(), num_data, beta)
    (-elbo).backward()
    return elbo.detach(), {k_: v.grad for k_, v in q.items()}


def _gpu(p, kind, x, z, sizes, num_data, beta, dtype):
    from deconditional_downscaling_b200 import ops
    d = x.shape[1]
    t = lambda v: v.to(device=DEV, dtype=dtype)
    if kind == "rbf":
        kinds, dims, raws = ["rbf"], [list(range(d))], [("raw_ls0", "raw_os0")]
    else:
        kinds, dims, raws = ["matern32", "rbf"], [[0, 1], list(range(2, d))], [("raw_ls0", "raw_os0"), ("raw_ls1", "raw_os1")]
    spec = ops.FlatSpec(kinds, dims, [t(softplus(p[a])) for a, _ in raws], [t(softplus(p[b])) for _, b in raws])
    offsets = torch.tensor([0] + torch.tensor(sizes).cumsum(0).tolist(), dtype=torch.int64, device=DEV)
    noise = t(softplus(p["raw_noise"]) + 1e-4)
    mc = t(p["mean_const"]) if "mean_const" in p else None
    elbo, g, info = ops.vbagg_elbo(spec, t(p["Z"]), t(x), offsets, t(z), t(p["var_mean"]), t(p["chol_var"]), noise,
                                   num_data, beta=beta, mean_const=mc, scale=-1.0)
    assert info.item() == 0
    # chain rule of the softplus constraints (gradients come back with respect to the VALUES)
    out = {"Z": g["Z"], "var_mean": g["var_mean"], "chol_var": g["chol_var"],
           "raw_noise": g["noise"] * torch.sigmoid(t(p["raw_noise"]))}
    for c, (a, b) in enumerate(raws):
        out[a] = g["dls"][c, :p[a].numel()] * torch.sigmoid(t(p[a]))
        out[b] = g["dos"][c] * torch.sigmoid(t(p[b]))
    if mc is not None:
        out["mean_const"] = g["mean_const"]
    return elbo, out


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 3e-3)])
@pytest.mark.parametrize("kind,M,n_bags,max_bag,d,mean_const", [("rbf", 96, 70, 12, 3, False), ("rbf", 200, 130, 9, 5, True),
                                                               ("additive", 130, 90, 15, 4, True), ("rbf", 5, 3, 2, 1, False)])
def test_elbo_and_every_gradient_match_the_oracle(dtype, tol, kind, M, n_bags, max_bag, d, mean_const):
    p, x, z, sizes = _problem(kind, M, n_bags, max_bag, d, seed=M + d, mean_const=mean_const)
    num_data, beta = 10.0 * len(z), 0.7
    ref, gref = _oracle(p, kind, x, z, sizes, num_data, beta)
    elbo, g = _gpu(p, kind, x, z, sizes, num_data, beta, dtype)
    assert abs(elbo.item() - ref.item()) <= tol * max(1.0, abs(ref.item())), (elbo.item(), ref.item())
    for name, gr in gref.items():
        got = g[name].double().cpu().reshape(gr.shape)
        if name == "chol_var":
            gr = gr.tril()       # the reference masks the factor (CholeskyVariationalDistribution): zero above
            assert got.triu(1).abs().max().item() == 0.0
        err = (got - gr).norm().item() / max(gr.norm().item(), 1e-12)
        assert err <= tol * 5, (name, err)


def test_value_only_call_and_scale():
    """need_grads=False runs the forward alone; scale multiplies every gradient."""
    from deconditional_downscaling_b200 import ops
    p, x, z, sizes = _problem("rbf", 64, 40, 8, 2, seed=3, mean_const=False)
    e1, g1 = _gpu(p, "rbf", x, z, sizes, 400.0, 1.0, torch.float64)
    t = lambda v: v.to(device=DEV, dtype=torch.float64)
    spec = ops.FlatSpec(["rbf"], [[0, 1]], [t(softplus(p["raw_ls0"]))], [t(softplus(p["raw_os0"]))])
    offsets = torch.tensor([0] + torch.tensor(sizes).cumsum(0).tolist(), dtype=torch.int64, device=DEV)
    noise = t(softplus(p["raw_noise"]) + 1e-4)
    e2, none, _ = ops.vbagg_elbo(spec, t(p["Z"]), t(x), offsets, t(z), t(p["var_mean"]), t(p["chol_var"]), noise, 400.0,
                                 need_grads=False)
    assert none is None and e1.item() == e2.item()
    _, g3, _ = ops.vbagg_elbo(spec, t(p["Z"]), t(x), offsets, t(z), t(p["var_mean"]), t(p["chol_var"]), noise, 400.0,
                              scale=2.5)
    assert torch.allclose(g3["var_mean"], -2.5 * g1["var_mean"], rtol=1e-12, atol=1e-14)
    assert torch.allclose(g3["Z"], -2.5 * g1["Z"], rtol=1e-10, atol=1e-13)


def test_failed_factorisation_is_reported_not_raised():
    """Duplicate inducing points and no jitter: the K_ZZ factorisation must fail through `info`, LAPACK style."""
    from deconditional_downscaling_b200 import ops
    p, x, z, sizes = _problem("rbf", 32, 20, 5, 2, seed=5, mean_const=False)
    t = lambda v: v.to(device=DEV, dtype=torch.float64)
    Z = t(p["Z"])
    Z[1] = Z[0]
    spec = ops.FlatSpec(["rbf"], [[0, 1]], [t(softplus(p["raw_ls0"]))], [t(softplus(p["raw_os0"]))])
    offsets = torch.tensor([0] + torch.tensor(sizes).cumsum(0).tolist(), dtype=torch.int64, device=DEV)
    _, _, info = ops.vbagg_elbo(spec, Z, t(x), offsets, t(z), t(p["var_mean"]), t(p["chol_var"]),
                                t(softplus(p["raw_noise"])), 100.0, need_grads=False, jitter=0.0)
    assert info.item() != 0


@pytest.mark.parametrize("mean", ["zero", "constant"])
def test_trainer_takes_the_composite_and_follows_the_layered_path(mean, monkeypatch):
    """ElboTrainer on a VariationalGP + VBaggGaussianLikelihood in FP64: three Adam steps through the C composite
    (default) and through the layered Python path (DCGP_VBAGG_FUSED=0) give the same losses and parameters."""
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as Md
    from deconditional_downscaling_b200 import _lib
    from deconditional_downscaling_b200.training import ElboTrainer
    torch.set_default_dtype(torch.float64)
    try:
        g = torch.Generator().manual_seed(11)
        sizes = torch.randint(2, 9, (60,), generator=g).tolist()
        x = torch.randn(sum(sizes), 3, generator=g, dtype=torch.float64).to(DEV)
        z = torch.randn(60, generator=g, dtype=torch.float64).to(DEV)
        Z0 = torch.randn(48, 3, generator=g, dtype=torch.float64)
        runs = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("DCGP_VBAGG_FUSED", flag)
            torch.manual_seed(123)      # q(u) is initialised with random noise on the first call
            model = Md.VariationalGP(inducing_points=Z0.clone(), mean_module=Mn.ConstantMean() if mean == "constant" else Mn.ZeroMean(),
                                     covar_module=K.ScaleKernel(K.RBFKernel(ard_num_dims=3))).to(DEV).double()
            lik = L.VBaggGaussianLikelihood().to(DEV).double()
            tr = ElboTrainer(model, lik, num_data=600, beta=1.0, lr=0.05, cuda_graph=False)
            l0 = _lib.load().dcgp_launch_count()
            losses = [tr.step(x, z, bags_sizes=sizes).item() for _ in range(3)]
            runs[flag] = (losses, {n: p.detach().clone() for n, p in list(model.named_parameters()) + list(lik.named_parameters())},
                          _lib.load().dcgp_launch_count() - l0)
        for a, b in zip(runs["1"][0], runs["0"][0]):
            assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (runs["1"][0], runs["0"][0])
        for n, p in runs["0"][1].items():
            assert (runs["1"][1][n] - p).abs().max().item() <= 1e-7, n
        assert runs["1"][2] != runs["0"][2]          # the two runs really took different paths
    finally:
        torch.set_default_dtype(torch.float32)


def test_fp32_trainer_with_the_fp64_island_follows_the_layered_path(monkeypatch):
    """FP32 model: the composite (FP64 K_ZZ island inside, default) against the layered path (FP64 island in Python)."""
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as Md
    from deconditional_downscaling_b200.training import ElboTrainer
    g = torch.Generator().manual_seed(12)
    sizes = torch.randint(2, 9, (200,), generator=g).tolist()
    x = torch.randn(sum(sizes), 3, generator=g).to(DEV)
    z = torch.randn(200, generator=g).to(DEV)
    Z0 = torch.randn(160, 3, generator=g)
    runs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DCGP_VBAGG_FUSED", flag)
        torch.manual_seed(321)
        model = Md.VariationalGP(inducing_points=Z0.clone(), mean_module=Mn.ZeroMean(),
                                 covar_module=K.ScaleKernel(K.RBFKernel(ard_num_dims=3))).to(DEV)
        lik = L.VBaggGaussianLikelihood().to(DEV)
        tr = ElboTrainer(model, lik, num_data=2000, beta=1.0, lr=0.02, cuda_graph=False)
        runs[flag] = [tr.step(x, z, bags_sizes=sizes).item() for _ in range(3)]
    for a, b in zip(runs["1"], runs["0"]):
        assert abs(a - b) <= 2e-3 * max(1.0, abs(b)), runs
