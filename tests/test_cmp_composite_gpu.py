"""dcgp_cmp_elbo_fwd / dcgp_cmp_elbo_bwd (the whole VariationalCMP minibatch ELBO with individuals noise and all its
gradients behind two C calls) against the CPU oracle port of the reference step (oracle/port.py::vcmp_loss: explicit root
of Lambda^{-1}, chol(Sigma_q), dense q(f) covariance - the reference's own algorithm,
src/likelihoods/cmp_likelihood.py:41-78), differentiated by torch autograd in FP64."""
import pytest
import torch
from torch.nn.functional import softplus

from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _problem(M, N, n, d, r, seed, additive):
    g = torch.Generator().manual_seed(seed)
    p = port.make_vcmp_params(M, d, r, dtype=torch.float64, seed=seed, additive_bag_kernel=additive)
    p["var_mean"] = 0.2 * torch.randn(M, generator=g, dtype=torch.float64)
    p["chol_var"] = torch.eye(M, dtype=torch.float64) + 0.05 * torch.randn(M, M, generator=g, dtype=torch.float64)
    p["k_raw_ls"] = p["k_raw_ls"] + 0.3 * torch.randn(d, generator=g, dtype=torch.float64)
    p["raw_noise"] = torch.tensor([0.2], dtype=torch.float64)
    p["noise_kernel_raw_os"] = torch.tensor(-0.4, dtype=torch.float64)
    x = torch.randn(N, d, generator=g, dtype=torch.float64)
    ext = torch.randn(N, r, generator=g, dtype=torch.float64)
    y = torch.randn(n, r, generator=g, dtype=torch.float64)
    z = torch.randn(n, generator=g, dtype=torch.float64)
    return p, x, ext, y, z


def _oracle(p, x, ext, y, z, lbda, num_data, beta):
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    loss = port.vcmp_loss(q, x, ext, y, z, lbda, num_data, beta, True)
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in q.items()}


def _gpu(p, x, ext, y, z, lbda, num_data, beta, dtype):
    from deconditional_downscaling_b200 import ops
    d, r = x.shape[1], ext.shape[1]
    t = lambda v: v.to(device=DEV, dtype=dtype)
    k_spec = ops.FlatSpec(["rbf"], [list(range(d))], [t(softplus(p["k_raw_ls"]))], [t(softplus(p["k_raw_os"]))])
    if "l1_raw_ls" in p:
        l_spec = ops.FlatSpec(["matern32", "rbf"], [[0, 1], list(range(2, r))],
                              [t(softplus(p["l0_raw_ls"])), t(softplus(p["l1_raw_ls"]))],
                              [t(softplus(p["l0_raw_os"])), t(softplus(p["l1_raw_os"]))])
        lraws = [("l0_raw_ls", "l0_raw_os"), ("l1_raw_ls", "l1_raw_os")]
    else:
        l_spec = ops.FlatSpec(["rbf"], [list(range(r))], [t(softplus(p["l0_raw_ls"]))], [t(softplus(p["l0_raw_os"]))])
        lraws = [("l0_raw_ls", "l0_raw_os")]
    noise = t(softplus(p["raw_noise"]) + 1e-4)
    ind = t(softplus(p["noise_kernel_raw_os"]))
    elbo, g, info = ops.cmp_elbo(k_spec, l_spec, t(p["Z"]), t(x), t(ext), t(y), t(z), t(p["var_mean"]), t(p["chol_var"]),
                                 noise, ind, lbda, num_data, beta=beta, scale=-1.0)
    assert info.tolist() == [0, 0, 0]
    sg = lambda name: torch.sigmoid(t(p[name]))
    out = {"Z": g["Z"], "var_mean": g["var_mean"], "chol_var": g["chol_var"], "raw_noise": g["noise"] * sg("raw_noise"),
           "noise_kernel_raw_os": g["ind_noise"] * sg("noise_kernel_raw_os"),
           "k_raw_ls": g["dk_ls"][0, :d] * sg("k_raw_ls"), "k_raw_os": g["dk_os"][0] * sg("k_raw_os")}
    for c, (a, b) in enumerate(lraws):
        out[a] = g["dl_ls"][c, :p[a].numel()] * sg(a)
        out[b] = g["dl_os"][c] * sg(b)
    return -elbo, out


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-7), (torch.float32, 3e-3)])
@pytest.mark.parametrize("M,N,n,d,r,additive", [(48, 300, 40, 3, 3, True), (130, 700, 150, 5, 3, True), (33, 257, 19, 2, 2, False)])
def test_loss_and_every_gradient_match_the_oracle(dtype, tol, M, N, n, d, r, additive):
    p, x, ext, y, z = _problem(M, N, n, d, r, seed=M + n, additive=additive)
    lbda, num_data, beta = 1e-3, 10.0 * n, 0.8
    ref, gref = _oracle(p, x, ext, y, z, lbda, num_data, beta)
    loss, g = _gpu(p, x, ext, y, z, lbda, num_data, beta, dtype)
    assert abs(loss.item() - ref.item()) <= tol * max(1.0, abs(ref.item())), (loss.item(), ref.item())
    for name, gr in gref.items():
        got = g[name].double().cpu().reshape(gr.shape)
        if name == "chol_var":
            gr = gr.tril()
            assert got.triu(1).abs().max().item() == 0.0
        err = (got - gr).norm().item() / max(gr.norm().item(), 1e-12)
        # bag-kernel hyper-parameters: sums of two opposing parts (through l(y~, y) and through Lambda^{-1})
        lim = tol * (40 if (dtype == torch.float32 and name.startswith("l")) else 5)
        assert err <= lim, (name, err)


def test_failed_lambda_factorisation_is_reported():
    """Two identical individuals' bag values and lbda = 0: Lambda is singular, info[0] must say so."""
    from deconditional_downscaling_b200 import ops
    p, x, ext, y, z = _problem(16, 140, 12, 2, 2, seed=4, additive=False)
    ext[5] = ext[4]
    t = lambda v: v.to(device=DEV, dtype=torch.float64)
    k_spec = ops.FlatSpec(["rbf"], [[0, 1]], [t(softplus(p["k_raw_ls"]))], [t(softplus(p["k_raw_os"]))])
    l_spec = ops.FlatSpec(["rbf"], [[0, 1]], [t(softplus(p["l0_raw_ls"]))], [t(softplus(p["l0_raw_os"]))])
    _, _, info = ops.cmp_elbo(k_spec, l_spec, t(p["Z"]), t(x), t(ext), t(y), t(z), t(p["var_mean"]), t(p["chol_var"]),
                              t(softplus(p["raw_noise"])), t(softplus(p["noise_kernel_raw_os"])), 0.0, 100.0,
                              need_grads=False)
    assert info[0].item() != 0


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-8), (torch.float32, 2e-3)])
def test_trainer_takes_the_composite_and_follows_the_layered_path(dtype, tol, monkeypatch):
    """ElboTrainer on a VariationalCMP (additive bag kernel, constant mean, individuals noise): three Adam steps through the
    C composite (default) and through the layered Python path (DCGP_CMP_FUSED=0) give the same losses and parameters."""
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as Md
    from deconditional_downscaling_b200 import _lib
    from deconditional_downscaling_b200.training import ElboTrainer
    g = torch.Generator().manual_seed(17)
    N, n, M, d, r = 600, 64, 40, 3, 3
    x = torch.randn(N, d, generator=g).to(DEV, dtype)
    ext = torch.randn(N, r, generator=g).to(DEV, dtype)
    y = torch.randn(n, r, generator=g).to(DEV, dtype)
    z = torch.randn(n, generator=g).to(DEV, dtype)
    Z0 = torch.randn(M, d, generator=g)
    runs = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("DCGP_CMP_FUSED", flag)
        torch.manual_seed(5)
        k = K.ScaleKernel(K.RBFKernel(ard_num_dims=d))
        l = K.ScaleKernel(K.MaternKernel(nu=1.5, ard_num_dims=2, active_dims=[0, 1])) + \
            K.ScaleKernel(K.RBFKernel(ard_num_dims=1, active_dims=[2]))
        model = Md.VariationalCMP(inducing_points=Z0.clone(), individuals_mean=Mn.ConstantMean(), individuals_kernel=k,
                                  bag_kernel=l, lbda=1e-3, use_individuals_noise=True).to(DEV).to(dtype)
        lik = L.CMPLikelihood(use_individuals_noise=True).to(DEV).to(dtype)
        tr = ElboTrainer(model, lik, num_data=640, beta=1.0, lr=0.05, cuda_graph=False)
        l0 = _lib.load().dcgp_launch_count()
        losses = [tr.step(x, z, bags_values=y, extended_bags_values=ext).item() for _ in range(3)]
        runs[flag] = (losses, {nm: p.detach().clone() for nm, p in list(model.named_parameters()) + list(lik.named_parameters())},
                      _lib.load().dcgp_launch_count() - l0)
    for a, b in zip(runs["1"][0], runs["0"][0]):
        assert abs(a - b) <= tol * max(1.0, abs(b)), (runs["1"][0], runs["0"][0])
    for nm, p in runs["0"][1].items():
        assert (runs["1"][1][nm] - p).abs().max().item() <= (1e-6 if dtype == torch.float64 else 5e-3), nm
    assert runs["1"][2] != runs["0"][2]          # the two runs really took different paths


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-7), (torch.float32, 3e-3)])
def test_constant_individuals_mean_and_its_gradient(dtype, tol):
    """individuals_mean = ConstantMean(c): mu(x) enters q(f).mean (variational_strategy.py:47-52) and the residual
    z - A^T mu_q (cmp_likelihood.py:70)."""
    from oracle import restate as R
    from deconditional_downscaling_b200 import ops
    p, x, ext, y, z = _problem(40, 260, 30, 3, 3, seed=9, additive=True)
    lbda, num_data, beta, c0 = 1e-3, 300.0, 1.0, 0.35
    q = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    q["mean_const"] = torch.tensor([c0], dtype=torch.float64, requires_grad=True)
    k, l = port.vcmp_specs(q, 3)
    mu_q, Sigma_q = R.svgp_q(k, q["Z"], x, q["var_mean"], q["chol_var"], mu_x=q["mean_const"].expand(x.shape[0]))
    Rt, B = R.elbo_computation_parameters(l, y, ext, lbda)
    ell = R.cmp_ell_with_noise(z, mu_q, Sigma_q, Rt, B, softplus(q["noise_kernel_raw_os"]), q["raw_noise"])
    ref = -R.bag_elbo(ell, R.whitened_kl(q["var_mean"], q["chol_var"]), z.numel(), num_data, beta)
    ref.backward()
    t = lambda v: v.to(device=DEV, dtype=dtype)
    k_spec = ops.FlatSpec(["rbf"], [[0, 1, 2]], [t(softplus(p["k_raw_ls"]))], [t(softplus(p["k_raw_os"]))])
    l_spec = ops.FlatSpec(["matern32", "rbf"], [[0, 1], [2]], [t(softplus(p["l0_raw_ls"])), t(softplus(p["l1_raw_ls"]))],
                          [t(softplus(p["l0_raw_os"])), t(softplus(p["l1_raw_os"]))])
    elbo, g, info = ops.cmp_elbo(k_spec, l_spec, t(p["Z"]), t(x), t(ext), t(y), t(z), t(p["var_mean"]), t(p["chol_var"]),
                                 t(softplus(p["raw_noise"]) + 1e-4), t(softplus(p["noise_kernel_raw_os"])), lbda, num_data,
                                 beta=beta, mean_const=t(torch.tensor([c0])), scale=-1.0)
    assert info.tolist() == [0, 0, 0]
    assert abs(-elbo.item() - ref.item()) <= tol * max(1.0, abs(ref.item()))
    assert abs(g["mean_const"].item() - q["mean_const"].grad.item()) <= 5 * tol * max(abs(q["mean_const"].grad.item()), 1e-3)
    err = (g["var_mean"].double().cpu() - q["var_mean"].grad).norm() / q["var_mean"].grad.norm()
    assert err.item() <= 5 * tol


def test_argument_errors_are_reported_as_negative_codes():
    """C-ABI error convention: -k for a bad argument, no CUDA work issued."""
    import ctypes
    from deconditional_downscaling_b200 import _lib
    lib = _lib.load()
    assert lib.dcgp_cmp_elbo_workspace(0, 4, 4, 2, 0) == 0
    assert lib.dcgp_vbagg_elbo_workspace(4, 0, 2, 0) == 0
    empty = _lib.CmpProblem()
    assert lib.dcgp_cmp_elbo_fwd(ctypes.byref(empty), None, None, None, 0, 0, 0, None) == -1
    empty_v = _lib.VbaggProblem()
    assert lib.dcgp_vbagg_elbo_fwd(ctypes.byref(empty_v), None, None, None, 0, 0, 0, None) == -1
    # a workspace that is too small is refused before anything is launched
    from deconditional_downscaling_b200 import ops
    p, x, ext, y, z = _problem(16, 140, 12, 2, 2, seed=4, additive=False)
    t = lambda v: v.to(device=DEV, dtype=torch.float64)
    k_spec = ops.FlatSpec(["rbf"], [[0, 1]], [t(softplus(p["k_raw_ls"]))], [t(softplus(p["k_raw_os"]))]).c_spec()
    l_spec = ops.FlatSpec(["rbf"], [[0, 1]], [t(softplus(p["l0_raw_ls"]))], [t(softplus(p["l0_raw_os"]))]).c_spec()
    Z, xd, ed, yd, zd, m, Ls = t(p["Z"]), t(x), t(ext), t(y), t(z), t(p["var_mean"]), t(p["chol_var"])
    noise, ind = t(torch.ones(1)), t(torch.ones(1))
    prob = _lib.CmpProblem(ctypes.pointer(k_spec), ctypes.pointer(l_spec), Z.data_ptr(), 16, 2, xd.data_ptr(), 140, 2,
                           ed.data_ptr(), 2, yd.data_ptr(), 12, 2, zd.data_ptr(), m.data_ptr(), Ls.data_ptr(), 16,
                           noise.data_ptr(), ind.data_ptr(), None, 1e-3, 100.0, 1.0, 1e-3, 1e-4)
    ws = torch.empty(1024, dtype=torch.uint8, device=DEV)
    out = torch.empty(1, dtype=torch.float64, device=DEV)
    info = torch.zeros(3, dtype=torch.int32, device=DEV)
    l0 = lib.dcgp_launch_count()
    rc = lib.dcgp_cmp_elbo_fwd(ctypes.byref(prob), out.data_ptr(), info.data_ptr(), ws.data_ptr(), 1024, 0, 0, None)
    assert rc == -5 and lib.dcgp_launch_count() == l0


def test_trainer_step_with_a_singular_lambda_never_corrupts_the_parameters():
    """lbda = 0 and duplicated bag values make Lambda singular: the composite reports it through `info`, the device-side
    update is vetoed, and the trainer's checked pass (layered path, psd_safe_cholesky-style jitter retries) either
    completes the step or raises NotPSDError - in both cases no parameter may turn NaN."""
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as Md
    from deconditional_downscaling_b200.functional import NotPSDError
    from deconditional_downscaling_b200.training import ElboTrainer
    g = torch.Generator().manual_seed(3)
    N, n, M, d, r = 400, 32, 24, 2, 2
    x = torch.randn(N, d, generator=g, dtype=torch.float64).to(DEV)
    ext = torch.randn(N, r, generator=g, dtype=torch.float64)
    ext[N // 2:] = ext[: N // 2]                      # every bag value twice: rank-deficient Gram
    ext = ext.to(DEV)
    y = torch.randn(n, r, generator=g, dtype=torch.float64).to(DEV)
    z = torch.randn(n, generator=g, dtype=torch.float64).to(DEV)
    torch.manual_seed(7)
    model = Md.VariationalCMP(inducing_points=torch.randn(M, d, generator=g, dtype=torch.float64), individuals_mean=Mn.ZeroMean(),
                              individuals_kernel=K.ScaleKernel(K.RBFKernel(ard_num_dims=d)),
                              bag_kernel=K.ScaleKernel(K.RBFKernel(ard_num_dims=r)), lbda=0.0,
                              use_individuals_noise=True).to(DEV).double()
    lik = L.CMPLikelihood(use_individuals_noise=True).to(DEV).double()
    tr = ElboTrainer(model, lik, num_data=320, beta=1.0, lr=0.05, cuda_graph=False)
    before = {nm: p.detach().clone() for nm, p in model.named_parameters()}
    try:
        loss = tr.step(x, z, bags_values=y, extended_bags_values=ext)
        assert torch.isfinite(loss).item()
    except NotPSDError:
        for nm, p in model.named_parameters():      # a failed step leaves the parameters as they were
            if nm in before and before[nm].shape == p.shape and "variational" not in nm:
                assert torch.equal(p.detach(), before[nm]), nm
    for nm, p in model.named_parameters():
        assert torch.isfinite(p).all().item(), nm
