"""Name-compatibility layer: the `gpytorch` / `src` namespaces the unchanged
runners import, the builder registry and seeding decorator, and the checkpoint wire format (state_dict keys and
shapes recorded from the reference by tests/golden/make_state_dict_keys.py).  No GPU compute here."""
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture()
def compat():
    from deconditional_downscaling_b200 import compat as c
    saved = {k: v for k, v in sys.modules.items() if k == "gpytorch" or k.startswith("gpytorch.") or k == "src" or k.startswith("src.")}
    for k in saved:
        del sys.modules[k]
    c.install()
    yield c
    c.uninstall()
    sys.modules.update(saved)


def test_runner_imports_resolve(compat):
    # every import statement the reference's experiments/ and src/ trees make (grep over /root/reference)
    import gpytorch
    from gpytorch import distributions, kernels, lazy, means, settings, variational  # noqa: F401
    from gpytorch.distributions import MultivariateNormal  # noqa: F401
    from gpytorch.kernels import ScaleKernel  # noqa: F401
    from gpytorch.lazy import DiagLazyTensor, MatmulLazyTensor, SumLazyTensor  # noqa: F401
    from gpytorch.likelihoods import GaussianLikelihood  # noqa: F401
    from gpytorch.mlls.marginal_log_likelihood import MarginalLogLikelihood  # noqa: F401
    from gpytorch.models import ApproximateGP, ExactGP  # noqa: F401
    from gpytorch.models.exact_prediction_strategies import DefaultPredictionStrategy  # noqa: F401
    from gpytorch.settings import trace_mode  # noqa: F401
    from gpytorch.utils.errors import CachingError  # noqa: F401
    from gpytorch.utils.memoize import cached, clear_cache_hook, pop_from_cache_ignore_args  # noqa: F401
    from gpytorch.variational import CholeskyVariationalDistribution, VariationalStrategy  # noqa: F401
    from src.kernels import RFFKernel  # noqa: F401
    from src.likelihoods import CMPLikelihood, VBaggGaussianLikelihood  # noqa: F401
    from src.mlls import BagVariationalELBO  # noqa: F401
    from src.models import BaggedGP, ExactCMP, ExactGP as SrcExactGP, VariationalCMP, VariationalGP  # noqa: F401
    from src.utils import Registry, setseed  # noqa: F401
    for attr in ("kernels.RBFKernel", "kernels.MaternKernel", "kernels.RFFKernel", "means.ZeroMean", "mlls.ExactMarginalLogLikelihood",
                 "mlls.VariationalELBO", "likelihoods.Likelihood", "utils.errors.NotPSDError", "settings.fast_computations",
                 "settings.fast_pred_var", "settings.detach_test_caches", "lazy.LazyTensor", "lazy.RootLazyTensor",
                 "lazy.LowRankRootLazyTensor", "lazy.lazify", "lazy.delazify", "distributions.base_distributions.Normal"):
        obj = gpytorch
        for part in attr.split("."):
            obj = getattr(obj, part)
    k = gpytorch.kernels.ScaleKernel(gpytorch.kernels.MaternKernel(nu=1.5, ard_num_dims=2, active_dims=[0, 1])) + \
        gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=1, active_dims=[2]))
    assert len(k.kernels) == 2
    with gpytorch.settings.fast_computations(log_prob=False, covar_root_decomposition=False), gpytorch.settings.fast_pred_var():
        assert gpytorch.settings.fast_pred_var.on()
    assert gpytorch.settings.fast_pred_var.off() and gpytorch.settings.detach_test_caches.on()


def test_install_refuses_to_shadow_an_imported_module(compat):
    compat.uninstall()
    sys.modules["src"] = type(sys)("src")
    try:
        with pytest.raises(RuntimeError):
            compat.install(gpytorch=False)
        compat.install(gpytorch=False, force=True)
        from src.utils import Registry  # noqa: F401
    finally:
        compat.uninstall()
        sys.modules.pop("src", None)


def test_registry_and_setseed():
    from deconditional_downscaling_b200.utils import Registry, setseed
    reg = Registry()

    @reg.register("fn")
    def build_fn(a):
        return a + 1

    @reg.register("cls")
    class Thing:
        @classmethod
        def build(cls, a):
            return ("thing", a)

    assert reg["fn"](1) == 2 and reg["cls"](3) == ("thing", 3)
    reg.register("again", build_fn)
    with pytest.raises(AssertionError):
        reg.register("fn", build_fn)
    with pytest.raises(TypeError):
        reg.register("bad", 3)

    @setseed("random", "numpy", "torch")
    def draw(seed=None):
        return random.random(), float(np.random.rand()), float(torch.rand(1))

    assert draw(seed=7) == draw(seed=7)
    assert draw(seed=7) != draw(seed=8)
    with pytest.raises(AssertionError):
        setseed("jax")


def test_state_dict_wire_format_matches_reference_checkpoints():
    from deconditional_downscaling_b200 import kernels as K, likelihoods as L, means as Mn, models as M
    with open(os.path.join(HERE, "golden", "state_dict_keys.json")) as f:
        ref = json.load(f)
    sr = lambda d: K.ScaleKernel(K.RBFKernel(ard_num_dims=d))  # noqa: E731
    torch.manual_seed(0)
    x, ext, y, z = torch.randn(40, 3), torch.randn(40, 1), torch.randn(5, 1), torch.randn(5)
    mine = {
        "exact_cmp": M.ExactCMP(train_individuals=x, extended_train_bags=ext, train_bags=y, train_aggregate_targets=z,
                                individuals_mean=Mn.ZeroMean(), individuals_kernel=sr(3), bag_kernel=sr(1), lbda=0.1,
                                likelihood=L.GaussianLikelihood(), use_individuals_noise=True),
        "variational_cmp": M.VariationalCMP(inducing_points=torch.randn(7, 3), individuals_mean=Mn.ZeroMean(),
                                            individuals_kernel=sr(3), bag_kernel=sr(1), lbda=0.1, use_individuals_noise=True),
        "cmp_likelihood": L.CMPLikelihood(use_individuals_noise=True),
        "variational_gp": M.VariationalGP(inducing_points=torch.randn(7, 3), mean_module=Mn.ZeroMean(), covar_module=sr(3)),
        "vbagg_likelihood": L.VBaggGaussianLikelihood(),
        "additive_kernel": K.ScaleKernel(K.MaternKernel(nu=1.5, ard_num_dims=2, active_dims=[0, 1])) + sr(1),
    }
    for name, module in mine.items():
        got = {k: list(v.shape) for k, v in module.state_dict().items()}
        assert got == ref[name], name
        # a reference-format checkpoint loads, and what is saved loads back
        fake = {k: torch.full(s, 0.25) if s else torch.tensor(0.25) for k, s in ref[name].items()}
        module.load_state_dict(fake)
        assert all(torch.all(v == 0.25) for v in module.state_dict().values() if v.is_floating_point())
    lik = L.CMPLikelihood(use_individuals_noise=True)
    assert float(lik.noise_covar.raw_noise_constraint.transform(torch.zeros(1))) == pytest.approx(math_softplus0() + 1e-4)


def math_softplus0():
    import math
    return math.log(2.0)
