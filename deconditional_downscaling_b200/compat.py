"""Name-compatibility layer for the unchanged experiment runners.

The reference's trainers do `import gpytorch` and `from src.models import ExactCMP, ...`
(experiments/swiss_roll/models/exact_cmp.py:4-6, experiments/downscaling/models/variational_cmp.py:1-12).
`install()` publishes this package under those two names: a `gpytorch` namespace holding the classes the
runners touch (kernels, means, likelihoods, mlls, models, distributions, variational, lazy, settings,
utils.errors, utils.memoize) and a `src` namespace (kernels, means, likelihoods, mlls, models, variational,
utils), so that the runner scripts execute on libdcgp without edits.  A real gpytorch installation is never
shadowed unless `force=True`.
"""
import importlib.util
import sys
import types

from . import distributions, functional, kernels, lazy, likelihoods, means, mlls, models, module, settings, utils, variational
from .models import base as _base

_INSTALLED = {}


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__package__ = name.rpartition(".")[0] or name
    return m


def _memoize_module():
    """gpytorch.utils.memoize: the reference's prediction strategies cache through these
    (src/models/prediction_strategies/cmp_prediction_strategy.py:2-3); here caching lives in the model objects,
    the decorators only have to be importable and behave as plain per-instance caches."""
    import functools

    def cached(method=None, name=None, ignore_args=False):
        if method is None:
            return functools.partial(cached, name=name, ignore_args=ignore_args)
        key = name or method.__name__

        @functools.wraps(method)
        def wrapper(self, *a, **k):
            store = self.__dict__.setdefault("_memoize_cache", {})
            if key not in store:
                store[key] = method(self, *a, **k)
            return store[key]
        return wrapper

    def clear_cache_hook(obj, *a, **k):
        obj.__dict__.pop("_memoize_cache", None)

    def pop_from_cache_ignore_args(obj, name):
        try:
            return obj.__dict__.get("_memoize_cache", {}).pop(name)
        except KeyError:
            raise functional.CachingError(f"{name} not in cache")

    def add_to_cache(obj, name, val, *a, **k):
        obj.__dict__.setdefault("_memoize_cache", {})[name] = val
        return val

    return _mod("gpytorch.utils.memoize", cached=cached, clear_cache_hook=clear_cache_hook,
                pop_from_cache_ignore_args=pop_from_cache_ignore_args, add_to_cache=add_to_cache)


def gpytorch_namespace():
    k = _mod("gpytorch.kernels", Kernel=kernels.Kernel, RBFKernel=kernels.RBFKernel, MaternKernel=kernels.MaternKernel,
             ScaleKernel=kernels.ScaleKernel, AdditiveKernel=kernels.AdditiveKernel, RFFKernel=kernels.GPyTorchRFFKernel)
    mn = _mod("gpytorch.means", Mean=means.Mean, ZeroMean=means.ZeroMean, ConstantMean=means.ConstantMean)
    lk = _mod("gpytorch.likelihoods", Likelihood=likelihoods.Likelihood, GaussianLikelihood=likelihoods.GaussianLikelihood)
    mll_base = _mod("gpytorch.mlls.marginal_log_likelihood", MarginalLogLikelihood=mlls.MarginalLogLikelihood)
    ml = _mod("gpytorch.mlls", MarginalLogLikelihood=mlls.MarginalLogLikelihood, ExactMarginalLogLikelihood=mlls.ExactMarginalLogLikelihood,
              VariationalELBO=mlls.VariationalELBO, marginal_log_likelihood=mll_base)
    eps = _mod("gpytorch.models.exact_prediction_strategies", DefaultPredictionStrategy=_base.DefaultPredictionStrategy)
    md = _mod("gpytorch.models", GP=_base.GP, ExactGP=_base.ExactGPBase, ApproximateGP=_base.ApproximateGP,
              exact_prediction_strategies=eps)
    bd = _mod("gpytorch.distributions.base_distributions", Normal=__import__("torch").distributions.Normal)
    ds = _mod("gpytorch.distributions", MultivariateNormal=distributions.MultivariateNormal, base_distributions=bd)
    vr = _mod("gpytorch.variational", VariationalStrategy=variational.VariationalStrategy,
              CholeskyVariationalDistribution=variational.CholeskyVariationalDistribution)
    lz = _mod("gpytorch.lazy", LazyTensor=lazy.DenseMatrix, NonLazyTensor=lazy.DenseMatrix, DiagLazyTensor=lazy.DiagLazyTensor,
              SumLazyTensor=lazy.SumLazyTensor, MatmulLazyTensor=lazy.MatmulLazyTensor, RootLazyTensor=lazy.RootLazyTensor,
              LowRankRootLazyTensor=lazy.RootLazyTensor, lazify=lazy.lazify, delazify=lazy.delazify)
    er = _mod("gpytorch.utils.errors", NotPSDError=functional.NotPSDError, CachingError=functional.CachingError)
    ch = _mod("gpytorch.utils.cholesky", psd_safe_cholesky=functional.psd_safe_cholesky)
    ut = _mod("gpytorch.utils", errors=er, memoize=_memoize_module(), cholesky=ch)
    st = settings
    root = _mod("gpytorch", kernels=k, means=mn, likelihoods=lk, mlls=ml, models=md, distributions=ds, variational=vr, lazy=lz,
                utils=ut, settings=st, Module=module.Module, __version__="1.4.0+dcgp", lazify=lazy.lazify, delazify=lazy.delazify,
                __path__=[])
    subs = {"gpytorch": root, "gpytorch.kernels": k, "gpytorch.means": mn, "gpytorch.likelihoods": lk, "gpytorch.mlls": ml,
            "gpytorch.mlls.marginal_log_likelihood": mll_base, "gpytorch.models": md,
            "gpytorch.models.exact_prediction_strategies": eps, "gpytorch.distributions": ds,
            "gpytorch.distributions.base_distributions": bd, "gpytorch.variational": vr, "gpytorch.lazy": lz,
            "gpytorch.utils": ut, "gpytorch.utils.errors": er, "gpytorch.utils.memoize": ut.memoize,
            "gpytorch.utils.cholesky": ch, "gpytorch.settings": st}
    for m in (md, ml, ds, ut):
        m.__path__ = []
    return subs


def src_namespace():
    sk = _mod("src.kernels", CMEAggregateKernel=kernels.CMEAggregateKernel, DeltaKernel=kernels.DeltaKernel, RFFKernel=kernels.RFFKernel)
    sm = _mod("src.means", CMEAggregateMean=means.CMEAggregateMean)
    sl = _mod("src.likelihoods", CMPLikelihood=likelihoods.CMPLikelihood, VBaggGaussianLikelihood=likelihoods.VBaggGaussianLikelihood)
    se = _mod("src.mlls", BagVariationalELBO=mlls.BagVariationalELBO)
    so = _mod("src.models", ExactGP=models.ExactGP, VariationalGP=models.VariationalGP, ExactCMP=models.ExactCMP,
              VariationalCMP=models.VariationalCMP, BaggedGP=models.BaggedGP, CMP=models.CMP)
    sv = _mod("src.variational", VariationalStrategy=variational.VariationalStrategy)
    su = _mod("src.utils", Registry=utils.Registry, setseed=utils.setseed)
    root = _mod("src", kernels=sk, means=sm, likelihoods=sl, mlls=se, models=so, variational=sv, utils=su, __path__=[])
    return {"src": root, "src.kernels": sk, "src.means": sm, "src.likelihoods": sl, "src.mlls": se, "src.models": so,
            "src.variational": sv, "src.utils": su}


def install(gpytorch=True, src=True, force=False):
    """Register the namespaces in sys.modules.  Returns the dict of modules that were installed."""
    done = {}
    if gpytorch:
        real = "gpytorch" in sys.modules and "gpytorch" not in _INSTALLED
        if not real and "gpytorch" not in _INSTALLED:
            try:
                real = importlib.util.find_spec("gpytorch") is not None
            except (ImportError, ValueError):
                real = False
        if real and not force:
            raise RuntimeError("a real gpytorch is importable; pass force=True to shadow it with the libdcgp classes")
        done.update(gpytorch_namespace())
    if src:
        if "src" in sys.modules and "src" not in _INSTALLED and not force:
            raise RuntimeError("a module named `src` is already imported; pass force=True to replace it")
        done.update(src_namespace())
    sys.modules.update(done)
    _INSTALLED.update(done)
    return done


def uninstall():
    for name, m in list(_INSTALLED.items()):
        if sys.modules.get(name) is m:
            del sys.modules[name]
    _INSTALLED.clear()
