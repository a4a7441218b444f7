"""Generates tests/golden/state_dict_keys.json: names and shapes of every state_dict entry of the REFERENCE's
modules (unmodified /root/reference/src imported over the restated gpytorch 1.4.0 subset, oracle/gpytorch_min),
i.e. the checkpoint wire format of `state.pt` (experiments/swiss_roll/models/exact_cmp.py:121-124).
Run in the build container only:  python tests/golden/make_state_dict_keys.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.ref_loader import load_reference_src  # noqa: E402

load_reference_src()
import gpytorch  # noqa: E402  (the shim)
import src.likelihoods  # noqa: E402
import src.models  # noqa: E402


def shapes(module):
    return {k: list(v.shape) for k, v in module.state_dict().items()}


def scale_rbf(d):
    return gpytorch.kernels.ScaleKernel(gpytorch.kernels.RBFKernel(ard_num_dims=d))


def main():
    torch.manual_seed(0)
    x, ext, y, z = torch.randn(40, 3), torch.randn(40, 1), torch.randn(5, 1), torch.randn(5)
    out = {}
    m = src.models.ExactCMP(train_individuals=x, extended_train_bags=ext, train_bags=y, train_aggregate_targets=z,
                            individuals_mean=gpytorch.means.ZeroMean(), individuals_kernel=scale_rbf(3), bag_kernel=scale_rbf(1),
                            lbda=0.1, likelihood=gpytorch.likelihoods.GaussianLikelihood(), use_individuals_noise=True)
    out["exact_cmp"] = shapes(m)
    m = src.models.VariationalCMP(inducing_points=torch.randn(7, 3), individuals_mean=gpytorch.means.ZeroMean(),
                                  individuals_kernel=scale_rbf(3), bag_kernel=scale_rbf(1), lbda=0.1, use_individuals_noise=True)
    out["variational_cmp"] = shapes(m)
    out["cmp_likelihood"] = shapes(src.likelihoods.CMPLikelihood(use_individuals_noise=True))
    m = src.models.VariationalGP(inducing_points=torch.randn(7, 3), mean_module=gpytorch.means.ZeroMean(), covar_module=scale_rbf(3))
    out["variational_gp"] = shapes(m)
    out["vbagg_likelihood"] = shapes(src.likelihoods.VBaggGaussianLikelihood())
    add = gpytorch.kernels.ScaleKernel(gpytorch.kernels.MaternKernel(nu=1.5, ard_num_dims=2, active_dims=[0, 1])) + scale_rbf(1)
    out["additive_kernel"] = shapes(add)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "state_dict_keys.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
