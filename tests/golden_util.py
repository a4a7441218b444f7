"""Helpers shared by the oracle tests and the GPU parity tests: load the
committed golden vectors (tests/golden/*.npz, produced by make_golden.py from
the reference's own src/ code) and rebuild oracle KernelSpecs from them."""
import os
import numpy as np
import torch

from oracle.restate import Component, KernelSpec

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name, dtype=torch.float64):
    raw = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in raw.files:
        a = raw[k]
        if a.dtype.kind == "f":
            out[k] = torch.from_numpy(a.copy()).to(dtype)
        elif a.ndim == 0:
            out[k] = a.item()
        else:
            out[k] = a
    return out


def p(g, key, prefix="param__"):
    return g[prefix + key]


def scaled(g, base, kind, dims=None, weights=None, prefix="param__", requires_grad=False):
    """Component for a ScaleKernel(<base kernel>) stored under `base` (e.g.
    'model__bag_kernel__kernels__0')."""
    ls = g[prefix + base + "__base_kernel__raw_lengthscale"].reshape(-1).clone()
    os_ = g[prefix + base + "__raw_outputscale"].reshape(()).clone()
    if requires_grad:
        ls.requires_grad_(True)
        os_.requires_grad_(True)
    return Component(kind=kind, raw_lengthscale=ls, raw_outputscale=os_, dims=dims, weights=weights)


def spec(*components):
    return KernelSpec(list(components))
