"""Tiny helper for tools/kbench.py: builds a FlatSpec without touching oracle/."""
import torch
from deconditional_downscaling_b200 import ops


def rbf_spec(d, dtype, device="cuda:0", lengthscale=1.0, outputscale=1.0, kind="rbf"):
    return ops.FlatSpec([kind], [list(range(d))], [torch.full((d,), lengthscale, dtype=dtype, device=device)],
                        [torch.full((1,), outputscale, dtype=dtype, device=device)])
