"""Differentiable random-Fourier-feature map (libdcgp rff kernels)."""
import torch
from torch.autograd import Function

from . import ops


class _RFFFn(Function):
    @staticmethod
    def forward(ctx, x, W, lengthscale):
        ctx.save_for_backward(x, W, lengthscale)
        return ops.rff_features(x, W, lengthscale.detach())

    @staticmethod
    def backward(ctx, dphi):
        x, W, ls = ctx.saved_tensors
        dx, dls = ops.rff_features_bwd(x, W, ls.detach(), dphi.contiguous(), need_dx=ctx.needs_input_grad[0])
        return dx, None, dls.reshape(ls.shape) if ctx.needs_input_grad[2] else None


def rff_features(x, W, lengthscale):
    """[cos(x W / l), sin(x W / l)] (src/kernels/rff_kernel.py:82-89); W is a buffer (no grad)."""
    return _RFFFn.apply(x, W, lengthscale)
