"""nn.Module with gpytorch's `initialize(**kwargs)` convention (SURVEY.md Appendix A,
`Parameter transforms`): raw parameters are set directly, dotted names recurse, other
names go through property setters."""
import torch
from torch import nn


class Module(nn.Module):
    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if isinstance(val, int):
                val = float(val)
            if "." in name:
                head, tail = name.split(".", 1)
                getattr(self, head).initialize(**{tail: val})
            elif name in self._parameters:
                p = self._parameters[name]
                with torch.no_grad():
                    if torch.is_tensor(val):
                        val = val.to(p)
                        p.copy_(val.expand_as(p) if val.numel() != p.numel() or val.dim() <= p.dim() and _expandable(val, p) else val.reshape(p.shape))
                    else:
                        p.fill_(val)
            else:
                setattr(self, name, val)
        return self

    def named_priors(self):
        return iter(())

    def added_loss_terms(self):
        return iter(())

    def hyperparameters(self):
        for _, p in self.named_parameters():
            yield p


def _expandable(val, p):
    try:
        torch.broadcast_shapes(val.shape, p.shape)
        return torch.broadcast_shapes(val.shape, p.shape) == p.shape
    except RuntimeError:
        return False
