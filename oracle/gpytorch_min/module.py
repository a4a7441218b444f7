"""gpytorch.Module restated (test infrastructure): nn.Module + `initialize`,
constraints and (empty) prior / added-loss iterators.  gpytorch==1.4.0
semantics recalled in SURVEY.md Appendix A ("Parameter transforms")."""
import torch
from torch import nn


class Module(nn.Module):
    def __init__(self):
        super().__init__()
        self._added_loss_terms = {}

    # ---- constraints -------------------------------------------------
    def register_constraint(self, param_name, constraint):
        self.add_module(param_name + "_constraint", constraint)

    def constraint_for_parameter_name(self, name):
        return getattr(self, name + "_constraint", None)

    # ---- initialize(raw_x=...) --------------------------------------
    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if isinstance(val, int):
                val = float(val)
            if "." in name:
                head, tail = name.split(".", 1)
                getattr(self, head).initialize(**{tail: val})
                continue
            if name in self._parameters:
                p = self._parameters[name]
                if torch.is_tensor(val):
                    try:
                        p.data.copy_(val.expand_as(p))
                    except RuntimeError:
                        p.data.copy_(val.reshape(p.shape))
                else:
                    p.data.fill_(val)
            else:
                # goes through a property setter (e.g. `noise`, `lengthscale`)
                setattr(self, name, val)
        return self

    # ---- priors / added losses (none on the hot path) -----------------
    def named_priors(self):
        return iter(())

    def added_loss_terms(self):
        return iter(())

    def hyperparameters(self):
        for _, p in self.named_parameters():
            yield p

    def __call__(self, *inputs, **kwargs):
        return self.forward(*inputs, **kwargs)
