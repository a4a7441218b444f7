"""gpytorch.settings stand-ins for the four switches the reference touches: `fast_pred_var` and
`fast_computations` around prediction / NLL code (experiments/swiss_roll/core/metrics.py:24),
`detach_test_caches` (src/models/exact_cmp.py:141-143) and `trace_mode` (src/variational/variational_strategy.py:4,56).  This implementation has one numerical path (Cholesky on the
GPU, never Lanczos / CG), so they only have to exist, nest and report their state."""
import contextlib


class _Flag(contextlib.ContextDecorator):
    _default = False

    def __init__(self, state=True):
        self.state, self._prev = bool(state), None

    def __enter__(self):
        self._prev = type(self).on()
        type(self)._value = self.state
        return self

    def __exit__(self, *exc):
        type(self)._value = self._prev
        return False

    @classmethod
    def on(cls):
        return bool(cls.__dict__.get("_value", cls._default))

    @classmethod
    def off(cls):
        return not cls.on()


class fast_pred_var(_Flag):
    pass


class detach_test_caches(_Flag):
    _default = True


class trace_mode(_Flag):
    pass


class fast_computations(contextlib.ContextDecorator):
    """fast_computations(covar_root_decomposition, log_prob, solves): all three are always "exact" here."""

    def __init__(self, covar_root_decomposition=True, log_prob=True, solves=True):
        self.args = (covar_root_decomposition, log_prob, solves)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
