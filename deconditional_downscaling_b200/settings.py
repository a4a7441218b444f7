"""gpytorch.settings stand-ins.  The reference toggles these around prediction / NLL code
(experiments/swiss_roll/core/metrics.py:24, src/models/exact_cmp.py:141-143); this implementation has one
numerical path (Cholesky on the GPU, never Lanczos / CG), so they only have to exist, nest and report
their state."""
import contextlib


class _Flag(contextlib.ContextDecorator):
    _default = False

    def __init__(self, state=True):
        self.state, self._prev = bool(state), None

    def __enter__(self):
        self._prev = type(self)._state()
        type(self)._set(self.state)
        return self

    def __exit__(self, *exc):
        type(self)._set(self._prev)
        return False

    @classmethod
    def _state(cls):
        return cls.__dict__.get("_value", cls._default)

    @classmethod
    def _set(cls, v):
        cls._value = v

    @classmethod
    def on(cls):
        return bool(cls._state())

    @classmethod
    def off(cls):
        return not cls.on()


class fast_pred_var(_Flag):
    pass


class fast_pred_samples(_Flag):
    pass


class detach_test_caches(_Flag):
    _default = True


class trace_mode(_Flag):
    pass


class skip_posterior_variances(_Flag):
    pass


class debug(_Flag):
    _default = True


class fast_computations(contextlib.ContextDecorator):
    """fast_computations(covar_root_decomposition, log_prob, solves): all three are always "exact" here."""

    def __init__(self, covar_root_decomposition=True, log_prob=True, solves=True):
        self.args = (covar_root_decomposition, log_prob, solves)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _Value(contextlib.ContextDecorator):
    _default = None

    def __init__(self, value):
        self.v, self._prev = value, None

    def __enter__(self):
        self._prev = type(self).value()
        type(self)._value = self.v
        return self

    def __exit__(self, *exc):
        type(self)._value = self._prev
        return False

    @classmethod
    def value(cls):
        return cls.__dict__.get("_value", cls._default)


class max_cholesky_size(_Value):
    _default = 800


class cholesky_jitter(_Value):
    _default = None


class max_root_decomposition_size(_Value):
    _default = 100
