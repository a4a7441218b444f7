"""gpytorch.settings restated (test infrastructure).  Only the flags the
reference touches; defaults from SURVEY.md Appendix A ("settings defaults").
The shim always runs the Cholesky branch, so `fast_computations` and
`max_cholesky_size` are accepted and ignored."""


class _feature_flag:
    _state = False

    def __init__(self, state=True):
        self.prev = self.__class__._state
        self.state = state

    @classmethod
    def on(cls):
        return cls._state

    @classmethod
    def off(cls):
        return not cls._state

    def __enter__(self):
        self.__class__._state = self.state

    def __exit__(self, *args):
        self.__class__._state = self.prev
        return False


class fast_pred_var(_feature_flag):
    _state = False


class detach_test_caches(_feature_flag):
    _state = True


class trace_mode(_feature_flag):
    _state = False


class debug(_feature_flag):
    _state = True


class lazily_evaluate_kernels(_feature_flag):
    _state = True


class fast_computations:
    def __init__(self, covar_root_decomposition=True, log_prob=True, solves=True):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *args):
        return False


class _value_context:
    _value = None

    def __init__(self, value):
        self.prev = self.__class__._value
        self.new = value

    @classmethod
    def value(cls, *a):
        return cls._value

    def __enter__(self):
        self.__class__._value = self.new

    def __exit__(self, *args):
        self.__class__._value = self.prev
        return False


class max_cholesky_size(_value_context):
    _value = 800


class max_root_decomposition_size(_value_context):
    _value = 100


class cholesky_jitter(_value_context):
    @classmethod
    def value(cls, dtype=None):
        import torch
        return 1e-8 if dtype == torch.float64 else 1e-6
