"""gpytorch.utils.errors restated (test infrastructure)."""


class NotPSDError(RuntimeError):
    pass


class NanError(RuntimeError):
    pass


class CachingError(RuntimeError):
    pass
