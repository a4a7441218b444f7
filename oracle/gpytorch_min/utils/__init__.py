"""gpytorch.utils restated (test infrastructure)."""
from . import errors, memoize, cholesky  # noqa: F401
