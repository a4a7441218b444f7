"""gpytorch.utils.cholesky.psd_safe_cholesky restated (test infrastructure).
Semantics: try cholesky(A); on failure retry with
A + jitter*10^i*I, i=0..2, base jitter 1e-6 (fp32) / 1e-8 (fp64); raise
NotPSDError if all fail."""
import warnings
import torch
from .errors import NotPSDError, NanError


def psd_safe_cholesky(A, upper=False, jitter=None, max_tries=3):
    L, info = torch.linalg.cholesky_ex(A)
    if not bool(info.any()):
        return L.transpose(-1, -2) if upper else L
    if torch.isnan(A).any():
        raise NanError("cholesky: input has NaNs")
    if jitter is None:
        jitter = 1e-8 if A.dtype == torch.float64 else 1e-6
    Aprime = A.clone()
    jitter_prev = 0.0
    for i in range(max_tries):
        jitter_new = jitter * (10 ** i)
        Aprime.diagonal(dim1=-2, dim2=-1).add_(jitter_new - jitter_prev)
        jitter_prev = jitter_new
        L, info = torch.linalg.cholesky_ex(Aprime)
        if not bool(info.any()):
            warnings.warn(f"A not p.d., added jitter of {jitter_new:.1e} to the diagonal", RuntimeWarning)
            return L.transpose(-1, -2) if upper else L
    raise NotPSDError(f"Matrix not positive definite after repeatedly adding jitter up to {jitter_new:.1e}.")
