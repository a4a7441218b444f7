"""Evaluation-side helpers: posterior summaries over large prediction sets and the
chunked negative log-likelihood of the reference's trainers.

* `posterior_mean_variance(predict, x, chunk)`: q(f) / exact-posterior mean and marginal variance over an
  arbitrarily large set of individuals (the 254 898-pixel HR grid of
  experiments/downscaling/models/variational_cmp.py:256-274) in row chunks, so that the M x N* whitened cross
  covariance never exceeds `chunk` columns.
* `chunked_nll(predict, x, t, chunk_size)`: experiments/swiss_roll/core/metrics.py:19-32 - mean over chunks of
  -log N(t_chunk | mean, dense covariance) / chunk_size, NaN when a chunk's covariance is not positive definite.
Both only call the model's own `predict` / `__call__`; every heavy operation inside runs on libdcgp."""
import math

import torch

from .functional import NotPSDError


def posterior_mean_variance(predict, x, chunk=65536):
    """`predict(x_chunk)` -> distribution with `.mean` and `.variance`.  Returns (mean, variance) over all rows."""
    means, variances = [], []
    with torch.no_grad():
        for xc in x.split(chunk):
            post = predict(xc)
            means.append(post.mean)
            variances.append(post.variance)
    return torch.cat(means), torch.cat(variances)


def chunked_nll(predict, x, t, chunk_size):
    try:
        nlls = []
        with torch.no_grad():
            for xc, tc in zip(x.split(chunk_size), t.split(chunk_size)):
                post = predict(xc)
                nlls.append(float((-post.log_prob(tc) / chunk_size).item()))
        return sum(nlls) / len(nlls)
    except NotPSDError:
        return math.nan
