"""Host restatement of exp_nonpos (csrc/common.cuh), the FP64 exp of the Gram kernels: same constants, same operation
order (numpy float64 has no fused multiply-add; the reduction constants are chosen so that n * hi is exact, which is
what the fma would give), checked against numpy's correctly rounded-to-1-ulp exp over the range the kernels use."""
import re
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def constants():
    """Read the literals out of the CUDA source (in order of appearance), so that this test checks what is compiled."""
    src = open(os.path.join(ROOT, "deconditional_downscaling_b200", "csrc", "common.cuh")).read()
    body = src[src.index("double exp_nonpos(double x) {"):]
    body = body[:body.index("template <> __device__ __forceinline__ double t_exp<double>")]
    body = re.sub(r"//[^\n]*", "", body)
    lit = [float(v) for v in re.findall(r"(?<![\w.])(\d+\.\d+(?:e[+-]?\d+)?)", body)]
    # x * L2E4 | -nd * hi | -nd * lo | p = c9 | eight fma constants c8 .. c1 | c0 | table (4) | cut-off | 0.0
    l2e4, hi, lo = lit[0], lit[1], lit[2]
    coef = lit[3:13]
    tab = lit[13:17]
    assert lit[17] == 700.0
    return l2e4, hi, lo, coef, tab


def exp_nonpos(x):
    l2e4, hi, lo, coef, tab = constants()
    x = np.asarray(x, dtype=np.float64)
    n = np.rint(x * l2e4)
    t = (x - n * hi) - n * lo
    p = np.full_like(t, coef[0])
    for c in coef[1:]:
        p = p * t + c
    ni = n.astype(np.int64)
    out = p * np.array(tab)[ni & 3] * np.ldexp(1.0, ni >> 2)
    return np.where(x < -700.0, 0.0, out)


def test_constants_are_what_the_derivation_says():
    l2e4, hi, lo, coef, tab = constants()
    assert abs(l2e4 - 4 / np.log(2)) < 1e-15
    assert abs((hi + lo) - np.log(2) / 4) < 1e-17 and (np.float64(hi).view(np.uint64) & np.uint64((1 << 27) - 1)) == 0
    assert len(coef) == 10 and coef[-1] == 1.0 and coef[-2] == 1.0 and abs(coef[0] - 1 / 362880) < 1e-20
    assert len(tab) == 4 and all(abs(tab[k] - 2 ** (k / 4)) < 1e-15 for k in range(4))


def test_exp_nonpos_is_within_three_ulp_of_exp():
    rng = np.random.default_rng(0)
    x = -np.abs(np.concatenate([rng.uniform(0, 1, 300000), rng.uniform(0, 50, 300000), rng.uniform(0, 699, 200000),
                                rng.exponential(1e-3, 100000), [0.0, 1e-300, 1e-17, 0.5 * np.log(2) / 4, 699.999]]))
    got, ref = exp_nonpos(x), np.exp(x)
    ulp = np.abs(got - ref) / np.spacing(ref)
    assert ulp.max() <= 3.0, ulp.max()          # 2 ulp against exact arithmetic + numpy's own 1 ulp
    assert exp_nonpos(np.array([0.0]))[0] == 1.0
    assert (exp_nonpos(np.array([-700.5, -1e4, -1e300])) == 0).all()
