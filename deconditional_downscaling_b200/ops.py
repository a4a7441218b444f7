"""Thin Python launchers over the C ABI (include/dcgp.h) and the
torch.autograd.Function wrappers with hand-written backward passes.

torch is plumbing here: it owns device memory and the current stream; every
O(N^2)/O(N^3) operation is a libdcgp kernel.  No function in this module has a
CPU or pure-torch fallback: non-CUDA tensors raise.
"""
import ctypes
import os
import weakref
from typing import List, Optional, Sequence

import torch

from . import _lib
from ._lib import F32, F64, GEMM_LOWER, GRAM_SAME, KernelSpec, MAX_DIMS, check

KINDS = {"rbf": _lib.KIND_RBF, "matern32": _lib.KIND_MATERN32, "matern12": _lib.KIND_MATERN12,
         "matern52": _lib.KIND_MATERN52}


# FP32 contraction precision.  Measured on the bench workload (tools/acc_probe.py, profiles/):
# single-pass TF32 inside the Cholesky factorisation and the triangular solves of the bag
# covariance (cond ~ 1e4) puts the ELBO 5e-3 off the FP64 value, outside the 1e-3 tolerance;
# with those on error-compensated 3xTF32 and plain products on single-pass TF32 the ELBO is
# 5e-5 off.  Hence the default "auto":
#   "fast" = single-pass TF32 everywhere, "x3" = 3xTF32 (~FP32 accuracy) everywhere,
#   "auto" = 3xTF32 for dcgp_potrf / dcgp_trsm / dcgp_potrs and for small GEMMs,
#            single-pass TF32 (tcgen05 kernel) for the large products.
TF32_MODE = "auto"
USE_TCGEN05 = True   # False keeps FP32 contractions on the mma.sync kernel (A/B comparisons)
X3_GEMM_LIMIT = 512 ** 3
TC_ALIGN_MIN = 1 << 33   # FP32 products with m n k at least this large get 16-byte aligned operand rows (see _tma_rows)
X3_SOLVE_LIMIT = 1 << 62  # dcgp_trsm / dcgp_potrs: always compensated in "auto"
X3_POTRF_LIMIT = 1 << 62  # dcgp_potrf


def _x3(t, small):
    notc = 0 if USE_TCGEN05 else _lib.GEMM_NO_TCGEN05
    if t.dtype != torch.float32 or TF32_MODE == "fast":
        return notc
    return (_lib.GEMM_TF32X3 if (TF32_MODE == "x3" or small) else 0) | notc


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float64:
        return F64
    if t.dtype == torch.float32:
        return F32
    raise TypeError(f"libdcgp supports float64 and float32, got {t.dtype}")


def _req(*tensors):
    """CUDA tensors only (no CPU fallback), all on the CURRENT device: the launchers pass the current device's stream
    and libdcgp keeps per-device state (kernel attributes, look-ahead lanes) for the device that is current at the
    call - a tensor living elsewhere would be launched on the wrong device."""
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise _lib.DcgpError("libdcgp ops need CUDA tensors (there is no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise _lib.DcgpError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: "
                                 "wrap the call in `with torch.cuda.device(t.device)`")


def _same_dtype(what, *tensors):
    dts = {t.dtype for t in tensors if t is not None}
    if len(dts) > 1:
        raise TypeError(f"{what}: operands must share one dtype, got {sorted(str(d) for d in dts)}")


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _rowmajor(t: torch.Tensor) -> torch.Tensor:
    """2-D tensor whose last dim is contiguous (row stride arbitrary)."""
    if t.dim() == 1:
        t = t.unsqueeze(-1)
    if t.stride(-1) != 1 or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        t = t.contiguous()
    return t


def _inplace_ok(t: torch.Tensor):
    if t.dim() != 2 or (t.size(1) > 1 and t.stride(1) != 1) or (t.size(0) > 1 and t.stride(0) < t.size(1)):
        raise ValueError("in-place libdcgp ops need a 2-D row-major tensor")


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.size(0) > 1 else max(t.size(1), 1)


# ---- operands of single-pass TF32 products, rounded to nearest ------------------------------------------------
# tcgen05.mma reads FP32 operands by truncation: on raw data every single-pass product comes out about 7e-4 too small
# (measured: tests/test_parity_scale_gpu.py before this rounding - gradients of the config-5 minibatch 2e-3 .. 7e-3 off
# the FP64 oracle after three or four chained products).  Rounding the operands to the nearest TF32 value first (what
# cuBLAS TF32 does inside its kernels) makes the products unbiased.  A tensor is rounded once and the copy reused
# while the tensor object is alive and unmodified (A = Lambda^{-1} B and A~ feed half a dozen products each).
ROUND_TF32 = os.environ.get("DCGP_ROUND_TF32", "1") != "0"
# (k(x1, x2) + c I) W of FP32 models on the fused tcgen05 kernel of csrc/gram_mm.cu (Gram tiles generated in shared
# memory, never written); "0" keeps the panel path: Gram rows into a workspace + GEMM
FUSED_GRAM_MATMUL = os.environ.get("DCGP_FUSED_GRAM_MATMUL", "1") != "0"
_RNA = {}


def round_tf32(X):
    """Copy of the FP32 matrix X with every value rounded to the nearest TF32 value (dcgp_round_tf32)."""
    X = _rowmajor(X)
    _req(X)
    if X.dtype != torch.float32:
        raise TypeError("round_tf32 is an FP32 operation")
    # rows stay 16-byte aligned whatever the column count (TMA operand of the tcgen05 kernel)
    out = torch.empty(X.size(0), -(-X.size(1) // 4) * 4, dtype=X.dtype, device=X.device)[:, :X.size(1)]
    check(_lib.load().dcgp_round_tf32(_ptr(X), _ptr(out), X.size(0), X.size(1), _ld(X), _ld(out), _stream()), "dcgp_round_tf32")
    return out


def _rna(t):
    """round_tf32(t), cached per tensor object (weak reference + version counter: a freed or modified tensor never
    hits; a new view of the same memory simply misses)."""
    if not ROUND_TF32:
        return t
    key = id(t)
    hit = _RNA.get(key)
    if hit is not None and hit[0]() is t and hit[1] == t._version:
        return hit[2]
    out = round_tf32(t)
    if len(_RNA) > 256:
        for k in [k for k, v in _RNA.items() if v[0]() is None]:
            del _RNA[k]
    _RNA[key] = (weakref.ref(t, lambda _, k=key: _RNA.pop(k, None)), t._version, out)
    return out


def _single_pass_tc(dtype, m, n, k, flags):
    """Mirrors the dispatch of dcgp_gemm (csrc/gemm.cu): FP32, no 3xTF32, a shape the tcgen05 kernel takes."""
    return (dtype == torch.float32 and USE_TCGEN05 and not (flags & (_lib.GEMM_TF32X3 | _lib.GEMM_NO_TCGEN05)) and m >= 128
            and n >= 128 and k >= 64 and ((m + 127) // 128) * ((n + 127) // 128) >= 32)


class FlatSpec:
    """Additive kernel = list of (kind, dims, lengthscale[nd], outputscale[]) with the
    hyper-parameters as device tensors (already softplus-transformed)."""

    def __init__(self, kinds: Sequence[str], dims: Sequence[Sequence[int]], lengthscales: Sequence[torch.Tensor],
                 outputscales: Sequence[torch.Tensor]):
        self.kinds, self.dims = list(kinds), [list(d) for d in dims]
        self.lengthscales = [l.reshape(-1).contiguous() for l in lengthscales]
        self.outputscales = [o.reshape(1).contiguous() for o in outputscales]
        if not (1 <= len(self.kinds) <= _lib.MAX_COMPONENTS):
            raise ValueError("1..4 additive components supported")

    def cast(self, dtype):
        return FlatSpec(self.kinds, self.dims, [l.to(dtype) for l in self.lengthscales],
                        [o.to(dtype) for o in self.outputscales])

    def detach(self):
        return FlatSpec(self.kinds, self.dims, [l.detach() for l in self.lengthscales],
                        [o.detach() for o in self.outputscales])

    def c_spec(self) -> KernelSpec:
        s = KernelSpec()
        s.n_components = len(self.kinds)
        for c, (kind, dims, ls, os_) in enumerate(zip(self.kinds, self.dims, self.lengthscales, self.outputscales)):
            _req(ls, os_)
            if ls.numel() != len(dims):
                raise ValueError("lengthscale size must match active dims")
            s.comp[c].kind = KINDS[kind]
            s.comp[c].ndims = len(dims)
            for j, d in enumerate(dims):
                s.comp[c].dims[j] = int(d)
            s.comp[c].lengthscale = ls.data_ptr()
            s.comp[c].outputscale = os_.data_ptr()
        return s


# --------------------------------------------------------------------------------------
# raw launchers (no autograd)
# --------------------------------------------------------------------------------------
def gram(spec: FlatSpec, x1, x2=None, diag_dev=None, diag_const=0.0, out=None):
    same = x2 is None
    x1 = _rowmajor(x1)
    x2 = x1 if same else _rowmajor(x2)
    _req(x1, x2, diag_dev)
    n1, n2 = x1.size(0), x2.size(0)
    K = out if out is not None else torch.empty(n1, n2, dtype=x1.dtype, device=x1.device)
    cs = spec.c_spec()
    rc = _lib.load().dcgp_gram(ctypes.byref(cs), _ptr(x1), n1, _ld(x1), _ptr(x2), n2, _ld(x2), _ptr(K), _ld(K),
                               _ptr(diag_dev), float(diag_const), _dt(x1), GRAM_SAME if same else 0, _stream())
    check(rc, "dcgp_gram")
    return K


def gram_matmul(spec: FlatSpec, x1, x2, W, diag_dev=None, diag_const=0.0):
    """(k(x1, x2) [+ nugget when x2 is None]) @ W with the Gram matrix held only as bounded row panels."""
    same = x2 is None
    x1 = _rowmajor(x1)
    x2 = x1 if same else _rowmajor(x2)
    W = _rowmajor(W)
    _req(x1, x2, W, diag_dev)
    n1, n2, m = x1.size(0), x2.size(0), W.size(1)
    if W.size(0) != n2:
        raise ValueError(f"gram_matmul: W has {W.size(0)} rows, the Gram matrix {n2} columns")
    out = torch.empty(n1, m, dtype=x1.dtype, device=x1.device)
    lib = _lib.load()
    cs = spec.c_spec()
    flags = (GRAM_SAME if same else 0) | _x3(x1, n1 * n2 * m <= X3_GEMM_LIMIT)
    fused = False
    if x1.dtype == torch.float32 and USE_TCGEN05 and not (flags & (_lib.GEMM_TF32X3 | _lib.GEMM_NO_TCGEN05)):
        if FUSED_GRAM_MATMUL and n1 >= 128 and n2 >= 64 and 32 <= m <= 512 and n1 * n2 >= (1 << 22):
            W = _tma_rows(W)   # the fused kernel (csrc/gram_mm.cu) reads W through TMA: 16-byte aligned rows
            fused = True       # mirrors dcgp_gram_matmul_fused_f32's own conditions: no panel workspace needed
        if fused or _single_pass_tc(x1.dtype, n1, m, n2, flags):
            W = _rna(W)        # the Gram tiles / panels are rounded to nearest TF32 by their generator
    if not fused:
        flags |= _lib.GRAM_NO_FUSE
    ws = _workspace(16 if fused else lib.dcgp_gram_matmul_workspace(n1, n2, _dt(x1)), x1.device)
    rc = lib.dcgp_gram_matmul(ctypes.byref(cs), _ptr(x1), n1, _ld(x1), _ptr(x2), n2, _ld(x2), _ptr(diag_dev), float(diag_const),
                              _ptr(W), m, _ld(W), _ptr(out), _ld(out), _ptr(ws), ws.numel(), _dt(x1), flags, _stream())
    check(rc, "dcgp_gram_matmul")
    return out


def syrk(A, trans=False, alpha=1.0, beta=0.0, C=None):
    """Lower triangle of alpha * op(A) op(A)^T + beta * C (the strict upper triangle of C is left untouched)."""
    A = _rowmajor(A)
    _req(A, C)
    n, k = (A.size(1), A.size(0)) if trans else (A.size(0), A.size(1))
    if C is None:
        C = torch.zeros(n, n, dtype=A.dtype, device=A.device)
        beta = 0.0
    rc = _lib.load().dcgp_syrk(int(trans), n, k, float(alpha), _ptr(A), _ld(A), float(beta), _ptr(C), _ld(C), _dt(A),
                               _x3(A, n * n * k <= X3_GEMM_LIMIT), _stream())
    check(rc, "dcgp_syrk")
    return C


def _zero_theta(nc, dtype, device):
    """Adjoint accumulators of the hyper-parameters, (dls [nc, MAX_DIMS], dos [nc]), zeroed by ONE fill."""
    buf = torch.zeros(nc * (MAX_DIMS + 1), dtype=dtype, device=device)
    return buf[:nc * MAX_DIMS].view(nc, MAX_DIMS), buf[nc * MAX_DIMS:]


def gram_bwd(spec: FlatSpec, x1, x2, G, need_dx1=False):
    """Returns (dls [ncomp, MAX_DIMS], dos [ncomp], dx1 or None)."""
    same = x2 is None
    x1 = _rowmajor(x1)
    x2 = x1 if same else _rowmajor(x2)
    G = _rowmajor(G)
    _req(x1, x2, G)
    nc = len(spec.kinds)
    dls, dos = _zero_theta(nc, x1.dtype, x1.device)
    dx1 = torch.zeros_like(x1) if need_dx1 else None
    cs = spec.c_spec()
    rc = _lib.load().dcgp_gram_bwd(ctypes.byref(cs), _ptr(x1), x1.size(0), _ld(x1), _ptr(x2), x2.size(0), _ld(x2),
                                   _ptr(G), _ld(G), _ptr(dls), _ptr(dos), _ptr(dx1), _ld(dx1) if need_dx1 else 0,
                                   _dt(x1), GRAM_SAME if same else 0, _stream())
    check(rc, "dcgp_gram_bwd")
    return dls, dos, dx1


def _tma_rows(t):
    """t itself if its rows start on 16-byte boundaries, else a copy laid out with the row stride rounded up to 4 floats.
    TMA cannot address unaligned rows, so e.g. an N x 327 operand would send a large product to the scalar-load
    mma.sync kernel (N = 32768: 51 instead of 390 TF/s); the O(rows x cols) copy is noise next to such a product."""
    if (_ld(t) & 3) == 0 and (t.data_ptr() & 15) == 0:
        return t
    buf = torch.empty(t.size(0), -(-t.size(1) // 4) * 4, dtype=t.dtype, device=t.device)[:, :t.size(1)]
    buf.copy_(t)
    return buf


def gemm(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None, lower=False, no_splitk=False):
    """C = alpha * op(A) @ op(B) + beta * C."""
    A, B = _rowmajor(A), _rowmajor(B)
    _req(A, B, C)
    _same_dtype("gemm", A, B, C)
    m, k = (A.size(1), A.size(0)) if transa else (A.size(0), A.size(1))
    k2, n = (B.size(1), B.size(0)) if transb else (B.size(0), B.size(1))
    if k != k2:
        raise ValueError(f"gemm inner dims differ: {k} vs {k2}")
    if A.dtype == torch.float32 and USE_TCGEN05 and m * n * k >= TC_ALIGN_MIN:
        A, B = _tma_rows(A), _tma_rows(B)
    if C is None:
        C = torch.empty(m, n, dtype=A.dtype, device=A.device)
        beta = 0.0
    flags = (GEMM_LOWER if lower else 0) | (_lib.GEMM_NO_SPLITK if no_splitk else 0) | _x3(A, m * n * k <= X3_GEMM_LIMIT)
    if _single_pass_tc(A.dtype, m, n, k, flags):
        A, B = _rna(A), _rna(B)
    rc = _lib.load().dcgp_gemm(int(transa), int(transb), m, n, k, float(alpha), _ptr(A), _ld(A), _ptr(B), _ld(B),
                               float(beta), _ptr(C), _ld(C), _dt(A), flags, _stream())
    check(rc, "dcgp_gemm")
    return C


def split_lo(X):
    """x - trunc_tf32(x): the part of an FP32 matrix that the tcgen05 TF32 operand read drops."""
    X = _rowmajor(X)
    _req(X)
    if X.dtype != torch.float32:
        raise TypeError("split_lo is an FP32 operation")
    out = torch.empty(X.size(0), X.size(1), dtype=X.dtype, device=X.device)
    check(_lib.load().dcgp_split_lo(_ptr(X), _ptr(out), X.size(0), X.size(1), _ld(X), _ld(out), _stream()), "dcgp_split_lo")
    return out


def gemm_x3(A, B, transa=False, transb=False, alpha=1.0, beta=0.0, C=None, Alo=None, Blo=None):
    """FP32 product at ~FP32 accuracy: 3xTF32 on the tcgen05 kernel from hi/lo operand pairs (the remainders
    are computed here unless the caller already holds them)."""
    A, B = _rowmajor(A), _rowmajor(B)
    _req(A, B, C)
    _same_dtype("gemm_x3", A, B, C, Alo, Blo)
    if A.dtype != torch.float32:
        return gemm(A, B, transa=transa, transb=transb, alpha=alpha, beta=beta, C=C)
    m, k = (A.size(1), A.size(0)) if transa else (A.size(0), A.size(1))
    k2, n = (B.size(1), B.size(0)) if transb else (B.size(0), B.size(1))
    if k != k2:
        raise ValueError(f"gemm inner dims differ: {k} vs {k2}")
    if C is None:
        C = torch.empty(m, n, dtype=A.dtype, device=A.device)
        beta = 0.0
    Alo = split_lo(A) if Alo is None else _rowmajor(Alo)
    Blo = split_lo(B) if Blo is None else _rowmajor(Blo)
    flags = _lib.GEMM_TF32X3 | (0 if USE_TCGEN05 else _lib.GEMM_NO_TCGEN05)
    rc = _lib.load().dcgp_gemm_ex(int(transa), int(transb), m, n, k, float(alpha), _ptr(A), _ld(A), _ptr(Alo), _ptr(B), _ld(B),
                                  _ptr(Blo), float(beta), _ptr(C), _ld(C), None, _dt(A), flags, _stream())
    check(rc, "dcgp_gemm_ex")
    return C


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


def potrf_(A, info=None):
    """In-place lower Cholesky of the (row-major) square matrix A.  Returns the device info tensor."""
    _inplace_ok(A)
    _req(A)
    n = A.size(0)
    if info is None:
        info = torch.zeros(1, dtype=torch.int32, device=A.device)
    lib = _lib.load()
    ws = _workspace(lib.dcgp_potrf_workspace(n, _dt(A)), A.device)
    rc = lib.dcgp_potrf(_ptr(A), n, _ld(A), _ptr(info), _ptr(ws), ws.numel(), _dt(A), _x3(A, n <= X3_POTRF_LIMIT), _stream())
    check(rc, "dcgp_potrf")
    return info


def solve_plan(L):
    """Prepared form of a Cholesky factor for repeated solves (dcgp_solve_plan); None for small factors."""
    L = _rowmajor(L)
    _req(L)
    lib = _lib.load()
    nbytes = lib.dcgp_solve_plan_bytes(L.size(0), _ld(L), _dt(L))
    if nbytes == 0:
        return None
    plan = _workspace(nbytes, L.device)
    rc = lib.dcgp_solve_plan(_ptr(L), L.size(0), _ld(L), _ptr(plan), plan.numel(), _dt(L), _x3(L, L.size(0) <= X3_SOLVE_LIMIT), _stream())
    check(rc, "dcgp_solve_plan")
    return plan


def _solve(trans, L, B, plan):
    L = _rowmajor(L)
    _inplace_ok(B)
    _req(L, B)
    _same_dtype("triangular solve", L, B)
    n, m = B.size(0), B.size(1)
    lib = _lib.load()
    # sized for B's row stride: the 512-block sweeps work in B's layout (a view with ld > m would otherwise drop to the
    # 128-block path: FP64 n = 8192, m = 81, ld = 84 took 5.9 instead of 1.9 ms)
    ws = _workspace(lib.dcgp_trsm_workspace(n, max(m, _ld(B)), _dt(L)), L.device)
    flags = _x3(L, n <= X3_SOLVE_LIMIT)
    if plan is not None:
        rc = lib.dcgp_trsm_planned(trans, _ptr(L), n, _ld(L), _ptr(plan), plan.numel(), _ptr(B), m, _ld(B), _ptr(ws), ws.numel(),
                                   _dt(L), flags, _stream())
        check(rc, "dcgp_trsm_planned")
    else:
        rc = lib.dcgp_trsm(trans, _ptr(L), n, _ld(L), _ptr(B), m, _ld(B), _ptr(ws), ws.numel(), _dt(L), flags, _stream())
        check(rc, "dcgp_trsm")
    return B


def trsm_(L, B, trans=False, plan=None):
    """In-place solve op(L) X = B, B is (n, m) row-major."""
    return _solve(int(trans), L, B, plan)


def potrs_(L, B, plan=None):
    """In-place (L L^T) X = B."""
    return _solve(2, L, B, plan)


def trtri(L):
    """W = L^{-1} of a lower-triangular factor (dense, zero above the diagonal)."""
    L = _rowmajor(L)
    _req(L)
    n = L.size(0)
    W = torch.empty(n, n, dtype=L.dtype, device=L.device)
    lib = _lib.load()
    ws = _workspace(lib.dcgp_trtri_workspace(n, _dt(L)), L.device)
    rc = lib.dcgp_trtri(_ptr(L), n, _ld(L), _ptr(W), n, _ptr(ws), ws.numel(), _dt(L), _x3(L, n <= X3_SOLVE_LIMIT), _stream())
    check(rc, "dcgp_trtri")
    return W


def logdet_chol(L):
    L = _rowmajor(L)
    _req(L)
    out = torch.empty(1, dtype=L.dtype, device=L.device)
    check(_lib.load().dcgp_logdet_chol(_ptr(L), L.size(0), _ld(L), _ptr(out), _dt(L), _stream()), "dcgp_logdet_chol")
    return out[0]


def bag_sum(X, offsets):
    X = _rowmajor(X)
    _req(X, offsets)
    nb = offsets.numel() - 1
    out = torch.empty(nb, X.size(1), dtype=X.dtype, device=X.device)
    check(_lib.load().dcgp_bag_sum(_ptr(X), _ld(X), X.size(1), _ptr(offsets), nb, _ptr(out), _ld(out), _dt(X),
                                   _stream()), "dcgp_bag_sum")
    return out


def gram_bagsum(spec, z, x, offsets):
    z, x = _rowmajor(z), _rowmajor(x)
    _req(z, x, offsets)
    nb = offsets.numel() - 1
    out = torch.empty(nb, z.size(0), dtype=z.dtype, device=z.device)
    cs = spec.c_spec()
    check(_lib.load().dcgp_gram_bagsum(ctypes.byref(cs), _ptr(z), z.size(0), _ld(z), _ptr(x), _ld(x), _ptr(offsets),
                                       nb, _ptr(out), _ld(out), _dt(z), 0, _stream()), "dcgp_gram_bagsum")
    return out


def gram_bagsum_bwd(spec, z, x, offsets, G, need_dz=True):
    z, x, G = _rowmajor(z), _rowmajor(x), _rowmajor(G)
    _req(z, x, offsets, G)
    nb = offsets.numel() - 1
    nc = len(spec.kinds)
    dls, dos = _zero_theta(nc, z.dtype, z.device)
    dz = torch.zeros_like(z) if need_dz else None
    cs = spec.c_spec()
    check(_lib.load().dcgp_gram_bagsum_bwd(ctypes.byref(cs), _ptr(z), z.size(0), _ld(z), _ptr(x), _ld(x),
                                           _ptr(offsets), nb, _ptr(G), _ld(G), _ptr(dls), _ptr(dos), _ptr(dz),
                                           _ld(dz) if need_dz else 0, _dt(z), 0, _stream()), "dcgp_gram_bagsum_bwd")
    return dls, dos, dz


def gram_bag_blocksum(spec, x, offsets):
    x = _rowmajor(x)
    _req(x, offsets)
    nb = offsets.numel() - 1
    out = torch.empty(nb, dtype=x.dtype, device=x.device)
    cs = spec.c_spec()
    check(_lib.load().dcgp_gram_bag_blocksum(ctypes.byref(cs), _ptr(x), _ld(x), _ptr(offsets), nb, _ptr(out), _dt(x),
                                             0, _stream()), "dcgp_gram_bag_blocksum")
    return out


def gram_bag_blocksum_bwd(spec, x, offsets, g):
    x = _rowmajor(x)
    g = g.contiguous()
    _req(x, offsets, g)
    nb = offsets.numel() - 1
    nc = len(spec.kinds)
    dls, dos = _zero_theta(nc, x.dtype, x.device)
    cs = spec.c_spec()
    check(_lib.load().dcgp_gram_bag_blocksum_bwd(ctypes.byref(cs), _ptr(x), _ld(x), _ptr(offsets), nb, _ptr(g),
                                                 _ptr(dls), _ptr(dos), _dt(x), 0, _stream()),
          "dcgp_gram_bag_blocksum_bwd")
    return dls, dos


def whitened_kl(m, Ls, need_grad=False):
    m = m.contiguous()
    Ls = _rowmajor(Ls)
    _req(m, Ls)
    out = torch.zeros(1, dtype=m.dtype, device=m.device)
    dm = torch.empty_like(m) if need_grad else None
    dLs = torch.empty(Ls.size(0), Ls.size(1), dtype=m.dtype, device=m.device) if need_grad else None
    check(_lib.load().dcgp_whitened_kl(_ptr(m), _ptr(Ls), m.numel(), _ld(Ls), _ptr(out), _ptr(dm), _ptr(dLs),
                                       _ld(dLs) if need_grad else 0, _dt(m), _stream()), "dcgp_whitened_kl")
    return out[0], dm, dLs


def vbagg_ell(z, S, T, offsets, noise):
    """Returns (ell, dS, dT, dnoise) of the bag-level expected log-likelihood (dcgp_vbagg_ell)."""
    z, S, T = z.contiguous(), S.contiguous(), T.contiguous()
    noise = noise.reshape(1).contiguous()
    _req(z, S, T, noise)
    n = z.numel()
    ell = torch.empty(1, dtype=z.dtype, device=z.device)
    dS, dT, dn = torch.empty_like(S), torch.empty_like(T), torch.empty_like(noise)
    check(_lib.load().dcgp_vbagg_ell(_ptr(z), _ptr(S), _ptr(T), _ptr(offsets), n, _ptr(noise), _ptr(ell), _ptr(dS), _ptr(dT),
                                     _ptr(dn), _dt(z), _stream()), "dcgp_vbagg_ell")
    return ell[0], dS, dT, dn


def rff_features(x, W, lengthscale):
    """Phi = [cos(x W / l), sin(x W / l)]; x: N x d (active dims only), W: d x D."""
    x, W = _rowmajor(x), _rowmajor(W)
    ls = lengthscale.reshape(-1).contiguous()
    _req(x, W, ls)
    n, d, D = x.size(0), x.size(1), W.size(1)
    out = torch.empty(n, 2 * D, dtype=x.dtype, device=x.device)
    check(_lib.load().dcgp_rff_features(_ptr(x), n, _ld(x), d, _ptr(W), D, _ptr(ls), _ptr(out), _ld(out), _dt(x),
                                        _stream()), "dcgp_rff_features")
    return out


def rff_features_bwd(x, W, lengthscale, dphi, need_dx=True):
    x, W, dphi = _rowmajor(x), _rowmajor(W), _rowmajor(dphi)
    ls = lengthscale.reshape(-1).contiguous()
    _req(x, W, ls, dphi)
    n, d, D = x.size(0), x.size(1), W.size(1)
    dx = torch.zeros_like(x) if need_dx else None
    dls = torch.zeros(d, dtype=x.dtype, device=x.device)
    check(_lib.load().dcgp_rff_features_bwd(_ptr(x), n, _ld(x), d, _ptr(W), D, _ptr(ls), _ptr(dphi), _ld(dphi),
                                            _ptr(dx), _ld(dx) if need_dx else 0, _ptr(dls), _dt(x), _stream()),
          "dcgp_rff_features_bwd")
    return dx, dls


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, grad_scale=1.0, veto=None, zero_grads=False):
    """Fused multi-tensor Adam over flat buffers (dcgp_adam_step); `step` is a device float64 counter, `veto` optional
    device int32 words (non-zero -> the update is skipped on the device)."""
    _req(params, grads, exp_avg, exp_avg_sq, step, veto)
    if not (params.dtype == grads.dtype == exp_avg.dtype == exp_avg_sq.dtype):
        raise TypeError("adam_step: parameter, gradient and moment buffers must share one dtype")
    if step.dtype != torch.float64 or (veto is not None and veto.dtype != torch.int32):
        raise TypeError("adam_step: step must be float64 and veto int32")
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not t.is_contiguous() or t.numel() != params.numel():
            raise ValueError("adam_step needs contiguous flat buffers of one length")
    with torch.cuda.device(params.device):
        rc = _lib.load().dcgp_adam_step(_ptr(params), _ptr(grads), _ptr(exp_avg), _ptr(exp_avg_sq), _ptr(step), params.numel(),
                                        float(lr), float(beta1), float(beta2), float(eps), float(grad_scale), _ptr(veto),
                                        0 if veto is None else veto.numel(), 1 if zero_grads else 0, _dt(params),
                                        _stream(params.device))
    check(rc, "dcgp_adam_step")


def vbagg_elbo(spec: FlatSpec, Z, x, offsets, z, var_mean, chol_var, noise, num_data, beta=1.0, mean_const=None,
               scale=-1.0, need_grads=True, jitter=1e-3, var_jitter=1e-4, info=None):
    """Whole VBagg minibatch ELBO through dcgp_vbagg_elbo_fwd / _bwd: returns (elbo [device scalar], grads dict or None,
    info).  grads = scale * d ELBO / d {Z, lengthscales, outputscales, var_mean, chol_var, noise, mean_const}, with
    respect to the hyper-parameter VALUES (the softplus chain of the raw parameters is the caller's)."""
    Z, x = _rowmajor(Z), _rowmajor(x)
    z, var_mean, chol_var = z.contiguous(), var_mean.contiguous(), _rowmajor(chol_var)
    noise = noise.reshape(1).contiguous()
    _req(Z, x, offsets, z, var_mean, chol_var, noise, mean_const)
    _same_dtype("vbagg_elbo", Z, x, z, var_mean, chol_var, noise)
    lib = _lib.load()
    M, nb, dt, dev = Z.size(0), offsets.numel() - 1, Z.dtype, Z.device
    cs = spec.c_spec()
    p = _lib.VbaggProblem(ctypes.pointer(cs), _ptr(Z), M, _ld(Z), _ptr(x), _ld(x), _ptr(offsets), nb, _ptr(z),
                          _ptr(var_mean), _ptr(chol_var), _ld(chol_var), _ptr(noise), _ptr(mean_const), float(num_data),
                          float(beta), float(jitter), float(var_jitter))
    nbytes = lib.dcgp_vbagg_elbo_workspace(M, nb, _ld(Z), _dt(Z))
    ws = _workspace(nbytes, dev)
    elbo = torch.empty(1, dtype=dt, device=dev)
    if info is None:
        info = torch.zeros(1, dtype=torch.int32, device=dev)
    check(lib.dcgp_vbagg_elbo_fwd(ctypes.byref(p), _ptr(elbo), _ptr(info), _ptr(ws), nbytes, _dt(Z), 0, _stream()),
          "dcgp_vbagg_elbo_fwd")
    if not need_grads:
        return elbo[0], None, info
    nc = len(spec.kinds)
    g = {"Z": torch.empty_like(Z), "dls": torch.empty(nc, MAX_DIMS, dtype=dt, device=dev),
         "dos": torch.empty(nc, dtype=dt, device=dev), "var_mean": torch.empty_like(var_mean),
         "chol_var": torch.empty(M, M, dtype=dt, device=dev), "noise": torch.empty(1, dtype=dt, device=dev),
         "mean_const": torch.empty(1, dtype=dt, device=dev) if mean_const is not None else None}
    cg = _lib.VbaggGrads(_ptr(g["Z"]), _ld(g["Z"]), _ptr(g["dls"]), _ptr(g["dos"]), _ptr(g["var_mean"]),
                         _ptr(g["chol_var"]), M, _ptr(g["noise"]), _ptr(g["mean_const"]))
    check(lib.dcgp_vbagg_elbo_bwd(ctypes.byref(p), float(scale), ctypes.byref(cg), _ptr(ws), nbytes, _dt(Z), 0, _stream()),
          "dcgp_vbagg_elbo_bwd")
    return elbo[0], g, info


def cmp_elbo(k_spec: FlatSpec, l_spec: FlatSpec, Z, x, ext, y, z, var_mean, chol_var, noise, ind_noise, lbda, num_data,
             beta=1.0, mean_const=None, scale=-1.0, need_grads=True, jitter=1e-3, var_jitter=1e-4, info=None):
    """Whole VariationalCMP minibatch ELBO (with individuals noise) through dcgp_cmp_elbo_fwd / _bwd: returns
    (elbo [device scalar], grads dict or None, info [3 ints: Lambda, K_ZZ, C]).  grads = scale * d ELBO / d {Z, k and l
    lengthscales / outputscales, var_mean, chol_var, noise, ind_noise, mean_const} with respect to the VALUES."""
    Z, x, ext, y = _rowmajor(Z), _rowmajor(x), _rowmajor(ext), _rowmajor(y)
    z, var_mean, chol_var = z.contiguous(), var_mean.contiguous(), _rowmajor(chol_var)
    noise, ind_noise = noise.reshape(1).contiguous(), ind_noise.reshape(1).contiguous()
    _req(Z, x, ext, y, z, var_mean, chol_var, noise, ind_noise, mean_const)
    _same_dtype("cmp_elbo", Z, x, ext, y, z, var_mean, chol_var, noise, ind_noise)
    lib = _lib.load()
    M, N, n, dt, dev = Z.size(0), x.size(0), y.size(0), Z.dtype, Z.device
    if ext.size(0) != N or z.numel() != n:
        raise ValueError("ext must have one row per individual and z one entry per bag")
    ck, cl = k_spec.c_spec(), l_spec.c_spec()
    p = _lib.CmpProblem(ctypes.pointer(ck), ctypes.pointer(cl), _ptr(Z), M, _ld(Z), _ptr(x), N, _ld(x), _ptr(ext), _ld(ext),
                        _ptr(y), n, _ld(y), _ptr(z), _ptr(var_mean), _ptr(chol_var), _ld(chol_var), _ptr(noise),
                        _ptr(ind_noise), _ptr(mean_const), float(lbda), float(num_data), float(beta), float(jitter),
                        float(var_jitter))
    nbytes = lib.dcgp_cmp_elbo_workspace(N, n, M, _ld(Z), _dt(Z))
    ws = _workspace(nbytes, dev)
    elbo = torch.empty(1, dtype=dt, device=dev)
    if info is None:
        info = torch.zeros(3, dtype=torch.int32, device=dev)
    check(lib.dcgp_cmp_elbo_fwd(ctypes.byref(p), _ptr(elbo), _ptr(info), _ptr(ws), nbytes, _dt(Z), 0, _stream()),
          "dcgp_cmp_elbo_fwd")
    if not need_grads:
        return elbo[0], None, info
    nk, nl = len(k_spec.kinds), len(l_spec.kinds)
    g = {"Z": torch.empty_like(Z), "dk_ls": torch.empty(nk, MAX_DIMS, dtype=dt, device=dev),
         "dk_os": torch.empty(nk, dtype=dt, device=dev), "dl_ls": torch.empty(nl, MAX_DIMS, dtype=dt, device=dev),
         "dl_os": torch.empty(nl, dtype=dt, device=dev), "var_mean": torch.empty_like(var_mean),
         "chol_var": torch.empty(M, M, dtype=dt, device=dev), "noise": torch.empty(1, dtype=dt, device=dev),
         "ind_noise": torch.empty(1, dtype=dt, device=dev),
         "mean_const": torch.empty(1, dtype=dt, device=dev) if mean_const is not None else None}
    cg = _lib.CmpGrads(_ptr(g["Z"]), _ld(g["Z"]), _ptr(g["dk_ls"]), _ptr(g["dk_os"]), _ptr(g["dl_ls"]), _ptr(g["dl_os"]),
                       _ptr(g["var_mean"]), _ptr(g["chol_var"]), M, _ptr(g["noise"]), _ptr(g["ind_noise"]),
                       _ptr(g["mean_const"]))
    check(lib.dcgp_cmp_elbo_bwd(ctypes.byref(p), float(scale), ctypes.byref(cg), _ptr(ws), nbytes, _dt(Z), 0, _stream()),
          "dcgp_cmp_elbo_bwd")
    return elbo[0], g, info


def gram_bwd_lowrank(spec: FlatSpec, x1, x2, P, Q, alpha=1.0, fused=False):
    """(dls [ncomp, MAX_DIMS], dos [ncomp]) = sum_ij G_ij d k(x1_i, x2_j) / d theta for G = alpha P Q^T (dcgp_gram_bwd_lowrank);
    x2 None means x1.  fused=True (FP32): G is contracted inside the tcgen05 product and never stored - pass TF32-rounded
    operands (round_tf32); shapes / kernels the fused kernel does not take fall back to the two-call route."""
    x1 = _rowmajor(x1)
    x2 = x1 if x2 is None else _rowmajor(x2)
    P, Q = _rowmajor(P), _rowmajor(Q)
    _req(x1, x2, P, Q)
    _same_dtype("gram_bwd_lowrank", x1, x2, P, Q)
    lib = _lib.load()
    n1, n2, k = x1.size(0), x2.size(0), P.size(1)
    if P.size(0) != n1 or Q.size(0) != n2 or Q.size(1) != k:
        raise ValueError("P must be n1 x k and Q n2 x k")
    nc = len(spec.kinds)
    dls, dos = _zero_theta(nc, x1.dtype, x1.device)
    cs = spec.c_spec()
    nbytes = lib.dcgp_gram_bwd_lowrank_workspace(n1, n2, _dt(x1))
    flags = _lib.GRAM_BWD_FUSE if (fused and x1.dtype == torch.float32) else 0
    call = lambda ws: lib.dcgp_gram_bwd_lowrank(ctypes.byref(cs), _ptr(x1), n1, _ld(x1), _ptr(x2), n2, _ld(x2), _ptr(P), _ld(P),
                                                _ptr(Q), _ld(Q), k, float(alpha), _ptr(dls), _ptr(dos), _ptr(ws), ws.numel(),
                                                _dt(x1), flags, _stream())
    rc = call(_workspace(16, x1.device)) if flags else -16
    if rc == -16:      # two-call route (or the fused kernel declined the shape): it needs room for G
        rc = call(_workspace(nbytes, x1.device))
    check(rc, "dcgp_gram_bwd_lowrank")
    return dls, dos
