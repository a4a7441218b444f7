"""ctypes binding of libdcgp.so (the C ABI declared in include/dcgp.h).

There is no CPU fallback: importing the ops without the built library raises.
torch is used only to obtain device pointers and the current CUDA stream.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_double, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# DCGP_LIB selects another build of the same library (e.g. the -DDCGP_LEAF_TIMING one of tools/leaf_timing.py)
LIB_PATH = os.environ.get("DCGP_LIB") or os.path.join(_HERE, "lib", "libdcgp.so")

F64, F32 = 0, 1
KIND_RBF, KIND_MATERN32, KIND_MATERN12, KIND_MATERN52 = 0, 1, 2, 3
MAX_COMPONENTS, MAX_DIMS = 4, 8
GRAM_SAME, GRAM_OUT_F64, GRAM_NO_FUSE, GRAM_BWD_FUSE = 1, 2, 256, 512
GEMM_LOWER, GEMM_NO_SPLITK, GEMM_TF32X3, GEMM_NO_TCGEN05 = 1, 2, 4, 8


class Component(Structure):
    _fields_ = [("kind", c_int32), ("ndims", c_int32), ("dims", c_int32 * MAX_DIMS),
                ("lengthscale", c_void_p), ("outputscale", c_void_p)]


class KernelSpec(Structure):
    _fields_ = [("n_components", c_int32), ("reserved", c_int32), ("comp", Component * MAX_COMPONENTS)]


class VbaggProblem(Structure):
    """dcgp_vbagg_problem (include/dcgp.h)."""
    _fields_ = [("spec", POINTER(KernelSpec)), ("Z", c_void_p), ("M", c_int64), ("ldz", c_int64), ("x", c_void_p),
                ("ldx", c_int64), ("offsets", c_void_p), ("n_bags", c_int64), ("z", c_void_p), ("var_mean", c_void_p),
                ("chol_var", c_void_p), ("ldc", c_int64), ("noise", c_void_p), ("mean_const", c_void_p),
                ("num_data", c_double), ("beta", c_double), ("jitter", c_double), ("var_jitter", c_double)]


class VbaggGrads(Structure):
    """dcgp_vbagg_grads (include/dcgp.h)."""
    _fields_ = [("dZ", c_void_p), ("lddz", c_int64), ("dls", c_void_p), ("dos", c_void_p), ("dvar_mean", c_void_p),
                ("dchol_var", c_void_p), ("lddc", c_int64), ("dnoise", c_void_p), ("dmean_const", c_void_p)]


class CmpProblem(Structure):
    """dcgp_cmp_problem (include/dcgp.h)."""
    _fields_ = [("k_spec", POINTER(KernelSpec)), ("l_spec", POINTER(KernelSpec)), ("Z", c_void_p), ("M", c_int64),
                ("ldz", c_int64), ("x", c_void_p), ("N", c_int64), ("ldx", c_int64), ("ext", c_void_p), ("ldext", c_int64),
                ("y", c_void_p), ("n", c_int64), ("ldy", c_int64), ("z", c_void_p), ("var_mean", c_void_p),
                ("chol_var", c_void_p), ("ldc", c_int64), ("noise", c_void_p), ("ind_noise", c_void_p),
                ("mean_const", c_void_p), ("lbda", c_double), ("num_data", c_double), ("beta", c_double),
                ("jitter", c_double), ("var_jitter", c_double)]


class CmpGrads(Structure):
    """dcgp_cmp_grads (include/dcgp.h)."""
    _fields_ = [("dZ", c_void_p), ("lddz", c_int64), ("dk_ls", c_void_p), ("dk_os", c_void_p), ("dl_ls", c_void_p),
                ("dl_os", c_void_p), ("dvar_mean", c_void_p), ("dchol_var", c_void_p), ("lddc", c_int64),
                ("dnoise", c_void_p), ("dind_noise", c_void_p), ("dmean_const", c_void_p)]


class DcgpError(RuntimeError):
    pass


_SIGNATURES = {
    "dcgp_version": (c_int, []),
    "dcgp_arch": (c_int, []),
    "dcgp_launch_count": (ctypes.c_ulonglong, []),
    "dcgp_gram": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                          c_int64, c_void_p, c_double, c_int, c_int, c_void_p]),
    "dcgp_gram_bwd": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                              c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "dcgp_gram_bwd_lowrank_workspace": (c_size_t, [c_int64, c_int64, c_int]),
    "dcgp_gram_bwd_lowrank": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                      c_int64, c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p, c_void_p, c_size_t, c_int,
                                      c_int, c_void_p]),
    "dcgp_gemm": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_double, c_void_p, c_int64, c_void_p, c_int64,
                          c_double, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "dcgp_split_lo": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "dcgp_round_tf32": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "dcgp_gemm_ex": (c_int, [c_int, c_int, c_int64, c_int64, c_int64, c_double, c_void_p, c_int64, c_void_p, c_void_p,
                             c_int64, c_void_p, c_double, c_void_p, c_int64, c_void_p, c_int, c_int, c_void_p]),
    "dcgp_syrk": (c_int, [c_int, c_int64, c_int64, c_double, c_void_p, c_int64, c_double, c_void_p, c_int64, c_int, c_int,
                          c_void_p]),
    "dcgp_gram_matmul_workspace": (c_size_t, [c_int64, c_int64, c_int]),
    "dcgp_gram_matmul": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                 c_double, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_int, c_int,
                                 c_void_p]),
    "dcgp_vbagg_ell": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_int, c_void_p]),
    "dcgp_potrf_workspace": (c_size_t, [c_int64, c_int]),
    "dcgp_potrf": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "dcgp_trsm_workspace": (c_size_t, [c_int64, c_int64, c_int]),
    "dcgp_trsm": (c_int, [c_int, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_int,
                          c_int, c_void_p]),
    "dcgp_potrs": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_int, c_int,
                           c_void_p]),
    "dcgp_logdet_chol": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p]),
    "dcgp_solve_plan_bytes": (c_size_t, [c_int64, c_int64, c_int]),
    "dcgp_solve_plan": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "dcgp_trsm_planned": (c_int, [c_int, c_void_p, c_int64, c_int64, c_void_p, c_size_t, c_void_p, c_int64, c_int64,
                                  c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "dcgp_trtri_workspace": (c_size_t, [c_int64, c_int]),
    "dcgp_trtri": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "dcgp_bag_sum": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_void_p]),
    "dcgp_gram_bagsum": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                 c_int64, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "dcgp_gram_bagsum_bwd": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p,
                                     c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                     c_void_p]),
    "dcgp_gram_bag_blocksum": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int,
                                       c_int, c_void_p]),
    "dcgp_gram_bag_blocksum_bwd": (c_int, [POINTER(KernelSpec), c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                           c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "dcgp_rff_features": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                                  c_int, c_void_p]),
    "dcgp_rff_features_bwd": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p, c_void_p,
                                      c_int64, c_void_p, c_int64, c_void_p, c_int, c_void_p]),
    "dcgp_whitened_kl": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                 c_void_p]),
    "dcgp_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double,
                               c_double, c_double, c_void_p, c_int, c_int, c_int, c_void_p]),
    "dcgp_cmp_elbo_workspace": (c_size_t, [c_int64, c_int64, c_int64, c_int64, c_int]),
    "dcgp_cmp_elbo_fwd": (c_int, [POINTER(CmpProblem), c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "dcgp_cmp_elbo_bwd": (c_int, [POINTER(CmpProblem), c_double, POINTER(CmpGrads), c_void_p, c_size_t, c_int, c_int,
                                  c_void_p]),
    "dcgp_vbagg_elbo_workspace": (c_size_t, [c_int64, c_int64, c_int64, c_int]),
    "dcgp_vbagg_elbo_fwd": (c_int, [POINTER(VbaggProblem), c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "dcgp_vbagg_elbo_bwd": (c_int, [POINTER(VbaggProblem), c_double, POINTER(VbaggGrads), c_void_p, c_size_t, c_int, c_int,
                                    c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def load():
    """Loads libdcgp.so once; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DcgpError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or python deconditional_downscaling_b200/build.py). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc == 0:
        return
    if rc <= -1000:
        raise DcgpError(f"{what}: CUDA error {-rc - 1000}")
    raise DcgpError(f"{what}: invalid argument #{-rc}")
