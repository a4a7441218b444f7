// Blocked right-looking Cholesky and triangular solves (row-major, lower).
//
//   for each 128-wide block column j:
//     1. potrf_diag_kernel  : one CTA factors the 128x128 diagonal block in shared memory and
//                             inverts the triangular factor (stored in the workspace);
//     2. panel "TRSM"       : L21 = A21 * inv(L11)^T as a tensor-core GEMM (in place);
//     3. trailing "SYRK"    : A22 -= L21 L21^T on the lower tiles only (tensor-core GEMM).
//
// Steps 2-3 carry all O(n^3) work and run in dcgp_gemm (DMMA for FP64, TF32 for FP32); step 1
// is O(n * 128^2) on CUDA cores.  Triangular solves with many right-hand sides use the same
// inverted diagonal blocks: X_i = inv(L_ii) B_i followed by a rank-128 GEMM update of all
// remaining block rows (right-looking, so every update fills the machine).
//
// LAPACK-style failure reporting: *info (device) receives the 1-based index of the first
// non-positive pivot; nothing here synchronises the stream.
#include "common.cuh"

int dcgp_gemm_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype, int flags,
                   cudaStream_t st);


namespace {

constexpr int NB = 128;       // block size
constexpr int SP = NB + 1;    // shared-memory pitch (conflict-free column walks)
constexpr int NTHREADS = 512;

// Factor the nb x nb block held in S (lower triangle) in place: right-looking, whole CTA.
template <typename T>
__device__ void block_potrf(T* S, int nb, int* info, int64_t j0) {
    for (int j = 0; j < nb; ++j) {
        __shared__ T pivot;
        if (threadIdx.x == 0) {
            T d = S[j * SP + j];
            if (!(d > T(0))) {
                if (info && *info == 0) *info = (int)(j0 + j + 1);
            }
            pivot = t_sqrt<T>(d);
            S[j * SP + j] = pivot;
        }
        __syncthreads();
        const T inv = T(1) / pivot;
        for (int i = j + 1 + threadIdx.x; i < nb; i += NTHREADS) S[i * SP + j] *= inv;
        __syncthreads();
        // trailing update of the lower triangle: (i, k) with j < k <= i < nb
        {
            const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
            for (int i = j + 1 + ty; i < nb; i += NTHREADS / 32) {
                const T lij = S[i * SP + j];
                for (int k = j + 1 + tx; k <= i; k += 32) S[i * SP + k] = fma(-lij, S[k * SP + j], S[i * SP + k]);
            }
        }
        __syncthreads();
    }
}

// Invert the lower-triangular factor held in S.  X = L^{-1} is lower triangular; its strict
// lower part is written into the (free) strict upper triangle of S transposed
// (X[i][c] -> S[c][i], i > c); the diagonal is 1 / L_ii.  Four lanes cooperate on a column.
template <typename T>
__device__ void block_trtri(T* S, int nb) {
    const int c = threadIdx.x >> 2, q = threadIdx.x & 3;
    const unsigned gmask = 0xFu << ((threadIdx.x & 31) & ~3);  // the 4 lanes sharing column c
    if (c < nb) {
        for (int i = c + 1; i < nb; ++i) {
            // X_ic = -( L_ic * X_cc + sum_{k=c+1}^{i-1} L_ik X_kc ) / L_ii
            T s = T(0);
            for (int k = c + 1 + q; k < i; k += 4) s = fma(S[i * SP + k], S[c * SP + k], s);
            s += __shfl_xor_sync(gmask, s, 1);
            s += __shfl_xor_sync(gmask, s, 2);
            if (q == 0) S[c * SP + i] = -(S[i * SP + c] / S[c * SP + c] + s) / S[i * SP + i];
            __syncwarp(gmask);
        }
    }
    __syncthreads();
}

template <typename T>
__device__ void store_inverse(const T* S, int nb, T* inv) {  // inv: NB x NB row-major (lower, zero above)
    for (int e = threadIdx.x; e < nb * nb; e += NTHREADS) {
        const int i = e / nb, c = e % nb;
        T v = T(0);
        if (i == c) v = T(1) / S[i * SP + i];
        else if (i > c) v = S[c * SP + i];
        inv[i * NB + c] = v;
    }
}

template <typename T>
__global__ void __launch_bounds__(NTHREADS) potrf_diag_kernel(T* __restrict__ A, int64_t lda, int nb,
                                                              T* __restrict__ inv, int* info, int64_t j0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
    for (int e = threadIdx.x; e < nb * nb; e += NTHREADS) {
        const int i = e / nb, k = e % nb;
        if (k <= i) S[i * SP + k] = A[(int64_t)i * lda + k];
    }
    __syncthreads();
    block_potrf<T>(S, nb, info, j0);
    for (int e = threadIdx.x; e < nb * nb; e += NTHREADS) {
        const int i = e / nb, k = e % nb;
        if (k <= i) A[(int64_t)i * lda + k] = S[i * SP + k];
    }
    __syncthreads();
    block_trtri<T>(S, nb);
    store_inverse<T>(S, nb, inv);
}

// batched inversion of the diagonal blocks of an existing factor
template <typename T>
__global__ void __launch_bounds__(NTHREADS) trtri_blocks_kernel(const T* __restrict__ L, int64_t ldl, int64_t n,
                                                                T* __restrict__ inv) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
    const int64_t i0 = (int64_t)blockIdx.x * NB;
    const int nb = (int)min((int64_t)NB, n - i0);
    const T* Lb = L + i0 * ldl + i0;
    for (int e = threadIdx.x; e < nb * nb; e += NTHREADS) {
        const int i = e / nb, k = e % nb;
        if (k <= i) S[i * SP + k] = Lb[(int64_t)i * ldl + k];
    }
    __syncthreads();
    block_trtri<T>(S, nb);
    store_inverse<T>(S, nb, inv + (int64_t)blockIdx.x * NB * NB);
}

template <typename T>
__global__ void __launch_bounds__(256) logdet_kernel(const T* __restrict__ L, int64_t n, int64_t ldl,
                                                     T* __restrict__ out) {
    __shared__ T red[8];
    T acc = T(0);
    for (int64_t i = threadIdx.x; i < n; i += 256) acc += t_log<T>(L[i * ldl + i]);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = T(0);
        for (int w = 0; w < 8; ++w) t += red[w];
        out[0] = T(2) * t;
    }
}

template <typename T>
int set_smem_attr(const void* fn) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NB * SP * sizeof(T)));
    return e == cudaSuccess ? 0 : DCGP_CUDA_ERR(e);
}

template <typename T>
int potrf_impl(T* A, int64_t n, int64_t lda, int* info, T* ws, int dtype, int xf, cudaStream_t st) {
    const size_t smem = (size_t)NB * SP * sizeof(T);
    int rc = set_smem_attr<T>((const void*)potrf_diag_kernel<T>);
    if (rc) return rc;
    // Two-level blocking: an NBO-wide panel is factored with NB-wide steps whose updates stay
    // inside the panel (K = NB, O(n NBO^2) flops); the rest of the matrix is then updated ONCE per
    // panel with K = NBO, so ~(1 - NBO/n) of all flops run in deep-K tensor-core GEMMs whose C
    // tile traffic is amortised over NBO/16 k-steps.
    const int64_t NBO = (n >= 4096) ? 512 : ((n >= 1024) ? 256 : NB);
    for (int64_t J0 = 0; J0 < n; J0 += NBO) {
        const int64_t jb = (n - J0 < NBO) ? (n - J0) : NBO;
        const int64_t Jend = J0 + jb;
        for (int64_t j0 = J0; j0 < Jend; j0 += NB) {
            const int nb = (int)((Jend - j0 < NB) ? (Jend - j0) : NB);
            T* inv = ws + (j0 / NB) * NB * NB;
            potrf_diag_kernel<T><<<1, NTHREADS, smem, st>>>(A + j0 * lda + j0, lda, nb, inv, info, j0);
            DCGP_CHECK_LAUNCH();
            const int64_t rem = n - j0 - nb;
            if (rem > 0) {
                T* A21 = A + (j0 + nb) * lda + j0;
                rc = dcgp_gemm_impl(0, 1, rem, nb, nb, 1.0, A21, lda, inv, NB, 0.0, A21, lda, dtype,
                                    DCGP_GEMM_NO_SPLITK | xf, st);
                if (rc) return rc;
                const int64_t pc = Jend - (j0 + nb);  // panel columns still to be updated
                if (pc > 0) {
                    rc = dcgp_gemm_impl(0, 1, rem, pc, nb, -1.0, A21, lda, A21, lda, 1.0,
                                        A + (j0 + nb) * lda + (j0 + nb), lda, dtype,
                                        DCGP_GEMM_LOWER | DCGP_GEMM_NO_SPLITK | xf, st);
                    if (rc) return rc;
                }
            }
        }
        const int64_t rem = n - Jend;
        if (rem > 0) {
            T* P = A + Jend * lda + J0;  // rem x jb panel of L
            rc = dcgp_gemm_impl(0, 1, rem, rem, jb, -1.0, P, lda, P, lda, 1.0, A + Jend * lda + Jend, lda, dtype,
                                DCGP_GEMM_LOWER | DCGP_GEMM_NO_SPLITK | xf, st);
            if (rc) return rc;
        }
    }
    return 0;
}

template <typename T>
int trsm_impl(int trans, const T* L, int64_t n, int64_t ldl, T* B, int64_t m, int64_t ldb, T* ws, int dtype, int xf,
              cudaStream_t st) {
    const size_t smem = (size_t)NB * SP * sizeof(T);
    int rc = set_smem_attr<T>((const void*)trtri_blocks_kernel<T>);
    if (rc) return rc;
    const int64_t nblk = (n + NB - 1) / NB;
    trtri_blocks_kernel<T><<<(unsigned)nblk, NTHREADS, smem, st>>>(L, ldl, n, ws);
    DCGP_CHECK_LAUNCH();
    // Two-level blocking as in potrf: NB-wide substitution steps update only the rows of their
    // NBO-wide outer block; all other rows are updated once per outer block with K = NBO.
    const int64_t NBO = (n >= 4096) ? 512 : ((n >= 1024) ? 256 : NB);
    const int64_t nouter = (n + NBO - 1) / NBO;
    if (!trans) {
        for (int64_t ob = 0; ob < nouter; ++ob) {
            const int64_t I0 = ob * NBO;
            const int64_t Iend = (n < I0 + NBO) ? n : I0 + NBO;
            for (int64_t i0 = I0; i0 < Iend; i0 += NB) {
                const int64_t nb = (Iend - i0 < NB) ? (Iend - i0) : NB;
                T* Bi = B + i0 * ldb;
                rc = dcgp_gemm_impl(0, 0, nb, m, nb, 1.0, ws + (i0 / NB) * NB * NB, NB, Bi, ldb, 0.0, Bi, ldb, dtype,
                                    DCGP_GEMM_NO_SPLITK | xf, st);
                if (rc) return rc;
                const int64_t rin = Iend - i0 - nb;  // rows below, inside the outer block
                if (rin > 0) {
                    rc = dcgp_gemm_impl(0, 0, rin, m, nb, -1.0, L + (i0 + nb) * ldl + i0, ldl, Bi, ldb, 1.0,
                                        B + (i0 + nb) * ldb, ldb, dtype, DCGP_GEMM_NO_SPLITK | xf, st);
                    if (rc) return rc;
                }
            }
            const int64_t rem = n - Iend;
            if (rem > 0) {
                rc = dcgp_gemm_impl(0, 0, rem, m, Iend - I0, -1.0, L + Iend * ldl + I0, ldl, B + I0 * ldb, ldb, 1.0,
                                    B + Iend * ldb, ldb, dtype, DCGP_GEMM_NO_SPLITK | xf, st);
                if (rc) return rc;
            }
        }
    } else {
        for (int64_t ob = nouter - 1; ob >= 0; --ob) {
            const int64_t I0 = ob * NBO;
            const int64_t Iend = (n < I0 + NBO) ? n : I0 + NBO;
            const int64_t nin = (Iend - I0 + NB - 1) / NB;
            for (int64_t ib = nin - 1; ib >= 0; --ib) {
                const int64_t i0 = I0 + ib * NB;
                const int64_t nb = (Iend - i0 < NB) ? (Iend - i0) : NB;
                T* Bi = B + i0 * ldb;
                rc = dcgp_gemm_impl(1, 0, nb, m, nb, 1.0, ws + (i0 / NB) * NB * NB, NB, Bi, ldb, 0.0, Bi, ldb, dtype,
                                    DCGP_GEMM_NO_SPLITK | xf, st);
                if (rc) return rc;
                const int64_t rin = i0 - I0;  // rows above, inside the outer block
                if (rin > 0) {
                    rc = dcgp_gemm_impl(1, 0, rin, m, nb, -1.0, L + i0 * ldl + I0, ldl, Bi, ldb, 1.0, B + I0 * ldb, ldb,
                                        dtype, DCGP_GEMM_NO_SPLITK | xf, st);
                    if (rc) return rc;
                }
            }
            if (I0 > 0) {
                rc = dcgp_gemm_impl(1, 0, I0, m, Iend - I0, -1.0, L + I0 * ldl, ldl, B + I0 * ldb, ldb, 1.0, B, ldb,
                                    dtype, DCGP_GEMM_NO_SPLITK | xf, st);
                if (rc) return rc;
            }
        }
    }
    return 0;
}

size_t inv_ws_bytes(int64_t n, int dtype) {
    const int64_t nblk = (n + NB - 1) / NB;
    return (size_t)nblk * NB * NB * (dtype == DCGP_F64 ? 8 : 4);
}

}  // namespace

extern "C" size_t dcgp_potrf_workspace(int64_t n, int dtype) { return n > 0 ? inv_ws_bytes(n, dtype) : 0; }

extern "C" int dcgp_potrf(void* A, int64_t n, int64_t lda, int* info, void* workspace, size_t workspace_bytes,
                          int dtype, int flags, void* stream) {
    if (n < 0) return -2;
    if (n == 0) return 0;
    if (!A) return -1;
    if (lda < n) return -3;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -7;
    if (!workspace || workspace_bytes < inv_ws_bytes(n, dtype)) return -5;
    cudaStream_t st = (cudaStream_t)stream;
    if (info) {
        cudaError_t e = cudaMemsetAsync(info, 0, sizeof(int), st);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
    }
    if (dtype == DCGP_F64) return potrf_impl<double>((double*)A, n, lda, info, (double*)workspace, dtype, flags & DCGP_GEMM_TF32X3, st);
    return potrf_impl<float>((float*)A, n, lda, info, (float*)workspace, dtype, flags & DCGP_GEMM_TF32X3, st);
}

extern "C" size_t dcgp_trsm_workspace(int64_t n, int64_t m, int dtype) { return n > 0 ? inv_ws_bytes(n, dtype) : 0; }

extern "C" int dcgp_trsm(int trans, const void* L, int64_t n, int64_t ldl, void* B, int64_t m, int64_t ldb,
                         void* workspace, size_t workspace_bytes, int dtype, int flags, void* stream) {
    if (n < 0) return -3;
    if (m < 0) return -6;
    if (n == 0 || m == 0) return 0;
    if (!L || ldl < n) return -2;
    if (!B || ldb < m) return -5;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -10;
    if (!workspace || workspace_bytes < inv_ws_bytes(n, dtype)) return -8;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DCGP_F64)
        return trsm_impl<double>(trans, (const double*)L, n, ldl, (double*)B, m, ldb, (double*)workspace, dtype, flags & DCGP_GEMM_TF32X3, st);
    return trsm_impl<float>(trans, (const float*)L, n, ldl, (float*)B, m, ldb, (float*)workspace, dtype, flags & DCGP_GEMM_TF32X3, st);
}

extern "C" int dcgp_potrs(const void* L, int64_t n, int64_t ldl, void* B, int64_t m, int64_t ldb, void* workspace,
                          size_t workspace_bytes, int dtype, int flags, void* stream) {
    int rc = dcgp_trsm(0, L, n, ldl, B, m, ldb, workspace, workspace_bytes, dtype, flags, stream);
    if (rc) return rc;
    return dcgp_trsm(1, L, n, ldl, B, m, ldb, workspace, workspace_bytes, dtype, flags, stream);
}

extern "C" int dcgp_logdet_chol(const void* L, int64_t n, int64_t ldl, void* out, int dtype, void* stream) {
    if (!L || !out) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DCGP_F64) logdet_kernel<double><<<1, 256, 0, st>>>((const double*)L, n, ldl, (double*)out);
    else if (dtype == DCGP_F32) logdet_kernel<float><<<1, 256, 0, st>>>((const float*)L, n, ldl, (float*)out);
    else return -5;
    DCGP_CHECK_LAUNCH();
    return 0;
}
