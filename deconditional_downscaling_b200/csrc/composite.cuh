// Shared pieces of the composite entry points (vbagg.cu, cmp.cu): small dense-matrix kernels, the 3xTF32 product helper and
// error macros.  Include after common.cuh.
#pragma once

// internal entries of gemm.cu (3xTF32 on the tcgen05 kernel needs the TF32 remainders of both operands)
int dcgp_gemm_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype, int flags,
                   cudaStream_t st, const void* Alo, const void* Blo, void* Clo);
int dcgp_split_lo_impl(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, cudaStream_t st);

namespace composite {


template <typename T>
__global__ void tril_copy_kernel(const T* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t n) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[i * n + j] = (j <= i) ? src[i * lds + j] : T(0);
}

template <typename T>
__global__ void add_diag_kernel(T* __restrict__ A, int64_t n, T c) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[i * n + i] += c;
}

// dst = alpha * tril(src) (+ beta * add on the lower triangle), zero above;  halve_diag: diagonal times 1/2
template <typename T>
__global__ void tril_axpby_kernel(const T* __restrict__ src, int64_t n, T alpha, const T* __restrict__ add, int64_t lda,
                                  T beta, int halve_diag, T* __restrict__ dst, int64_t ldd) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    T v = T(0);
    if (j <= i) {
        v = alpha * src[i * n + j];
        if (halve_diag && i == j) v *= T(0.5);
        if (add) v += beta * add[i * lda + j];
    }
    dst[i * ldd + j] = v;
}

// dst = (src + src^T) / 2
template <typename T>
__global__ void symmetrize_kernel(const T* __restrict__ src, int64_t n, T* __restrict__ dst) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) dst[i * n + j] = T(0.5) * (src[i * n + j] + src[j * n + i]);
}


inline size_t up(size_t v) { return (v + 31) & ~(size_t)31; }

#define DCGP_CU_RC(x)                                     \
    do {                                                 \
        cudaError_t e_ = (x);                            \
        if (e_ != cudaSuccess) return DCGP_CUDA_ERR(e_); \
    } while (0)

#define RC(x)               \
    do {                    \
        int rc_ = (x);      \
        if (rc_) return rc_; \
    } while (0)

// C = alpha op(A) op(B) + beta C with packed operands.  FP32: 3xTF32; the remainders of A / B are taken from Alo / Blo
// when the caller keeps them, else split here into the scratch slots s1 / s2 (the product then runs on the tcgen05
// kernel; dcgp_gemm_impl falls back to the mma.sync 3xTF32 kernel for shapes that kernel declines).
template <typename T>
int gemm3(int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const T* A, int64_t lda, const T* Alo, const T* B,
          int64_t ldb, const T* Blo, double beta, T* C, int64_t ldc, T* s1, T* s2, int dtype, int flags, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        if (m >= 128 && n >= 32 && k >= 64) {
            if (!Alo) {
                RC(dcgp_split_lo_impl(A, s1, ta ? k : m, ta ? m : k, lda, lda, st));
                Alo = s1;
            }
            if (!Blo) {
                RC(dcgp_split_lo_impl(B, s2, tb ? n : k, tb ? k : n, ldb, ldb, st));
                Blo = s2;
            }
        } else {
            Alo = Blo = nullptr;
        }
    } else {
        Alo = Blo = nullptr;
    }
    return dcgp_gemm_impl(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, dtype, flags, st, Alo, Blo, nullptr);
}

// Single-pass TF32 product for the large (individuals-sized) contractions: tcgen05 reads FP32 operands by truncation, so
// both operands are first rounded to nearest TF32 (Ar / Br: copies the caller keeps, else made here in s1 / s2) - the
// policy of the Python layer (ops._rna).  FP64, and shapes the tcgen05 kernel does not take, go to gemm3.
template <typename T>
int gemm1(int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const T* A, int64_t lda, const T* Ar, const T* B,
          int64_t ldb, const T* Br, double beta, T* C, int64_t ldc, T* s1, T* s2, int dtype, int flags, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        const int64_t tiles = ((m + 127) / 128) * ((n + 127) / 128);
        if (m >= 128 && n >= 128 && k >= 64 && tiles >= 32 && (lda % 4) == 0 && (ldb % 4) == 0 && (ldc % 4) == 0) {
            if (!Ar) {
                RC(dcgp_round_tf32(A, s1, ta ? k : m, ta ? m : k, lda, lda, st));
                Ar = s1;
            }
            if (!Br) {
                RC(dcgp_round_tf32(B, s2, tb ? n : k, tb ? k : n, ldb, ldb, st));
                Br = s2;
            }
            return dcgp_gemm_impl(ta, tb, m, n, k, alpha, Ar, lda, Br, ldb, beta, C, ldc, dtype, flags & ~DCGP_GEMM_TF32X3, st,
                                  nullptr, nullptr, nullptr);
        }
    }
    return gemm3<T>(ta, tb, m, n, k, alpha, A, lda, nullptr, B, ldb, nullptr, beta, C, ldc, s1, s2, dtype, flags, st);
}

template <typename T>
int split_keep(const T* src, T* dst, int64_t rows, int64_t cols, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) return dcgp_split_lo_impl(src, dst, rows, cols, cols, cols, st);
    return 0;
}


// ---- bandwidth-bound helpers (the GEMM kernels are the wrong tool for one right-hand side or for a dot product) ----
// out[0] += sum_i a[i] * b[i]   (grid-stride, one atomic per block; out zeroed by the caller)
template <typename T>
__global__ void __launch_bounds__(256) dot_kernel(const T* __restrict__ a, const T* __restrict__ b, int64_t n, T* __restrict__ out) {
    T acc = T(0);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc = fma(a[i], b[i], acc);
    acc = warp_sum(acc);
    __shared__ T red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = T(0);
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out, t);
    }
}

template <typename T>
int dot(const T* a, const T* b, int64_t n, T* out, cudaStream_t st) {
    int64_t blocks = (n + 256 * 8 - 1) / (256 * 8);
    if (blocks > 592) blocks = 592;
    if (blocks < 1) blocks = 1;
    dot_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(a, b, n, out);
    DCGP_CHECK_LAUNCH();
    return 0;
}

// y = A x for row-major A (m x k, row stride lda): one warp per row
template <typename T>
__global__ void __launch_bounds__(256) gemv_n_kernel(const T* __restrict__ A, int64_t lda, int64_t m, int64_t k,
                                                     const T* __restrict__ x, T* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= m) return;
    T acc = T(0);
    for (int64_t j = lane; j < k; j += 32) acc = fma(A[row * lda + j], x[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[row] = acc;
}

// y += A^T x for row-major A (k x m): a thread owns a column, a block a chunk of 64 rows (256-row chunks left the grid at
// 256 blocks for the 1024 x 8192 operands of the ELBO step: 78 us per call); y zeroed by the caller
template <typename T>
__global__ void __launch_bounds__(128) gemv_t_kernel(const T* __restrict__ A, int64_t lda, int64_t k, int64_t m,
                                                     const T* __restrict__ x, T* __restrict__ y) {
    const int64_t col = (int64_t)blockIdx.x * 128 + threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.y * 64, r1 = r0 + 64 < k ? r0 + 64 : k;
    __shared__ T xs[64];
    if (threadIdx.x < 64) xs[threadIdx.x] = (r0 + threadIdx.x < k) ? x[r0 + threadIdx.x] : T(0);
    __syncthreads();
    if (col >= m) return;
    T acc = T(0);
    for (int64_t i = r0; i < r1; ++i) acc = fma(A[i * lda + col], xs[i - r0], acc);
    atomicAdd(&y[col], acc);
}

// y = op(A) x: ta = 0: A is m x k; ta = 1: A is k x m (stored), y has m entries
template <typename T>
int gemv(int ta, int64_t m, int64_t k, const T* A, int64_t lda, const T* x, T* y, cudaStream_t st) {
    if (ta == 0) {
        gemv_n_kernel<T><<<(unsigned)((m + 7) / 8), 256, 0, st>>>(A, lda, m, k, x, y);
    } else {
        DCGP_CU_RC(cudaMemsetAsync(y, 0, (size_t)m * sizeof(T), st));
        const dim3 grid((unsigned)((m + 127) / 128), (unsigned)((k + 63) / 64));
        gemv_t_kernel<T><<<grid, 128, 0, st>>>(A, lda, k, m, x, y);
    }
    DCGP_CHECK_LAUNCH();
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// FP64 island of FP32 problems (src/variational/variational_strategy.py:35-44: the reference factors K_ZZ in double):
// K_ZZ + jitter I is evaluated, factored and inverted in FP64 on the side lane (off the critical path), W = L_Z^{-1}
// is handed to the FP32 products as an exact hi + lo pair, and the adjoint of the factor runs in FP64 again.
// ---------------------------------------------------------------------------------------------------------------
static __global__ void f32_to_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (double)src[i];
}

// hyper-parameter values of a spec, FP32 -> FP64: slot c * 9 + j (lengthscales), c * 9 + 8 (outputscale)
static __global__ void spec_to_f64_kernel(const __grid_constant__ dcgp_kernel_spec sp, double* __restrict__ dst) {
    const int c = blockIdx.x, j = threadIdx.x;
    if (c >= sp.n_components) return;
    if (j < sp.comp[c].ndims) dst[c * 9 + j] = (double)((const float*)sp.comp[c].lengthscale)[j];
    if (j == 8) dst[c * 9 + 8] = (double)((const float*)sp.comp[c].outputscale)[0];
}

// hi = float(x), lo = float(x - hi)
static __global__ void f64_to_hilo_kernel(const double* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = src[i];
    const float h = (float)x;
    hi[i] = h;
    lo[i] = (float)(x - (double)h);
}

// dst (rows x cols, row stride ldd, FP32) += src (FP64, row stride lds); dst may be NULL
static __global__ void acc_f64_kernel(const double* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t rows,
                               int64_t cols) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (dst && i < rows && j < cols) dst[i * ldd + j] += (float)src[i * lds + j];
}

struct Island {   // byte offsets of the FP64 region
    size_t Z, hyp, L, W, t1, t2, t3, dZ, dth, end;
};

inline Island island_layout(size_t base, int64_t M, int64_t ldz) {
    Island a;
    size_t o = (base + 255) & ~(size_t)255;
    const size_t mm = (((size_t)M * M * 8) + 255) & ~(size_t)255, mz = (((size_t)M * ldz * 8) + 255) & ~(size_t)255;
    a.Z = o; o += mz;
    a.hyp = o; o += 512;
    a.L = o; o += mm;
    a.W = o; o += mm;
    a.t1 = o; o += mm;
    a.t2 = o; o += mm;
    a.t3 = o; o += mm;
    a.dZ = o; o += mz;
    a.dth = o; o += 512;      // dls64 [4][8], dos64 [4]
    a.end = o;
    return a;
}

inline dcgp_kernel_spec spec_f64(const dcgp_kernel_spec* sp, const double* hyp) {
    dcgp_kernel_spec s = *sp;
    for (int c = 0; c < s.n_components; ++c) {
        s.comp[c].lengthscale = hyp + c * 9;
        s.comp[c].outputscale = hyp + c * 9 + 8;
    }
    return s;
}


}  // namespace composite
