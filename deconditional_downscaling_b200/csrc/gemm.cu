// Dense contraction C = alpha * op(A) op(B) + beta * C, row-major, arbitrary transposes.
//
// FP64: tcgen05 has no f64 kind, so double precision runs on the legacy tensor path
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  128x128x16 CTA tile, 8 warps (2x4), 64x32 warp
// tile = 32 DMMA per k-step with 128 accumulator registers per thread.  Operands stream
// global -> shared with cp.async (LDGSTS) through a 3-stage ring, one __syncthreads per k-tile;
// they are always laid out K-contiguous in shared memory with a pitch of 20 doubles, which makes
// every fragment read bank-conflict free for both half-warps (transposed operands are scattered
// into that layout by 8-byte cp.async).
//
// FP32 storage contracts on TF32 tensor cores (mma.sync.m16n8k8.tf32, FP32 accumulation).
// Two precisions: single-pass TF32 (cvt.rna at fragment load) and error-compensated "3xTF32"
// (a = a_hi + a_lo, three MMAs: lo*hi + hi*lo + hi*hi) which recovers ~FP32 accuracy.
//
// DCGP_GEMM_LOWER restricts the grid to tiles that intersect the lower triangle (SYRK
// trailing update of the blocked Cholesky).
#include "common.cuh"
#include "mma.cuh"
#include <type_traits>

namespace {

struct GemmParams {
    int64_t m, n, k;
    const void* A;
    int64_t sam, sak;  // element strides of op(A)[m, k]
    const void* B;
    int64_t sbn, sbk;  // element strides of op(B)[k, n]
    void* C;
    int64_t ldc;
    double alpha, beta;
    int lower;   // bit0: lower tiles only; bit1: never split K (C aliases an input)
    int ksplit;  // number of K slices (gridDim.z); >1 => atomicAdd epilogue on pre-scaled C
    int vecA, vecB;  // 16-byte cp.async allowed for the K-contiguous layout of A / B
    int batch;       // > 1: blockIdx.z enumerates independent problems (no split-K)
    int64_t strideA, strideB, strideC;  // element strides between batch members
    void* Clo;       // optional (FP32): receives c - trunc_tf32(c) for every stored element of C (same ld)
};

template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem, const void* gmem, int src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(s), "l"(gmem), "n"(BYTES), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// k-tile depth / shared pitch / pipeline depth per type; the CTA tile (BM x BN) is a kernel template
// parameter: 128 x 128 for problems that fill the machine, 64 x 64 for the many small, latency-bound
// contractions inside the blocked factorisations (4x more CTAs, 1/4 of the per-CTA critical path).
template <typename T> struct Tile;
#ifndef DCGP_F64_BK
#define DCGP_F64_BK 16
#endif
#ifndef DCGP_F64_STAGES
#define DCGP_F64_STAGES 3
#endif
template <> struct Tile<double> {
    static constexpr int BK = DCGP_F64_BK, PK = DCGP_F64_BK + 4, STAGES = DCGP_F64_STAGES;
};
template <> struct Tile<float> {
    static constexpr int BK = 32, PK = 36, STAGES = 3;
};

// issue the cp.async copies of one operand tile: ROWS x BK of op(X)[row, k] -> S[row][k] (pitch PK)
template <typename T, int ROWS, int BK, int PK>
__device__ __forceinline__ void issue_tile(T* __restrict__ S, const T* __restrict__ X, int64_t srow, int64_t sk,
                                           int64_t row0, int64_t nrows, int64_t k0, int64_t kend, int vec) {
    constexpr int VE = 16 / sizeof(T);  // elements per 16-byte chunk
    if (sk == 1 && vec) {
        constexpr int CPR = BK / VE;  // chunks per row
        constexpr int PER = ROWS * CPR / 256;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = threadIdx.x + i * 256;
            const int r = e / CPR, kk = (e % CPR) * VE;
            const int64_t gr = row0 + r, gk = k0 + kk;
            int bytes = 0;
            if (gr < nrows && gk < kend) bytes = (int)min((int64_t)VE, kend - gk) * (int)sizeof(T);
            const T* src = bytes ? X + gr * srow + gk : X;
            cp_async<16>(S + r * PK + kk, src, bytes);
        }
    } else {
        constexpr int PER = ROWS * BK / 256;
        const bool kcontig = (sk == 1);
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int e = threadIdx.x + i * 256;
            int r, kk;
            if (kcontig) {
                r = e / BK;
                kk = e % BK;
            } else {
                kk = e / ROWS;
                r = e % ROWS;
            }
            const int64_t gr = row0 + r, gk = k0 + kk;
            const bool ok = gr < nrows && gk < kend;
            const T* src = ok ? X + gr * srow + gk * sk : X;
            cp_async<(int)sizeof(T)>(S + r * PK + kk, src, ok ? (int)sizeof(T) : 0);
        }
    }
}

template <typename T, bool X3, int BM, int BN>
__global__ void __launch_bounds__(256, (BM == 128 ? 1 : 2)) gemm_kernel(const __grid_constant__ GemmParams p) {
    constexpr int BK = Tile<T>::BK, PK = Tile<T>::PK, ST = Tile<T>::STAGES;
    constexpr int WM = BM / 2, WN = BN / 4;  // 2 x 4 warps
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* As = reinterpret_cast<T*>(smem_raw);  // [ST][BM][PK]
    T* Bs = As + ST * BM * PK;               // [ST][BN][PK]

    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    if ((p.lower & 1) && n0 > m0 + BM - 1) return;
    // K slice of this CTA
    const int64_t ktiles = (p.k + BK - 1) / BK;
    const int64_t per = (ktiles + p.ksplit - 1) / p.ksplit;
    const int64_t kt0 = (p.batch > 1) ? 0 : (int64_t)blockIdx.z * per;
    const int64_t kt1 = min(ktiles, kt0 + per);
    if (kt0 >= kt1 && p.ksplit > 1) return;
    const int64_t kend = min(p.k, kt1 * BK);
    const int nkt = (int)max((int64_t)0, kt1 - kt0);

    const int64_t bz = (p.batch > 1) ? (int64_t)blockIdx.z : 0;
    const T* A = (const T*)p.A + bz * p.strideA;
    const T* B = (const T*)p.B + bz * p.strideB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile WM x WN
    const int g = lane >> 2, t = lane & 3;

    constexpr int NACC = (sizeof(T) == 8) ? 2 : 4;
    constexpr int RT = (sizeof(T) == 8) ? 8 : 16;  // rows per m-tile
    constexpr int MT = WM / RT;                    // m-tiles per warp
    constexpr int NT = WN / 8;                     // n-tiles per warp (8 cols each)
    T acc[MT][NT][NACC];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[i][j][q] = T(0);

    // prologue: ST-1 tiles in flight
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) {
        if (s < nkt) {
            issue_tile<T, BM, BK, PK>(As + s * BM * PK, A, p.sam, p.sak, m0, p.m, (kt0 + s) * BK, kend, p.vecA);
            issue_tile<T, BN, BK, PK>(Bs + s * BN * PK, B, p.sbn, p.sbk, n0, p.n, (kt0 + s) * BK, kend, p.vecB);
        }
        cp_async_commit();
    }
    for (int it = 0; it < nkt; ++it) {
        cp_async_wait<ST - 2>();
        __syncthreads();  // tile `it` has landed for every thread; everyone is done with tile it-1's buffer
        {
            const int nx = it + ST - 1;
            if (nx < nkt) {
                const int sb = nx % ST;
                issue_tile<T, BM, BK, PK>(As + sb * BM * PK, A, p.sam, p.sak, m0, p.m, (kt0 + nx) * BK, kend, p.vecA);
                issue_tile<T, BN, BK, PK>(Bs + sb * BN * PK, B, p.sbn, p.sbk, n0, p.n, (kt0 + nx) * BK, kend, p.vecB);
            }
            cp_async_commit();
        }
        const int buf = it % ST;
        const T* as = As + buf * BM * PK + (wm * WM) * PK;
        const T* bs = Bs + buf * BN * PK + (wn * WN) * PK;
        if constexpr (sizeof(T) == 8) {
#ifdef DCGP_F64_PREFETCH
            // fragments of step ks + 1 are loaded before the DMMAs of step ks issue
            double af[2][MT], bf[2][NT];
#pragma unroll
            for (int i = 0; i < MT; ++i) af[0][i] = as[(i * 8 + g) * PK + t];
#pragma unroll
            for (int j = 0; j < NT; ++j) bf[0][j] = bs[(j * 8 + g) * PK + t];
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                const int cur = ks & 1, nxt = cur ^ 1;
                if (ks + 1 < BK / 4) {
#pragma unroll
                    for (int i = 0; i < MT; ++i) af[nxt][i] = as[(i * 8 + g) * PK + (ks + 1) * 4 + t];
#pragma unroll
                    for (int j = 0; j < NT; ++j) bf[nxt][j] = bs[(j * 8 + g) * PK + (ks + 1) * 4 + t];
                }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) mma_f64(acc[i][j], af[cur][i], bf[cur][j]);
            }
#else
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                double af[MT], bf[NT];
#pragma unroll
                for (int i = 0; i < MT; ++i) af[i] = as[(i * 8 + g) * PK + ks * 4 + t];
#pragma unroll
                for (int j = 0; j < NT; ++j) bf[j] = bs[(j * 8 + g) * PK + ks * 4 + t];
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) mma_f64(acc[i][j], af[i], bf[j]);
            }
#endif
        } else {
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                uint32_t ah[MT][4], bh[NT][2];
                uint32_t al[X3 ? MT : 1][4], bl[X3 ? NT : 1][2];
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    const float* base = as + (i * 16 + g) * PK + ks * 8 + t;
                    const float v[4] = {base[0], base[8 * PK], base[4], base[8 * PK + 4]};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if constexpr (X3) split_tf32(v[q], ah[i][q], al[i][q]);
                        else ah[i][q] = to_tf32(v[q]);
                    }
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float* base = bs + (j * 8 + g) * PK + ks * 8 + t;
                    const float v[2] = {base[0], base[4]};
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if constexpr (X3) split_tf32(v[q], bh[j][q], bl[j][q]);
                        else bh[j][q] = to_tf32(v[q]);
                    }
                }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) {
                        if constexpr (X3) {
                            mma_tf32(acc[i][j], al[i], bh[j]);
                            mma_tf32(acc[i][j], ah[i], bl[j]);
                        }
                        mma_tf32(acc[i][j], ah[i], bh[j]);
                    }
            }
        }
    }
    cp_async_wait<0>();

    // epilogue: each thread owns column pairs (2t, 2t+1) -> 16-byte (f64) / 8-byte (f32) accesses
    T* C = (T*)p.C + bz * p.strideC;
    const T alpha = (T)p.alpha, beta = (T)p.beta;
    const bool atomic = p.ksplit > 1;
    const bool pair_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(C) % (2 * sizeof(T))) == 0);
    using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
#pragma unroll
    for (int i = 0; i < MT; ++i) {
#pragma unroll
        for (int h = 0; h < NACC / 2; ++h) {
            const int64_t row = m0 + wm * WM + i * RT + g + h * 8;
            if (row >= p.m) continue;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int64_t col = n0 + wn * WN + j * 8 + 2 * t;
                T* dst = C + row * p.ldc + col;
                const T v0 = alpha * acc[i][j][h * 2], v1 = alpha * acc[i][j][h * 2 + 1];
                if (atomic) {
                    if (col < p.n) atomicAdd(dst, v0);
                    if (col + 1 < p.n) atomicAdd(dst + 1, v1);
                } else if (pair_ok && col + 1 < p.n) {
                    T2 o;
                    if (beta == T(0)) {
                        o.x = v0;
                        o.y = v1;
                    } else {
                        const T2 c = *reinterpret_cast<const T2*>(dst);
                        o.x = fma(beta, c.x, v0);
                        o.y = fma(beta, c.y, v1);
                    }
                    *reinterpret_cast<T2*>(dst) = o;
                    if constexpr (sizeof(T) == 4) {
                        if (p.Clo) {  // TF32 remainders of the finished values (3xTF32 operands of later products)
                            float* lo = (float*)p.Clo + row * p.ldc + col;
                            lo[0] = o.x - __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u);
                            lo[1] = o.y - __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u);
                        }
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (col + q < p.n) {
                            const T vq = q ? v1 : v0;
                            const T o = (beta == T(0)) ? vq : fma(beta, dst[q], vq);
                            dst[q] = o;
                            if constexpr (sizeof(T) == 4) {
                                if (p.Clo)
                                    ((float*)p.Clo)[row * p.ldc + col + q] =
                                        o - __uint_as_float(__float_as_uint(o) & 0xFFFFE000u);
                            }
                        }
                    }
                }
            }
        }
    }
}

// C *= beta (or zero) ahead of a split-K accumulation
template <typename T>
__global__ void scale_kernel(T* C, int64_t m, int64_t n, int64_t ldc, T beta, int lower, int bm, int bn) {
    const int64_t total = m * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / n, c = e % n;
        if (lower && (c / bn) * bn > (r / bm) * bm + bm - 1) continue;
        T* d = C + r * ldc + c;
        *d = (beta == T(0)) ? T(0) : beta * *d;
    }
}

template <typename T>
int vec_ok(const void* X, int64_t ld) {
    return (reinterpret_cast<uintptr_t>(X) % 16 == 0) && ((ld * (int64_t)sizeof(T)) % 16 == 0);
}

template <typename T, bool X3, int BM, int BN>
int launch_gemm_tile(GemmParams p, cudaStream_t st) {
    constexpr int BK = Tile<T>::BK, PK = Tile<T>::PK, ST = Tile<T>::STAGES;
    const size_t smem = (size_t)ST * (BM + BN) * PK * sizeof(T);
    cudaError_t e = cudaFuncSetAttribute(gemm_kernel<T, X3, BM, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
    const int64_t gm = (p.m + BM - 1) / BM, gn = (p.n + BN - 1) / BN;
    // split K when the output grid cannot fill the machine and K is deep
    int64_t tiles = (p.lower & 1) ? gm * (gm + 1) / 2 : gm * gn;
    int ksplit = 1;
    const int64_t ktiles = (p.k + BK - 1) / BK;
    if (!(p.lower & 2) && tiles < num_sms() && ktiles >= 16) {
        ksplit = (int)min((int64_t)((2 * num_sms() + tiles - 1) / tiles), ktiles / 8);
        if (ksplit < 1) ksplit = 1;
        if (ksplit > 64) ksplit = 64;
    }
    if (p.batch > 1) ksplit = 1;
    p.ksplit = ksplit;
    if (ksplit > 1) {
        int64_t total = p.m * p.n;
        unsigned g = (unsigned)min((int64_t)(num_sms() * 8), (total + 255) / 256);
        scale_kernel<T><<<g, 256, 0, st>>>((T*)p.C, p.m, p.n, p.ldc, (T)p.beta, p.lower & 1, BM, BN);
    }
    dim3 grid((unsigned)gn, (unsigned)gm, (unsigned)(p.batch > 1 ? p.batch : ksplit));
    gemm_kernel<T, X3, BM, BN><<<grid, 256, smem, st>>>(p);
    DCGP_CHECK_LAUNCH();
    return 0;
}

template <typename T, bool X3>
int launch_gemm(GemmParams p, cudaStream_t st) {
    if (p.lower & 4) {
        // C aliases an operand (triangular-solve leaves): one CTA must own every column (row) of the
        // aliased operand it reads, i.e. a single tile along the short dimension
        if (p.n <= 128) return launch_gemm_tile<T, X3, 64, 128>(p, st);
        if (p.m <= 128) return launch_gemm_tile<T, X3, 128, 64>(p, st);
        return -16;
    }
    const int64_t t128 = ((p.m + 127) / 128) * ((p.n + 127) / 128) * (p.batch > 1 ? p.batch : 1);
    if (t128 < num_sms() / 2) return launch_gemm_tile<T, X3, 64, 64>(p, st);
    return launch_gemm_tile<T, X3, 128, 128>(p, st);
}

}  // namespace

int dcgp_gemm_tc_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                      const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int flags, const void* Alo,
                      const void* Blo, void* Clo, cudaStream_t st);

// internal entry used by the Cholesky / TRSM drivers
// Alo / Blo: optional x - trunc_tf32(x) companions of A / B (same layout) that let a 3xTF32 request run on
// the tcgen05 kernel; without them compensated contractions use the mma.sync kernel.
int dcgp_gemm_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype, int flags,
                   cudaStream_t st, const void* Alo, const void* Blo, void* Clo) {
    if (m < 0) return -3;
    if (n < 0) return -4;
    if (k < 0) return -5;
    if (m == 0 || n == 0) return 0;
    if (!C || ldc < n) return -12;
    if (k > 0 && (!A || !B)) return -7;
    GemmParams p;
    p.m = m; p.n = n; p.k = k;
    p.A = A; p.B = B; p.C = C; p.ldc = ldc;
    // op(A)[i, kk]: notrans A is m x k (row stride lda); trans A is k x m
    p.sam = transa ? 1 : lda;
    p.sak = transa ? lda : 1;
    // op(B)[kk, j]: notrans B is k x n; trans B is n x k
    p.sbn = transb ? ldb : 1;
    p.sbk = transb ? 1 : ldb;
    p.alpha = alpha; p.beta = beta;
    // bit0: lower-only tiles; bit1: never split K; bit2: C aliases an operand (in-place leaf)
    p.lower = ((flags & DCGP_GEMM_LOWER) ? 1 : 0) | ((flags & (DCGP_GEMM_NO_SPLITK | DCGP_GEMM_INPLACE)) ? 2 : 0) |
              ((flags & DCGP_GEMM_INPLACE) ? 4 : 0);
    p.ksplit = 1;
    p.batch = 1;
    p.strideA = p.strideB = p.strideC = 0;
    p.Clo = Clo;  // FP32: the epilogue also publishes the TF32 remainders of C (operands of later 3xTF32 products)
    if (Clo) p.lower |= 2;  // ... which the split-K (atomic) epilogue cannot do
    if (dtype == DCGP_F64) {
        p.vecA = vec_ok<double>(A, lda);
        p.vecB = vec_ok<double>(B, ldb);
        return launch_gemm<double, false>(p, st);
    }
    if (dtype == DCGP_F32) {
        p.vecA = vec_ok<float>(A, lda);
        p.vecB = vec_ok<float>(B, ldb);
        // large TF32 contractions go to the tcgen05 / TMA / TMEM kernel (gemm_tc.cu): single-pass, or
        // 3xTF32 when the caller supplies the lo companions; small or split-K-shaped problems stay on
        // the mma.sync kernel below
        const bool x3 = (flags & DCGP_GEMM_TF32X3) != 0;
        const int64_t tc_tiles = ((m + 127) / 128) * ((n + 127) / 128);  // narrow (128 x 128) tiles
        // (3xTF32 with remainder operands: the mma.sync alternative is ~4x slower, so even a mostly empty 128-wide
        // tile column - a few dozen right-hand sides - is worth it; TMA zero-fills the missing columns)
        if (!(flags & (DCGP_GEMM_NO_TCGEN05 | DCGP_GEMM_INPLACE)) && m >= 128 && n >= ((x3 && Alo && Blo) ? 32 : 128) && k >= 64 &&
            (!x3 || (Alo && Blo)) && tc_tiles >= (x3 ? 4 : 32)) {
            const int rc = dcgp_gemm_tc_impl(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags,
                                             x3 ? Alo : nullptr, x3 ? Blo : nullptr, Clo, st);
            if (rc <= 0) return rc;
        }
        if (x3) return launch_gemm<float, true>(p, st);
        return launch_gemm<float, false>(p, st);
    }
    return -14;
}

// Batched variant for the small independent products of the blocked triangular inversion (mma.sync kernel).
int dcgp_gemm_batched_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                           int64_t lda, int64_t strideA, const void* B, int64_t ldb, int64_t strideB, double beta,
                           void* C, int64_t ldc, int64_t strideC, int batch, int dtype, int flags, cudaStream_t st) {
    if (m <= 0 || n <= 0 || batch <= 0) return 0;
    GemmParams p;
    p.m = m; p.n = n; p.k = k;
    p.A = A; p.B = B; p.C = C; p.ldc = ldc;
    p.sam = transa ? 1 : lda;
    p.sak = transa ? lda : 1;
    p.sbn = transb ? ldb : 1;
    p.sbk = transb ? 1 : ldb;
    p.alpha = alpha; p.beta = beta;
    p.lower = 2;
    p.ksplit = 1;
    p.batch = batch;
    p.strideA = strideA; p.strideB = strideB; p.strideC = strideC;
    p.Clo = nullptr;
    if (dtype == DCGP_F64) {
        p.vecA = vec_ok<double>(A, lda) && (strideA % 2 == 0);
        p.vecB = vec_ok<double>(B, ldb) && (strideB % 2 == 0);
        return launch_gemm<double, false>(p, st);
    }
    p.vecA = vec_ok<float>(A, lda) && (strideA % 4 == 0);
    p.vecB = vec_ok<float>(B, ldb) && (strideB % 4 == 0);
    if (flags & DCGP_GEMM_TF32X3) return launch_gemm<float, true>(p, st);
    return launch_gemm<float, false>(p, st);
}

int dcgp_split_lo_impl(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, cudaStream_t st);

extern "C" int dcgp_split_lo(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd,
                             void* stream) {
    if (rows < 0 || cols < 0) return -3;
    if (rows == 0 || cols == 0) return 0;
    if (!src || !dst || lds < cols || ldd < cols) return -1;
    return dcgp_split_lo_impl(src, dst, rows, cols, lds, ldd, (cudaStream_t)stream);
}

extern "C" int dcgp_gemm_ex(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                            int64_t lda, const void* Alo, const void* B, int64_t ldb, const void* Blo, double beta,
                            void* C, int64_t ldc, void* Clo, int dtype, int flags, void* stream) {
    if (dtype != DCGP_F32 && (Alo || Blo || Clo)) return -17;
    return dcgp_gemm_impl(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, dtype, flags,
                          (cudaStream_t)stream, Alo, Blo, Clo);
}

// C (lower triangle) = alpha * op(A) op(A)^T + beta * C;  trans = 0: A is n x k, trans = 1: A is k x n
extern "C" int dcgp_syrk(int trans, int64_t n, int64_t k, double alpha, const void* A, int64_t lda, double beta, void* C,
                         int64_t ldc, int dtype, int flags, void* stream) {
    return dcgp_gemm_impl(trans ? 1 : 0, trans ? 0 : 1, n, n, k, alpha, A, lda, A, lda, beta, C, ldc, dtype,
                          (flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05 | DCGP_GEMM_NO_SPLITK)) | DCGP_GEMM_LOWER,
                          (cudaStream_t)stream, nullptr, nullptr, nullptr);
}

extern "C" int dcgp_gemm(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                         int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype,
                         int flags, void* stream) {
    return dcgp_gemm_impl(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, dtype, flags,
                          (cudaStream_t)stream, nullptr, nullptr, nullptr);
}
