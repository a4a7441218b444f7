// Dense contraction C = alpha * op(A) op(B) + beta * C, row-major, arbitrary transposes.
//
// FP64: tcgen05 has no f64 kind, so double precision runs on the legacy tensor path
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4).  128x128x16 CTA tile, 8 warps (2x4), 64x32 warp
// tile = 32 DMMA per k-step with 128 accumulator registers per thread; operands are staged
// global -> registers -> shared (double-buffered, one __syncthreads per k-tile) and always
// laid out K-contiguous in shared memory with a pitch of 20 doubles, which makes every
// fragment read bank-conflict free for both half-warps.
//
// FP32 storage contracts on TF32 tensor cores (mma.sync.m16n8k8.tf32, operands rounded to
// TF32 with cvt.rna when they are staged into shared memory, FP32 accumulation).
//
// DCGP_GEMM_LOWER restricts the grid to tiles that intersect the lower triangle (SYRK
// trailing update of the blocked Cholesky).
#include "common.cuh"

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
    int lower;
    int ksplit;  // number of K slices (gridDim.z); >1 => atomicAdd epilogue on pre-scaled C
};

__device__ __forceinline__ void mma_f64(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

template <typename T> struct Tile;
template <> struct Tile<double> {
    static constexpr int BM = 128, BN = 128, BK = 16, PK = 20;
};
template <> struct Tile<float> {
    static constexpr int BM = 128, BN = 128, BK = 32, PK = 36;
};

// stage one operand tile (ROWS x BK of op(X)[row, k]) global -> registers
template <typename T, int ROWS, int BK>
__device__ __forceinline__ void load_tile(const T* __restrict__ X, int64_t srow, int64_t sk, int64_t row0, int64_t nrows,
                                          int64_t k0, int64_t kend, T (&reg)[ROWS * BK / 256]) {
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
        reg[i] = (gr < nrows && gk < kend) ? X[gr * srow + gk * sk] : T(0);
    }
}

template <typename T, int ROWS, int BK, int PK>
__device__ __forceinline__ void store_tile(T* __restrict__ S, int64_t sk, const T (&reg)[ROWS * BK / 256]) {
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
        if constexpr (sizeof(T) == 4) S[r * PK + kk] = to_tf32(reg[i]);
        else S[r * PK + kk] = reg[i];
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 1) gemm_kernel(const __grid_constant__ GemmParams p) {
    constexpr int BM = Tile<T>::BM, BN = Tile<T>::BN, BK = Tile<T>::BK, PK = Tile<T>::PK;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* As = reinterpret_cast<T*>(smem_raw);  // [2][BM][PK]
    T* Bs = As + 2 * BM * PK;                // [2][BN][PK]

    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    if ((p.lower & 1) && n0 > m0 + BM - 1) return;
    // K slice of this CTA
    const int64_t ktiles = (p.k + BK - 1) / BK;
    const int64_t per = (ktiles + p.ksplit - 1) / p.ksplit;
    const int64_t kt0 = (int64_t)blockIdx.z * per;
    const int64_t kt1 = min(ktiles, kt0 + per);
    if (kt0 >= kt1 && p.ksplit > 1) return;
    const int64_t kend = min(p.k, kt1 * BK);

    const T* A = (const T*)p.A;
    const T* B = (const T*)p.B;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile 64 x 32
    const int g = lane >> 2, t = lane & 3;

    constexpr int NACC = (sizeof(T) == 8) ? 2 : 4;
    constexpr int MT = (sizeof(T) == 8) ? 8 : 4;  // m-tiles per warp (8 or 16 rows each)
    constexpr int NT = 4;                         // n-tiles per warp (8 cols each)
    T acc[MT][NT][NACC];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < NACC; ++q) acc[i][j][q] = T(0);

    T ra[BM * BK / 256], rb[BN * BK / 256];
    if (kt0 < kt1) {
        load_tile<T, BM, BK>(A, p.sam, p.sak, m0, p.m, kt0 * BK, kend, ra);
        load_tile<T, BN, BK>(B, p.sbn, p.sbk, n0, p.n, kt0 * BK, kend, rb);
        store_tile<T, BM, BK, PK>(As, p.sak, ra);
        store_tile<T, BN, BK, PK>(Bs, p.sbk, rb);
    }
    __syncthreads();
    for (int64_t kt = kt0; kt < kt1; ++kt) {
        const int buf = (int)((kt - kt0) & 1);
        const bool more = kt + 1 < kt1;
        if (more) {
            load_tile<T, BM, BK>(A, p.sam, p.sak, m0, p.m, (kt + 1) * BK, kend, ra);
            load_tile<T, BN, BK>(B, p.sbn, p.sbk, n0, p.n, (kt + 1) * BK, kend, rb);
        }
        const T* as = As + buf * BM * PK + (wm * 64) * PK;
        const T* bs = Bs + buf * BN * PK + (wn * 32) * PK;
        if constexpr (sizeof(T) == 8) {
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
        } else {
#pragma unroll
            for (int ks = 0; ks < BK / 8; ++ks) {
                uint32_t af[MT][4], bf[NT][2];
#pragma unroll
                for (int i = 0; i < MT; ++i) {
                    const float* base = as + (i * 16 + g) * PK + ks * 8 + t;
                    af[i][0] = __float_as_uint(base[0]);
                    af[i][1] = __float_as_uint(base[8 * PK]);
                    af[i][2] = __float_as_uint(base[4]);
                    af[i][3] = __float_as_uint(base[8 * PK + 4]);
                }
#pragma unroll
                for (int j = 0; j < NT; ++j) {
                    const float* base = bs + (j * 8 + g) * PK + ks * 8 + t;
                    bf[j][0] = __float_as_uint(base[0]);
                    bf[j][1] = __float_as_uint(base[4]);
                }
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int j = 0; j < NT; ++j) mma_tf32(acc[i][j], af[i], bf[j]);
            }
        }
        if (more) {
            store_tile<T, BM, BK, PK>(As + (buf ^ 1) * BM * PK, p.sak, ra);
            store_tile<T, BN, BK, PK>(Bs + (buf ^ 1) * BN * PK, p.sbk, rb);
        }
        __syncthreads();
    }

    // epilogue
    T* C = (T*)p.C;
    const T alpha = (T)p.alpha, beta = (T)p.beta;
    const bool atomic = p.ksplit > 1;
    constexpr int RT = (sizeof(T) == 8) ? 8 : 16;  // rows per m-tile
#pragma unroll
    for (int i = 0; i < MT; ++i) {
#pragma unroll
        for (int h = 0; h < NACC / 2; ++h) {
            const int64_t row = m0 + wm * 64 + i * RT + g + h * 8;
            if (row >= p.m) continue;
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const int64_t col = n0 + wn * 32 + j * 8 + 2 * t;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (col + q < p.n) {
                        T* dst = C + row * p.ldc + col + q;
                        const T v = alpha * acc[i][j][h * 2 + q];
                        if (atomic) atomicAdd(dst, v);
                        else *dst = (beta == T(0)) ? v : fma(beta, *dst, v);
                    }
                }
            }
        }
    }
}

// C *= beta (or zero) ahead of a split-K accumulation
template <typename T>
__global__ void scale_kernel(T* C, int64_t m, int64_t n, int64_t ldc, T beta, int lower) {
    const int64_t total = m * n;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / n, c = e % n;
        if (lower && (c / 128) * 128 > (r / 128) * 128 + 127) continue;
        T* d = C + r * ldc + c;
        *d = (beta == T(0)) ? T(0) : beta * *d;
    }
}

template <typename T>
int launch_gemm(GemmParams p, cudaStream_t st) {
    constexpr int BM = Tile<T>::BM, BN = Tile<T>::BN, BK = Tile<T>::BK, PK = Tile<T>::PK;
    const size_t smem = (size_t)2 * (BM + BN) * PK * sizeof(T);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        attr_set = true;
    }
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
    p.ksplit = ksplit;
    if (ksplit > 1) {
        int64_t total = p.m * p.n;
        unsigned g = (unsigned)min((int64_t)(num_sms() * 8), (total + 255) / 256);
        scale_kernel<T><<<g, 256, 0, st>>>((T*)p.C, p.m, p.n, p.ldc, (T)p.beta, p.lower & 1);
    }
    dim3 grid((unsigned)gn, (unsigned)gm, (unsigned)ksplit);
    gemm_kernel<T><<<grid, 256, smem, st>>>(p);
    DCGP_CHECK_LAUNCH();
    return 0;
}

}  // namespace

// internal entry used by the Cholesky / TRSM drivers (same translation-unit-independent ABI)
int dcgp_gemm_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype, int flags,
                   cudaStream_t st) {
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
    // bit0: lower-only tiles; bit1: caller aliases C with an input, never split K
    p.lower = ((flags & DCGP_GEMM_LOWER) ? 1 : 0) | ((flags & DCGP_GEMM_NO_SPLITK) ? 2 : 0);
    p.ksplit = 1;
    if (dtype == DCGP_F64) return launch_gemm<double>(p, st);
    if (dtype == DCGP_F32) return launch_gemm<float>(p, st);
    return -14;
}

extern "C" int dcgp_gemm(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                         int64_t lda, const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype,
                         int flags, void* stream) {
    return dcgp_gemm_impl(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, dtype, flags,
                          (cudaStream_t)stream);
}
