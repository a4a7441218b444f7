// Microbenchmark of the serial pivot chain of the factorisation leaves (csrc/chol.cu: diag_step): variants of the
// in-register Cholesky of a 16 x 16 (or 8 x 8) diagonal block by ONE warp, timed with clock64 inside a single CTA.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/leafbench tools/leafbench.cu   (build container)
//   tools/leafbench                                                                                      (GPU box)
//
// Every variant factors the eight 16 x 16 diagonal blocks of one 128 x 128 SPD tile held in shared memory (pitch
// 129, as the leaf kernels hold it), one after the other, exactly as the chain of a leaf does; printed: cycles per
// block (best of 5) and the error of the factors against a host FP64 Cholesky.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int NB = 128, SP = NB + 1, SB = 16;

template <typename T> __device__ __forceinline__ T rsq_newton(T d);
template <> __device__ __forceinline__ float rsq_newton<float>(float d) {
    float y = rsqrtf(d);
    return y * fmaf(-0.5f * d, y * y, 1.5f);
}
template <> __device__ __forceinline__ double rsq_newton<double>(double d) {
    double y = (double)rsqrtf((float)d);
    y = y * fma(-0.5 * d, y * y, 1.5);
    return y * fma(-0.5 * d, y * y, 1.5);
}
// flush-to-zero hardware seed: no denormal fix-up instructions around MUFU.RSQ
template <typename T> __device__ __forceinline__ T rsq_ftz(T d);
template <> __device__ __forceinline__ float rsq_ftz<float>(float d) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    return y * fmaf(-0.5f * d, y * y, 1.5f);
}
template <> __device__ __forceinline__ double rsq_ftz<double>(double d) {
    float yf;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"((float)d));
    double y = (double)yf;
    y = y * fma(-0.5 * d, y * y, 1.5);
    return y * fma(-0.5 * d, y * y, 1.5);
}

#define FULL 0xffffffffu

// V0: the shipped diag_step (lane r = row r, 16 columns in registers, lanes >= 16 identity padding)
// V1: PIV2 (next pivot broadcast from lane c+1's own registers)
// V2: V0 with the ftz seed
template <typename T, int VAR>
__device__ __forceinline__ void diag_v012(T* S, T* rinv, int p0) {
    const int r = threadIdx.x & 31;
    T row[SB];
    T myinv = T(0);
#pragma unroll
    for (int c = 0; c < SB; ++c) row[c] = (r < SB && c <= r) ? S[(p0 + r) * SP + p0 + c] : ((r == c) ? T(1) : T(0));
    T dnext = (VAR == 1) ? __shfl_sync(FULL, row[0], 0) : T(0);
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        const T d = (VAR == 1) ? dnext : __shfl_sync(FULL, row[c], c);
        const T y = (VAR == 2) ? rsq_ftz<T>(d) : rsq_newton<T>(d);
        const T scaled = row[c] * y;
        row[c] = (r == c) ? d * y : scaled;
        if (r == c) myinv = y;
        const T lrc = row[c];
        if (VAR == 1 && c + 1 < SB) dnext = __shfl_sync(FULL, fma(-lrc, lrc, row[c + 1]), c + 1);
#pragma unroll
        for (int k = c + 1; k < SB; ++k) {
            const T lkc = __shfl_sync(FULL, row[c], k);
            row[k] = fma(-lrc, lkc, row[k]);
        }
    }
#pragma unroll
    for (int c = 0; c < SB; ++c)
        if (r < SB && c <= r) S[(p0 + r) * SP + p0 + c] = row[c];
    if (r < SB) rinv[p0 + r] = myinv;
}

// V3: two columns per latency round.  With a = A[c][c], b = A[c+1][c], e = A[c+1][c+1] (all already updated):
//   y1 = 1/sqrt(a), y2' = 1/sqrt(a e - b^2) - two INDEPENDENT rsqrt - and 1/sqrt(e - b^2/a) = (a y1) y2'.
// W = block width (16 or 8); ROWS32: lanes >= W carry real rows below the block (rows p0 + r) instead of padding.
template <typename T, int W, bool ROWS32, bool FTZ>
__device__ __forceinline__ void diag_pair(T* S, T* rinv, int p0, int nrows) {
    const int r = threadIdx.x & 31;
    T row[W];
    T myinv = T(0);
    const bool real = ROWS32 ? (r < nrows) : (r < W);
#pragma unroll
    for (int c = 0; c < W; ++c) row[c] = (real && (c <= r)) ? S[(p0 + r) * SP + p0 + c] : ((r == c) ? T(1) : T(0));
#pragma unroll
    for (int c = 0; c < W; c += 2) {
        const T a = __shfl_sync(FULL, row[c], c);
        const T b = __shfl_sync(FULL, row[c], c + 1);
        const T e = __shfl_sync(FULL, row[c + 1], c + 1);
        const T det = fma(a, e, -(b * b));
        const T y1 = FTZ ? rsq_ftz<T>(a) : rsq_newton<T>(a);
        const T yd = FTZ ? rsq_ftz<T>(det) : rsq_newton<T>(det);
        const T sa = a * y1;            // sqrt(a)
        const T y2 = sa * yd;           // 1 / sqrt(e - b^2 / a)
        const T l21 = b * y1;
        const T l1 = (r == c) ? sa : row[c] * y1;
        const T t = fma(-l1, l21, row[c + 1]);
        const T l2 = (r == c + 1) ? (det * yd) * y1 : t * y2;   // sqrt(det / a) on the diagonal
        row[c] = l1;
        row[c + 1] = l2;
        if (r == c) myinv = y1;
        if (r == c + 1) myinv = y2;
#pragma unroll
        for (int k = c + 2; k < W; ++k) {
            const T lk1 = __shfl_sync(FULL, l1, k);
            const T lk2 = __shfl_sync(FULL, l2, k);
            row[k] = fma(-l2, lk2, fma(-l1, lk1, row[k]));
        }
    }
#pragma unroll
    for (int c = 0; c < W; ++c)
        if (real && (c <= r || r >= W)) S[(p0 + r) * SP + p0 + c] = row[c];
    if (r < W) rinv[p0 + r] = myinv;
}

// V6: one column per round, but the column of L travels through shared memory (1 store + broadcast vector loads)
// instead of W shuffles per lane
template <typename T, int W>
__device__ __forceinline__ void diag_smem(T* S, T* rinv, T* colbuf, int p0) {
    const int r = threadIdx.x & 31;
    T row[W];
    T myinv = T(0);
#pragma unroll
    for (int c = 0; c < W; ++c) row[c] = (r < W && c <= r) ? S[(p0 + r) * SP + p0 + c] : ((r == c) ? T(1) : T(0));
#pragma unroll
    for (int c = 0; c < W; ++c) {
        const T d = __shfl_sync(FULL, row[c], c);
        const T y = rsq_ftz<T>(d);
        const T lrc = (r == c) ? d * y : row[c] * y;
        row[c] = lrc;
        if (r == c) myinv = y;
        colbuf[(c & 1) * 32 + r] = lrc;
        __syncwarp();
#pragma unroll
        for (int k = c + 1; k < W; ++k) row[k] = fma(-lrc, colbuf[(c & 1) * 32 + k], row[k]);
    }
#pragma unroll
    for (int c = 0; c < W; ++c)
        if (r < W && c <= r) S[(p0 + r) * SP + p0 + c] = row[c];
    if (r < W) rinv[p0 + r] = myinv;
}

template <typename T, int VAR>
__global__ void __launch_bounds__(512) bench(const T* __restrict__ A, T* __restrict__ L, long long* cyc) {
    extern __shared__ __align__(16) unsigned char raw[];
    T* S = reinterpret_cast<T*>(raw);
    __shared__ T rinv[NB];
    __shared__ T colbuf[64];
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) S[(e >> 7) * SP + (e & 127)] = A[e];
    __syncthreads();
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        t0 = clock64();
        if (VAR <= 2) {
            for (int p0 = 0; p0 < NB; p0 += SB) diag_v012<T, VAR>(S, rinv, p0);
        } else if (VAR == 3) {
            for (int p0 = 0; p0 < NB; p0 += SB) diag_pair<T, 16, false, false>(S, rinv, p0, 16);
        } else if (VAR == 4) {
            for (int p0 = 0; p0 < NB; p0 += SB) diag_pair<T, 16, false, true>(S, rinv, p0, 16);
        } else if (VAR == 5) {   // 32 rows ride along (block + the 16 rows below it); the last block has none below
            for (int p0 = 0; p0 < NB; p0 += SB) diag_pair<T, 16, true, true>(S, rinv, p0, min(32, NB - p0));
        } else if (VAR == 6) {
            for (int p0 = 0; p0 < NB; p0 += SB) diag_smem<T, 16>(S, rinv, colbuf, p0);
        } else if (VAR == 7) {   // 8-wide blocks, 32 rows: sixteen blocks per tile (cycles are still reported per 16 columns)
            for (int p0 = 0; p0 < NB; p0 += 8) diag_pair<T, 8, true, true>(S, rinv, p0, min(32, NB - p0));
        } else if (VAR == 8) {   // 8-wide blocks, block rows only
            for (int p0 = 0; p0 < NB; p0 += 8) diag_pair<T, 8, false, true>(S, rinv, p0, 8);
        }
        t1 = clock64();
    }
    __syncthreads();
    for (int e = threadIdx.x; e < NB * NB; e += blockDim.x) L[e] = S[(e >> 7) * SP + (e & 127)];
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <typename T, int VAR>
void run(const char* name, const std::vector<double>& A64, int width, bool rows32) {
    std::vector<T> hA(NB * NB);
    for (int i = 0; i < NB * NB; ++i) hA[i] = (T)A64[i];
    T *dA, *dL;
    long long* dc;
    cudaMalloc(&dA, NB * NB * sizeof(T));
    cudaMalloc(&dL, NB * NB * sizeof(T));
    cudaMalloc(&dc, 8);
    cudaMemcpy(dA, hA.data(), NB * NB * sizeof(T), cudaMemcpyHostToDevice);
    const size_t smem = (size_t)NB * SP * sizeof(T);
    cudaFuncSetAttribute(bench<T, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long best = 1LL << 60;
    for (int it = 0; it < 5; ++it) {
        bench<T, VAR><<<1, 512, smem>>>(dA, dL, dc);
        long long c;
        cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        if (c < best) best = c;
    }
    cudaError_t err = cudaDeviceSynchronize();
    std::vector<T> hL(NB * NB);
    cudaMemcpy(hL.data(), dL, NB * NB * sizeof(T), cudaMemcpyDeviceToHost);
    // reference: Cholesky of every width x width diagonal block (FP64), and the solved rows below when rows32
    double maxerr = 0;
    for (int p0 = 0; p0 < NB; p0 += width) {
        std::vector<double> Lr(width * width, 0.0);
        for (int j = 0; j < width; ++j) {
            double s = A64[(p0 + j) * NB + p0 + j];
            for (int k = 0; k < j; ++k) s -= Lr[j * width + k] * Lr[j * width + k];
            Lr[j * width + j] = std::sqrt(s);
            for (int i = j + 1; i < width; ++i) {
                double t = A64[(p0 + i) * NB + p0 + j];
                for (int k = 0; k < j; ++k) t -= Lr[i * width + k] * Lr[j * width + k];
                Lr[i * width + j] = t / Lr[j * width + j];
            }
        }
        for (int i = 0; i < width; ++i)
            for (int j = 0; j <= i; ++j) maxerr = std::fmax(maxerr, std::fabs((double)hL[(p0 + i) * NB + p0 + j] - Lr[i * width + j]));
        if (rows32) {
            for (int i = p0 + width; i < std::min(p0 + 32, NB); ++i)
                for (int j = 0; j < width; ++j) {
                    double t = A64[i * NB + p0 + j];   // rows below see the ORIGINAL entries (blocks are independent here)
                    for (int k = 0; k < j; ++k) t -= (double)hL[i * NB + p0 + k] * Lr[j * width + k];
                    // entries of a row below were overwritten by an earlier (narrower-window) pass only if that pass
                    // covered them: with independent blocks they were not, so compare directly
                    maxerr = std::fmax(maxerr, std::fabs((double)hL[i * NB + p0 + j] - t / Lr[j * width + j]));
                }
        }
    }
    printf("%-44s %s  %8.0f cycles per 16 columns  (%6.2f us per 128-leaf chain at 1.965 GHz)  max err %.2e  %s\n", name,
           sizeof(T) == 8 ? "f64" : "f32", (double)best / 8.0, (double)best / 1.965e3, maxerr,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(dA);
    cudaFree(dL);
    cudaFree(dc);
}

int main() {
    // SPD tile whose diagonal blocks are well conditioned and whose below-block rows are small enough that the
    // "rows ride along" variants can be compared block by block
    std::vector<double> A(NB * NB);
    srand(1);
    std::vector<double> M(NB * NB);
    for (auto& v : M) v = (rand() / (double)RAND_MAX) - 0.5;
    for (int i = 0; i < NB; ++i)
        for (int j = 0; j < NB; ++j) {
            double s = 0;
            for (int k = 0; k < NB; ++k) s += M[i * NB + k] * M[j * NB + k];
            A[i * NB + j] = s / NB + (i == j ? 1.0 : 0.0);
        }
    run<float, 0>("V0 shipped diag_step", A, 16, false);
    run<float, 1>("V1 PIV2", A, 16, false);
    run<float, 2>("V2 ftz seed", A, 16, false);
    run<float, 3>("V3 two columns per round", A, 16, false);
    run<float, 4>("V4 two columns per round + ftz", A, 16, false);
    run<float, 5>("V5 V4 + 16 rows below ride along", A, 16, true);
    run<float, 6>("V6 column through shared memory + ftz", A, 16, false);
    run<float, 7>("V7 8-wide blocks, 32 rows, pairs + ftz", A, 8, true);
    run<float, 8>("V8 8-wide blocks, pairs + ftz", A, 8, false);
    run<double, 0>("V0 shipped diag_step", A, 16, false);
    run<double, 1>("V1 PIV2", A, 16, false);
    run<double, 2>("V2 ftz seed", A, 16, false);
    run<double, 3>("V3 two columns per round", A, 16, false);
    run<double, 4>("V4 two columns per round + ftz", A, 16, false);
    run<double, 5>("V5 V4 + 16 rows below ride along", A, 16, true);
    run<double, 6>("V6 column through shared memory + ftz", A, 16, false);
    run<double, 7>("V7 8-wide blocks, 32 rows, pairs + ftz", A, 8, true);
    run<double, 8>("V8 8-wide blocks, pairs + ftz", A, 8, false);
    return 0;
}
