// Recursive blocked Cholesky and triangular solves (row-major, lower).
//
//   chol(A[nb]):  nb <= 128 -> potrf_diag_kernel (one CTA: factor the block in shared memory and
//                              invert its triangular factor into the workspace)
//                 else      -> chol(A11); A21 <- A21 L11^{-T} (rtrsm); A22 -= A21 A21^T; chol(A22)
//   rtrsm(B, L[nt]): nt <= 128 -> B <- B inv(L)^T as one tensor-core GEMM (in place, K = nt)
//                    else      -> rtrsm(B1, La); B2 -= B1 Lc^T; rtrsm(B2, Lb)
//
// The recursion halves the block at every level, so all but ~n^2*128 of the n^3/3 flops run in
// tensor-core GEMMs whose inner dimension is as deep as the problem allows (n/2 at the top):
// DMMA (mma.sync m8n8k4.f64) for FP64, TF32 for FP32 - see gemm.cu.  The only CUDA-core work is
// the 128x128 diagonal block (O(n * 128^2) in total), which is register-blocked in 16-wide
// sub-panels so that it costs a few tens of microseconds on the critical path.
//
// Left-side solves op(L) X = B with many right-hand sides recurse the same way.
//
// LAPACK-style failure reporting: *info (device) receives the 1-based index of the first
// non-positive pivot; nothing here synchronises the stream.
#include "common.cuh"
#include "mma.cuh"
#include "tc.cuh"
#include <cstdio>
#include <cstdlib>
#include <unordered_map>
#include <vector>

int dcgp_gemm_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype, int flags,
                   cudaStream_t st, const void* Alo, const void* Blo, void* Clo);
int dcgp_split_lo_impl(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, cudaStream_t st);
int dcgp_gemm_batched_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A,
                           int64_t lda, int64_t strideA, const void* B, int64_t ldb, int64_t strideB, double beta,
                           void* C, int64_t ldc, int64_t strideC, int batch, int dtype, int flags, cudaStream_t st);

int dcgp_gemm_tc_ex(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                    const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int flags, const void* Alo, const void* Blo,
                    void* Clo, int64_t lower_off, int64_t clo_cols, cudaStream_t st);

void dcgp_tc_cap(int ctas);
int dcgp_potrf_link_tm(float* A, int64_t lda, int nb, float* inv, float* invlo, int* info, int64_t j0, const float* R,
                       const float* invp, cudaStream_t st);

namespace {

constexpr int NB = 128;     // leaf block size
constexpr int SP = NB + 1;  // shared-memory pitch (conflict-free row AND column walks)
constexpr int NTHREADS = 512;
constexpr int SB = 16;      // sub-panel width inside the leaf

// ---------------------------------------------------------------------------------------------
// 1/sqrt(d) without the IEEE sqrt / div subroutines (they dominate the serial pivot chain):
// hardware rsqrt seed + Newton steps y <- y (1.5 - 0.5 d y^2).  d <= 0 or NaN yields NaN.
template <typename T> __device__ __forceinline__ T fast_rsqrt(T d);
template <> __device__ __forceinline__ double fast_rsqrt<double>(double d) {
    // seed: 2 ulp of FP32 (2.4e-7); each Newton step squares the relative error (x 1.5): 8.6e-14, then 1.1e-26
    double y = (double)rsqrtf((float)d);
    y = y * fma(-0.5 * d, y * y, 1.5);
    return y * fma(-0.5 * d, y * y, 1.5);
}

// (1) of block_potrf, executed by ONE warp: factor the pw x pw diagonal sub-block at p0 in registers.  A single warp
// is latency-bound on instruction count, so the loop is predicate-free: lanes / columns outside the valid block
// behave as an identity padding and every lane updates all of its registers (entries above the diagonal are junk
// that is never stored).
// Measured in isolation on B200 (tools/leafbench.cu, cycles per 16 columns): this column-by-column form 3242 (FP32) /
// 3962 (FP64); broadcasting the next pivot from lane c+1's own registers ("PIV2", round 1) 3663 / 3661 - no gain,
// removed; the two-columns-per-round form below 1496 for FP32, but 3953 for FP64 (DFMA latency: kept column-wise).
template <typename T>
__device__ __forceinline__ void diag_step(T* S, T* rinv, int p0, int pw, int* info, int64_t j0) {
    const int lane = threadIdx.x & 31;
    const int r = lane;  // row within the sub-block
    T row[SB];
    T myinv = T(0);
    int bad = 0;
#pragma unroll
    for (int c = 0; c < SB; ++c) row[c] = (r < pw && c <= r) ? S[(p0 + r) * SP + p0 + c] : ((r == c) ? T(1) : T(0));
#pragma unroll
    for (int c = 0; c < SB; ++c) {
        const T d = __shfl_sync(0xffffffffu, row[c], c);  // pivot (already updated)
        if (!(d > T(0)) && bad == 0) bad = c + 1;
        const T y = fast_rsqrt<T>(d);
        const T scaled = row[c] * y;
        row[c] = (r == c) ? d * y : scaled;
        if (r == c) myinv = y;
        const T lrc = row[c];
#pragma unroll
        for (int k = c + 1; k < SB; ++k) {
            const T lkc = __shfl_sync(0xffffffffu, row[c], k);  // L[k][c]
            row[k] = fma(-lrc, lkc, row[k]);
        }
    }
#pragma unroll
    for (int c = 0; c < SB; ++c)
        if (r < pw && c <= r) S[(p0 + r) * SP + p0 + c] = row[c];
    if (r < pw) rinv[p0 + r] = myinv;
    if (lane == 0 && bad != 0 && bad <= pw && info && *info == 0) *info = (int)(j0 + p0 + bad);
}

// FP32: two columns per latency round.  With a = A[c][c], b = A[c+1][c], e = A[c+1][c+1] (all already updated),
//   y1 = 1/sqrt(a) and yd = 1/sqrt(a e - b^2) are INDEPENDENT reciprocal square roots (the second pivot is
//   (a e - b^2) / a, so 1/sqrt(pivot2) = sqrt(a) yd), and the two shuffle rounds of a column pair collapse into one.
// The hardware seed is taken flush-to-zero (no denormal fix-up instructions around MUFU.RSQ; a pivot below 1e-38 is a
// failed factorisation in FP32 anyway) and refined by one Newton step as before.
__device__ __forceinline__ float rsqrt_ftz_newton(float d) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    return y * fmaf(-0.5f * d, y * y, 1.5f);
}
template <>
__device__ __forceinline__ void diag_step<float>(float* S, float* rinv, int p0, int pw, int* info, int64_t j0) {
    const int lane = threadIdx.x & 31;
    const int r = lane;
    float row[SB];
    float myinv = 0.f;
    int bad = 0;
#pragma unroll
    for (int c = 0; c < SB; ++c) row[c] = (r < pw && c <= r) ? S[(p0 + r) * SP + p0 + c] : ((r == c) ? 1.f : 0.f);
#pragma unroll
    for (int c = 0; c < SB; c += 2) {
        const float a = __shfl_sync(0xffffffffu, row[c], c);
        const float b = __shfl_sync(0xffffffffu, row[c], c + 1);
        const float e = __shfl_sync(0xffffffffu, row[c + 1], c + 1);
        const float det = fmaf(a, e, -(b * b));
        if (bad == 0) {
            if (!(a > 0.f)) bad = c + 1;
            else if (!(det > 0.f)) bad = c + 2;
        }
        const float y1 = rsqrt_ftz_newton(a);
        const float yd = rsqrt_ftz_newton(det);
        const float sa = a * y1;            // sqrt(a)
        const float y2 = sa * yd;           // 1 / sqrt(e - b^2 / a)
        const float l21 = b * y1;
        const float l1 = (r == c) ? sa : row[c] * y1;
        const float t = fmaf(-l1, l21, row[c + 1]);
        const float l2 = (r == c + 1) ? (det * yd) * y1 : t * y2;   // sqrt(det / a) on the diagonal
        row[c] = l1;
        row[c + 1] = l2;
        if (r == c) myinv = y1;
        if (r == c + 1) myinv = y2;
#pragma unroll
        for (int k = c + 2; k < SB; ++k) {
            const float lk1 = __shfl_sync(0xffffffffu, l1, k);
            const float lk2 = __shfl_sync(0xffffffffu, l2, k);
            row[k] = fmaf(-l2, lk2, fmaf(-l1, lk1, row[k]));
        }
    }
#pragma unroll
    for (int c = 0; c < SB; ++c)
        if (r < pw && c <= r) S[(p0 + r) * SP + p0 + c] = row[c];
    if (r < pw) rinv[p0 + r] = myinv;
    if (lane == 0 && bad != 0 && bad <= pw && info && *info == 0) *info = (int)(j0 + p0 + bad);
}

// rank-pw update of the trailing lower triangle on the tensor cores: each participating warp takes 8x8 (DMMA) /
// 16x8 (3xTF32) output tiles, K = 16, operands read straight from the tile.  Column tiles [nj0, nj1) only;
// the tiles are dealt to warps wfirst .. wfirst + nwarps - 1.
template <typename T>
__device__ __forceinline__ void trailing_update(T* S, int nb, int p0, int pw, int nj0, int nj1, int wfirst, int nwarps) {
    constexpr int TR = WarpTile<T>::ROWS, TC = WarpTile<T>::COLS, NA = WarpTile<T>::NACC;
    const int warp = threadIdx.x >> 5;
    if (warp < wfirst || warp >= wfirst + nwarps) return;
    const int t0 = p0 + pw, rr = nb - t0;
    const int ntr = (rr + TR - 1) / TR, ntc = (rr + TC - 1) / TC;
    if (nj1 > ntc) nj1 = ntc;
    const int ncol = nj1 - nj0;
    if (ncol <= 0) return;
    for (int e = warp - wfirst; e < ntr * ncol; e += nwarps) {
        const int mi = e / ncol, nj = nj0 + (e - mi * ncol);
        const int row0 = t0 + mi * TR, col0 = t0 + nj * TC;
        if (col0 > row0 + TR - 1) continue;  // tile entirely above the diagonal
        T acc[4] = {T(0), T(0), T(0), T(0)};
        warp_tile_mma(
            acc, SB, [&](int rl, int k) { return (row0 + rl < nb && k < pw) ? S[(row0 + rl) * SP + p0 + k] : T(0); },
            [&](int k, int cl) { return (col0 + cl < nb && k < pw) ? S[(col0 + cl) * SP + p0 + k] : T(0); });
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            int rl, cl;
            warp_tile_coord<T>(q, rl, cl);
            const int row = row0 + rl, col = col0 + cl;
            if (row < nb && col <= row) S[row * SP + col] -= acc[q];
        }
    }
}

// ---- inverse riding along with the factorisation (FINV) ------------------------------------------------------
// X = L^{-1} is wanted next to L (the panel solves of the callers are products with it).  Row c of L^{-T} is the
// solution y of  y L^T = e_c  - the very operation step (2) of block_potrf applies to the rows below the pivots.  With
// FINV the unit rows e_c ride along as extra rows of the tile: they live in the unused upper part, shifted by one
// column, in the layout block_trtri produces ( X(i, c), i >= c  <->  S[c][i + 1] ), get their 16 columns solved in
// step (2) by threads the rows below leave idle, and their rank-16 updates in step (3b) by warps 1..15, which wait
// there for warp 0's pivot chain anyway (B200, one 128-block: pivots 2.4 us per sub-panel, block_trtri 17.6 us per
// leaf = 31 % of a link of the look-ahead chain; profiles/r01_leaf_timing_v31.txt).  Entries left of the diagonal of an
// extra row are identically zero and are neither loaded nor stored (that storage belongs to L).
template <typename T>
__device__ __forceinline__ void inv_rows_init(T* S, int nb) {   // called by warps 1 .. 15 (warp 0 is in diag_step)
    for (int e = (int)threadIdx.x - 32; e < NB * NB; e += NTHREADS - 32) {
        const int c = e >> 7, k = e & (NB - 1);
        if (c < nb && k < nb && k >= c) S[c * SP + k + 1] = (k == c) ? T(1) : T(0);
    }
}

// step (2) for extra row c < p0 + pw: y[p0 .. p0+pw) = residual * L_d^{-T}
template <typename T>
__device__ __forceinline__ void inv_rows_panel(T* S, const T* rinv, int p0, int pw) {
    const int c = NTHREADS - 1 - (int)threadIdx.x;   // from the top: the rows below use the low thread ids
    if (c >= p0 + pw) return;
    T x[SB];
#pragma unroll
    for (int kk = 0; kk < SB; ++kk) x[kk] = (kk < pw && p0 + kk >= c) ? S[c * SP + p0 + kk + 1] : T(0);
#pragma unroll
    for (int kk = 0; kk < SB; ++kk) {
        if (kk < pw) {
            T s = x[kk];
#pragma unroll
            for (int k = 0; k < SB; ++k)
                if (k < kk) s = fma(-x[k], S[(p0 + kk) * SP + p0 + k], s);
            x[kk] = s * rinv[p0 + kk];
        }
    }
#pragma unroll
    for (int kk = 0; kk < SB; ++kk)
        if (kk < pw && p0 + kk >= c) S[c * SP + p0 + kk + 1] = x[kk];
}

// step (3) for the extra rows [0, t0), t0 = p0 + pw: residual[c][k] -= sum_kk y[c][p0 + kk] L[k][p0 + kk], k in [t0, nb)
template <typename T>
__device__ __forceinline__ void inv_rows_update(T* S, int nb, int p0, int pw, int wfirst, int nwarps) {
    constexpr int TR = WarpTile<T>::ROWS, TC = WarpTile<T>::COLS, NA = WarpTile<T>::NACC;
    const int warp = threadIdx.x >> 5;
    if (warp < wfirst || warp >= wfirst + nwarps) return;
    const int t0 = p0 + pw;
    const int ntr = (t0 + TR - 1) / TR, ntc = (nb - t0 + TC - 1) / TC;
    for (int e = warp - wfirst; e < ntr * ntc; e += nwarps) {
        const int mi = e / ntc, nj = e - mi * ntc;
        const int row0 = mi * TR, col0 = t0 + nj * TC;
        T acc[4] = {T(0), T(0), T(0), T(0)};
        warp_tile_mma(
            acc, SB,
            [&](int rl, int k) {
                const int c = row0 + rl;
                return (c < t0 && k < pw && p0 + k >= c) ? S[c * SP + p0 + k + 1] : T(0);
            },
            [&](int k, int cl) { return (col0 + cl < nb && k < pw) ? S[(col0 + cl) * SP + p0 + k] : T(0); });
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            int rl, cl;
            warp_tile_coord<T>(q, rl, cl);
            const int c = row0 + rl, col = col0 + cl;
            if (c < t0 && col < nb) S[c * SP + col + 1] -= acc[q];
        }
    }
}

// Right-looking factorisation over 16-wide sub-panels with an in-tile look-ahead:
//   (1) warp 0 factors the 16x16 diagonal block (diag_step);
//   (2) every thread owning a row below solves its 16 entries against that block (division-free);
//   (3a) the NEXT sub-panel's 16 columns get their rank-16 update first (all warps, short);
//   (3b) warp 0 factors the next diagonal block while warps 1..15 update the rest of the trailing triangle.
// Three barriers per sub-panel; the serial pivot chain of (1) is hidden behind (3b) whenever that is longer.
template <typename T, int MODE = 0>   // MODE bit 0: FINV (the inverse rides along)
__device__ void block_potrf(T* S, T* rinv, int nb, int* info, int64_t j0) {
    constexpr bool FINV = (MODE & 1) != 0;
    const int tid = threadIdx.x, warp = tid >> 5;
    constexpr int TC = WarpTile<T>::COLS, NJ_NEXT = SB / TC, NW = NTHREADS / 32;
#ifdef DCGP_LEAF_TIMING
    long long acc1 = 0, acc2 = 0, acc3 = 0, tq = clock64();
#endif
    if (warp == 0) diag_step<T>(S, rinv, 0, min(SB, nb), info, j0);
    else if (FINV) inv_rows_init<T>(S, nb);
    __syncthreads();
    for (int p0 = 0; p0 < nb; p0 += SB) {
        const int pw = min(SB, nb - p0);
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); acc1 += t - tq; tq = t; }
#endif
        // (2) column sub-panel below: row i solves x L_d^T = a  (L_d = the 16x16 block just factored)
        for (int i = p0 + pw + tid; i < nb; i += NTHREADS) {
            T x[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) x[c] = (c < pw) ? S[i * SP + p0 + c] : T(0);
#pragma unroll
            for (int c = 0; c < SB; ++c) {
                if (c < pw) {
                    T s = x[c];
#pragma unroll
                    for (int k = 0; k < SB; ++k)
                        if (k < c) s = fma(-x[k], S[(p0 + c) * SP + p0 + k], s);
                    x[c] = s * rinv[p0 + c];
                }
            }
#pragma unroll
            for (int c = 0; c < SB; ++c)
                if (c < pw) S[i * SP + p0 + c] = x[c];
        }
        if (FINV) inv_rows_panel<T>(S, rinv, p0, pw);
        __syncthreads();
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); acc2 += t - tq; tq = t; }
#endif
        const int t0 = p0 + pw;
        if (t0 >= nb) break;
        // (3a) the next sub-panel's columns
        trailing_update<T>(S, nb, p0, pw, 0, NJ_NEXT, 0, NW);
        __syncthreads();
        // (3b) next pivots || rest of the trailing update (|| the extra rows of the inverse)
        if (warp == 0) diag_step<T>(S, rinv, t0, min(SB, nb - t0), info, j0);
        else {
            trailing_update<T>(S, nb, p0, pw, NJ_NEXT, 1 << 30, 1, NW - 1);
            if (FINV) inv_rows_update<T>(S, nb, p0, pw, 1, NW - 1);
        }
        __syncthreads();
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); acc3 += t - tq; tq = t; }
#endif
    }
#ifdef DCGP_LEAF_TIMING
    if (tid == 0) printf("  potrf steps: (first diag) %lld  panel %lld  trailing+next diag %lld\n", acc1, acc2, acc3);
#endif
}

// Invert the lower-triangular factor held in S by recursive doubling.  X = L^{-1} is kept in the
// unused upper part of the tile, transposed and shifted by one column:
//   X(i, c), i >= c   <->   S[c][i + 1]        (pitch 129, so column 128 exists)
// Level 0 inverts the 16x16 diagonal blocks (one column per thread, division-free with rinv[]);
// level b in {16, 32, 64} completes every 2b x 2b diagonal block from its two b x b halves:
//   X21 = -X22 (L21 X11)      as two tiled products (4x4 register tiles, all 256 threads).
// L21 is dead once T = L21 X11 is formed, so its storage receives X21 before it moves to XX.
// Requires the factor to have been saved elsewhere: the lower triangle of S is destroyed.
#define XX(i, c) S[(c) * SP + (i) + 1]
template <typename T>
__device__ void block_trtri(T* S, const T* rinv, int nb) {
    const int tid = threadIdx.x;
#ifdef DCGP_LEAF_TIMING
    long long tA = 0, tB = 0, tC = 0, tq = clock64(), t0c;
#endif
    for (int e = tid; e < nb; e += NTHREADS) {
        const int d0 = (e / SB) * SB, c = e;  // column c of the sub-block starting at d0
        const int dw = min(SB, nb - d0);
        T x[SB];
#pragma unroll
        for (int r = 0; r < SB; ++r) {
            const int i = d0 + r;
            T s = (i == c) ? T(1) : T(0);
            if (r < dw && i >= c) {
#pragma unroll
                for (int k = 0; k < SB; ++k)
                    if (k < r && d0 + k >= c) s = fma(-S[i * SP + d0 + k], x[k], s);
                x[r] = s * rinv[i];
            } else {
                x[r] = T(0);
            }
        }
#pragma unroll
        for (int r = 0; r < SB; ++r)
            if (r < dw && d0 + r >= c) XX(d0 + r, c) = x[r];
    }
    __syncthreads();
#ifdef DCGP_LEAF_TIMING
    { long long t = clock64(); t0c = t - tq; tq = t; }
#endif
    for (int b = SB; b < nb; b *= 2) {
        const int npairs = (nb + 2 * b - 1) / (2 * b);
        // step A: T = L21 X11 -> parked at XX(i, c), i in half 2, c in half 1 (tensor-core tiles)
        constexpr int TR = WarpTile<T>::ROWS, TC = WarpTile<T>::COLS, NA = WarpTile<T>::NACC;
        const int warp = tid >> 5;
        const int tpr = b / TR, tpc = b / TC;  // tiles per pair: rows x cols (b is a multiple of 16)
        for (int e = warp; e < npairs * tpr * tpc; e += NTHREADS / 32) {
            const int pr = e / (tpr * tpc), tt = e - pr * tpr * tpc;
            const int base = pr * 2 * b;
            const int row0 = base + b + (tt / tpc) * TR, col0 = base + (tt % tpc) * TC;
            if (row0 >= nb) continue;
            T acc[4] = {T(0), T(0), T(0), T(0)};
            warp_tile_mma(
                acc, b, [&](int rl, int k) { return (row0 + rl < nb) ? S[(row0 + rl) * SP + base + k] : T(0); },
                [&](int k, int cl) { return (base + k >= col0 + cl) ? XX(base + k, col0 + cl) : T(0); });
#pragma unroll
            for (int q = 0; q < NA; ++q) {
                int rl, cl;
                warp_tile_coord<T>(q, rl, cl);
                if (row0 + rl < nb) XX(row0 + rl, col0 + cl) = acc[q];
            }
        }
        __syncthreads();
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); tA += t - tq; tq = t; }
#endif
        // step B: X21 = -X22 T  -> written over the (dead) L21 block in the lower triangle
        for (int e = warp; e < npairs * tpr * tpc; e += NTHREADS / 32) {
            const int pr = e / (tpr * tpc), tt = e - pr * tpr * tpc;
            const int base = pr * 2 * b;
            const int row0 = base + b + (tt / tpc) * TR, col0 = base + (tt % tpc) * TC;
            if (row0 >= nb) continue;
            const int h2 = base + b;  // first row / column of the second half
            T acc[4] = {T(0), T(0), T(0), T(0)};
            warp_tile_mma(
                acc, b,
                [&](int rl, int k) {
                    const int row = row0 + rl, kk = h2 + k;
                    return (row < nb && kk <= row) ? XX(row, kk) : T(0);
                },
                [&](int k, int cl) { return (h2 + k < nb) ? XX(h2 + k, col0 + cl) : T(0); });
#pragma unroll
            for (int q = 0; q < NA; ++q) {
                int rl, cl;
                warp_tile_coord<T>(q, rl, cl);
                if (row0 + rl < nb) S[(row0 + rl) * SP + col0 + cl] = -acc[q];
            }
        }
        __syncthreads();
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); tB += t - tq; tq = t; }
#endif
        // step C: move X21 into the overlay
        for (int e = tid; e < npairs * b * b; e += NTHREADS) {
            const int pr = e / (b * b), tt = e - pr * b * b;
            const int i = pr * 2 * b + b + tt / b, c = pr * 2 * b + tt % b;
            if (i < nb) XX(i, c) = S[i * SP + c];
        }
        __syncthreads();
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); tC += t - tq; tq = t; }
#endif
    }
#ifdef DCGP_LEAF_TIMING
    if (tid == 0) printf("  trtri: level0 %lld  A %lld  B %lld  C %lld\n", t0c, tA, tB, tC);
#endif
}

template <typename T>
__device__ void store_inverse(const T* S, int nb, T* inv, int64_t ld, int ncols = NB) {  // nb x ncols block (lower, zero above)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < nb; i += NTHREADS / 32)
        for (int c = lane; c < ncols; c += 32) inv[(int64_t)i * ld + c] = (c <= i) ? XX(i, c) : T(0);
}
#undef XX

template <typename T>
__device__ void load_lower(T* S, T* rinv, const T* A, int64_t lda, int nb, bool want_rinv) {
    // flat mapping, 16 independent loads in flight per thread
    constexpr int UN = 16;
    for (int e0 = 0; e0 < NB * NB; e0 += NTHREADS * UN) {
        T v[UN];
#pragma unroll
        for (int j = 0; j < UN; ++j) {
            const int e = e0 + j * NTHREADS + threadIdx.x;
            const int i = e >> 7, k = e & (NB - 1);
            v[j] = (i < nb && k <= i) ? A[(int64_t)i * lda + k] : T(0);
        }
#pragma unroll
        for (int j = 0; j < UN; ++j) {
            const int e = e0 + j * NTHREADS + threadIdx.x;
            const int i = e >> 7, k = e & (NB - 1);
            if (i < nb && k <= i) {
                S[i * SP + k] = v[j];
                if (want_rinv && k == i) rinv[i] = T(1) / v[j];
            }
        }
    }
}

template <typename T, int MODE = 0>
__global__ void __launch_bounds__(NTHREADS) potrf_diag_kernel(T* __restrict__ A, int64_t lda, int nb,
                                                              T* __restrict__ inv, int* info, int64_t j0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
#ifdef DCGP_LEAF_TIMING
    long long t0 = clock64();
#endif
    __shared__ T rinv[NB];
    load_lower<T>(S, rinv, A, lda, nb, false);
    __syncthreads();
#ifdef DCGP_LEAF_TIMING
    long long t1 = clock64();
#endif
    block_potrf<T, MODE>(S, rinv, nb, info, j0);
#ifdef DCGP_LEAF_TIMING
    long long t2 = clock64();
#endif
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int i = warp; i < nb; i += NTHREADS / 32)
            for (int k = lane; k <= i; k += 32) A[(int64_t)i * lda + k] = S[i * SP + k];
    }
    __syncthreads();
#ifdef DCGP_LEAF_TIMING
    long long t3 = clock64();
#endif
    if (!(MODE & 1)) block_trtri<T>(S, rinv, nb);   // FINV: the inverse already sits in the overlay
#ifdef DCGP_LEAF_TIMING
    long long t4 = clock64();
#endif
    store_inverse<T>(S, nb, inv, NB);
#ifdef DCGP_LEAF_TIMING
    __syncthreads();
    if (threadIdx.x == 0)
        printf("leaf nb=%d load %lld potrf %lld store %lld trtri %lld storeinv %lld\n", nb, t1 - t0, t2 - t1, t3 - t2,
               t4 - t3, clock64() - t4);
#endif
}

// Look-ahead step kernel (FP32): the whole serial link between two diagonal blocks in one launch.
//   R  = A[j+1, j] (a private copy made before the bulk lane solves that panel in place), X = inv(L_jj):
//   P  = R X^T                                  (block row j+1 of L, kept on chip only)
//   S  = A[j+1, j+1] - P P^T                    (all earlier panels were applied by the bulk lane)
//   then as potrf_diag_kernel: S = L L^T, L -> A, inv(L) -> workspace.
// Both 128^3 products run on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM) as 3xTF32: the operands are written to shared memory in the canonical K-major 128-byte-swizzle
// layout by the loading threads themselves (no TMA: the data needs the hi/lo split anyway), three
// 64 KB operand buffers rotate through {R, X, R_lo}, {R, X, X_lo}, {P, P_lo, S}.
constexpr int STEP_BUF = 66 * 1024;  // >= 128 x 129 floats (the pitch-129 tile of the factorisation) and 1024-aligned
constexpr size_t STEP_SMEM = 3 * STEP_BUF + 1024;

// byte offset of element (row r, k) of a 128 x 128 FP32 K-major operand: four 16 KB k-blocks of 128-byte
// rows whose 16-byte chunks are XOR-swizzled with the row index (what TMA SWIZZLE_128B would produce)
__device__ __forceinline__ uint32_t canon_off(int r, int k) {
    return (uint32_t)((k >> 5) * 16384 + r * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2));
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// 16 MMAs: D[128 x 128] (+)= A[128 x 128] B[128 x 128]^T over K = 128 (both K-major canonical tiles)
__device__ __forceinline__ void issue_128(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, bool fresh) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
            umma_tf32(tmem_d, make_desc(a_addr + kb * 16384 + kk * 32, 16, 1024, 2),
                      make_desc(b_addr + kb * 16384 + kk * 32, 16, 1024, 2), idesc, (fresh && kb == 0 && kk == 0) ? 0u : 1u);
}

#define TMEM_LD32(r, taddr)                                                                                          \
    asm volatile(                                                                                                    \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                    \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                    \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                    \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),     \
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),    \
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                  \
        : "r"(taddr));                                                                                               \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

template <int MODE>
__global__ void __launch_bounds__(NTHREADS) potrf_step_kernel(float* __restrict__ A, int64_t lda, int nb,
                                                              float* __restrict__ inv, int* info, int64_t j0,
                                                              const float* __restrict__ R, const float* __restrict__ invp) {
    extern __shared__ unsigned char smem_dyn[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char* B0 = base;
    unsigned char* B1 = base + STEP_BUF;
    unsigned char* B2 = base + 2 * STEP_BUF;
    __shared__ float rinv[NB];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef DCGP_LEAF_TIMING
    long long tq0 = clock64();
#endif
    if (tid == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // R -> B0, R_lo -> B2, X -> B1 (canonical operand layout); 16-byte vectors, 8 lanes per 128-byte row segment
    for (int e = tid; e < NB * NB / 4; e += NTHREADS) {
        const int i = e >> 5, k = (e & 31) * 4;
        const float4 r = (i < nb) ? *reinterpret_cast<const float4*>(R + i * NB + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 x = *reinterpret_cast<const float4*>(invp + i * NB + k);
        const uint32_t off = canon_off(i, k);
        *reinterpret_cast<float4*>(B0 + off) = r;
        *reinterpret_cast<float4*>(B2 + off) = make_float4(tf32_lo(r.x), tf32_lo(r.y), tf32_lo(r.z), tf32_lo(r.w));
        *reinterpret_cast<float4*>(B1 + off) = x;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t a0 = smem_u32(B0), a1 = smem_u32(B1), a2 = smem_u32(B2);
#ifdef DCGP_LEAF_TIMING
    long long tq1 = clock64();
#endif
    // ---- P = R X^T: R_hi X_hi + R_lo X_hi, then (after X_lo replaces R_lo) R_hi X_lo
    if (tid == 0) {
        issue_128(tmem, a0, a1, true);
        issue_128(tmem, a2, a1, false);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int e = tid; e < NB * NB / 4; e += NTHREADS) {
        const float4 x = *reinterpret_cast<const float4*>(B1 + e * 16);  // elementwise: same offsets in both buffers
        *reinterpret_cast<float4*>(B2 + e * 16) = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        issue_128(tmem, a0, a2, false);
        umma_commit(&bar);
    }
    mbar_wait(&bar, 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- P (TMEM columns 0..127) -> operand tiles: P in B0, P_lo in B2.  Warp w reads lane quadrant w % 4,
    // columns 32 (w / 4) ..; a thread holds one row, 32 consecutive k = one k-block of that row.
    {
        const int row = (warp & 3) * 32 + lane, kb = warp >> 2;
        uint32_t r[32];
        TMEM_LD32(r, tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kb * 32));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 v = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]), __uint_as_float(r[4 * c + 2]),
                                         __uint_as_float(r[4 * c + 3]));
            const uint32_t off = (uint32_t)(kb * 16384 + row * 128 + ((c ^ (row & 7)) << 4));
            *reinterpret_cast<float4*>(B0 + off) = v;
            *reinterpret_cast<float4*>(B2 + off) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef DCGP_LEAF_TIMING
    long long tq2 = clock64();
#endif
    // ---- P P^T into TMEM columns 128..255 while every thread fetches the diagonal tile into B1 (X is dead)
    if (tid == 0) {
        issue_128(tmem + 128, a0, a0, true);
        issue_128(tmem + 128, a2, a0, false);
        issue_128(tmem + 128, a0, a2, false);
        umma_commit(&bar);
    }
    float* S = reinterpret_cast<float*>(B1);
    load_lower<float>(S, rinv, A, lda, nb, false);
    __syncthreads();
    mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
        const int row = (warp & 3) * 32 + lane, c0 = (warp >> 2) * 32;
        uint32_t r[32];
        TMEM_LD32(r, tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(128 + c0));
        if (row < nb) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (c0 + j <= row) S[row * SP + c0 + j] -= __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256));
    }
#ifdef DCGP_LEAF_TIMING
    long long tq3 = clock64();
#endif
    block_potrf<float, MODE>(S, rinv, nb, info, j0);
#ifdef DCGP_LEAF_TIMING
    long long tq4 = clock64();
#endif
    for (int i = warp; i < nb; i += NTHREADS / 32)
        for (int k = lane; k <= i; k += 32) A[(int64_t)i * lda + k] = S[i * SP + k];
    __syncthreads();
#ifdef DCGP_LEAF_TIMING
    long long tq5 = clock64();
#endif
    if (!(MODE & 1)) block_trtri<float>(S, rinv, nb);
#ifdef DCGP_LEAF_TIMING
    long long tq6 = clock64();
#endif
    store_inverse<float>(S, nb, inv, NB);
#ifdef DCGP_LEAF_TIMING
    __syncthreads();
    if (tid == 0 && j0 == 256)
        printf("step kernel: load %lld  P %lld  syrk %lld  potrf %lld  rest %lld\n", tq1 - tq0, tq2 - tq1, tq3 - tq2, tq4 - tq3,
               clock64() - tq4);
    if (tid == 0) printf("  tail: storeL %lld  trtri %lld  storeinv %lld\n", tq5 - tq4, tq6 - tq5, clock64() - tq6);
#endif
}

// batched inversion of the diagonal blocks of an existing factor
template <typename T>
// `group` 128-blocks form one (group*128)-wide output block with pitch group*128 (group = 1: the packed
// NB x NB workspace blocks of the factorisation; group = 4: the diagonal of the 512-wide inverse blocks)
__global__ void __launch_bounds__(NTHREADS) trtri_blocks_kernel(const T* __restrict__ L, int64_t ldl, int64_t n,
                                                                T* __restrict__ inv, int group, int64_t ldo = 0) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);
    const int64_t i0 = (int64_t)blockIdx.x * NB;
    const int nb = (int)min((int64_t)NB, n - i0);
    __shared__ T rinv[NB];
    load_lower<T>(S, rinv, L + i0 * ldl + i0, ldl, nb, true);
    __syncthreads();
    block_trtri<T>(S, rinv, nb);
    if (ldo > 0) {  // straight onto the diagonal of an n x n matrix with pitch ldo
        store_inverse<T>(S, nb, inv + i0 * ldo + i0, ldo, nb);
        return;
    }
    const int64_t big = blockIdx.x / group, q = blockIdx.x % group, ld = (int64_t)group * NB;
    store_inverse<T>(S, nb, inv + big * ld * ld + q * NB * ld + q * NB, ld);
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
constexpr size_t leaf_smem() { return (size_t)NB * SP * sizeof(T); }

// Leaf variant (MODE of block_potrf): DCGP_LEAF_FINV=0 restores the separate block_trtri pass after the factorisation
// (round-1 default); with 1 (default since round 2: FP32 8192^2 factorisation 4.67 -> 4.22 ms, ELBO step 12.14 -> 11.53 ms,
// profiles/r02_finv_ab.txt) the leaves build the inverse of their factor inside the pivot loop (inv_rows_*).
inline int leaf_mode() {
    static const int mode = (getenv("DCGP_LEAF_FINV") && atoi(getenv("DCGP_LEAF_FINV")) == 0) ? 0 : 1;
    return mode;
}
#define DCGP_LEAF_DISPATCH(CALL) \
    switch (leaf_mode()) {       \
        case 0: CALL(0); break;  \
        default: CALL(1); break; \
    }

template <typename T>
int set_smem_attr(const void* fn) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)leaf_smem<T>());
    return e == cudaSuccess ? 0 : DCGP_CUDA_ERR(e);
}

inline int64_t split_point(int64_t nb) {  // first half, a multiple of the leaf size
    int64_t h = ((nb / 2 + NB - 1) / NB) * NB;
    return h >= nb ? nb - NB : h;
}

template <typename T>
struct Ctx {
    T* A;
    int64_t lda;
    T* ws;
    int* info;
    int dtype, xf;
    cudaStream_t st;
    T* Alo;  // FP32 3xTF32 only: x - trunc_tf32(x) of the finished off-diagonal blocks of L (layout of A), or null
    int64_t la_pw = 0;  // look-ahead panel width: 0 = choose by policy, -1 = plain recursion
    float* rcopy = nullptr;  // FP32: 2 x 128 x 128 scratch of the fused look-ahead path
};

// B = A[r0 : r0+nr, t0 : t0+nt]  <-  B * L[t0 : t0+nt, t0 : t0+nt]^{-T}
template <typename T>
int rtrsm_rec(const Ctx<T>& c, int64_t r0, int64_t nr, int64_t t0, int64_t nt) {
    if (nt <= NB) {
        T* Bp = c.A + r0 * c.lda + t0;
        int rc = dcgp_gemm_impl(0, 1, nr, nt, nt, 1.0, Bp, c.lda, c.ws + (t0 / NB) * NB * NB, NB, 0.0, Bp, c.lda,
                                c.dtype, DCGP_GEMM_INPLACE | c.xf, c.st, nullptr, nullptr,
                                // these entries of L are final: the epilogue also publishes their TF32
                                // remainders for the tcgen05 3xTF32 products
                                c.Alo ? c.Alo + r0 * c.lda + t0 : nullptr);
        return rc;
    }
    const int64_t t1 = split_point(nt), t2 = nt - t1;
    int rc = rtrsm_rec(c, r0, nr, t0, t1);
    if (rc) return rc;
    // B2 -= B1 * Lc^T,  Lc = L[t0+t1 : t0+nt, t0 : t0+t1]
    const int64_t oa = r0 * c.lda + t0, ob = (t0 + t1) * c.lda + t0;
    rc = dcgp_gemm_impl(0, 1, nr, t2, t1, -1.0, c.A + oa, c.lda, c.A + ob, c.lda, 1.0, c.A + r0 * c.lda + t0 + t1, c.lda,
                        c.dtype, c.xf, c.st, c.Alo ? c.Alo + oa : nullptr,
                        c.Alo ? c.Alo + ob : nullptr, nullptr);
    if (rc) return rc;
    return rtrsm_rec(c, r0, nr, t0 + t1, t2);
}

template <typename T>
struct Ctx;
int chol_la_fused(const Ctx<float>& c, int64_t j0, int64_t nb, float* rcopy);
template <typename T>
int chol_la(const Ctx<T>& c, int64_t j0, int64_t nb);
inline int64_t la_panel(int dtype, int64_t nb);

template <typename T>
int chol_rec(const Ctx<T>& c, int64_t j0, int64_t nb) {
    if (nb <= NB) {
#define CALL(M)                                                                                              \
    potrf_diag_kernel<T, M><<<1, NTHREADS, leaf_smem<T>(), c.st>>>(c.A + j0 * c.lda + j0, c.lda, (int)nb,      \
                                                                   c.ws + (j0 / NB) * NB * NB, c.info, j0)
        DCGP_LEAF_DISPATCH(CALL);
#undef CALL
        DCGP_CHECK_LAUNCH();
        return 0;
    }
    // (Halving FP32 blocks above 4096 by the recursion first - two look-ahead factorisations with L2-resident trailing
    // matrices around one deep-K solve and update - was measured on B200 with the TMEM link kernel: 8192^2 3.80 -> 4.34 ms,
    // reconstruction error 1.8e-6 -> 2.2e-5; not kept.)
    if (c.la_pw >= 0) {
        const int64_t pw = c.la_pw ? c.la_pw : la_panel(c.dtype, nb);
        if (pw && nb > 2 * pw) {
            Ctx<T> cl = c;
            cl.la_pw = pw;
            if constexpr (sizeof(T) == 4) {
                if (pw == NB && c.Alo && c.rcopy && (c.lda & 3) == 0 && (reinterpret_cast<uintptr_t>(c.A) & 15) == 0)
                    return chol_la_fused(cl, j0, nb, c.rcopy);
            }
            return chol_la(cl, j0, nb);
        }
    }
    const int64_t n1 = split_point(nb), n2 = nb - n1;
    int rc = chol_rec(c, j0, n1);
    if (rc) return rc;
    rc = rtrsm_rec(c, j0 + n1, n2, j0, n1);  // A21 <- A21 L11^{-T}
    if (rc) return rc;
    const int64_t o21 = (j0 + n1) * c.lda + j0;
    T* A21 = c.A + o21;
    const T* A21lo = c.Alo ? c.Alo + o21 : nullptr;
    rc = dcgp_gemm_impl(0, 1, n2, n2, n1, -1.0, A21, c.lda, A21, c.lda, 1.0, c.A + (j0 + n1) * c.lda + j0 + n1, c.lda,
                        c.dtype, DCGP_GEMM_LOWER | c.xf, c.st, A21lo, A21lo, nullptr);
    if (rc) return rc;
    return chol_rec(c, j0 + n1, n2);
}

// ---------------------------------------------------------------------------------------------
// Look-ahead right-looking factorisation of one diagonal block (the base case of chol_rec).
//
// The recursion above runs every kernel on one stream, so the 128-wide leaves (one CTA each) sit
// in series with all the GEMMs between them.  Here the leaf chain gets its own lane:
//   stream A (high priority): leaf(j) -> row block j+1 <- row * inv_j^T -> tile(j+1,j+1) -= P P^T -> leaf(j+1)
//   stream B (low priority) : rows below j+1 <- rows * inv_j^T; column j+1 and the trailing square -= P P^T
// so the rank-128 trailing update of panel j overlaps leaf j+1.  Both lanes are forked from and joined back
// into the caller's stream with events only, which keeps the call capturable in a CUDA graph.
struct LookAhead {
    cudaStream_t sh = nullptr, sb = nullptr, sc = nullptr;  // chain (highest priority), bulk (lowest) and helper lanes
    cudaEvent_t leaf[4] = {}, row[4] = {}, upd[4] = {}, tr[4] = {}, hc[4] = {}, fork = nullptr, join = nullptr;
    int dev = -1;
    int init() {
        int d = 0;
        cudaGetDevice(&d);
        if (d == dev && sb) return 0;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaError_t e = cudaStreamCreateWithPriority(&sh, cudaStreamNonBlocking, hi);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        e = cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        e = cudaStreamCreateWithPriority(&sc, cudaStreamNonBlocking, hi + 1 <= lo ? hi + 1 : lo);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        for (int i = 0; i < 4; ++i) {
            cudaEventCreateWithFlags(&leaf[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&row[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&upd[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&tr[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&hc[i], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
        e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        dev = d;
        return 0;
    }
};
// One lane set per caller stream: factorisations issued from different streams (e.g. K_ZZ on a side stream while
// Lambda factors on the main one) must not queue behind each other on shared internal streams.
// (Unbounded: a lane set is two streams and fourteen events; sharing one between two caller streams - what a bounded map
// did in round 1 for a ninth stream - would let two concurrent factorisations wait on each other's events.)
thread_local std::unordered_map<cudaStream_t, LookAhead> g_lanes;
inline LookAhead& lanes_for(cudaStream_t st) { return g_lanes[st]; }

#ifdef DCGP_TRACE
struct Trace {  // debug builds only: device timeline of the two lanes
    struct Rec { const char* what; int64_t step; cudaEvent_t ev; };
    std::vector<Rec> recs;
    cudaEvent_t base;
    void begin(cudaStream_t st) { recs.clear(); cudaEventCreate(&base); cudaEventRecord(base, st); }
    void mark(const char* what, int64_t step, cudaStream_t st) {
        cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); recs.push_back({what, step, e});
    }
    void dump() {
        cudaDeviceSynchronize();
        for (auto& r : recs) { float ms = 0; cudaEventElapsedTime(&ms, base, r.ev); printf("TRACE %s %lld %.1f\n", r.what, (long long)r.step, ms * 1e3f); cudaEventDestroy(r.ev); }
        cudaEventDestroy(base);
    }
};
thread_local Trace g_trace;
#define TRACE_BEGIN(st) g_trace.begin(st)
#define TRACE_MARK(w, s, st) g_trace.mark(w, s, st)
#define TRACE_DUMP() g_trace.dump()
#else
#define TRACE_BEGIN(st)
#define TRACE_MARK(w, s, st)
#define TRACE_DUMP()
#endif

// Panel width of the look-ahead factorisation of an nb-wide diagonal block (0 = plain recursion).
// Narrow panels shorten the serial chain but make the trailing updates shallow (K = panel width).
inline int64_t la_panel(int dtype, int64_t nb) {
    static int64_t force[2] = {-2, -2};  // DCGP_LA_PW_F64 / DCGP_LA_PW_F32: tuning override (0 disables)
    if (force[0] == -2) {
        const char* e;
        force[1] = (e = getenv("DCGP_LA_PW_F32")) ? atoll(e) : -1;
        force[0] = (e = getenv("DCGP_LA_PW_F64")) ? atoll(e) : -1;
    }
    const int i = dtype == DCGP_F64 ? 0 : 1;
    int64_t pw;
    if (force[i] >= 0) pw = force[i];
    else if (i == 1) pw = nb <= 8192 ? 128 : (nb <= 16384 ? 256 : 512);      // measured on B200 (tools/potrf_sweep.py)
    else pw = nb <= 5120 ? 128 : (nb <= 7168 ? 256 : (nb <= 12288 ? 512 : 1024));
    return (pw > 0 && nb > 2 * pw) ? pw : 0;
}

#define DCGP_CU(x)                                           \
    do {                                                     \
        cudaError_t e_ = (x);                                \
        if (e_ != cudaSuccess) return DCGP_CUDA_ERR(e_);     \
    } while (0)

// FP32 / 3xTF32 flavour with 128-wide panels: the chain lane is ONE kernel per block (potrf_step_kernel);
// the bulk lane solves the whole panel in place (block row j+1 included) and applies it to everything
// right of it in a single tcgen05 launch (lower tiles only, offset diagonal), then hands the chain lane a
// private copy of the next block row so that the two lanes never touch the same memory.
int chol_la_fused(const Ctx<float>& c, int64_t j0, int64_t nb, float* rcopy) {
    LookAhead& x = lanes_for(c.st);
    int rc = x.init();
    if (rc) return rc;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaSuccess;
#define CALL(M) e = cudaFuncSetAttribute(potrf_step_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM)
        DCGP_LEAF_DISPATCH(CALL);
#undef CALL
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        attr = true;
    }
    const int64_t q = (nb + NB - 1) / NB, end = j0 + nb;
    const cudaStream_t sa = x.sh, sb = x.sb;
    auto copy_row = [&](int64_t brow, int64_t bcol, int64_t rows, int slot, cudaStream_t st) {
        return cudaMemcpy2DAsync(rcopy + (size_t)slot * NB * NB, NB * sizeof(float), c.A + brow * c.lda + bcol,
                                 (size_t)c.lda * sizeof(float), NB * sizeof(float), (size_t)rows, cudaMemcpyDeviceToDevice, st);
    };
    {
        const int64_t b1 = j0 + NB, w1 = (end - b1 < NB) ? end - b1 : NB;
        DCGP_CU(copy_row(b1, j0, w1, 0, c.st));
        rc = dcgp_split_lo_impl(c.A + b1 * c.lda + j0, c.Alo + b1 * c.lda + j0, end - b1, NB, c.lda, c.lda, c.st);
        if (rc) return rc;
    }
    DCGP_CU(cudaEventRecord(x.fork, c.st));
    DCGP_CU(cudaStreamWaitEvent(sa, x.fork, 0));
    DCGP_CU(cudaStreamWaitEvent(sb, x.fork, 0));
    TRACE_BEGIN(c.st);
    // DCGP_LINK_TM (default 1): the links run on the TMEM-resident kernel of chol_tm.cu (tcgen05 rank-16 updates inside the
    // leaf; B200: FP32 8192^2 factorisation 4.18 -> 3.81 ms, 1024^2 0.46 -> 0.36 ms, profiles/r02_link_tm.txt); 0 = the
    // shared-memory potrf_step_kernel of round 1
    static const int link_tm = getenv("DCGP_LINK_TM") ? atoi(getenv("DCGP_LINK_TM")) : 1;
    // the TMEM link kernel also leaves the TF32 remainders of inv(L_jj) for the bulk lane's 3xTF32 panel solve: two slots
    // alternate (link j + 2 starts only after the bulk lane has finished with panel j)
    auto invlo_slot = [&](int64_t blk) { return rcopy + (size_t)(2 + (blk & 1)) * NB * NB; };
    if (link_tm) {
        rc = dcgp_potrf_link_tm(c.A + j0 * c.lda + j0, c.lda, NB, c.ws + (j0 / NB) * NB * NB, invlo_slot(0), c.info, j0, nullptr, nullptr, sa);
        if (rc) return rc;
    } else {
#define CALL(M)                                                                                                 \
    potrf_diag_kernel<float, M><<<1, NTHREADS, leaf_smem<float>(), sa>>>(c.A + j0 * c.lda + j0, c.lda, NB,        \
                                                                         c.ws + (j0 / NB) * NB * NB, c.info, j0)
        DCGP_LEAF_DISPATCH(CALL);
#undef CALL
        DCGP_CHECK_LAUNCH();
    }
    // DCGP_POTRF_H = h > 0: a helper lane takes the panel solve of panel j and the update of the NEXT panel's column
    // ("strip": rows below x 128 columns) on h SMs, the bulk lane keeps only the trailing square on the other SMs.  The
    // first half of an 8192^2 factorisation is bound by the bulk lane (panel solve 18 us + trailing update up to 86 us per
    // link against 40 us for the link itself); with the helper lane the panel solve and the strip run beside the
    // trailing update of the previous panel instead of in front of the current one.  Measured on B200 (h = 14 .. 18): 8192^2
    // 3.51 -> 3.31 ms stand-alone, but the trailing square on 132 SMs takes as long as the whole update took on 146 (it is
    // bound by the bytes crossing L2, not by SMs), and inside the ELBO step - where the K_ZZ lane runs beside the
    // factorisation - alternating A/B runs show no gain (101.2 vs 100.0 .. 101.6 steps/s): opt-in.
    static const int hsm = getenv("DCGP_POTRF_H") ? atoi(getenv("DCGP_POTRF_H")) : 0;
    if (hsm > 0 && link_tm && nb >= 48 * NB) {   // (4096: 1.25 -> 1.28 ms, 8192: 3.51 -> 3.32 ms)
        const cudaStream_t sc = x.sc;
        DCGP_CU(cudaStreamWaitEvent(sc, x.fork, 0));
        const int bulk_cap = num_sms() - 2 - hsm;
        for (int64_t j = 0; j + 1 < q; ++j) {
            const int64_t bj = j0 + j * NB;
            const int64_t b1 = bj + NB, w1 = (end - b1 < NB) ? end - b1 : NB, r0 = b1 + w1, nr = end - r0;
            const int e = (int)(j & 3), ep = (int)((j + 3) & 3);
            float* invj = c.ws + (bj / NB) * NB * NB;
            TRACE_MARK("A.leaf", j, sa);
            DCGP_CU(cudaEventRecord(x.leaf[e], sa));
            // ---- helper lane: panel solve of panel j (its column was completed by strip(j-1), issued on this lane)
            DCGP_CU(cudaStreamWaitEvent(sc, x.leaf[e], 0));
            float* P = c.A + b1 * c.lda + bj;
            float* Plo = c.Alo + b1 * c.lda + bj;
            rc = 1;
            dcgp_tc_cap(hsm);
            if (end - b1 >= NB) {
                rc = dcgp_gemm_tc_ex(0, 1, end - b1, NB, NB, 1.0, P, c.lda, invj, NB, 0.0, P, c.lda, 0, Plo, invlo_slot(j), Plo, 0, 0, sc);
                if (rc < 0) { dcgp_tc_cap(0); return rc; }
            }
            dcgp_tc_cap(0);
            if (rc == 1) {
                rc = dcgp_gemm_impl(0, 1, end - b1, NB, NB, 1.0, P, c.lda, invj, NB, 0.0, P, c.lda, c.dtype, DCGP_GEMM_INPLACE | c.xf,
                                    sc, nullptr, nullptr, Plo);
                if (rc) return rc;
            }
            TRACE_MARK("B.trsm", j, sc);
            DCGP_CU(cudaEventRecord(x.tr[e], sc));
            if (nr > 0) {
                float* P2 = c.A + r0 * c.lda + bj;
                const float* P2lo = c.Alo + r0 * c.lda + bj;
                // ---- bulk lane: trailing square, rows and columns from r0
                DCGP_CU(cudaStreamWaitEvent(sb, x.tr[e], 0));
                rc = 1;
                if (nr >= NB) {
                    dcgp_tc_cap(bulk_cap);
                    rc = dcgp_gemm_tc_ex(0, 1, nr, nr, NB, -1.0, P2, c.lda, P2, c.lda, 1.0, c.A + r0 * c.lda + r0, c.lda, DCGP_GEMM_LOWER,
                                         P2lo, P2lo, nullptr, 0, 0, sb);
                    dcgp_tc_cap(0);
                    if (rc < 0) return rc;
                }
                if (rc == 1) {
                    rc = dcgp_gemm_impl(0, 1, nr, nr, NB, -1.0, P2, c.lda, P2, c.lda, 1.0, c.A + r0 * c.lda + r0, c.lda, c.dtype,
                                        DCGP_GEMM_LOWER | c.xf, sb, P2lo, P2lo, nullptr);
                    if (rc) return rc;
                }
                TRACE_MARK("B.trail", j, sb);
                DCGP_CU(cudaEventRecord(x.upd[e], sb));
                // ---- helper lane: strip = column of panel j+1 below its diagonal block (last written by the trailing
                // square of panel j-1), then the private copy of block row j+2 for the chain lane
                if (j > 0) DCGP_CU(cudaStreamWaitEvent(sc, x.upd[ep], 0));
                rc = 1;
                if (nr >= NB) {
                    dcgp_tc_cap(hsm);
                    rc = dcgp_gemm_tc_ex(0, 1, nr, w1, NB, -1.0, P2, c.lda, P, c.lda, 1.0, c.A + r0 * c.lda + b1, c.lda, 0, P2lo, Plo,
                                         c.Alo + r0 * c.lda + b1, 0, 0, sc);
                    dcgp_tc_cap(0);
                    if (rc < 0) return rc;
                }
                if (rc == 1) {
                    rc = dcgp_gemm_impl(0, 1, nr, w1, NB, -1.0, P2, c.lda, P, c.lda, 1.0, c.A + r0 * c.lda + b1, c.lda, c.dtype, c.xf, sc,
                                        P2lo, Plo, c.Alo + r0 * c.lda + b1);
                    if (rc) return rc;
                }
                const int64_t w2 = (end - r0 < NB) ? end - r0 : NB;
                DCGP_CU(copy_row(r0, b1, w2, (int)((j + 1) & 1), sc));
                TRACE_MARK("B.col", j, sc);
                DCGP_CU(cudaEventRecord(x.hc[e], sc));
            }
            // ---- chain lane: block j+1 (diagonal block from the trailing square of panel j-1, block row from the helper)
            if (j > 0) {
                DCGP_CU(cudaStreamWaitEvent(sa, x.upd[ep], 0));
                DCGP_CU(cudaStreamWaitEvent(sa, x.hc[ep], 0));
            }
            TRACE_MARK("A.waited", j, sa);
            rc = dcgp_potrf_link_tm(c.A + b1 * c.lda + b1, c.lda, (int)w1, c.ws + (b1 / NB) * NB * NB, invlo_slot(j + 1), c.info, b1,
                                    rcopy + (size_t)(j & 1) * NB * NB, invj, sa);
            if (rc) return rc;
        }
        DCGP_CU(cudaEventRecord(x.join, sa));
        DCGP_CU(cudaStreamWaitEvent(c.st, x.join, 0));
        DCGP_CU(cudaEventRecord(x.fork, sb));
        DCGP_CU(cudaStreamWaitEvent(c.st, x.fork, 0));
        DCGP_CU(cudaEventRecord(x.row[0], sc));
        DCGP_CU(cudaStreamWaitEvent(c.st, x.row[0], 0));
        TRACE_DUMP();
        return 0;
    }
    for (int64_t j = 0; j + 1 < q; ++j) {
        const int64_t bj = j0 + j * NB;
        const int64_t b1 = bj + NB, w1 = (end - b1 < NB) ? end - b1 : NB, r0 = b1 + w1, nr = end - r0;
        const int e = (int)(j & 3), ep = (int)((j + 3) & 3);
        float* invj = c.ws + (bj / NB) * NB * NB;
        TRACE_MARK("A.leaf", j, sa);
        DCGP_CU(cudaEventRecord(x.leaf[e], sa));
        // ---- bulk lane: panel j
        DCGP_CU(cudaStreamWaitEvent(sb, x.leaf[e], 0));
        TRACE_MARK("B.start", j, sb);
        float* P = c.A + b1 * c.lda + bj;          // rows b1 .. end of panel j
        float* Plo = c.Alo + b1 * c.lda + bj;
        // in place on the tcgen05 kernel: a tile owns all 128 columns of its rows, which it has read in full
        // before its epilogue overwrites them (the remainders of the panel came from the previous update)
        rc = 1;
        if (end - b1 >= NB) {
            float* invlo = invlo_slot(j);
            if (!link_tm) {
                rc = dcgp_split_lo_impl(invj, invlo, NB, NB, NB, NB, sb);
                if (rc) return rc;
            }
            rc = dcgp_gemm_tc_ex(0, 1, end - b1, NB, NB, 1.0, P, c.lda, invj, NB, 0.0, P, c.lda, 0, Plo, invlo, Plo, 0, 0, sb);
            if (rc < 0) return rc;
        }
        if (rc == 1) {
            rc = dcgp_gemm_impl(0, 1, end - b1, NB, NB, 1.0, P, c.lda, invj, NB, 0.0, P, c.lda, c.dtype, DCGP_GEMM_INPLACE | c.xf,
                                sb, nullptr, nullptr, Plo);
            if (rc) return rc;
        }
        TRACE_MARK("B.trsm", j, sb);
        if (nr > 0) {
            float* P2 = c.A + r0 * c.lda + bj;
            const float* P2lo = c.Alo + r0 * c.lda + bj;
            rc = 1;
            if (nr >= NB)
                rc = dcgp_gemm_tc_ex(0, 1, nr, w1 + nr, NB, -1.0, P2, c.lda, P, c.lda, 1.0, c.A + r0 * c.lda + b1, c.lda,
                                     DCGP_GEMM_LOWER, P2lo, Plo, c.Alo + r0 * c.lda + b1, w1, w1, sb);  // Clo: next panel only
            if (rc < 0) return rc;
            if (rc == 1) {  // shape / alignment not taken by the tcgen05 kernel
                rc = dcgp_gemm_impl(0, 1, nr, w1, NB, -1.0, P2, c.lda, P, c.lda, 1.0, c.A + r0 * c.lda + b1, c.lda, c.dtype, c.xf,
                                    sb, P2lo, Plo, c.Alo + r0 * c.lda + b1);
                if (rc) return rc;
                rc = dcgp_gemm_impl(0, 1, nr, nr, NB, -1.0, P2, c.lda, P2, c.lda, 1.0, c.A + r0 * c.lda + r0, c.lda, c.dtype,
                                    DCGP_GEMM_LOWER | c.xf, sb, P2lo, P2lo, nullptr);
                if (rc) return rc;
            }
            TRACE_MARK("B.trail", j, sb);
            const int64_t w2 = (end - r0 < NB) ? end - r0 : NB;
            DCGP_CU(copy_row(r0, b1, w2, (int)((j + 1) & 1), sb));
            DCGP_CU(cudaEventRecord(x.upd[e], sb));
        }
        // ---- chain lane: block j+1 (its inputs were completed by the bulk update of panel j-1)
        if (j > 0) DCGP_CU(cudaStreamWaitEvent(sa, x.upd[ep], 0));
        TRACE_MARK("A.waited", j, sa);
        if (link_tm) {
            rc = dcgp_potrf_link_tm(c.A + b1 * c.lda + b1, c.lda, (int)w1, c.ws + (b1 / NB) * NB * NB, invlo_slot(j + 1), c.info, b1,
                                    rcopy + (size_t)(j & 1) * NB * NB, invj, sa);
            if (rc) return rc;
        } else {
#define CALL(M)                                                                                                  \
    potrf_step_kernel<M><<<1, NTHREADS, STEP_SMEM, sa>>>(c.A + b1 * c.lda + b1, c.lda, (int)w1,                    \
                                                         c.ws + (b1 / NB) * NB * NB, c.info, b1,                   \
                                                         rcopy + (size_t)(j & 1) * NB * NB, invj)
            DCGP_LEAF_DISPATCH(CALL);
#undef CALL
            DCGP_CHECK_LAUNCH();
        }
    }
    // join both lanes: the last panel solve on the bulk lane has no later consumer on the chain lane
    DCGP_CU(cudaEventRecord(x.join, sa));
    DCGP_CU(cudaStreamWaitEvent(c.st, x.join, 0));
    DCGP_CU(cudaEventRecord(x.fork, sb));
    DCGP_CU(cudaStreamWaitEvent(c.st, x.fork, 0));
    TRACE_DUMP();
    return 0;
}

template <typename T>
int chol_la(const Ctx<T>& c, int64_t j0, int64_t nb) {
    LookAhead& x = lanes_for(c.st);
    int rc = x.init();
    if (rc) return rc;
    const int64_t pw = c.la_pw, q = (nb + pw - 1) / pw, end = j0 + nb;
    const cudaStream_t sa = x.sh, sb = x.sb;
    Ctx<T> ca = c, cb = c;
    ca.la_pw = cb.la_pw = -1;  // panels themselves: plain recursion on their own stream
    ca.st = sa;
    cb.st = sb;
    DCGP_CU(cudaEventRecord(x.fork, c.st));
    DCGP_CU(cudaStreamWaitEvent(sa, x.fork, 0));
    DCGP_CU(cudaStreamWaitEvent(sb, x.fork, 0));
    TRACE_BEGIN(c.st);
    for (int64_t j = 0; j < q; ++j) {
        const int64_t bj = j0 + j * pw, wj = (end - bj < pw) ? end - bj : pw;
        rc = chol_rec(ca, bj, wj);
        if (rc) return rc;
        TRACE_MARK("A.leaf", j, sa);
        if (j == q - 1) break;
        const int e = (int)(j & 3), ep = (int)((j + 3) & 3);
        DCGP_CU(cudaEventRecord(x.leaf[e], sa));
        const int64_t b1 = bj + pw, w1 = (end - b1 < pw) ? end - b1 : pw, r0 = b1 + w1, nr = end - r0;
        // stream A: the next block row and its diagonal block (needs the trailing update of panel j-1)
        if (j > 0) DCGP_CU(cudaStreamWaitEvent(sa, x.upd[ep], 0));
        TRACE_MARK("A.waited", j, sa);
        rc = rtrsm_rec(ca, b1, w1, bj, pw);
        if (rc) return rc;
        TRACE_MARK("A.row", j, sa);
        DCGP_CU(cudaEventRecord(x.row[e], sa));
        T* P1 = c.A + b1 * c.lda + bj;
        const T* P1lo = c.Alo ? c.Alo + b1 * c.lda + bj : nullptr;
        rc = dcgp_gemm_impl(0, 1, w1, w1, pw, -1.0, P1, c.lda, P1, c.lda, 1.0, c.A + b1 * c.lda + b1, c.lda, c.dtype,
                            DCGP_GEMM_LOWER | c.xf, sa, P1lo, P1lo, nullptr);
        if (rc) return rc;
        TRACE_MARK("A.syrk", j, sa);
        if (nr <= 0) continue;
        // stream B: everything below
        DCGP_CU(cudaStreamWaitEvent(sb, x.leaf[e], 0));
        TRACE_MARK("B.start", j, sb);
        rc = rtrsm_rec(cb, r0, nr, bj, pw);
        if (rc) return rc;
        TRACE_MARK("B.trsm", j, sb);
        T* P2 = c.A + r0 * c.lda + bj;
        const T* P2lo = c.Alo ? c.Alo + r0 * c.lda + bj : nullptr;
        DCGP_CU(cudaStreamWaitEvent(sb, x.row[e], 0));
        rc = dcgp_gemm_impl(0, 1, nr, w1, pw, -1.0, P2, c.lda, P1, c.lda, 1.0, c.A + r0 * c.lda + b1, c.lda, c.dtype, c.xf,
                            sb, P2lo, P1lo, nullptr);
        if (rc) return rc;
        TRACE_MARK("B.col", j, sb);
        rc = dcgp_gemm_impl(0, 1, nr, nr, pw, -1.0, P2, c.lda, P2, c.lda, 1.0, c.A + r0 * c.lda + r0, c.lda, c.dtype,
                            DCGP_GEMM_LOWER | c.xf, sb, P2lo, P2lo, nullptr);
        if (rc) return rc;
        TRACE_MARK("B.trail", j, sb);
        DCGP_CU(cudaEventRecord(x.upd[e], sb));
    }
    // the chain lane has waited for every bulk update it needs, which is all of them (q >= 3)
    DCGP_CU(cudaEventRecord(x.join, sa));
    DCGP_CU(cudaStreamWaitEvent(c.st, x.join, 0));
    TRACE_DUMP();
    return 0;
}

size_t inv_ws_bytes(int64_t n, int dtype) {
    const int64_t nblk = (n + NB - 1) / NB;
    return (size_t)nblk * NB * NB * (dtype == DCGP_F64 ? 8 : 4);
}

template <typename T>
int potrf_impl(T* A, int64_t n, int64_t lda, int* info, T* ws, size_t ws_bytes, int dtype, int xf, cudaStream_t st) {
    int rc = 0;
#define CALL(M) rc = set_smem_attr<T>((const void*)potrf_diag_kernel<T, M>)
    DCGP_LEAF_DISPATCH(CALL);
#undef CALL
    if (rc) return rc;
    T* lo = nullptr;
    const size_t rcopy_bytes = 4 * NB * NB * 4;  // two block-row copies + TF32 remainders of two inverted blocks
    if (sizeof(T) == 4 && xf && n > 2 * NB && ws_bytes >= inv_ws_bytes(n, dtype) + rcopy_bytes + (size_t)n * lda * 4)
        lo = ws + (inv_ws_bytes(n, dtype) + rcopy_bytes) / sizeof(T);
    Ctx<T> c{A, lda, ws, info, dtype, xf, st, lo};
    if (lo) c.rcopy = reinterpret_cast<float*>(ws) + inv_ws_bytes(n, dtype) / 4;
    return chol_rec(c, 0, n);
}

// left-side solve with L[i0 : i0+nn, i0 : i0+nn] on the rows i0 .. i0+nn of B (n x m)
template <typename T>
struct LCtx {
    const T* L;
    int64_t ldl;
    T* B;
    int64_t m, ldb;
    T* ws;
    int dtype, xf;
    cudaStream_t st;
    const T* Llo;  // remainders of L (layout of L) and of the solved rows of B (layout of B), or null
    T* Blo;
};

template <typename T>
int ltrsm_rec(const LCtx<T>& c, int trans, int64_t i0, int64_t nn) {
    if (nn <= NB) {
        T* Bi = c.B + i0 * c.ldb;
        int rc = dcgp_gemm_impl(trans, 0, nn, c.m, nn, 1.0, c.ws + (i0 / NB) * NB * NB, NB, Bi, c.ldb, 0.0, Bi, c.ldb,
                                c.dtype, DCGP_GEMM_INPLACE | c.xf, c.st, nullptr, nullptr,
                                c.Blo ? c.Blo + i0 * c.ldb : nullptr);
        return rc;
    }
    const int64_t n1 = split_point(nn), n2 = nn - n1;
    const int64_t oc = (i0 + n1) * c.ldl + i0;
    const T* Lc = c.L + oc;  // n2 x n1
    const T* Lclo = c.Llo ? c.Llo + oc : nullptr;
    int rc;
    if (!trans) {
        rc = ltrsm_rec(c, trans, i0, n1);
        if (rc) return rc;
        rc = dcgp_gemm_impl(0, 0, n2, c.m, n1, -1.0, Lc, c.ldl, c.B + i0 * c.ldb, c.ldb, 1.0, c.B + (i0 + n1) * c.ldb,
                            c.ldb, c.dtype, c.xf, c.st, Lclo,
                            c.Blo ? c.Blo + i0 * c.ldb : nullptr, nullptr);
        if (rc) return rc;
        return ltrsm_rec(c, trans, i0 + n1, n2);
    }
    rc = ltrsm_rec(c, trans, i0 + n1, n2);
    if (rc) return rc;
    rc = dcgp_gemm_impl(1, 0, n1, c.m, n2, -1.0, Lc, c.ldl, c.B + (i0 + n1) * c.ldb, c.ldb, 1.0, c.B + i0 * c.ldb, c.ldb,
                        c.dtype, c.xf, c.st, Lclo, c.Blo ? c.Blo + (i0 + n1) * c.ldb : nullptr, nullptr);
    if (rc) return rc;
    return ltrsm_rec(c, trans, i0, n1);
}

template <typename T>
int trsm128_impl(int trans, const T* L, int64_t n, int64_t ldl, T* B, int64_t m, int64_t ldb, T* ws, size_t ws_bytes,
                 int dtype, int xf, cudaStream_t st) {
    int rc = set_smem_attr<T>((const void*)trtri_blocks_kernel<T>);
    if (rc) return rc;
    const int64_t nblk = (n + NB - 1) / NB;
    trtri_blocks_kernel<T><<<(unsigned)nblk, NTHREADS, leaf_smem<T>(), st>>>(L, ldl, n, ws, 1);
    DCGP_CHECK_LAUNCH();
    T *llo = nullptr, *blo = nullptr;
    const size_t inv = inv_ws_bytes(n, dtype);
    if (sizeof(T) == 4 && xf && n > 2 * NB && m >= 160 && ws_bytes >= inv + ((size_t)n * ldl + (size_t)n * ldb) * 4) {
        llo = ws + inv / sizeof(T);
        blo = llo + (size_t)n * ldl;
        rc = dcgp_split_lo_impl(L, llo, n, n, ldl, ldl, st);
        if (rc) return rc;
    }
    LCtx<T> c{L, ldl, B, m, ldb, ws, dtype, xf, st, llo, blo};
    return ltrsm_rec(c, trans, 0, n);
}

// ---------------------------------------------------------------------------------------------
// Solves with many right-hand sides on 512-wide inverted diagonal blocks.  The dependent chain of a
// substitution sweep shrinks from n/128 to n/512 steps, every step being a tensor-core GEMM:
//   (a) 128-wide diagonal inverses (trtri_blocks_kernel) written on the diagonal of 512-wide blocks;
//   (b) two doubling levels (128 -> 256 -> 512), X21 = -X22 (L21 X11), batched over all blocks;
//   (c) recursion on 512-row leaves: X_i = op(inv512_i) B_i (out of place) + deep-K updates.
// ---------------------------------------------------------------------------------------------
// Width of the inverted diagonal blocks.  512 by default; DCGP_NBI=1024 halves the number of steps in the serial
// chain of a sweep but the wider leaf products and the extra doubling level eat the gain (measured on B200,
// n = 8192, m = 1024: sweeps 0.82 -> 0.72 ms, plan build 0.38 -> 0.59 ms, errors 2x larger).
// Chosen at every public entry point, so plans, workspaces and solves always agree.
constexpr int NBI_MIN = 512;
constexpr int64_t NBI_DEFAULT = 1024;  // measured on the ELBO step (n = 8192, m = 1024): 512 -> 10.74, 1024 -> 10.52, 2048 -> 10.55 ms
thread_local int64_t NBI = NBI_MIN;
inline void pick_nbi(int64_t n) {
    // DCGP_NBI = 512 | 1024 | 2048 forces the width where the matrix has at least four such blocks
    static const int64_t forced = getenv("DCGP_NBI") ? atoll(getenv("DCGP_NBI")) : 0;
    int64_t w = forced ? forced : NBI_DEFAULT;
    if (w != 512 && w != 1024 && w != 2048) w = 512;
    while (w > NBI_MIN && n < 4 * w) w >>= 1;
    NBI = w;
}

template <typename T>
struct L5Ctx {
    const T* L;
    int64_t ldl;
    T* B;
    int64_t m, ldb;
    T* inv;   // [nbig][NBI][NBI]
    T* X;     // solved rows (layout of B)
    int dtype, xf;
    cudaStream_t st;
    const T* Llo;
    T* Xlo;
    int64_t n;
    T* Blo = nullptr;          // look-ahead sweep: remainders of the not yet solved rows of B
    const T* invlo = nullptr;  // ... and of the inverted diagonal blocks
    // two-step look-ahead products of the plan (build_la2): block b's row of L times the inverted diagonal blocks of its
    // two left neighbours (fwd), the inverted diagonal blocks of its two lower neighbours times their blocks in column b
    // (bwd), and their TF32 remainders
    const T *fwd = nullptr, *bwd = nullptr, *fwdlo = nullptr, *bwdlo = nullptr;
};

template <typename T>
int build_inv512(const T* L, int64_t n, int64_t ldl, T* inv, T* tmp, int dtype, int xf, cudaStream_t st) {
    const int64_t nbig = (n + NBI - 1) / NBI, nfull = n / NBI;
    cudaError_t e = cudaMemsetAsync(inv, 0, (size_t)nbig * NBI * NBI * sizeof(T), st);
    if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
    int rc = set_smem_attr<T>((const void*)trtri_blocks_kernel<T>);
    if (rc) return rc;
    const int64_t nblk = (n + NB - 1) / NB;
    trtri_blocks_kernel<T><<<(unsigned)nblk, NTHREADS, leaf_smem<T>(), st>>>(L, ldl, n, inv, NBI / NB);
    DCGP_CHECK_LAUNCH();
    const int64_t sL = (int64_t)NBI * (ldl + 1), sX = (int64_t)NBI * NBI;
    for (int b = NB; b < NBI && b < 512; b *= 2) {   // wider levels: build_inv_wide
        for (int j = 0; j < NBI / (2 * b); ++j) {
            const int64_t o1 = (int64_t)j * 2 * b, o2 = o1 + b;  // first / second half inside the 512 block
            // full blocks, batched:  T = L21 X11 ;  X21 = -X22 T
            if (nfull > 0) {
                rc = dcgp_gemm_batched_impl(0, 0, b, b, b, 1.0, L + o2 * ldl + o1, ldl, sL, inv + o1 * NBI + o1, NBI, sX, 0.0,
                                            tmp + o2 * NBI + o1, NBI, sX, (int)nfull, dtype, xf, st);
                if (rc) return rc;
                rc = dcgp_gemm_batched_impl(0, 0, b, b, b, -1.0, inv + o2 * NBI + o2, NBI, sX, tmp + o2 * NBI + o1, NBI, sX,
                                            0.0, inv + o2 * NBI + o1, NBI, sX, (int)nfull, dtype, xf, st);
                if (rc) return rc;
            }
            // ragged last block
            const int64_t rem = n - nfull * NBI;
            if (rem > o2) {
                const int64_t r2 = (rem - o2 < b) ? (rem - o2) : b;  // valid rows of the second half
                const T* Lp = L + nfull * sL;
                T* ip = inv + nfull * sX;
                T* tp = tmp + nfull * sX;
                rc = dcgp_gemm_batched_impl(0, 0, r2, b, b, 1.0, Lp + o2 * ldl + o1, ldl, 0, ip + o1 * NBI + o1, NBI, 0, 0.0,
                                            tp + o2 * NBI + o1, NBI, 0, 1, dtype, xf, st);
                if (rc) return rc;
                rc = dcgp_gemm_batched_impl(0, 0, r2, b, r2, -1.0, ip + o2 * NBI + o2, NBI, 0, tp + o2 * NBI + o1, NBI, 0, 0.0,
                                            ip + o2 * NBI + o1, NBI, 0, 1, dtype, xf, st);
                if (rc) return rc;
            }
        }
    }
    return 0;
}

// Doubling levels b >= 512 of the inverted diagonal blocks (NBI = 1024, 2048) as individual tensor-core products on the two
// lanes: the batched mma.sync kernel above runs 3xTF32 at ~30 TF/s, which made wide blocks a loss in round 1 (NBI = 1024:
// sweeps 0.82 -> 0.72 ms, plan 0.38 -> 0.59 ms).  T = L21 X11 and X21 = -X22 T with hi / lo operands: the remainders of L
// (Llo), of the blocks inverted so far (invlo, kept up to date through the epilogues) and of T.
template <typename T>
int build_inv_wide(const T* L, const T* Llo, int64_t n, int64_t ldl, T* inv, T* invlo, T* Tb, T* Tblo, int dtype, int xf,
                   cudaStream_t st) {
    LookAhead& x = lanes_for(st);
    int rc = x.init();
    if (rc) return rc;
    const int64_t nbig = (n + NBI - 1) / NBI, sX = (int64_t)NBI * NBI;
    const bool lo = Llo && invlo && Tblo;
    for (int64_t b = 512; b < NBI; b *= 2) {
        DCGP_CU(cudaEventRecord(x.fork, st));
        DCGP_CU(cudaStreamWaitEvent(x.sh, x.fork, 0));
        DCGP_CU(cudaStreamWaitEvent(x.sb, x.fork, 0));
        int turn = 0;
        for (int64_t kb = 0; kb < nbig; ++kb) {
            const int64_t row0 = kb * NBI, rows = (n - row0 < NBI) ? n - row0 : NBI;
            for (int64_t j = 0; j < NBI / (2 * b); ++j) {
                const int64_t o1 = j * 2 * b, o2 = o1 + b;
                if (rows <= o2) break;
                const int64_t r2 = (rows - o2 < b) ? rows - o2 : b;
                cudaStream_t ls = (turn++ & 1) ? x.sb : x.sh;
                const int64_t oL = (row0 + o2) * ldl + row0 + o1, o11 = kb * sX + o1 * NBI + o1, o22 = kb * sX + o2 * NBI + o2,
                              o21 = kb * sX + o2 * NBI + o1;
                rc = dcgp_gemm_impl(0, 0, r2, b, b, 1.0, L + oL, ldl, inv + o11, NBI, 0.0, Tb + o21, NBI, dtype, xf, ls,
                                    lo ? Llo + oL : nullptr, lo ? invlo + o11 : nullptr, lo ? Tblo + o21 : nullptr);
                if (rc) return rc;
                rc = dcgp_gemm_impl(0, 0, r2, b, r2, -1.0, inv + o22, NBI, Tb + o21, NBI, 0.0, inv + o21, NBI, dtype,
                                    xf | DCGP_GEMM_A_LOWER, ls, lo ? invlo + o22 : nullptr, lo ? Tblo + o21 : nullptr,
                                    lo ? invlo + o21 : nullptr);
                if (rc) return rc;
            }
        }
        DCGP_CU(cudaEventRecord(x.join, x.sh));
        DCGP_CU(cudaStreamWaitEvent(st, x.join, 0));
        DCGP_CU(cudaEventRecord(x.fork, x.sb));
        DCGP_CU(cudaStreamWaitEvent(st, x.fork, 0));
    }
    return 0;
}

inline int64_t split_point512(int64_t nn) {
    int64_t h = ((nn / 2 + NBI - 1) / NBI) * NBI;
    return h >= nn ? nn - NBI : h;
}

template <typename T>
int ltrsm512_rec(const L5Ctx<T>& c, int trans, int64_t i0, int64_t nn) {
    if (nn <= NBI) {
        // X_i = op(inv512_i) B_i, out of place; the epilogue publishes the TF32 remainders of X_i
        return dcgp_gemm_impl(trans, 0, nn, c.m, nn, 1.0, c.inv + (i0 / NBI) * NBI * NBI, NBI, c.B + i0 * c.ldb, c.ldb, 0.0,
                              c.X + i0 * c.ldb, c.ldb, c.dtype, c.xf | DCGP_GEMM_NO_TCGEN05, c.st, nullptr, nullptr,
                              c.Xlo ? c.Xlo + i0 * c.ldb : nullptr);
    }
    const int64_t n1 = split_point512(nn), n2 = nn - n1;
    const int64_t oc = (i0 + n1) * c.ldl + i0;
    const T* Lc = c.L + oc;  // n2 x n1
    const T* Lclo = c.Llo ? c.Llo + oc : nullptr;
    int rc;
    if (!trans) {
        rc = ltrsm512_rec(c, trans, i0, n1);
        if (rc) return rc;
        rc = dcgp_gemm_impl(0, 0, n2, c.m, n1, -1.0, Lc, c.ldl, c.X + i0 * c.ldb, c.ldb, 1.0, c.B + (i0 + n1) * c.ldb, c.ldb,
                            c.dtype, c.xf, c.st, Lclo, c.Xlo ? c.Xlo + i0 * c.ldb : nullptr, nullptr);
        if (rc) return rc;
        return ltrsm512_rec(c, trans, i0 + n1, n2);
    }
    rc = ltrsm512_rec(c, trans, i0 + n1, n2);
    if (rc) return rc;
    rc = dcgp_gemm_impl(1, 0, n1, c.m, n2, -1.0, Lc, c.ldl, c.X + (i0 + n1) * c.ldb, c.ldb, 1.0, c.B + i0 * c.ldb, c.ldb,
                        c.dtype, c.xf, c.st, Lclo, c.Xlo ? c.Xlo + (i0 + n1) * c.ldb : nullptr, nullptr);
    if (rc) return rc;
    return ltrsm512_rec(c, trans, i0, n1);
}

// Right-looking sweep over the 512-row blocks on the two look-ahead lanes (see chol_la):
//   chain: X_i = op(inv_i) B_i ; the neighbouring block B_{i+-1} -= op(L_{.,.}) X_i
//   bulk : all remaining blocks -= op(L) X_i        (wide, fills the machine, overlaps the chain)
// The recursion above serialises GEMMs whose 512..4096-row outputs cover a fraction of the SMs.
template <typename T>
int ltrsm512_la(const L5Ctx<T>& c, int trans) {
    LookAhead& x = lanes_for(c.st);
    int rc = x.init();
    if (rc) return rc;
    const int64_t n = c.n, q = (n + NBI - 1) / NBI;
    const cudaStream_t sa = x.sh, sb = x.sb;
    // 3xTF32 on the tcgen05 kernel needs the TF32 remainders of every operand: L and the inverted blocks
    // are split once by the caller; X_i and the block the chain touches next get theirs from the
    // epilogues (Clo); only the first block of the sweep has to be split here.
    const bool lo = c.Llo && c.Xlo && c.Blo && c.invlo;
    if (lo) {
        const int64_t r = trans ? (q - 1) * NBI : 0, w = (n - r < NBI) ? n - r : NBI;
        rc = dcgp_split_lo_impl(c.B + r * c.ldb, c.Blo + r * c.ldb, w, c.m, c.ldb, c.ldb, c.st);
        if (rc) return rc;
    }
    DCGP_CU(cudaEventRecord(x.fork, c.st));
    DCGP_CU(cudaStreamWaitEvent(sa, x.fork, 0));
    DCGP_CU(cudaStreamWaitEvent(sb, x.fork, 0));
    TRACE_BEGIN(c.st);
    for (int64_t s = 0; s < q; ++s) {
        const int64_t i = trans ? q - 1 - s : s;                    // block solved at this step
        const int64_t r = i * NBI, w = (n - r < NBI) ? n - r : NBI;
        const int e = (int)(s & 3), ep = (int)((s + 3) & 3);
        T* Xi = c.X + r * c.ldb;
        T* Xilo = lo ? c.Xlo + r * c.ldb : nullptr;
        // (the inverted block is lower triangular, its transpose upper: the tcgen05 kernel skips the zero k-tiles)
        rc = dcgp_gemm_impl(trans, 0, w, c.m, w, 1.0, c.inv + i * NBI * NBI, NBI, c.B + r * c.ldb, c.ldb, 0.0, Xi, c.ldb,
                            c.dtype, c.xf | (trans ? DCGP_GEMM_A_UPPER : DCGP_GEMM_A_LOWER), sa,
                            lo ? c.invlo + i * NBI * NBI : nullptr, lo ? c.Blo + r * c.ldb : nullptr, Xilo);
        if (rc) return rc;
        TRACE_MARK("A.leaf", s, sa);
        if (s == q - 1) break;
        DCGP_CU(cudaEventRecord(x.leaf[e], sa));
        if (s > 0) DCGP_CU(cudaStreamWaitEvent(sa, x.upd[ep], 0));
        TRACE_MARK("A.waited", s, sa);
        if (!trans) {
            const int64_t r1 = r + NBI, w1 = (n - r1 < NBI) ? n - r1 : NBI, r2 = r1 + w1, nr = n - r2;
            rc = dcgp_gemm_impl(0, 0, w1, c.m, w, -1.0, c.L + r1 * c.ldl + r, c.ldl, Xi, c.ldb, 1.0, c.B + r1 * c.ldb, c.ldb,
                                c.dtype, c.xf, sa, lo ? c.Llo + r1 * c.ldl + r : nullptr, Xilo, lo ? c.Blo + r1 * c.ldb : nullptr);
            if (rc) return rc;
            TRACE_MARK("A.near", s, sa);
            if (nr <= 0) continue;
            DCGP_CU(cudaStreamWaitEvent(sb, x.leaf[e], 0));
            TRACE_MARK("B.start", s, sb);
            rc = dcgp_gemm_impl(0, 0, nr, c.m, w, -1.0, c.L + r2 * c.ldl + r, c.ldl, Xi, c.ldb, 1.0, c.B + r2 * c.ldb, c.ldb,
                                c.dtype, c.xf, sb, lo ? c.Llo + r2 * c.ldl + r : nullptr, Xilo, nullptr);
        } else {
            const int64_t r1 = r - NBI, nr = r1;  // neighbour block i-1 (always full), rows [0, r1) remain
            rc = dcgp_gemm_impl(1, 0, NBI, c.m, w, -1.0, c.L + r * c.ldl + r1, c.ldl, Xi, c.ldb, 1.0, c.B + r1 * c.ldb, c.ldb,
                                c.dtype, c.xf, sa, lo ? c.Llo + r * c.ldl + r1 : nullptr, Xilo, lo ? c.Blo + r1 * c.ldb : nullptr);
            if (rc) return rc;
            if (nr <= 0) continue;
            DCGP_CU(cudaStreamWaitEvent(sb, x.leaf[e], 0));
            rc = dcgp_gemm_impl(1, 0, nr, c.m, w, -1.0, c.L + r * c.ldl, c.ldl, Xi, c.ldb, 1.0, c.B, c.ldb, c.dtype, c.xf, sb,
                                lo ? c.Llo + r * c.ldl : nullptr, Xilo, nullptr);
        }
        if (rc) return rc;
        TRACE_MARK("B.trail", s, sb);
        DCGP_CU(cudaEventRecord(x.upd[e], sb));
    }
    DCGP_CU(cudaEventRecord(x.join, sa));
    DCGP_CU(cudaStreamWaitEvent(c.st, x.join, 0));
    TRACE_DUMP();
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Two-step look-ahead sweep: ONE dependent GEMM per block on the chain lane.
//
// In ltrsm512_la the chain does two dependent products per 512-block, X_i = inv_i B_i and B_{i+1} -= L_{i+1,i} X_i
// (measured on B200, n = 8192, m = 1024: ~45 us per block, 16 blocks per sweep, four sweeps per ELBO step - the chain,
// not the flops, is what a solve costs).  Both products only need B_i, so with
//     F_b = [ L_{b,b-2} inv_{b-2} | L_{b,b-1} inv_{b-1} ]              (prepared once per factor, build_la2)
// block b becomes final through a single product over the two blocks before it,
//     B_b -= F_b [ B_{b-2} ; B_{b-1} ]                                  (contiguous rows of B, K = 1024)
// provided every block further left has already been applied - which the bulk lane does two steps behind:
//     bulk(s):  X_s = inv_s B_s ;  B_{s+3 ...} -= L_{s+3 ..., s} X_s.
// Writers of block b in time order: bulk(0 .. b-3), then chain(b-1) alone (it waits for bulk(b-3)), so the chain's
// epilogue also leaves the final TF32 remainders of the block.  The transposed sweep mirrors this with
//     G_b = [ inv_{b+1} L_{b+1,b} ; inv_{b+2} L_{b+2,b} ]               B_b -= G_b^T [ B_{b+1} ; B_{b+2} ].
template <typename T>
int ltrsm512_la2(const L5Ctx<T>& c, int trans) {
    LookAhead& x = lanes_for(c.st);
    int rc = x.init();
    if (rc) return rc;
    const int64_t n = c.n, q = (n + NBI - 1) / NBI;
    const cudaStream_t sa = x.sh, sb = x.sb;
    const bool lo = c.Llo && c.Xlo && c.Blo && c.invlo && c.fwdlo && c.bwdlo;
    auto width = [&](int64_t i) { return (n - i * NBI < NBI) ? n - i * NBI : (int64_t)NBI; };
    if (lo) {  // the first block of the sweep is final as given: its remainders are not produced by any epilogue
        const int64_t i = trans ? q - 1 : 0;
        rc = dcgp_split_lo_impl(c.B + i * NBI * c.ldb, c.Blo + i * NBI * c.ldb, width(i), c.m, c.ldb, c.ldb, c.st);
        if (rc) return rc;
    }
    DCGP_CU(cudaEventRecord(x.fork, c.st));
    DCGP_CU(cudaStreamWaitEvent(sa, x.fork, 0));
    DCGP_CU(cudaStreamWaitEvent(sb, x.fork, 0));
    TRACE_BEGIN(c.st);
    for (int64_t s = 0; s < q; ++s) {
        const int64_t i = trans ? q - 1 - s : s;  // block that is final at the start of this step
        const int e = (int)(s & 3);
        // ---- chain lane: make the next block final
        if (s + 1 < q) {
            if (s >= 2) DCGP_CU(cudaStreamWaitEvent(sa, x.upd[(int)((s - 2) & 3)], 0));
            TRACE_MARK("A.waited", s, sa);
            if (!trans) {
                const int64_t b = i + 1, k0 = (i >= 1 ? i - 1 : 0) * NBI, kk = (i + 1) * NBI - k0;
                rc = dcgp_gemm_impl(0, 0, width(b), c.m, kk, -1.0, c.fwd + b * 2 * NBI * NBI, 2 * NBI, c.B + k0 * c.ldb, c.ldb,
                                    1.0, c.B + b * NBI * c.ldb, c.ldb, c.dtype, c.xf, sa,
                                    lo ? c.fwdlo + b * 2 * NBI * NBI : nullptr, lo ? c.Blo + k0 * c.ldb : nullptr,
                                    lo ? c.Blo + b * NBI * c.ldb : nullptr);
            } else {
                const int64_t b = i - 1, k0 = i * NBI, k1 = (i + 2 < q ? i + 2 : q), kk = (k1 * NBI < n ? k1 * NBI : n) - k0;
                rc = dcgp_gemm_impl(1, 0, NBI, c.m, kk, -1.0, c.bwd + b * 2 * NBI * NBI, NBI, c.B + k0 * c.ldb, c.ldb, 1.0,
                                    c.B + b * NBI * c.ldb, c.ldb, c.dtype, c.xf, sa,
                                    lo ? c.bwdlo + b * 2 * NBI * NBI : nullptr, lo ? c.Blo + k0 * c.ldb : nullptr,
                                    lo ? c.Blo + b * NBI * c.ldb : nullptr);
            }
            if (rc) return rc;
            TRACE_MARK("A.leaf", s, sa);
            DCGP_CU(cudaEventRecord(x.leaf[e], sa));
        }
        // ---- bulk lane: X_i, then everything three or more blocks away
        if (s >= 1) DCGP_CU(cudaStreamWaitEvent(sb, x.leaf[(int)((s - 1) & 3)], 0));
        TRACE_MARK("B.start", s, sb);
        const int64_t r = i * NBI, w = width(i);
        T* Xi = c.X + r * c.ldb;
        T* Xilo = lo ? c.Xlo + r * c.ldb : nullptr;
        rc = dcgp_gemm_impl(trans, 0, w, c.m, w, 1.0, c.inv + i * NBI * NBI, NBI, c.B + r * c.ldb, c.ldb, 0.0, Xi, c.ldb,
                            c.dtype, c.xf | (trans ? DCGP_GEMM_A_UPPER : DCGP_GEMM_A_LOWER), sb,
                            lo ? c.invlo + i * NBI * NBI : nullptr, lo ? c.Blo + r * c.ldb : nullptr, Xilo);
        if (rc) return rc;
        if (!trans) {
            const int64_t r3 = (i + 3) * NBI;
            if (r3 < n) {
                rc = dcgp_gemm_impl(0, 0, n - r3, c.m, w, -1.0, c.L + r3 * c.ldl + r, c.ldl, Xi, c.ldb, 1.0, c.B + r3 * c.ldb,
                                    c.ldb, c.dtype, c.xf, sb, lo ? c.Llo + r3 * c.ldl + r : nullptr, Xilo, nullptr);
                if (rc) return rc;
            }
        } else {
            const int64_t nr = (i - 2) * NBI;  // rows [0, nr) are three or more blocks above
            if (nr > 0) {
                rc = dcgp_gemm_impl(1, 0, nr, c.m, w, -1.0, c.L + r * c.ldl, c.ldl, Xi, c.ldb, 1.0, c.B, c.ldb, c.dtype, c.xf,
                                    sb, lo ? c.Llo + r * c.ldl : nullptr, Xilo, nullptr);
                if (rc) return rc;
            }
        }
        TRACE_MARK("B.trail", s, sb);
        DCGP_CU(cudaEventRecord(x.upd[e], sb));
    }
    DCGP_CU(cudaEventRecord(x.join, sa));
    DCGP_CU(cudaStreamWaitEvent(c.st, x.join, 0));
    DCGP_CU(cudaEventRecord(x.fork, sb));
    DCGP_CU(cudaStreamWaitEvent(c.st, x.fork, 0));
    TRACE_DUMP();
    return 0;
}

// F_b (b = 1 .. q-1) and G_b (b = 0 .. q-2) of ltrsm512_la2 from the factor and its inverted diagonal blocks: up to four
// 512^3 products per block, independent of each other, dealt to the two lanes.
template <typename T>
int build_la2(const T* L, int64_t n, int64_t ldl, const T* inv, const T* invlo, const T* Llo, T* fwd, T* bwd, T* fwdlo,
              T* bwdlo, int dtype, int xf, cudaStream_t st) {
    LookAhead& x = lanes_for(st);
    int rc = x.init();
    if (rc) return rc;
    const int64_t q = (n + NBI - 1) / NBI;
    auto width = [&](int64_t i) { return (n - i * NBI < NBI) ? n - i * NBI : (int64_t)NBI; };
    const bool lo = invlo && Llo && fwdlo && bwdlo;
    DCGP_CU(cudaEventRecord(x.fork, st));
    DCGP_CU(cudaStreamWaitEvent(x.sh, x.fork, 0));
    DCGP_CU(cudaStreamWaitEvent(x.sb, x.fork, 0));
    int turn = 0;
    for (int64_t b = 0; b < q; ++b) {
        for (int64_t j = (b >= 2 ? b - 2 : 0); j < b; ++j) {   // F_b: columns of block j
            const int64_t off = b * 2 * NBI * NBI + (j - (b >= 2 ? b - 2 : 0)) * NBI;
            const int64_t ol = b * NBI * ldl + j * NBI;
            rc = dcgp_gemm_impl(0, 0, width(b), NBI, NBI, 1.0, L + ol, ldl, inv + j * NBI * NBI, NBI, 0.0, fwd + off, 2 * NBI,
                                dtype, xf, (turn++ & 1) ? x.sb : x.sh, lo ? Llo + ol : nullptr,
                                lo ? invlo + j * NBI * NBI : nullptr, lo ? fwdlo + off : nullptr);
            if (rc) return rc;
        }
        for (int64_t i = b + 1; i < q && i <= b + 2; ++i) {    // G_b: rows of block i
            const int64_t off = b * 2 * NBI * NBI + (i - b - 1) * NBI * NBI;
            const int64_t ol = i * NBI * ldl + b * NBI;
            rc = dcgp_gemm_impl(0, 0, width(i), NBI, width(i), 1.0, inv + i * NBI * NBI, NBI, L + ol, ldl, 0.0, bwd + off, NBI,
                                dtype, xf | DCGP_GEMM_A_LOWER, (turn++ & 1) ? x.sb : x.sh, lo ? invlo + i * NBI * NBI : nullptr,
                                lo ? Llo + ol : nullptr, lo ? bwdlo + off : nullptr);
            if (rc) return rc;
        }
    }
    DCGP_CU(cudaEventRecord(x.join, x.sh));
    DCGP_CU(cudaStreamWaitEvent(st, x.join, 0));
    DCGP_CU(cudaEventRecord(x.fork, x.sb));
    DCGP_CU(cudaStreamWaitEvent(st, x.fork, 0));
    return 0;
}

// DCGP_TRSM_LA: 1 (default) = look-ahead sweeps with two dependent GEMMs per block (ltrsm512_la), 2 = the two-step
// look-ahead sweeps above (one dependent GEMM per block, products prepared in the plan), 0 = recursive sweeps.
// Measured on B200 (profiles/r02_trsm_la2.txt, n = 8192, m = 1024, FP32): planned potrs 1.45 ms (1) vs 1.58 ms (2), plan
// build 0.32 vs 0.88 ms - the chain GEMM of mode 2 has twice the depth on the same 32 tiles, so a block costs what
// the two shallower products cost; it needs a split-K tcgen05 kernel to pay off and stays opt-in until then.
inline int trsm_la_mode() {
    static const int mode = getenv("DCGP_TRSM_LA") ? atoi(getenv("DCGP_TRSM_LA")) : 1;
    return mode;
}

// A solve plan holds everything about L that a sweep needs beyond L itself, so that several solves with
// one factor (forward and adjoint of the same Lambda^{-1}) prepare it once:
//   [inv512 blocks][their TF32 remainders (scratch of the inversion)][TF32 remainders of L]
//   [DCGP_TRSM_LA=2 only: products F_b, G_b of ltrsm512_la2, 2 x nbig blocks of 2 NBI x NBI][FP32: their TF32 remainders]
size_t plan_bytes_impl(int64_t n, int64_t ldl, int dtype) {
    const size_t es = dtype == DCGP_F64 ? 8 : 4;
    const size_t nbig = (size_t)(n + NBI - 1) / NBI;
    return 2 * nbig * NBI * NBI * es + (dtype == DCGP_F32 ? (size_t)n * ldl * 4 : 0) +
           (trsm_la_mode() >= 2 ? (dtype == DCGP_F32 ? 2 : 1) * 2 * nbig * 2 * NBI * NBI * es : 0) +
           (NBI > 512 ? 2 * nbig * NBI * NBI * es : 0);   // wide blocks: scratch of build_inv_wide (T and its remainders)
}
template <typename T>
inline T* plan_la2(const T* plan, int64_t n, int64_t ldl, int dtype) {
    const size_t nbig = (size_t)(n + NBI - 1) / NBI;
    return const_cast<T*>(plan) + 2 * nbig * NBI * NBI + (dtype == DCGP_F32 ? (size_t)n * ldl : 0);
}
template <typename T>
inline T* plan_wide_scratch(const T* plan, int64_t n, int64_t ldl, int dtype) {
    const size_t nbig = (size_t)(n + NBI - 1) / NBI;
    return plan_la2<T>(plan, n, ldl, dtype) + (trsm_la_mode() >= 2 ? (dtype == DCGP_F32 ? 2 : 1) * 2 * nbig * 2 * NBI * NBI : 0);
}

// X (+ FP32: remainders of X and of B, and a copy of B with a TMA-compatible leading dimension - a multiple of
// 4 floats - for right-hand sides whose own layout is not)
size_t trsm512_work_bytes(int64_t n, int64_t m_ld, int dtype) {
    if (dtype == DCGP_F64) return (size_t)n * m_ld * 8;
    return (size_t)n * ((m_ld + 3) & ~(int64_t)3) * 4 * 4;
}

size_t trsm512_ws_bytes(int64_t n, int64_t ldl, int64_t m_ld, int dtype) {
    return plan_bytes_impl(n, ldl, dtype) + trsm512_work_bytes(n, m_ld, dtype);
}

template <typename T>
int plan_build(const T* L, int64_t n, int64_t ldl, T* plan, int dtype, int xf, cudaStream_t st) {
    const int64_t nbig = (n + NBI - 1) / NBI;
    T* inv = plan;
    T* tmp = inv + nbig * NBI * NBI;
    int rc = build_inv512<T>(L, n, ldl, inv, tmp, dtype, xf, st);   // levels below 512 (scratch: tmp)
    if (rc) return rc;
    const bool lo = sizeof(T) == 4 && xf;
    if (lo) {
        // the doubling scratch is dead once the 512-wide inverses exist: it receives their remainders
        rc = dcgp_split_lo_impl(inv, tmp, nbig * NBI, NBI, NBI, NBI, st);
        if (rc) return rc;
        rc = dcgp_split_lo_impl(L, tmp + nbig * NBI * NBI, n, n, ldl, ldl, st);
        if (rc) return rc;
    }
    if (NBI > 512) {
        T* Tb = plan_wide_scratch<T>(plan, n, ldl, dtype);
        rc = build_inv_wide<T>(L, lo ? tmp + nbig * NBI * NBI : nullptr, n, ldl, inv, lo ? tmp : nullptr, Tb,
                               lo ? Tb + nbig * NBI * NBI : nullptr, dtype, xf, st);
        if (rc) return rc;
    }
    if (nbig >= 4 && trsm_la_mode() >= 2) {   // the sweeps that use them (trsm_impl)
        const bool lo = sizeof(T) == 4 && xf;
        T* fwd = plan_la2<T>(plan, n, ldl, dtype);
        T* bwd = fwd + nbig * 2 * NBI * NBI;
        rc = build_la2<T>(L, n, ldl, inv, lo ? tmp : nullptr, lo ? tmp + nbig * NBI * NBI : nullptr, fwd, bwd,
                          lo ? bwd + nbig * 2 * NBI * NBI : nullptr, lo ? bwd + 2 * nbig * 2 * NBI * NBI : nullptr, dtype, xf, st);
    }
    return rc;
}

// trans: 0 = L X = B, 1 = L^T X = B, 2 = both in sequence ((L L^T) X = B, shares the inverted blocks)
// plan: prepared by plan_build, or null (then it is built at the head of the workspace)
template <typename T>
int trsm_impl(int trans, const T* L, int64_t n, int64_t ldl, T* B, int64_t m, int64_t ldb, T* ws, size_t ws_bytes,
              int dtype, int xf, cudaStream_t st, const T* plan = nullptr) {
    const bool big = n >= 2 * NBI_MIN && m >= 16 &&
                     ws_bytes >= (plan ? trsm512_work_bytes(n, ldb, dtype) : trsm512_ws_bytes(n, ldl, ldb, dtype));
    if (!big) {
        if (ws_bytes < inv_ws_bytes(n, dtype)) return -8;
        if (trans != 2) return trsm128_impl<T>(trans, L, n, ldl, B, m, ldb, ws, ws_bytes, dtype, xf, st);
        int rc = trsm128_impl<T>(0, L, n, ldl, B, m, ldb, ws, ws_bytes, dtype, xf, st);
        if (rc) return rc;
        return trsm128_impl<T>(1, L, n, ldl, B, m, ldb, ws, ws_bytes, dtype, xf, st);
    }
    const int64_t nbig = (n + NBI - 1) / NBI;
    int rc;
    if (!plan) {
        rc = plan_build<T>(L, n, ldl, ws, dtype, xf, st);
        if (rc) return rc;
        plan = ws;
        ws += plan_bytes_impl(n, ldl, dtype) / sizeof(T);
    }
    const T* inv = plan;
    // FP32 3xTF32 runs on the tcgen05 kernel, whose TMA maps need 16-byte aligned rows: right-hand sides with another
    // layout (e.g. 50 or 81 columns, packed) are solved in a padded copy
    T* Bw = B;
    int64_t ldw = ldb;
    const bool pad = sizeof(T) == 4 && xf && ((ldb & 3) || (reinterpret_cast<uintptr_t>(B) & 15));
    if (pad) ldw = (m + 3) & ~(int64_t)3;
    T* X = ws;
    if (pad) {
        Bw = X + 3 * (size_t)n * ldw;
        cudaError_t e = cudaMemcpy2DAsync(Bw, (size_t)ldw * sizeof(T), B, (size_t)ldb * sizeof(T), (size_t)m * sizeof(T),
                                          (size_t)n, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
    }
    L5Ctx<T> c{L, ldl, Bw, m, ldw, const_cast<T*>(inv), X, dtype, xf, st, nullptr, nullptr, n};
    if (sizeof(T) == 4 && xf) {
        c.invlo = inv + nbig * NBI * NBI;
        c.Llo = c.invlo + nbig * NBI * NBI;
        c.Xlo = X + (size_t)n * ldw;
        c.Blo = c.Xlo + (size_t)n * ldw;
    }
    if (nbig >= 4 && trsm_la_mode() >= 2) {
        c.fwd = plan_la2<T>(plan, n, ldl, dtype);
        c.bwd = c.fwd + nbig * 2 * NBI * NBI;
        if (sizeof(T) == 4 && xf) {
            c.fwdlo = c.bwd + nbig * 2 * NBI * NBI;
            c.bwdlo = c.fwdlo + nbig * 2 * NBI * NBI;
        }
    }
    const int use_la = trsm_la_mode();
    const int last = (trans == 0 ? 0 : 1);
    for (int sweep = (trans == 1 ? 1 : 0); sweep <= last; ++sweep) {
        rc = (use_la && nbig >= 4) ? (use_la >= 2 ? ltrsm512_la2(c, sweep) : ltrsm512_la(c, sweep)) : ltrsm512_rec(c, sweep, 0, n);
        if (rc) return rc;
        // the solved rows become the right-hand side of the next sweep / the result
        T* dst = (sweep == last) ? B : Bw;
        const int64_t ldd = (sweep == last) ? ldb : ldw;
        cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)ldd * sizeof(T), X, (size_t)ldw * sizeof(T), (size_t)m * sizeof(T),
                                          (size_t)n, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
    }
    return 0;
}

// W = L^{-1} (lower, zero above) by blocked recursive doubling: 128-wide diagonal inverses, then for
// b = 128, 256, ... every 2b-wide diagonal block is completed from its halves, X21 = -X22 (L21 X11),
// batched over the blocks of a level.  2 log2(n / 128) GEMM launches in total.
template <typename T>
int trtri_impl(const T* L, int64_t n, int64_t ldl, T* W, int64_t ldw, T* tmp, int dtype, int xf, cudaStream_t st) {
    cudaError_t e = cudaMemset2DAsync(W, (size_t)ldw * sizeof(T), 0, (size_t)n * sizeof(T), (size_t)n, st);
    if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
    int rc = set_smem_attr<T>((const void*)trtri_blocks_kernel<T>);
    if (rc) return rc;
    trtri_blocks_kernel<T><<<(unsigned)((n + NB - 1) / NB), NTHREADS, leaf_smem<T>(), st>>>(L, ldl, n, W, 1, ldw);
    DCGP_CHECK_LAUNCH();
    for (int64_t b = NB; b < n; b *= 2) {
        const int64_t nfull = n / (2 * b);                 // pairs with both halves complete
        const int64_t sL = 2 * b * (ldl + 1), sW = 2 * b * (ldw + 1);
        if (nfull > 0) {
            rc = dcgp_gemm_batched_impl(0, 0, b, b, b, 1.0, L + b * ldl, ldl, sL, W, ldw, sW, 0.0, tmp + b * ldw, ldw, sW,
                                        (int)nfull, dtype, xf, st);
            if (rc) return rc;
            rc = dcgp_gemm_batched_impl(0, 0, b, b, b, -1.0, W + b * (ldw + 1), ldw, sW, tmp + b * ldw, ldw, sW, 0.0,
                                        W + b * ldw, ldw, sW, (int)nfull, dtype, xf, st);
            if (rc) return rc;
        }
        const int64_t o1 = nfull * 2 * b, o2 = o1 + b;     // ragged last pair
        if (o2 < n) {
            const int64_t r2 = n - o2;
            rc = dcgp_gemm_batched_impl(0, 0, r2, b, b, 1.0, L + o2 * ldl + o1, ldl, 0, W + o1 * (ldw + 1), ldw, 0, 0.0,
                                        tmp + o2 * ldw + o1, ldw, 0, 1, dtype, xf, st);
            if (rc) return rc;
            rc = dcgp_gemm_batched_impl(0, 0, r2, b, r2, -1.0, W + o2 * (ldw + 1), ldw, 0, tmp + o2 * ldw + o1, ldw, 0, 0.0,
                                        W + o2 * ldw + o1, ldw, 0, 1, dtype, xf, st);
            if (rc) return rc;
        }
    }
    return 0;
}

}  // namespace

// FP32 workspaces also hold the TF32-remainder matrices of the 3xTF32 tcgen05 path (assuming packed
// leading dimensions; with padded ones the drivers fall back to the mma.sync 3xTF32 kernel).
extern "C" size_t dcgp_potrf_workspace(int64_t n, int dtype) {
    if (n <= 0) return 0;
    return inv_ws_bytes(n, dtype) + (dtype == DCGP_F32 ? (size_t)n * n * 4 + 4 * NB * NB * 4 : 0);
}

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
    if (dtype == DCGP_F64)
        return potrf_impl<double>((double*)A, n, lda, info, (double*)workspace, workspace_bytes, dtype,
                                  flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05), st);
    return potrf_impl<float>((float*)A, n, lda, info, (float*)workspace, workspace_bytes, dtype,
                             flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05), st);
}

extern "C" size_t dcgp_trsm_workspace(int64_t n, int64_t m, int dtype) {
    if (n <= 0) return 0;
    pick_nbi(n);
    const size_t small = inv_ws_bytes(n, dtype) + (dtype == DCGP_F32 ? ((size_t)n * n + (size_t)n * m) * 4 : 0);
    const size_t big = trsm512_ws_bytes(n, n, m, dtype);
    return small > big ? small : big;
}

extern "C" int dcgp_trsm(int trans, const void* L, int64_t n, int64_t ldl, void* B, int64_t m, int64_t ldb,
                         void* workspace, size_t workspace_bytes, int dtype, int flags, void* stream) {
    if (n < 0) return -3;
    if (m < 0) return -6;
    if (n == 0 || m == 0) return 0;
    if (!L || ldl < n) return -2;
    if (!B || ldb < m) return -5;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -10;
    if (!workspace || workspace_bytes < inv_ws_bytes(n, dtype)) return -8;
    pick_nbi(n);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DCGP_F64)
        return trsm_impl<double>(trans, (const double*)L, n, ldl, (double*)B, m, ldb, (double*)workspace,
                                 workspace_bytes, dtype, flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05), st);
    return trsm_impl<float>(trans, (const float*)L, n, ldl, (float*)B, m, ldb, (float*)workspace, workspace_bytes,
                            dtype, flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05), st);
}

extern "C" int dcgp_potrs(const void* L, int64_t n, int64_t ldl, void* B, int64_t m, int64_t ldb, void* workspace,
                          size_t workspace_bytes, int dtype, int flags, void* stream) {
    return dcgp_trsm(2, L, n, ldl, B, m, ldb, workspace, workspace_bytes, dtype, flags, stream);
}

extern "C" size_t dcgp_solve_plan_bytes(int64_t n, int64_t ldl, int dtype) {
    if (n < 2 * NBI_MIN) return 0;  // small factors are solved on 128-wide blocks, nothing to prepare
    pick_nbi(n);
    return plan_bytes_impl(n, ldl, dtype);
}

extern "C" int dcgp_solve_plan(const void* L, int64_t n, int64_t ldl, void* plan, size_t plan_bytes, int dtype, int flags,
                               void* stream) {
    if (n < 2 * NBI_MIN) return 0;
    pick_nbi(n);
    if (!L || ldl < n) return -2;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -6;
    if (!plan || plan_bytes < plan_bytes_impl(n, ldl, dtype)) return -5;
    cudaStream_t st = (cudaStream_t)stream;
    const int xf = flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05);
    if (dtype == DCGP_F64) return plan_build<double>((const double*)L, n, ldl, (double*)plan, dtype, xf, st);
    return plan_build<float>((const float*)L, n, ldl, (float*)plan, dtype, xf, st);
}

extern "C" int dcgp_trsm_planned(int trans, const void* L, int64_t n, int64_t ldl, const void* plan, size_t plan_bytes,
                                 void* B, int64_t m, int64_t ldb, void* workspace, size_t workspace_bytes, int dtype,
                                 int flags, void* stream) {
    pick_nbi(n);
    if (!plan || n < 2 * NBI_MIN || m < 16 || plan_bytes < plan_bytes_impl(n, ldl, dtype))
        return dcgp_trsm(trans, L, n, ldl, B, m, ldb, workspace, workspace_bytes, dtype, flags, stream);
    if (!L || ldl < n) return -2;
    if (!B || ldb < m) return -7;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -12;
    if (!workspace || workspace_bytes < trsm512_work_bytes(n, ldb, dtype)) return -10;
    cudaStream_t st = (cudaStream_t)stream;
    const int xf = flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05);
    if (dtype == DCGP_F64)
        return trsm_impl<double>(trans, (const double*)L, n, ldl, (double*)B, m, ldb, (double*)workspace, workspace_bytes,
                                 dtype, xf, st, (const double*)plan);
    return trsm_impl<float>(trans, (const float*)L, n, ldl, (float*)B, m, ldb, (float*)workspace, workspace_bytes, dtype, xf,
                            st, (const float*)plan);
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

extern "C" size_t dcgp_trtri_workspace(int64_t n, int dtype) {
    if (n <= 0) return 0;
    return (size_t)n * n * (dtype == DCGP_F64 ? 8 : 4);
}

extern "C" int dcgp_trtri(const void* L, int64_t n, int64_t ldl, void* W, int64_t ldw, void* workspace,
                          size_t workspace_bytes, int dtype, int flags, void* stream) {
    if (n < 0) return -2;
    if (n == 0) return 0;
    if (!L || !W) return -1;
    if (ldl < n || ldw < n) return -3;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -8;
    if (ldw != n) return -5;  // the scratch matrix shares W's packed layout
    if (!workspace || workspace_bytes < dcgp_trtri_workspace(n, dtype)) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    const int xf = flags & DCGP_GEMM_TF32X3;
    if (dtype == DCGP_F64)
        return trtri_impl<double>((const double*)L, n, ldl, (double*)W, ldw, (double*)workspace, dtype, xf, st);
    return trtri_impl<float>((const float*)L, n, ldl, (float*)W, ldw, (float*)workspace, dtype, xf, st);
}
