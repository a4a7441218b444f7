// Warp-level tensor-core primitives shared by gemm.cu and chol.cu (legacy mma.sync path:
// DMMA for FP64 - tcgen05 has no f64 kind - and TF32 for FP32, optionally 3xTF32-compensated).
#pragma once
#include <stdint.h>

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

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

// One warp accumulates a small product tile on the tensor cores from element functors:
//   acc += sum_{k < K} fa(row, k) * fb(k, col)
// FP64: 8 x 8 tile, acc[0..1] = (row g, cols 2t, 2t+1), K % 4 == 0.
// FP32: 16 x 8 tile (3xTF32, ~FP32 accuracy), acc[0..3] = (g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1), K % 8 == 0.
// (g = lane / 4, t = lane % 4.)  Functors return 0 outside the valid region.
//
// The order in which k is consumed is free (it is a sum), so inside a chunk of up to 4 MMA steps lane
// group t takes a contiguous k-range instead of the interleaved one of the fragment layout: for
// operands in shared memory with an odd pitch the bank of (row g, k) is then g + 8 t + s - distinct
// over the warp - where the natural order (g + t + 4 s) collides four ways.
template <typename FA, typename FB>
__device__ __forceinline__ void warp_tile_mma(double (&acc)[4], int K, FA fa, FB fb) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double c2[2] = {acc[0], acc[1]};
    for (int kb = 0; kb < K; kb += 16) {
        const int ns = min(4, (K - kb) >> 2);
        for (int s = 0; s < ns; ++s) {
            const int k = kb + ns * t + s;
            mma_f64(c2, fa(g, k), fb(k, g));
        }
    }
    acc[0] = c2[0];
    acc[1] = c2[1];
}

template <typename FA, typename FB>
__device__ __forceinline__ void warp_tile_mma(float (&acc)[4], int K, FA fa, FB fb) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int kb = 0; kb < K; kb += 32) {
        const int ns = min(4, (K - kb) >> 3);
        for (int s = 0; s < ns; ++s) {
            const int ka = kb + 2 * ns * t + s, kc = ka + ns;
            const float av[4] = {fa(g, ka), fa(g + 8, ka), fa(g, kc), fa(g + 8, kc)};
            const float bv[2] = {fb(ka, g), fb(kc, g)};
            uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
            for (int q = 0; q < 4; ++q) split_tf32(av[q], ah[q], al[q]);
#pragma unroll
            for (int q = 0; q < 2; ++q) split_tf32(bv[q], bh[q], bl[q]);
            mma_tf32(acc, al, bh);
            mma_tf32(acc, ah, bl);
            mma_tf32(acc, ah, bh);
        }
    }
}

// Register-blocked 3xTF32 product: one warp accumulates a (16 MT) x (8 NT) block, splitting every operand
// element once per k-step and reusing it across the block (3 MT NT MMAs per 4 MT + 2 NT loads).  K % 32 == 0.
template <int MT, int NT, typename FA, typename FB>
__device__ __forceinline__ void warp_block_mma(float (&acc)[MT][NT][4], int K, FA fa, FB fb) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int kb = 0; kb < K; kb += 32) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int ka = kb + 8 * t + s, kc = ka + 4;
            uint32_t ah[MT][4], al[MT][4], bh[NT][2], bl[NT][2];
#pragma unroll
            for (int mi = 0; mi < MT; ++mi) {
                split_tf32(fa(16 * mi + g, ka), ah[mi][0], al[mi][0]);
                split_tf32(fa(16 * mi + g + 8, ka), ah[mi][1], al[mi][1]);
                split_tf32(fa(16 * mi + g, kc), ah[mi][2], al[mi][2]);
                split_tf32(fa(16 * mi + g + 8, kc), ah[mi][3], al[mi][3]);
            }
#pragma unroll
            for (int ni = 0; ni < NT; ++ni) {
                split_tf32(fb(ka, 8 * ni + g), bh[ni][0], bl[ni][0]);
                split_tf32(fb(kc, 8 * ni + g), bh[ni][1], bl[ni][1]);
            }
#pragma unroll
            for (int mi = 0; mi < MT; ++mi)
#pragma unroll
                for (int ni = 0; ni < NT; ++ni) {
                    mma_tf32(acc[mi][ni], al[mi], bh[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bl[ni]);
                    mma_tf32(acc[mi][ni], ah[mi], bh[ni]);
                }
        }
    }
}

template <typename T> struct WarpTile;
template <> struct WarpTile<double> { static constexpr int ROWS = 8, COLS = 8, NACC = 2; };
template <> struct WarpTile<float> { static constexpr int ROWS = 16, COLS = 8, NACC = 4; };

// (row, col) of accumulator element q of this lane inside the tile
template <typename T>
__device__ __forceinline__ void warp_tile_coord(int q, int& row, int& col) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    row = g + ((q >> 1) << 3);
    col = 2 * t + (q & 1);
}
