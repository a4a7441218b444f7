// FP32 chain link of the look-ahead Cholesky factorisation with the 128 x 128 tile RESIDENT IN TENSOR MEMORY.
//
// One launch = one link of the serial chain of dcgp_potrf (chol.cu, chol_la_fused): given the block row R left of the
// diagonal block (already updated by the bulk lane) and X = inv(L_prev),
//     P = R X^T,   S = A_diag - P P^T,   S = L L^T,   Y = L^{-1},
// and L -> A, Y -> the workspace (+ its TF32 remainder for the bulk lane's 3xTF32 panel solve).
//
// What round 1's potrf_step_kernel did with S in shared memory and mma.sync rank-16 updates inside a 16-warp
// block_potrf (38 us of a 57 us link with the inverse riding along: the trailing updates of the factorisation and of the
// inverse rows are ~4 MFLOP-MMAs x 3 passes on the legacy tensor pipe of ONE SM), this kernel does on tcgen05:
//   * S (128 lanes x 128 columns) and the rows of the inverse (Y: lane c = row c of L^{-T}, starts as the identity) are
//     two FP32 accumulator tiles in TMEM for the whole kernel;
//   * per 16-column sub-panel: one warp pulls the 16 x 16 diagonal block out of TMEM (tcgen05.ld: one row per lane),
//     factors it in registers (two columns per shuffle round, chol.cu diag_step<float>), publishes L_d; every row of S
//     below and every started row of Y solves its 16 entries against L_d in registers (one thread per row), writes them
//     to global memory (they are final) and - split into TF32 hi / lo - into a K-major SWIZZLE_128B operand tile;
//   * the rank-16 updates S -= P16 P16^T and Y -= Y16 P16^T are tcgen05.mma with a negated A operand accumulating into
//     the TMEM tiles: first the 16 columns the next sub-panel needs (N = 16: the only MMAs the chain waits for), then the
//     rest of the trailing columns, which complete under the next sub-panel's pivots.
#include "common.cuh"
#include "mma.cuh"
#include "tc.cuh"
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int NB = 128, SB = 16, NT = 512;
constexpr int TBUF = 64 * 1024;                 // one 128 x 128 FP32 operand tile (four 16 KB k-blocks)
constexpr size_t TM_SMEM = 3 * TBUF + 1024;

__device__ __forceinline__ uint32_t canon_off(int r, int k) {
    return (uint32_t)((k >> 5) * 16384 + r * 128 + ((((k & 31) >> 2) ^ (r & 7)) << 4) + ((k & 3) << 2));
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_lo(float x) { return x - tf32_hi(x); }

// D[128 x n] (+)= (-)A[128 x 8 ksteps] B[n x 8 ksteps]^T, both operands K-major canonical tiles; b_row0 = first row of B
__device__ __forceinline__ void issue_mma(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int b_row0, int n, int ksteps,
                                          bool negate, bool fresh) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (negate ? (1u << 13) : 0u) | ((uint32_t)(n >> 3) << 17) |
                           ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t koff = (uint32_t)((ks >> 2) * 16384 + (ks & 3) * 32);
        umma_tf32(tmem_d, make_desc(a_addr + koff, 16, 1024, 2), make_desc(b_addr + (uint32_t)b_row0 * 128 + koff, 16, 1024, 2),
                  idesc, (fresh && ks == 0) ? 0u : 1u);
    }
}

#define TM_LD16(r, taddr)                                                                                               \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), \
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])             \
                 : "r"(taddr));                                                                                         \
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

#define TM_LD32(r, taddr)                                                                                            \
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

#define TM_ST32(taddr, r)                                                                                            \
    asm volatile(                                                                                                    \
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                              \
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "                                   \
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"                           \
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), \
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), \
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), \
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                \
        : "memory")

__device__ __forceinline__ float rsqrt_ftz_nr(float d) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(d));
    return y * fmaf(-0.5f * d, y * y, 1.5f);
}

// HAS_PREV = false: first block of a factorisation (no P, no update: S = A_diag)
template <bool HAS_PREV>
__global__ void __launch_bounds__(NT, 1) potrf_link_tm_kernel(float* __restrict__ A, int64_t lda, int nb, float* __restrict__ inv,
                                                              float* __restrict__ invlo, int* info, int64_t j0,
                                                              const float* __restrict__ R, const float* __restrict__ invp) {
    extern __shared__ unsigned char smem_dyn[];
    // (pointer arithmetic, not an integer round trip: the compiler keeps the shared address space and emits LDS / STS)
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* B0 = base;
    unsigned char* B1 = base + TBUF;
    unsigned char* B2 = base + 2 * TBUF;
    __shared__ __align__(16) float Ld[2][SB][SB];
    __shared__ __align__(16) float Tm[2][SB][SB];   // L_d^{-T} of the current 16 x 16 diagonal block
    __shared__ uint64_t bar, bar_u, bar_r;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef DCGP_LEAF_TIMING
    long long tq[8] = {clock64(), 0, 0, 0, 0, 0, 0, 0}, ta = 0, tb = 0, tc = 0, tl, tb1 = 0, tb2 = 0, tb3 = 0, tb4 = 0, tc1 = 0, tc2 = 0, tm;
#define TMARK(i) tq[i] = clock64()
#else
#define TMARK(i)
#endif

    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar_u, 2);   // two issuing threads (S and Y updates) commit to each of these
        mbar_init(&bar_r, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (HAS_PREV) {
        // R -> B0, R_lo -> B2, X -> B1 (canonical operand layout); 16-byte vectors, 8 lanes per 128-byte row segment
        for (int e = tid; e < NB * NB / 4; e += NT) {
            const int i = e >> 5, k = (e & 31) * 4;
            const float4 r = (i < nb) ? *reinterpret_cast<const float4*>(R + i * NB + k) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 x = *reinterpret_cast<const float4*>(invp + i * NB + k);
            const uint32_t off = canon_off(i, k);
            *reinterpret_cast<float4*>(B0 + off) = r;
            *reinterpret_cast<float4*>(B2 + off) = make_float4(tf32_lo(r.x), tf32_lo(r.y), tf32_lo(r.z), tf32_lo(r.w));
            *reinterpret_cast<float4*>(B1 + off) = x;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t a0 = smem_u32(B0), a1 = smem_u32(B1), a2 = smem_u32(B2);
    const uint32_t TS = tmem, TY = tmem + 128, TP = tmem + 256;
    uint32_t ph = 0;   // phase of `bar`
    TMARK(1);          // operands of P in shared memory

    if (HAS_PREV && tid == 0) {   // P = R X^T into TP: R_hi X_hi + R_lo X_hi now, R_hi X_lo after the swap below
        issue_mma(TP, a0, a1, 0, 128, 16, false, true);
        issue_mma(TP, a2, a1, 0, 128, 16, false, false);
        umma_commit(&bar);
    }
    // ---- meanwhile: the diagonal tile -> TS (lane = row; identity padding outside the valid lower triangle; warps 0-3
    // fill columns 0..63 of their rows, warps 8-11 - same TMEM lane quadrants - columns 64..127) and I -> TY (warps 4-7)
    if (warp < 12) {
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        if (warp < 4 || warp >= 8) {
            // 16-byte loads of the row's lower part (the caller guarantees lda % 4 == 0 and a 16-byte aligned A), all of a
            // 32-column chunk in flight before the first use
            const float4* arow = reinterpret_cast<const float4*>(A + (int64_t)row * lda);
            const int cbeg = warp < 4 ? 0 : 64;
#pragma unroll 1
            for (int c0 = cbeg; c0 < cbeg + 64; c0 += 32) {
                float4 q4[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    q4[c] = (row < nb && c0 + 4 * c <= row) ? arow[(c0 >> 2) + c] : make_float4(0.f, 0.f, 0.f, 0.f);
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float e4[4] = {q4[c].x, q4[c].y, q4[c].z, q4[c].w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = c0 + 4 * c + u;
                        v[4 * c + u] = __float_as_uint((row < nb && k <= row) ? e4[u] : ((k == row) ? 1.f : 0.f));
                    }
                }
                TM_ST32(TS + lane_base + (uint32_t)c0, v);
            }
        } else {
#pragma unroll 1
            for (int c0 = 0; c0 < NB; c0 += 32) {
                uint32_t v[32];
#pragma unroll
                for (int c = 0; c < 32; ++c) v[c] = __float_as_uint((c0 + c == row) ? 1.f : 0.f);
                TM_ST32(TY + lane_base + (uint32_t)c0, v);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    TMARK(2);          // diagonal tile and identity in TMEM
    if (HAS_PREV) {
        mbar_wait(&bar, ph);
        ph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int e = tid; e < NB * NB / 4; e += NT) {
            const float4 x = *reinterpret_cast<const float4*>(B1 + e * 16);  // elementwise: same offsets in both buffers
            *reinterpret_cast<float4*>(B2 + e * 16) = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            issue_mma(TP, a0, a2, 0, 128, 16, false, false);
            umma_commit(&bar);
        }
        mbar_wait(&bar, ph);
        ph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // P (TP) -> operand tiles: P_hi in B0, P_lo in B2.  Warp w reads lane quadrant w % 4, columns 32 (w / 4) ..
        {
            const int row = (warp & 3) * 32 + lane, kb = warp >> 2;
            uint32_t r[32];
            TM_LD32(r, TP + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kb * 32));
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
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TMARK(3);          // P complete and back in shared memory as hi / lo operands
    if (HAS_PREV) {
        if (tid == 0) {   // TS -= P P^T (negated A operand, accumulating onto the diagonal tile)
            issue_mma(TS, a0, a0, 0, 128, 16, true, false);
            issue_mma(TS, a2, a0, 0, 128, 16, true, false);
            issue_mma(TS, a0, a2, 0, 128, 16, true, false);
            umma_commit(&bar);
        }
        mbar_wait(&bar, ph);
        ph ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ---- factorisation + inverse, 16 columns at a time.  Operand sets (P16 hi / lo, Y16 hi / lo: 128 rows x 128 B each)
    // alternate between B0 (+0 .. +64 KB) and B1; B0's first use (p = 0) comes after the update above has finished with it.
    // Every row thread (S rows: warps 0-3, Y rows: warps 4-7) pulls the 16 entries of its row in the current sub-panel out
    // of TMEM once the narrow ("urgent") update of the previous sub-panel has landed; the pivot warp factors the diagonal
    // block out of the same registers.  (A tcgen05.ld issued behind queued MMAs waits for them - 1600 .. 2900 cycles per
    // sub-panel behind the wide update, B200 clock64.  Loading one sub-panel ahead, before the wide update is queued,
    // needs one more CTA barrier per sub-panel and measured slower: link 44 .. 46 us against 40 .. 42 us,
    // profiles/r02_link_tm.txt.)
    uint32_t n_u = 0, n_r = 0;
    const bool isY = warp >= 4;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t tm_row = (isY ? TY : TS) + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t v[16];
    const uint64_t dbase = make_desc(0, 16, 1024, 2);
    // six MMAs: D[:, c0 : c0 + n] -= A16 B16[c0 : c0 + n]^T as 3xTF32 (hi hi, lo hi, hi lo), K = 16 = two k-steps
    auto issue3x = [&](uint32_t tm_d, uint32_t ahi, uint32_t alo, uint32_t bhi, uint32_t blo, int c0, int n) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 13) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t dah = dbase | (uint64_t)((ahi >> 4) & 0x3FFF), dal = dbase | (uint64_t)((alo >> 4) & 0x3FFF);
        const uint64_t dbh = dbase | (uint64_t)(((bhi + (uint32_t)c0 * 128) >> 4) & 0x3FFF);
        const uint64_t dbl = dbase | (uint64_t)(((blo + (uint32_t)c0 * 128) >> 4) & 0x3FFF);
        umma_tf32(tm_d + (uint32_t)c0, dah, dbh, idesc, 1u);
        umma_tf32(tm_d + (uint32_t)c0, dah + 2, dbh + 2, idesc, 1u);     // + 32 bytes: the second k-step
        umma_tf32(tm_d + (uint32_t)c0, dal, dbh, idesc, 1u);
        umma_tf32(tm_d + (uint32_t)c0, dal + 2, dbh + 2, idesc, 1u);
        umma_tf32(tm_d + (uint32_t)c0, dah, dbl, idesc, 1u);
        umma_tf32(tm_d + (uint32_t)c0, dah + 2, dbl + 2, idesc, 1u);
    };
    TMARK(4);          // S = A_diag - P P^T in TMEM
    for (int p = 0; p * SB < nb; ++p) {
        const int p0 = p * SB, pw = min(SB, nb - p0), t0 = p0 + SB, set = p & 1;
#ifdef DCGP_LEAF_TIMING
        tl = clock64();
#endif
        // (a) pivots: the warp whose rows 16 (p & 1) .. hold the 16 x 16 diagonal block factors it out of its registers
        if (warp == (p >> 1)) {
            TM_LD16(v, tm_row + (uint32_t)p0);
            const int off = (p & 1) * 16, r = lane - off;
            const bool act = r >= 0 && r < pw, mine = r >= 0 && r < SB;
            // the 16 lanes of the other half carry the rows of the identity: what the elimination makes of them is
            // T = L_d^{-T} (row i = e_i L_d^{-T}), with which every other row solves its 16 entries as a product
            float rw[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c)
                rw[c] = mine ? ((act && c <= r) ? __uint_as_float(v[c]) : ((r == c) ? 1.f : 0.f)) : (((lane & 15) == c) ? 1.f : 0.f);
            int bad = 0;
#pragma unroll
            for (int c = 0; c < SB; c += 2) {
                const float a = __shfl_sync(0xffffffffu, rw[c], c + off);
                const float b = __shfl_sync(0xffffffffu, rw[c], c + 1 + off);
                const float e = __shfl_sync(0xffffffffu, rw[c + 1], c + 1 + off);
                const float det = fmaf(a, e, -(b * b));
                if (bad == 0) {
                    if (!(a > 0.f)) bad = c + 1;
                    else if (!(det > 0.f)) bad = c + 2;
                }
                const float y1 = rsqrt_ftz_nr(a), yd = rsqrt_ftz_nr(det);
                const float sa = a * y1, y2 = sa * yd, l21 = b * y1;
                const float l1 = (r == c) ? sa : rw[c] * y1;
                const float t = fmaf(-l1, l21, rw[c + 1]);
                const float l2 = (r == c + 1) ? (det * yd) * y1 : t * y2;
                rw[c] = l1;
                rw[c + 1] = l2;
#pragma unroll
                for (int k = c + 2; k < SB; ++k) {
                    const float lk1 = __shfl_sync(0xffffffffu, l1, k + off);
                    const float lk2 = __shfl_sync(0xffffffffu, l2, k + off);
                    rw[k] = fmaf(-l2, lk2, fmaf(-l1, lk1, rw[k]));
                }
            }
            if (mine) {
#pragma unroll
                for (int c = 0; c < SB; ++c) Ld[set][r][c] = (c <= r) ? rw[c] : 0.f;
            } else {
#pragma unroll
                for (int c = 0; c < SB; ++c) Tm[set][lane & 15][c] = (c >= (lane & 15)) ? rw[c] : 0.f;
            }
            if (lane == off && bad != 0 && bad <= pw && info && *info == 0) *info = (int)(j0 + p0 + bad);
        }
        __syncthreads();
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); ta += t - tl; tl = t; tm = t; }
#endif
        // (b) one thread per row: x = (its 16 entries) L_d^{-T}, published as the operands of the rank-16 updates.  (The
        // stores of the final values to global memory come after the MMAs are issued: fence.proxy.async waits for the
        // thread's outstanding memory operations.)
        float x[SB];
        if (warp < 8) {
            if (warp != (p >> 1)) { TM_LD16(v, tm_row + (uint32_t)p0); }
            const bool solve = isY ? (row < p0 + pw && row < nb) : (row >= t0 && row < nb);
            const bool diag = !isY && row >= p0 && row < p0 + pw;
            if (solve) {
                // x[c] = sum_{i <= c} a[i] T[i][c]: 136 independent multiply-adds, T rows as 16-byte broadcasts
#pragma unroll
                for (int c = 0; c < SB; ++c) x[c] = 0.f;
#pragma unroll
                for (int i = 0; i < SB; ++i) {
                    const float ai = __uint_as_float(v[i]);
#pragma unroll
                    for (int c4 = (i >> 2); c4 < 4; ++c4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(&Tm[set][i][4 * c4]);
                        x[4 * c4] = fmaf(ai, t4.x, x[4 * c4]);
                        x[4 * c4 + 1] = fmaf(ai, t4.y, x[4 * c4 + 1]);
                        x[4 * c4 + 2] = fmaf(ai, t4.z, x[4 * c4 + 2]);
                        x[4 * c4 + 3] = fmaf(ai, t4.w, x[4 * c4 + 3]);
                    }
                }
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (c >= pw) x[c] = 0.f;
            } else {
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 l4 = diag ? *reinterpret_cast<const float4*>(&Ld[set][diag ? row - p0 : 0][4 * c4]) : make_float4(0.f, 0.f, 0.f, 0.f);
                    x[4 * c4] = l4.x; x[4 * c4 + 1] = l4.y; x[4 * c4 + 2] = l4.z; x[4 * c4 + 3] = l4.w;
                }
            }
#ifdef DCGP_LEAF_TIMING
            { long long t = clock64(); tb2 += t - tm; tm = t; }
#endif
            // operand tiles: 16 floats = chunks 0 .. 3 of the row's 128-byte line
            unsigned char* hi = (set ? B1 : B0) + (isY ? 2 * 16384 : 0);
            unsigned char* lo = hi + 16384;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float4 h4 = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
                const uint32_t off = (uint32_t)(row * 128 + ((j ^ (row & 7)) << 4));
                *reinterpret_cast<float4*>(hi + off) = h4;
                *reinterpret_cast<float4*>(lo + off) = make_float4(tf32_lo(h4.x), tf32_lo(h4.y), tf32_lo(h4.z), tf32_lo(h4.w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#ifdef DCGP_LEAF_TIMING
            { long long t = clock64(); tb3 += t - tm; tm = t; }
#endif
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); tb += t - tl; tl = t; }
#endif
        const bool more = t0 < nb;
        const uint32_t sp = smem_u32(set ? B1 : B0);
        // (c) rank-16 updates, the next sub-panel's 16 columns ("urgent": the only MMAs the chain waits for) ahead of the
        // remaining trailing columns: S by thread 0, Y by thread 128 - two issuers, two arrivals per barrier
        if (more && tid == 0) {
            issue3x(TS, sp, sp + 16384, sp, sp + 16384, t0, SB);
            umma_commit(&bar_u);
            if (NB - t0 - SB > 0) {
                issue3x(TS, sp, sp + 16384, sp, sp + 16384, t0 + SB, NB - t0 - SB);
                umma_commit(&bar_r);
            }
        }
        if (more && tid == 128) {
            issue3x(TY, sp + 2 * 16384, sp + 3 * 16384, sp, sp + 16384, t0, SB);
            umma_commit(&bar_u);
            if (NB - t0 - SB > 0) {
                issue3x(TY, sp + 2 * 16384, sp + 3 * 16384, sp, sp + 16384, t0 + SB, NB - t0 - SB);
                umma_commit(&bar_r);
            }
        }
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); tc1 += t - tl; tm = t; }
#endif
        // final values -> global memory, under the MMAs
        if (warp < 8) {
            if (!isY) {
                if (row >= t0 && row < nb && pw == SB) {          // a full row below the block: four 16-byte stores
                    float4* dst = reinterpret_cast<float4*>(A + (int64_t)row * lda + p0);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) dst[c4] = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);
                } else if (row >= p0 && row < nb) {
                    float* dst = A + (int64_t)row * lda + p0;
#pragma unroll
                    for (int c = 0; c < SB; ++c)
                        if (c < pw && p0 + c <= row) dst[c] = x[c];
                }
            } else {
#pragma unroll
                for (int c = 0; c < SB; ++c) {
                    if (c < pw) {   // X(p0 + c, row), zero above the diagonal of X
                        const float val = (p0 + c >= row) ? x[c] : 0.f;
                        inv[(p0 + c) * NB + row] = val;
                        if (invlo) invlo[(p0 + c) * NB + row] = tf32_lo(val);
                    }
                }
            }
        }
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); tc2 += t - tm; tm = t; }
#endif
        if (!more) break;
        mbar_wait(&bar_u, n_u & 1u);
        ++n_u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (NB - t0 - SB > 0) ++n_r;
#ifdef DCGP_LEAF_TIMING
        { long long t = clock64(); tc += t - tl; tl = t; tb4 += t - tm; }
#endif
    }
    TMARK(5);
    if (n_r > 0) mbar_wait(&bar_r, (n_r - 1) & 1u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512));
    }
#ifdef DCGP_LEAF_TIMING
    if ((tid == 0 || tid == 200) && (j0 == 4096))
        printf("link_tm tid %d j0=%lld: operands %lld  tmem-fill %lld  P %lld  update %lld  loop %lld (pivots %lld  rows %lld [wait+ld %lld compute %lld sts+fence %lld rest-issue %lld]  mma+wait %lld [issue %lld stores %lld])  tail %lld\n",
               tid, (long long)j0, tq[1] - tq[0], tq[2] - tq[1], tq[3] - tq[2], tq[4] - tq[3], tq[5] - tq[4], ta, tb, tb1, tb2, tb3, tb4, tc, tc1, tc2, clock64() - tq[5]);
#endif
}

}  // namespace

// One chain link on stream st.  R / invp null: first block (no update).  invlo optional.
int dcgp_potrf_link_tm(float* A, int64_t lda, int nb, float* inv, float* invlo, int* info, int64_t j0, const float* R,
                       const float* invp, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(potrf_link_tm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(potrf_link_tm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TM_SMEM);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        attr = true;
    }
    if (R && invp) potrf_link_tm_kernel<true><<<1, NT, TM_SMEM, st>>>(A, lda, nb, inv, invlo, info, j0, R, invp);
    else potrf_link_tm_kernel<false><<<1, NT, TM_SMEM, st>>>(A, lda, nb, inv, invlo, info, j0, nullptr, nullptr);
    DCGP_CHECK_LAUNCH();
    return 0;
}
