// Blackwell-native TF32 contraction: C = alpha * op(A) op(B) + beta * C with FP32 storage.
//
// Persistent, warp-specialised CTAs (one per SM), each walking a strided list of 128 x TN output tiles:
//   * warp 0 (one elected lane): TMA producer - operands move global -> shared with
//     cp.async.bulk.tensor.2d (128-byte swizzle) through a 4-stage mbarrier ring that runs on across tiles;
//   * warp 1 (one elected lane): issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = TN, K = 8 per
//     instruction, 4 per 128-byte k-tile), releases stages with tcgen05.commit;
//   * the FP32 accumulators live in TMEM, double-buffered (2 x TN columns), so the MMAs of tile i+1 run
//     while warps 2..9 drain tile i: tcgen05.ld (32 lanes x 32 columns), a per-warp shared-memory
//     transpose so that global traffic is 128-byte-line coalesced, alpha/beta, optional TF32 remainder
//     output for the 3xTF32 consumers of C;
//   * 3xTF32 (A_lo B + A B_lo + A B from pre-split remainder matrices) as three MMA groups per k-tile;
//   * optional thread-block clusters of two CTAs on vertically adjacent tiles: each CTA fetches half of the shared
//     B tile and TMA-multicasts it to both (operand traffic per CTA -1/3), stages are released to both producers
//     with a multicast tcgen05.commit.
// Tiles are ordered in groups of 16 tile-rows (column-major inside a group) so that the ~148 tiles in
// flight share operand panels in L2.
//
// Both operand majors are supported through the instruction descriptor (a_major / b_major) and
// the matching canonical shared-memory layouts:
//   K-major  (row-major R x K in global): one TMA box {32 k, R rows}; rows are 128-byte lines,
//            8-row swizzle atoms every 1024 B (SBO); a k-step advances the start address by 32 B.
//   MN-major (row-major K x R in global): R/32 TMA boxes {32 r, 32 k}; each box is a 4 KB chunk of
//            32 k-rows x 128 B; LBO = 4096 B between chunks, SBO = 512 B between 4-k groups; a
//            k-step advances the start address by 1024 B.
// TMA zero-fills out-of-bounds elements, so ragged M / N / K need no special casing in the main
// loop; the epilogue guards its stores.
//
// Shapes this kernel does not take (tiny problems, split-K candidates, unaligned leading
// dimensions, 3xTF32 without remainder matrices) stay on the mma.sync kernel in gemm.cu.
#include "common.cuh"
#include "tc.cuh"
#include <cstdlib>

namespace {

constexpr int TM = 128, TK = 32;  // CTA tile rows; TK floats = one 128-byte swizzle line
constexpr int STAGES = 4;
constexpr int A_BYTES = TM * TK * 4;  // 16 KB
constexpr int EPI_WARPS = 8;
constexpr int NTHREADS_TC = 96 + 32 * EPI_WARPS;  // warp 0 TMA, warp 1 MMA (+TMEM alloc), warps 2-9 epilogue, warp 10 L2 prefetch
constexpr int STG_PITCH = 32;                     // epilogue transpose buffers: 32 x 32 floats per warp (XOR-swizzled)
constexpr int STG_BYTES = 32 * STG_PITCH * 4;
constexpr int GROUP_M = 16;
// FAT (3xTF32): one stage holds {A, A_lo, B, B_lo} of a real k-tile and feeds three MMA groups, so every operand
// tile crosses L2 -> shared memory once instead of twice (3 stages of 64 KB for 128-wide tiles, 2 of 96 KB for
// 256-wide ones).  3xTF32 products are bound by the per-SM operand streaming rate, not by the tensor pipe:
// measured on B200, potrs 8192 x 1024 1.98 -> 1.73 ms, 8192 x 8192 7.58 -> 6.96 ms.
// (Round 2 also tried making the A_lo / B_lo tiles of a stage in shared memory from the FP32 tiles - one or two extra warps
// computing x - trunc_tf32(x) behind the TMA - to halve the operand bytes of the rank-128 updates of the factorisation:
// correct, but slower everywhere (8192^2 potrf 3.52 -> 3.97 ms, planned potrs 1.25 -> 1.47 ms, the update itself 149 -> 175 us).
// ncu on that update: DRAM 49 %, L2 43 %, tensor pipe 43 % of peak, long-scoreboard stalls in the epilogue's reads of C -
// it is a latency-bound pipeline, not an operand-bandwidth-bound one.  Removed.  So was software-pipelining the epilogue's
// tcgen05.ld one chunk ahead (the load of chunk ch + 1 issued right after chunk ch left its registers): 147 -> 152 us.
// Reference points for that update (8192^2, K = 128): beta = 1 147 us single-pass and 3xTF32 alike, beta = 0 90 / 116 us;
// a plain in-place scale of the same matrix takes 83 us (tools/inplace_bw.py).)
template <int TN, bool FAT> constexpr int n_stages() { return FAT ? (TN == 128 ? 3 : 2) : STAGES; }
template <int TN, bool FAT> constexpr int stage_bytes() { return (FAT ? 2 : 1) * (A_BYTES + TN * TK * 4); }
template <int TN, bool FAT> constexpr size_t smem_tc() {
    return (size_t)n_stages<TN, FAT>() * stage_bytes<TN, FAT>() + (size_t)EPI_WARPS * STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
}

struct TcParams {
    int64_t m, n, k;
    float* C;
    float* Clo;  // optional: C - trunc_tf32(C)
    int64_t ldc;
    float alpha, beta;
    int lower;
    int a_mn, b_mn;  // operand is MN-major (stored K x R)
    int x3;          // error-compensated: 3 virtual k-tiles (A_lo B, A B_lo, A B) per real k-tile
    int tiles_m, tiles_n;
    int64_t lower_off;  // lower: keep tiles with col0 <= row0 + TM - 1 + lower_off
    int64_t clo_cols;   // Clo is written for columns < clo_cols only (0 = all)
    int a_tri;          // op(A) triangular: 1 lower (k <= row), 2 upper (k >= row) - k-tiles of zeros are skipped
    // Gram-adjoint epilogue (NC > 0): the product is an upstream gradient G = alpha op(A) op(B) of a Gram matrix
    // k(x1, x2) that is contracted with d k / d theta tile by tile and never stored
    const float* gx1;
    const float* gx2;
    int64_t gldx1, gldx2;
    float* gdls;        // [n_comp][DCGP_MAX_DIMS], accumulated
    float* gdos;        // [n_comp], accumulated
    DevSpec gsp;
};

// index -> (row, column) over a rows x cols grid, in groups of `group` rows (column-major inside a group) so that
// the CTAs in flight share operand panels in L2
__device__ __forceinline__ void grouped_coord(int t, int rows, int cols, int group, int& r, int& c) {
    const int per_group = group * cols;
    const int g = t / per_group, first = g * group;
    const int gsz = min(group, rows - first);
    const int q = t - g * per_group;
    r = first + q % gsz;
    c = q / gsz;
}

// i-th work item of this CTA: 0 = no more work, 1 = compute, 2 = nothing to do (every role skips it alike).
// CL = 1: tiles are dealt round-robin to CTAs.  CL = 2: clusters of two CTAs take vertically adjacent tile pairs
// (rank r gets tile row 2 pair + r) that share the B tile; a pair is skipped only when both tiles lie above the
// (offset) diagonal, `store` tells whether this CTA's own tile is wanted.
template <int TN, int CL>
__device__ __forceinline__ int work_item(const TcParams& p, int i, uint32_t rank, int& tm, int& tn, bool& store) {
    if constexpr (CL == 1) {
        const int t = blockIdx.x + i * gridDim.x;
        if (t >= p.tiles_m * p.tiles_n) return 0;
        grouped_coord(t, p.tiles_m, p.tiles_n, GROUP_M, tm, tn);
        store = !(p.lower && (int64_t)tn * TN > (int64_t)tm * TM + TM - 1 + p.lower_off);
        return store ? 1 : 2;
    } else {
        const int rows = (p.tiles_m + 1) / 2;
        const int t = (int)(blockIdx.x >> 1) + i * (int)(gridDim.x >> 1);
        if (t >= rows * p.tiles_n) return 0;
        int pr;
        grouped_coord(t, rows, p.tiles_n, GROUP_M / 2, pr, tn);
        tm = 2 * pr + (int)rank;
        store = !(p.lower && (int64_t)tn * TN > (int64_t)tm * TM + TM - 1 + p.lower_off);
        const bool pair = !(p.lower && (int64_t)tn * TN > (int64_t)(2 * pr + 1) * TM + TM - 1 + p.lower_off);
        return pair ? 1 : 2;
    }
}

// mapB / mapBlo: for CL = 2 and a K-major B their box holds TN / 2 rows (each CTA of the pair fetches one half
// of the shared B tile and multicasts it to both)
// real k-tiles [kt0, kt1) that can hold non-zeros of op(A) for tile row tm
__device__ __forceinline__ void k_range(const TcParams& p, int tm, int nkt, int& kt0, int& kt1) {
    kt0 = 0;
    kt1 = nkt;
    if (p.a_tri == 1) kt1 = min(nkt, ((tm + 1) * TM + TK - 1) / TK);
    if (p.a_tri == 2) kt0 = min(nkt - 1, (tm * TM) / TK);
}

template <int TN, bool FAT, int CL, int NC = 0, int DPC = 0>
__global__ void __launch_bounds__(NTHREADS_TC, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapAlo, const __grid_constant__ CUtensorMap mapBlo,
               const __grid_constant__ TcParams p) {
    constexpr int STAGE_BYTES = stage_bytes<TN, FAT>();
    constexpr int STAGES = n_stages<TN, FAT>();
    constexpr int B_BYTES = TN * TK * 4;
    constexpr int TMEM_COLS = 2 * TN;
    extern __shared__ unsigned char smem_raw[];
    // aligned by pointer arithmetic, not through an integer cast: the compiler then keeps every access below in the shared
    // state space (LDS / STS); with the cast the epilogue's transpose went through generic LD.E / ST.E (ncu source page of the
    // 3xTF32 rank-128 update, round 2: 149 -> 145 us at 8192^2, planned potrs 1.25 -> 1.21 ms)
    unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* stg_all = reinterpret_cast<float*>(tiles + (size_t)STAGES * STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(tiles + (size_t)STAGES * STAGE_BYTES + (size_t)EPI_WARPS * STG_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tmem_full = empty + STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkt = (int)((p.k + TK - 1) / TK);
    // 3xTF32: the tensor core reads FP32 operands truncated to TF32 (x_hi); with x_lo = x - x_hi given as
    // separate matrices, A B ~= A_lo B_hi + A_hi B_lo + A_hi B_hi, accumulated in the same TMEM tile.
    const uint32_t rank = (CL == 2) ? cluster_ctarank() : 0u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], CL);  // every CTA that writes into this stage waits for every CTA that reads it
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __shared__ SpecVals<float> gsv;   // Gram-adjoint epilogue only (68 bytes)
    if constexpr (NC > 0) load_spec_vals<float>(p.gsp, &gsv);
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CL == 2) cluster_sync_all();  // the peer's barriers exist before anything is multicast to them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer =====
            uint32_t it = 0;  // virtual k-tiles issued so far (ring position carries across tiles)
            for (int i = 0;; ++i) {
                int tm, tn;
                bool store;
                const int st = work_item<TN, CL>(p, i, rank, tm, tn, store);
                if (st == 0) break;
                if (st == 2) continue;
                const int m0 = tm * TM, n0 = tn * TN;
                int kt0, kt1;
                k_range(p, tm, nkt, kt0, kt1);
                const int nvt = (p.x3 && !FAT) ? 3 * (kt1 - kt0) : (kt1 - kt0);
                for (int v = 0; v < nvt; ++v, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);
                    unsigned char* sa = tiles + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(&full[s], STAGE_BYTES);
                    auto load_a = [&](unsigned char* dst, const CUtensorMap* ma, int k0) {
                        if (!p.a_mn) {
                            tma_load_2d(dst, ma, &full[s], k0, m0);
                        } else {
                            for (int c = 0; c < TM / 32; ++c) tma_load_2d(dst + c * 4096, ma, &full[s], m0 + 32 * c, k0);
                        }
                    };
                    auto load_b = [&](unsigned char* dst, const CUtensorMap* mb, int k0) {
                        if constexpr (CL == 2) {
                            // this CTA's half of the shared B tile, delivered to both CTAs of the pair
                            if (!p.b_mn) {
                                tma_load_2d_mc(dst + rank * (TN / 2) * 128, mb, &full[s], k0, n0 + (int)rank * (TN / 2), 3);
                            } else {
                                for (int c = (int)rank * (TN / 64); c < ((int)rank + 1) * (TN / 64); ++c)
                                    tma_load_2d_mc(dst + c * 4096, mb, &full[s], n0 + 32 * c, k0, 3);
                            }
                        } else if (!p.b_mn) {
                            tma_load_2d(dst, mb, &full[s], k0, n0);
                        } else {
                            for (int c = 0; c < TN / 32; ++c) tma_load_2d(dst + c * 4096, mb, &full[s], n0 + 32 * c, k0);
                        }
                    };
                    if constexpr (FAT) {
                        const int k0 = (kt0 + v) * TK;  // stage layout: A | A_lo | B | B_lo
                        load_a(sa, &mapA, k0);
                        load_a(sa + A_BYTES, &mapAlo, k0);
                        load_b(sa + 2 * A_BYTES, &mapB, k0);
                        load_b(sa + 2 * A_BYTES + B_BYTES, &mapBlo, k0);
                    } else {
                        const int kt = kt0 + (p.x3 ? v / 3 : v);
                        const int sel = p.x3 ? v % 3 : 2;  // 0: A_lo B, 1: A B_lo, 2: A B
                        load_a(sa, (sel == 0) ? &mapAlo : &mapA, kt * TK);
                        load_b(sa + A_BYTES, (sel == 1) ? &mapBlo : &mapB, kt * TK);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer =====
            // instruction descriptor: D = F32, A = B = TF32, majors, N >> 3, M >> 4
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                                   ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
            uint32_t it = 0, lt = 0;
            for (int i = 0;; ++i) {
                int tm, tn;
                bool store;
                const int st = work_item<TN, CL>(p, i, rank, tm, tn, store);
                if (st == 0) break;
                if (st == 2) continue;
                const uint32_t buf = lt & 1u;
                mbar_wait(&tmem_empty[buf], ((lt >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_base + buf * TN;
                int kt0, kt1;
                k_range(p, tm, nkt, kt0, kt1);
                const int nvt = (p.x3 && !FAT) ? 3 * (kt1 - kt0) : (kt1 - kt0);
                for (int v = 0; v < nvt; ++v, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1u;
                    mbar_wait(&full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t s0 = smem_u32(tiles + (size_t)s * STAGE_BYTES);
                    // MN-major 32-bit operands use the 32-byte-atom flavour of the 128-byte swizzle:
                    // atoms of 4 k-rows x 128 B (SBO = 512 B), 32-float chunks 4096 B apart (LBO)
                    auto group = [&](uint32_t sa, uint32_t sb, bool first) {
                        const uint32_t lbo = 4096, sbo = 512, kadv = 1024;
#pragma unroll
                        for (int kk = 0; kk < TK / 8; ++kk) {
                            const uint64_t ad = p.a_mn ? make_desc(sa + kk * kadv, lbo, sbo, 1) : make_desc(sa + kk * 32, 16, 1024, 2);
                            const uint64_t bd = p.b_mn ? make_desc(sb + kk * kadv, lbo, sbo, 1) : make_desc(sb + kk * 32, 16, 1024, 2);
                            umma_tf32(tacc, ad, bd, idesc, (first && kk == 0) ? 0u : 1u);
                        }
                    };
                    if constexpr (FAT) {
                        const uint32_t a = s0, alo = s0 + A_BYTES, b = s0 + 2 * A_BYTES, blo = b + B_BYTES;
                        group(alo, b, v == 0);
                        group(a, blo, false);
                        group(a, b, false);
                    } else {
                        group(s0, s0 + A_BYTES, v == 0);
                    }
                    // frees the stage once the MMAs above have read it (in both CTAs for a multicast pair)
                    if constexpr (CL == 2) umma_commit_mc(&empty[s], 3);
                    else umma_commit(&empty[s]);
                }
                umma_commit(&tmem_full[buf]);  // accumulator complete
                ++lt;
            }
        }
    } else if (warp == 2 + EPI_WARPS) {
        // ===== with beta != 0: pull the C tile of the tile whose MMAs are starting into L2, one tile ahead of the
        // epilogue, which cannot keep enough loads in flight by itself to cover DRAM latency =====
        if (NC == 0 && p.beta != 0.f) {
            constexpr int LPR = TN * 4 / 128;  // 128-byte lines per tile row
            uint32_t lt = 0;
            for (int i = 0;; ++i) {
                int tm, tn;
                bool store;
                const int st = work_item<TN, CL>(p, i, rank, tm, tn, store);
                if (st == 0) break;
                if (st == 2) continue;
                const int64_t m0 = (int64_t)tm * TM, n0 = (int64_t)tn * TN;
                if (lt > 0) mbar_wait(&tmem_full[(lt - 1) & 1u], ((lt - 1) >> 1) & 1u);
                if (!store) {
                    ++lt;
                    continue;
                }
                for (int e = lane; e < TM * LPR; e += 32) {
                    const int64_t row = m0 + e / LPR, col = n0 + (e % LPR) * 32;
                    if (row < p.m && col < p.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.C + row * p.ldc + col));
                }
                ++lt;
            }
        }
    } else {
        // ===== epilogue: warps 2..9; TMEM lane quadrant = warp % 4, column half = (warp - 2) / 4 =====
        // Per 32-column chunk: tcgen05.ld (thread = row) -> per-warp shared-memory transpose (32 x 32 floats, 16-byte
        // chunks XOR-swizzled with the row: conflict-free 128-bit writes by rows and reads by row segments) -> 8 lanes
        // cover one 128-byte row segment of C.  All global addressing is hoisted out of the chunk loop: with two
        // epilogue warps per scheduler the loop is bound by its instruction count, not by bandwidth.
        const int q = warp & 3, h = (warp - 2) >> 2;
        float4* stg4 = reinterpret_cast<float4*>(stg_all + (size_t)(warp - 2) * 32 * STG_PITCH);
        if constexpr (NC > 0) {
            // ===== Gram-adjoint epilogue: thread = row.  The scaled coordinates of this warp's TN / 2 columns sit in its
            // 4 KB staging buffer and are read as broadcasts; the D + 1 sums per component stay in registers over all
            // tiles of the CTA and reach memory as one atomic per value and warp. =====
            constexpr int TD = NC * DPC;
            static_assert(TD <= 8 && TN == 256, "the staging buffer holds 128 columns x 8 coordinates");
            float* cols = reinterpret_cast<float*>(stg4);
            float gl[TD], go[NC], ga[TD];
#pragma unroll
            for (int k = 0; k < TD; ++k) gl[k] = 0.f;
#pragma unroll
            for (int c = 0; c < NC; ++c) go[c] = 0.f;
            uint32_t lt = 0;
            for (int i = 0;; ++i) {
                int tm, tn;
                bool store;
                const int st = work_item<TN, CL>(p, i, rank, tm, tn, store);
                if (st == 0) break;
                if (st == 2) continue;
                const int64_t m0 = (int64_t)tm * TM, n0 = (int64_t)tn * TN;
                const uint32_t buf = lt & 1u;
                const int64_t row = m0 + q * 32 + lane;
                float a[TD];
#pragma unroll
                for (int k = 0; k < TD; ++k) a[k] = (row < p.m) ? p.gx1[row * p.gldx1 + p.gsp.col[k]] * gsv.inv_ls[k] : 0.f;
                __syncwarp();   // the previous tile's reads of `cols` are over
                for (int c = lane; c < TN / 2; c += 32) {
                    const int64_t col = n0 + h * (TN / 2) + c;
#pragma unroll
                    for (int k = 0; k < TD; ++k)
                        cols[c * TD + k] = (col < p.n) ? p.gx2[col * p.gldx2 + p.gsp.col[k]] * gsv.inv_ls[k] : 0.f;
                }
                __syncwarp();
                mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
                for (int ch = 0; ch < TN / 64; ++ch) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * TN + (uint32_t)(h * (TN / 2) + ch * 32);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                          "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (ch == TN / 64 - 1) {
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                    }
                    // rows / columns outside the matrix carry G = 0 (zero-filled operand tiles) and zero coordinates
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float b[TD];
#pragma unroll
                        for (int k = 0; k < TD; ++k) b[k] = cols[(ch * 32 + j) * TD + k];
                        const float g = p.alpha * __uint_as_float(r[j]);
                        accum_pair<float, NC, DPC, false, true>(p.gsp, gsv.os, a, b, g, g, gl, go, ga);
                    }
                }
                ++lt;
            }
#pragma unroll
            for (int k = 0; k < TD; ++k) {
                const float v = warp_sum(gl[k]);
                const int c = k / DPC, j = k % DPC;
                if (lane == 0 && p.gdls && c < p.gsp.n_comp && j < p.gsp.ndims[c])
                    atomicAdd(&p.gdls[c * DCGP_MAX_DIMS + j], v * gsv.inv_ls[k]);
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const float v = warp_sum(go[c]);
                if (lane == 0 && p.gdos && c < p.gsp.n_comp) atomicAdd(&p.gdos[c], v);
            }
        } else {
        const bool vec = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                         (!p.Clo || (reinterpret_cast<uintptr_t>(p.Clo) & 15) == 0);
        const int rsub = lane >> 3, c4 = (lane & 7) * 4;  // coalesced phase: 4 rows x 128 B per instruction
        const float alpha = p.alpha, beta = p.beta;
        const bool hasb = beta != 0.f;
        const int64_t rs = 4 * p.ldc;  // row step of the coalesced phase
        uint32_t lt = 0;
        for (int i = 0;; ++i) {
            int tm, tn;
            bool store;
            const int st = work_item<TN, CL>(p, i, rank, tm, tn, store);
            if (st == 0) break;
            if (st == 2) continue;
            const int64_t m0 = (int64_t)tm * TM, n0 = (int64_t)tn * TN;
            const uint32_t buf = lt & 1u;
            if (!store) {
                // partner-only tile of a pair (above the diagonal): accumulated for lock-step, never written
                mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                ++lt;
                continue;
            }
            const int64_t rbase = m0 + q * 32;
            const int64_t lrow0 = rbase + rsub;  // rows lrow0 + 4 i, i < 8, belong to this lane in the coalesced phase
            const int64_t left = p.m - lrow0;
            const int nrow = left <= 0 ? 0 : (left >= 29 ? 8 : (int)((left + 3) >> 2));
            float* const crow = p.C + lrow0 * p.ldc + n0 + c4;
            float* const lorow = p.Clo ? p.Clo + lrow0 * p.ldc + n0 + c4 : nullptr;
            const int ncol_lo = p.Clo ? (p.clo_cols == 0 ? 0x7fffffff : (int)(p.clo_cols - n0 - c4)) : 0;  // Clo for c0 < ncol_lo
            // old C is fetched one chunk ahead: the loads do not depend on the accumulator and stay in flight
            // across the tcgen05.ld / transpose / store work of the previous chunk
            float4 old[8], nxt[8];
            auto fetch = [&](int ch, float4 (&dst)[8]) {
                const int c0 = h * (TN / 2) + ch * 32;
                if (hasb && vec && (n0 + c0 + 32 <= p.n)) {
                    const float* src = crow + c0;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i] = (i < nrow) ? *reinterpret_cast<const float4*>(src + i * rs) : make_float4(0, 0, 0, 0);
                }
            };
            fetch(0, nxt);
#pragma unroll 1
            for (int ch = 0; ch < TN / 64; ++ch) {
                const int c0 = h * (TN / 2) + ch * 32;
                const bool fast = vec && (n0 + c0 + 32 <= p.n);
#pragma unroll
                for (int i = 0; i < 8; ++i) old[i] = nxt[i];
                if (ch + 1 < TN / 64) fetch(ch + 1, nxt);
                if (ch == 0) {
                    mbar_wait(&tmem_full[buf], (lt >> 1) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * TN + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                      "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                      "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                      "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                      "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (ch == TN / 64 - 1) {
                    // this warp's last read of the accumulator: hand the buffer back to the MMA warp
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty[buf])) : "memory");
                }
                // transpose: row `lane`, chunk c -> position c ^ (lane & 7)
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    stg4[lane * 8 + (c ^ (lane & 7))] = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]),
                                                                    __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
                __syncwarp();
                if (fast) {
                    float* dst = crow + c0;
                    const bool want_lo = c0 < ncol_lo;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rr = i * 4 + rsub;
                        const float4 sv = stg4[rr * 8 + ((lane & 7) ^ (rr & 7))];
                        float4 o;
                        if (hasb) {
                            o.x = fmaf(beta, old[i].x, alpha * sv.x);
                            o.y = fmaf(beta, old[i].y, alpha * sv.y);
                            o.z = fmaf(beta, old[i].z, alpha * sv.z);
                            o.w = fmaf(beta, old[i].w, alpha * sv.w);
                        } else {
                            o.x = alpha * sv.x; o.y = alpha * sv.y; o.z = alpha * sv.z; o.w = alpha * sv.w;
                        }
                        if (i < nrow) {
                            *reinterpret_cast<float4*>(dst + i * rs) = o;
                            if (want_lo) {
                                float4 l;
                                l.x = o.x - __uint_as_float(__float_as_uint(o.x) & 0xFFFFE000u);
                                l.y = o.y - __uint_as_float(__float_as_uint(o.y) & 0xFFFFE000u);
                                l.z = o.z - __uint_as_float(__float_as_uint(o.z) & 0xFFFFE000u);
                                l.w = o.w - __uint_as_float(__float_as_uint(o.w) & 0xFFFFE000u);
                                *reinterpret_cast<float4*>(lorow + c0 + i * rs) = l;
                            }
                        }
                    }
                } else {
                    // ragged / unaligned edge: one row per iteration, one column per lane
                    const float* stg = reinterpret_cast<const float*>(stg4);
                    const int64_t cc = n0 + c0 + lane;
                    for (int rr = 0; rr < 32; ++rr) {
                        const int64_t row = rbase + rr;
                        if (row < p.m && cc < p.n) {
                            float v = alpha * stg[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))];
                            float* dst = p.C + row * p.ldc + cc;
                            if (hasb) v = fmaf(beta, *dst, v);
                            *dst = v;
                            if (p.Clo && (p.clo_cols == 0 || cc < p.clo_cols)) p.Clo[row * p.ldc + cc] = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                        }
                    }
                }
                __syncwarp();
            }
            ++lt;
        }
        }   // store epilogue
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (CL == 2) cluster_sync_all();  // no CTA leaves while its peer can still signal or write into it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
    }
}

// lo = x - trunc_tf32(x): what the tensor core drops when it reads an FP32 operand as TF32
__global__ void split_lo_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int64_t cols,
                                int64_t lds, int64_t ldd) {
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols, c = e % cols;
        const float x = src[r * lds + c];
        dst[r * ldd + c] = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    }
}

// dst = x rounded to the nearest TF32 value (ties away from zero, cvt.rna): what a single-pass TF32 product should see
__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t rows, int64_t cols,
                                  int64_t lds, int64_t ldd) {
    const int64_t total = rows * cols;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / cols, c = e % cols;
        uint32_t v;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(v) : "f"(src[r * lds + c]));
        dst[r * ldd + c] = __uint_as_float(v);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}

// 2-D FP32 tensor map over a row-major matrix with `rows` rows of `cols` contiguous elements
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows,
             CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return 1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace

// tensor map of a row-major FP32 matrix for other translation units (gram_mm.cu): swz 0 = SWIZZLE_128B (K-major
// boxes), 1 = SWIZZLE_128B_ATOM_32B (MN-major 32-bit boxes)
int dcgp_tmap_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows, int swz) {
    return make_map(map, base, rows, cols, ld, box_cols, box_rows,
                    swz ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B);
}

int dcgp_split_lo_impl(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, cudaStream_t st) {
    if (rows <= 0 || cols <= 0) return 0;
    const int64_t total = rows * cols;
    unsigned grid = (unsigned)min((int64_t)(num_sms() * 16), (total + 255) / 256);
    split_lo_kernel<<<grid, 256, 0, st>>>((const float*)src, (float*)dst, rows, cols, lds, ldd);
    DCGP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dcgp_round_tf32(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, void* stream) {
    if (rows < 0 || cols < 0) return -3;
    if (rows == 0 || cols == 0) return 0;
    if (!src || !dst) return -1;
    if (lds < cols || ldd < cols) return -5;
    const int64_t total = rows * cols;
    unsigned grid = (unsigned)min((int64_t)(num_sms() * 16), (total + 255) / 256);
    round_tf32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, (float*)dst, rows, cols, lds, ldd);
    DCGP_CHECK_LAUNCH();
    return 0;
}

// Optional cap on the persistent grid of the calling thread's next launches (0 = none): the look-ahead factorisation runs
// its panel solves / strip updates on a helper lane beside the trailing update and gives each side its own share of the SMs.
thread_local int g_tc_cap = 0;
void dcgp_tc_cap(int ctas) { g_tc_cap = ctas; }

template <int TN, bool FAT, int CL, int NC = 0, int DPC = 0>
int launch_tc(const CUtensorMap& mapA, const CUtensorMap& mapB, const CUtensorMap& mapAlo, const CUtensorMap& mapBlo,
              TcParams& p, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<TN, FAT, CL, NC, DPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)smem_tc<TN, FAT>());
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        attr = true;
    }
    p.tiles_m = (int)((p.m + TM - 1) / TM);
    p.tiles_n = (int)((p.n + TN - 1) / TN);
    // work items: tiles, or vertically adjacent tile pairs for the multicast clusters
    const int64_t items = (CL == 2 ? (int64_t)((p.tiles_m + 1) / 2) : (int64_t)p.tiles_m) * p.tiles_n;
    // two SMs are left to the serial lanes (the single-CTA chain of the look-ahead factorisation / solves, and the
    // K_ZZ chain a captured ELBO step runs beside it): their kernels must not queue behind a persistent grid
    // occupying every SM
    static const int reserve = getenv("DCGP_TC_RESERVE") ? atoi(getenv("DCGP_TC_RESERVE")) : 2;
    int64_t cap = (num_sms() > 8 + reserve ? num_sms() - reserve : num_sms()) / CL;
    if (g_tc_cap > 0 && g_tc_cap / CL >= 1 && g_tc_cap / CL < cap) cap = g_tc_cap / CL;
    const unsigned grid = (unsigned)(items < cap ? items : cap) * CL;
    if constexpr (CL == 1) {
        gemm_tc_kernel<TN, FAT, CL, NC, DPC><<<grid, NTHREADS_TC, smem_tc<TN, FAT>(), st>>>(mapA, mapB, mapAlo, mapBlo, p);
    } else {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(NTHREADS_TC);
        cfg.dynamicSmemBytes = smem_tc<TN, FAT>();
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CL;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<TN, FAT, CL, NC, DPC>, mapA, mapB, mapAlo, mapBlo, p);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
    }
    DCGP_CHECK_LAUNCH();
    return 0;
}

// Returns 0 on success, 1 if the shape/alignment is not taken by this kernel (caller falls through
// to the mma.sync kernel), < 0 on CUDA errors.
// Alo / Blo (optional, same shape and leading dimension as A / B): x - trunc_tf32(x), enables 3xTF32.
// Clo (optional, layout of C): receives C - trunc_tf32(C).
int dcgp_gemm_tc_ex(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                    const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int flags, const void* Alo, const void* Blo,
                    void* Clo, int64_t lower_off, int64_t clo_cols, cudaStream_t st) {
    if (k < 1) return 1;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda & 3) || (ldb & 3)) return 1;
    if (m > 0x7fffffff || n > 0x7fffffff || k > 0x7fffffff) return 1;
    const int x3 = (Alo && Blo) ? 1 : 0;
    if (x3 && ((reinterpret_cast<uintptr_t>(Alo) & 15) || (reinterpret_cast<uintptr_t>(Blo) & 15))) return 1;
    // narrow tiles when the wide ones would not fill one wave of SMs
    const int64_t wide = ((m + TM - 1) / TM) * ((n + 255) / 256);
    const int tn = (2 * wide <= num_sms()) ? 128 : 256;
    // (64-wide tiles for products whose 128-wide tiles cover less than half of the SMs were measured: 512 x 1024 x 512
    // 15.9 -> 14.6 us per launch, but planned potrs and the ELBO step did not move - not kept.)
    // multicast pairs (two CTAs share every B tile) once there is more than a wave of wide tiles
    static const int cl_on = getenv("DCGP_TC_CLUSTER") ? atoi(getenv("DCGP_TC_CLUSTER")) : 2;
    // (measured on B200 at 8192^3: NN 650 -> 675, TN 597 -> 648 TF/s, but NT 692 -> 670: mode 1 = always, 2 = only when
    // B is MN-major, i.e. not transposed)
    const int cl = (cl_on && (cl_on == 1 || !transb) && tn == 256 && wide >= num_sms() && m >= 2 * TM &&
              !(flags & (DCGP_GEMM_A_LOWER | DCGP_GEMM_A_UPPER))) ? 2 : 1;  // pairs must share their k-range
    // (Multicast pairs on the 128-wide 3xTF32 tiles of the small products of the triangular sweeps were measured too:
    // 512 x 1024 x 512 stays at 16.0 us per launch, planned potrs 8192 x 1024 1.48 -> 1.60 ms - not kept.)
    CUtensorMap mapA, mapB, mapAlo, mapBlo;
    // op(A)[m, k]: notrans -> A is m x k (K-major); trans -> A is k x m (MN-major)
    // op(B)[k, n]: trans -> B is n x k (K-major); notrans -> B is k x n (MN-major)
    auto mk_a = [&](CUtensorMap* mp, const void* X) {
        return transa ? make_map(mp, X, k, m, lda, 32, TK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)
                      : make_map(mp, X, m, k, lda, TK, TM, CU_TENSOR_MAP_SWIZZLE_128B);
    };
    auto mk_b = [&](CUtensorMap* mp, const void* X) {
        return transb ? make_map(mp, X, n, k, ldb, TK, tn / cl, CU_TENSOR_MAP_SWIZZLE_128B)
                      : make_map(mp, X, k, n, ldb, 32, TK, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    };
    if (mk_a(&mapA, A) || mk_b(&mapB, B)) return 1;
    if (x3) {
        if (mk_a(&mapAlo, Alo) || mk_b(&mapBlo, Blo)) return 1;
    } else {
        mapAlo = mapA;
        mapBlo = mapB;
    }
    TcParams p;
    p.m = m; p.n = n; p.k = k;
    p.C = (float*)C; p.Clo = (float*)Clo; p.ldc = ldc;
    p.alpha = (float)alpha; p.beta = (float)beta;
    p.lower = (flags & DCGP_GEMM_LOWER) ? 1 : 0;
    p.a_mn = transa ? 1 : 0;
    p.b_mn = transb ? 0 : 1;
    p.x3 = x3;
    p.lower_off = lower_off;
    p.clo_cols = clo_cols;
    p.a_tri = (flags & DCGP_GEMM_A_LOWER) ? 1 : ((flags & DCGP_GEMM_A_UPPER) ? 2 : 0);
    p.gx1 = p.gx2 = nullptr; p.gldx1 = p.gldx2 = 0; p.gdls = p.gdos = nullptr; p.gsp = DevSpec{};
    static const int fat = getenv("DCGP_TC_FAT") ? atoi(getenv("DCGP_TC_FAT")) : 2;  // 0: off, 1: narrow tiles, 2: all
    if (tn == 128) {
        return (x3 && fat) ? launch_tc<128, true, 1>(mapA, mapB, mapAlo, mapBlo, p, st)
                           : launch_tc<128, false, 1>(mapA, mapB, mapAlo, mapBlo, p, st);
    }
    if (cl == 2)
        return (x3 && fat >= 2) ? launch_tc<256, true, 2>(mapA, mapB, mapAlo, mapBlo, p, st)
                                : launch_tc<256, false, 2>(mapA, mapB, mapAlo, mapBlo, p, st);
    return (x3 && fat >= 2) ? launch_tc<256, true, 1>(mapA, mapB, mapAlo, mapBlo, p, st)
                            : launch_tc<256, false, 1>(mapA, mapB, mapAlo, mapBlo, p, st);
}

// Gram adjoint for a low-rank upstream gradient: d theta += sum_ij G_ij d k(x1_i, x2_j) / d theta with G = alpha P Q^T
// (P: n1 x k, Q: n2 x k, row-major FP32, single-pass TF32: pass operands rounded to nearest).  G exists only as TMEM
// accumulator tiles: the epilogue of the tcgen05 kernel contracts every tile with the kernel derivatives on the spot.
// In the ELBO step the adjoints of Lambda and of K_xx are such products (dLambda = -X A^T, dK_xx = (A dG) A^T): this
// replaces an N x N x n product that wrote 268 MB and a Gram-adjoint kernel that read them back.
// Returns 0, 1 (shape / alignment / kernel layout not taken: more than 8 coordinate slots), < 0 on CUDA errors.
int dcgp_gram_bwd_lowrank_tc(const DevSpec& sp, const float* x1, int64_t n1, int64_t ldx1, const float* x2, int64_t n2,
                             int64_t ldx2, const float* P, int64_t ldp, const float* Q, int64_t ldq, int64_t k, double alpha,
                             float* dls, float* dos, cudaStream_t st) {
    if (sp.ncb * sp.dpc > 8 || k < 32 || n1 < 128 || n2 < 128) return 1;
    if ((reinterpret_cast<uintptr_t>(P) & 15) || (reinterpret_cast<uintptr_t>(Q) & 15) || (ldp & 3) || (ldq & 3)) return 1;
    if (n1 > 0x7fffffff || n2 > 0x7fffffff || k > 0x7fffffff) return 1;
    CUtensorMap mapA, mapB;
    if (make_map(&mapA, P, n1, k, ldp, TK, TM, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (make_map(&mapB, Q, n2, k, ldq, TK, 256, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    TcParams p;
    p.m = n1; p.n = n2; p.k = k;
    p.C = nullptr; p.Clo = nullptr; p.ldc = 0;
    p.alpha = (float)alpha; p.beta = 0.f;
    p.lower = 0; p.a_mn = 0; p.b_mn = 0; p.x3 = 0;
    p.lower_off = 0; p.clo_cols = 0; p.a_tri = 0;
    p.gx1 = x1; p.gx2 = x2; p.gldx1 = ldx1; p.gldx2 = ldx2;
    p.gdls = dls; p.gdos = dos; p.gsp = sp;
#define CALL(NCV, DPCV) return launch_tc<256, false, 1, NCV, DPCV>(mapA, mapB, mapA, mapB, p, st)
    if (sp.ncb == 1 && sp.dpc == 2) CALL(1, 2);
    if (sp.ncb == 1 && sp.dpc == 4) CALL(1, 4);
    if (sp.ncb == 1 && sp.dpc == 8) CALL(1, 8);
    if (sp.ncb == 2 && sp.dpc == 2) CALL(2, 2);
    if (sp.ncb == 2 && sp.dpc == 4) CALL(2, 4);
    if (sp.ncb == 4 && sp.dpc == 2) CALL(4, 2);
#undef CALL
    return 1;
}

int dcgp_gemm_tc_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                      const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int flags, const void* Alo,
                      const void* Blo, void* Clo, cudaStream_t st) {
    return dcgp_gemm_tc_ex(transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, flags, Alo, Blo, Clo, 0, 0, st);
}
