// Fused Gram-times-matrix on the 5th-generation tensor cores (FP32 storage, single-pass TF32):
//     out[n1 x m] = (k(x1, x2) [+ c I when x2 is x1]) W[n2 x m]
// with the Gram matrix NEVER written to global memory (src/kernels/cme_aggregate_kernel.py:58-70 - the K R products of
// the CME aggregate kernel -, src/models/cmp.py:87-104 and variational_strategy.py:57-66 evaluate a lazy kernel matrix
// and multiply it; SURVEY 7 step 4 / 8d: "K tiles generated on the fly, never written").
//
// One CTA owns 128 rows of the result and TN of its columns and runs down the n2 (contraction) dimension in k-tiles of
// 32 Gram columns:
//   * warps 2..17 (generators): evaluate the 128 x 32 Gram tile - lane = Gram column (its scaled x2 coordinates live in
//     registers for the k-tile), warp = row phase (scaled x1 coordinates of the CTA's 128 rows are staged in shared
//     memory once) - round every value to the nearest TF32 number (tcgen05 reads operands by truncation) and store it
//     straight into the canonical K-major SWIZZLE_128B operand layout of tcgen05.mma, then arrive on the stage's
//     mbarrier after a generic->async proxy fence;
//   * warp 0 (one lane): TMA loads of the matching 32 x TN tile of W (MN-major boxes) into the same stage;
//   * warp 1 (one lane): four tcgen05.mma.kind::tf32 (M = 128, N = TN, K = 8) per k-tile, accumulators in TMEM,
//     tcgen05.commit releases the stage to both producers;
//   * epilogue (warps 2..17 again): tcgen05.ld, + c W[row] on the diagonal-carrying rows, 16-byte stores.
// Algorithmic traffic: n1 m + n2 m values (+ coordinates); the N^2 Gram values exist only in shared memory.  Every
// column block of the result regenerates its Gram tiles (ceil(m / TN) times n1 n2 kernel evaluations in total): for the
// tall-skinny W of this path (m = number of bags <= 1024) that is 1 .. 4 passes of MUFU work under the MMAs.
#include "common.cuh"
#include "mma.cuh"
#include "tc.cuh"

int dcgp_tmap_f32(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_cols, int box_rows, int swz);

namespace {

constexpr int GM_TM = 128, GM_TK = 32, GM_GEN_WARPS = 16, GM_THREADS = 96 + 32 * GM_GEN_WARPS, GM_STAGES = 4;
constexpr int GM_A_BYTES = GM_TM * GM_TK * 4;

template <int TN> constexpr int gm_stage_bytes() { return GM_A_BYTES + TN * GM_TK * 4; }
template <int TN, int TD> constexpr size_t gm_smem() {
    return (size_t)GM_STAGES * gm_stage_bytes<TN>() + (size_t)GM_TM * (TD + 1) * 4 + (size_t)GM_STAGES * GM_TK * (TD + 4) * 4 +
           1024 /*align*/ + 256 /*barriers*/;
}

__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct GmParams {
    const float* x1;
    const float* x2;
    const float* W;
    float* out;
    const float* diag_dev;
    int64_t n1, n2, m, ldx1, ldx2, ldw, ldo;
    float diag_const;
    int same;
};

// RBF1: a single RBF term (the common case: ScaleKernel(RBFKernel(ard))) takes the short form of the evaluation - the
// first version spent 49 instructions per Gram entry (ncu, profiles/r02_ncu_gram_mm_v2.txt: issue slots 59 % busy,
// tensor pipe 20 %), i.e. the generators, not the MMAs, set the pace:
//     k = 2^(h1 + h2 + u1 . u2),  u = x / l * sqrt(log2 e),  h1 = log2 os - |u1|^2 / 2,  h2 = -|u2|^2 / 2
// (TD fused multiply-adds and one ex2; the expansion's cancellation error ~1e-7 |u|^2 is far below the TF32 rounding of
// the operand), with every address of the swizzled store folded into compile-time offsets, the rounding to TF32 done as
// an integer add + mask (the values are positive and finite), and - when the component has fewer dimensions than its
// template bucket (d = 3 in a bucket of 4, d = 5 in 8: MODE 2) - h1 riding in the free coordinate slot against a 1 on the
// x2 side, so that one 16-byte shared-memory load and TD multiply-adds produce the whole exponent.  v2 -> v4 -> v5 on
// B200, N = 16384, m = 328, d = 3: 1.44 -> 0.81 -> see profiles/r02_gram_mm.txt ms (2490 warp instructions per k-tile in
// v4, ncu: 620 issue cycles against 512 cycles of MMA work).
template <int NC, int DPC, int TN, int MODE>
__global__ void __launch_bounds__(GM_THREADS, 1)
gram_mm_tc_kernel(const __grid_constant__ DevSpec sp, const __grid_constant__ CUtensorMap mapW, const __grid_constant__ GmParams p) {
    constexpr int TD = NC * DPC;
    constexpr bool RBF1 = MODE != 0, SLOT = MODE == 2;   // SLOT: the component has fewer dims than its bucket
    constexpr int STAGE = gm_stage_bytes<TN>();
    extern __shared__ unsigned char smem_raw[];
    // (pointer arithmetic, not an integer round trip: the compiler keeps the shared address space and emits LDS / STS)
    unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* xs1 = reinterpret_cast<float*>(tiles + (size_t)GM_STAGES * STAGE);            // [128][TD] scaled x1 coordinates
    float* hs1 = xs1 + GM_TM * TD;                                                       // [128] RBF1: -|u1|^2 / 2
    constexpr int XC = TD + 4;                                                           // per column: TD coordinates, h2, pad
    float* xc = hs1 + GM_TM;                                                             // [stages][32][XC] scaled x2 coordinates
    uint64_t* full = reinterpret_cast<uint64_t*>(xc + GM_STAGES * GM_TK * XC);
    uint64_t* empty = full + GM_STAGES;
    uint64_t* cready = empty + GM_STAGES;
    uint64_t* acc_full = cready + GM_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    __shared__ SpecVals<float> sv;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t i0 = (int64_t)blockIdx.y * GM_TM, c0 = (int64_t)blockIdx.x * TN;
    const int nkt = (int)((p.n2 + GM_TK - 1) / GM_TK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < GM_STAGES; ++s) {
            mbar_init(&full[s], 1 + GM_GEN_WARPS);   // the TMA producer's expect_tx arrival + one arrival per generator warp
            mbar_init(&empty[s], 1);
            mbar_init(&cready[s], 1);
        }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    load_spec_vals<float>(sp, &sv);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ===== TMA producer: W[k0 : k0 + 32, c0 : c0 + TN] as TN / 32 boxes of {32 columns, 32 k-rows}
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % GM_STAGES;
                mbar_wait(&empty[s], ((kt / GM_STAGES) & 1u) ^ 1u);
                unsigned char* sb = tiles + (size_t)s * STAGE + GM_A_BYTES;
                mbar_expect_tx(&full[s], TN * GM_TK * 4);
                for (int c = 0; c < TN / 32; ++c) tma_load_2d(sb + c * 4096, &mapW, &full[s], (int)c0 + 32 * c, kt * GM_TK);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ===== MMA issuer: A K-major (generated), B MN-major (W as stored)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(GM_TM >> 4) << 24);
            for (int kt = 0; kt < nkt; ++kt) {
                const int s = kt % GM_STAGES;
                mbar_wait(&full[s], (kt / GM_STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t sa = smem_u32(tiles + (size_t)s * STAGE), sb = sa + GM_A_BYTES;
#pragma unroll
                for (int kk = 0; kk < GM_TK / 8; ++kk)
                    umma_tf32(tmem, make_desc(sa + kk * 32, 16, 1024, 2), make_desc(sb + kk * 1024, 4096, 512, 1), idesc,
                              (kt == 0 && kk == 0) ? 0u : 1u);
                umma_commit(&empty[s]);
            }
            umma_commit(acc_full);
        }
    } else if (warp == 2 + GM_GEN_WARPS) {
        // ===== coordinate loader: the scaled x2 coordinates (and -|u2|^2 / 2) of every k-tile go through shared memory.
        // Generator threads must not have global loads in flight when they execute fence.proxy.async (it waits for the
        // thread's outstanding memory operations: with the coordinates prefetched by the generators themselves a k-tile
        // took 1650 cycles on B200, three times the MMA time).
        const float cs = RBF1 ? 1.2011224087864498f : 1.f;
        for (int kt = 0; kt < nkt; ++kt) {
            const int s = kt % GM_STAGES;
            const int64_t j = (int64_t)kt * GM_TK + lane;
            float b[TD];
#pragma unroll
            for (int k = 0; k < TD; ++k) b[k] = (j < p.n2) ? p.x2[j * p.ldx2 + sp.col[k]] * (sv.inv_ls[k] * cs) : 0.f;
            float h2 = 0.f;
#pragma unroll
            for (int k = 0; k < TD; ++k) h2 = fmaf(b[k], b[k], h2);
            h2 = (j < p.n2) ? -0.5f * h2 : -1e30f;
            if (SLOT) b[sp.ndims[0]] = 1.f;          // meets h1 in the same slot of the x1 side
            mbar_wait(&empty[s], ((kt / GM_STAGES) & 1u) ^ 1u);
            float* dst = xc + ((size_t)s * GM_TK + lane) * XC;
#pragma unroll
            for (int k = 0; k < TD; ++k) dst[k] = b[k];
            dst[TD] = h2;
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&cready[s])) : "memory");
        }
    } else {
        // ===== generators: lane = Gram column inside the k-tile, gw = row phase (rows gw, gw + 16, ...)
        const int gw = warp - 2;
        const float cs = RBF1 ? 1.2011224087864498f : 1.f;    // sqrt(log2 e)
        for (int e = threadIdx.x - 64; e < GM_TM * TD; e += 32 * GM_GEN_WARPS) {
            const int r = e / TD, k = e % TD;
            xs1[r * TD + k] = (i0 + r < p.n1) ? p.x1[(i0 + r) * p.ldx1 + sp.col[k]] * (sv.inv_ls[k] * cs) : 0.f;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * GM_GEN_WARPS) : "memory");
        if (RBF1) {
            const float lg_os = log2f(sv.os[0]);
            for (int r = threadIdx.x - 64; r < GM_TM; r += 32 * GM_GEN_WARPS) {
                float h = 0.f;
#pragma unroll
                for (int k = 0; k < TD; ++k) h = fmaf(xs1[r * TD + k], xs1[r * TD + k], h);
                h = (i0 + r < p.n1) ? lg_os - 0.5f * h : -1e30f;       // rows past the end: 2^(-huge) = 0
                hs1[r] = h;
                if (SLOT) xs1[r * TD + sp.ndims[0]] = h;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(32 * GM_GEN_WARPS) : "memory");
        }
        // byte offset of (row gw + 16 i, column lane) in the canonical operand tile: (gw + 16 i) & 7 == gw & 7
        const uint32_t st_off = (uint32_t)(gw * 128 + ((((lane >> 2) ^ (gw & 7)) << 4) | ((lane & 3) << 2)));
        for (int kt = 0; kt < nkt; ++kt) {
            const int s = kt % GM_STAGES;
            const int64_t j = (int64_t)kt * GM_TK + lane;
            mbar_wait(&cready[s], (kt / GM_STAGES) & 1u);
            float b[TD];
            const float* src = xc + ((size_t)s * GM_TK + lane) * XC;
            if constexpr (TD % 4 == 0) {
#pragma unroll
                for (int k = 0; k < TD; k += 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(src + k);
                    b[k] = t4.x; b[k + 1] = t4.y; b[k + 2] = t4.z; b[k + 3] = t4.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < TD; ++k) b[k] = src[k];
            }
            const float h2 = src[TD];
            unsigned char* sa = tiles + (size_t)s * STAGE + st_off;   // (cready[s] was signalled after empty[s]: the stage is free)
#pragma unroll
            for (int i = 0; i < GM_TM / GM_GEN_WARPS; ++i) {
                const int r = gw + GM_GEN_WARPS * i;
                float a[TD];
                if constexpr (TD % 4 == 0) {
#pragma unroll
                    for (int k = 0; k < TD; k += 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(xs1 + r * TD + k);
                        a[k] = t4.x; a[k + 1] = t4.y; a[k + 2] = t4.z; a[k + 3] = t4.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < TD; ++k) a[k] = xs1[r * TD + k];
                }
                uint32_t bits;
                if (RBF1) {
                    float arg = SLOT ? h2 : hs1[r] + h2;
#pragma unroll
                    for (int k = 0; k < TD; ++k) arg = fmaf(a[k], b[k], arg);
                    // round to nearest TF32 (ties away): positive finite values, so an integer add and a mask do it
                    bits = (__float_as_uint(ex2_ftz(arg)) + 0x1000u) & 0xFFFFE000u;
                } else {
                    float v = eval_pair<float, NC, DPC>(sp, sv.os, a, b);
                    if (j >= p.n2 || i0 + r >= p.n1) v = 0.f;
                    bits = to_tf32(v);
                }
                *reinterpret_cast<uint32_t*>(sa + i * (GM_GEN_WARPS * 128)) = bits;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[s])) : "memory");
        }
        // ===== epilogue: TMEM lane quadrant = warp % 4, column quarter = gw / 4
        mbar_wait(acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int q = warp & 3, h = gw >> 2;   // h = 0 .. 3: a quarter of the TN columns each
        const int64_t row = i0 + q * 32 + lane;
        const float dd = p.same ? p.diag_const + (p.diag_dev ? p.diag_dev[0] : 0.f) : 0.f;
        const bool vec = ((p.ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0) &&
                         ((p.ldw & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.W) & 15) == 0);
        for (int cc = h * (TN / 4); cc < (h + 1) * (TN / 4); cc += 32) {
            uint32_t r[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)cc));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int64_t col = c0 + cc;
            if (row < p.n1 && col < p.m) {
                float* dst = p.out + row * p.ldo + col;
                const float* wrow = p.W + row * p.ldw + col;   // same: n1 == n2, W row `row` exists
                if (vec && col + 32 <= p.m) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 o = make_float4(__uint_as_float(r[4 * c]), __uint_as_float(r[4 * c + 1]),
                                               __uint_as_float(r[4 * c + 2]), __uint_as_float(r[4 * c + 3]));
                        if (dd != 0.f) {
                            const float4 w = *reinterpret_cast<const float4*>(wrow + 4 * c);
                            o.x = fmaf(dd, w.x, o.x); o.y = fmaf(dd, w.y, o.y); o.z = fmaf(dd, w.z, o.z); o.w = fmaf(dd, w.w, o.w);
                        }
                        *reinterpret_cast<float4*>(dst + 4 * c) = o;
                    }
                } else {
                    for (int c = 0; c < 32 && col + c < p.m; ++c) {
                        float o = __uint_as_float(r[c]);
                        if (dd != 0.f) o = fmaf(dd, wrow[c], o);
                        dst[c] = o;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TN));
    }
}

template <int NC, int DPC, int TN, int MODE>
int launch_gm1(const DevSpec& sp, const CUtensorMap& mapW, const GmParams& p, cudaStream_t st) {
    constexpr size_t smem = gm_smem<TN, NC * DPC>();
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(gram_mm_tc_kernel<NC, DPC, TN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return DCGP_CUDA_ERR(e);
        attr = true;
    }
    dim3 grid((unsigned)((p.m + TN - 1) / TN), (unsigned)((p.n1 + GM_TM - 1) / GM_TM));
    gram_mm_tc_kernel<NC, DPC, TN, MODE><<<grid, GM_THREADS, smem, st>>>(sp, mapW, p);
    DCGP_CHECK_LAUNCH();
    return 0;
}
template <int NC, int DPC, int TN>
int launch_gm(const DevSpec& sp, const CUtensorMap& mapW, const GmParams& p, cudaStream_t st) {
    if constexpr (NC == 1) {
        if (sp.kind[0] == DCGP_KIND_RBF)
            return sp.ndims[0] < DPC ? launch_gm1<NC, DPC, TN, 2>(sp, mapW, p, st) : launch_gm1<NC, DPC, TN, 1>(sp, mapW, p, st);
    }
    return launch_gm1<NC, DPC, TN, 0>(sp, mapW, p, st);
}

}  // namespace

// Returns 0 when the fused kernel took the product, 1 when the shape / layout is not its (caller falls back to
// panel generation + GEMM), < 0 on errors.
int dcgp_gram_matmul_fused_f32(const DevSpec& sp, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2,
                               int64_t ldx2, const void* diag_dev, double diag_const, const void* W, int64_t m, int64_t ldw,
                               void* out, int64_t ldo, int same, cudaStream_t st) {
    if (n1 < 128 || n2 < 64 || m < 32) return 1;                       // not worth a 128-row tensor-core tile
    // every 256-column block of the result regenerates the Gram tiles: up to two blocks the fused kernel matches or beats
    // panel generation + GEMM (B200, profiles/r02_gram_mm.txt: N = 32768, m = 328: 2.75 vs 2.97 ms; N = 16384: 0.74 vs 0.72),
    // at four (m = 1024, d = 5) it loses (0.48 vs 0.36 ms) - wider products keep the panel path
    if (m > 512) return 1;
    if ((ldw & 3) || (reinterpret_cast<uintptr_t>(W) & 15)) return 1;  // TMA needs 16-byte aligned rows of W
    if (n2 > 0x7fffffff || m > 0x7fffffff) return 1;
    if ((int64_t)n1 * n2 < ((int64_t)1 << 22)) return 1;
    CUtensorMap mapW;
    if (dcgp_tmap_f32(&mapW, W, n2, m, ldw, 32, GM_TK, 1)) return 1;
    GmParams p;
    p.x1 = (const float*)x1; p.x2 = (const float*)x2; p.W = (const float*)W; p.out = (float*)out;
    p.diag_dev = (const float*)diag_dev;
    p.n1 = n1; p.n2 = n2; p.m = m; p.ldx1 = ldx1; p.ldx2 = ldx2; p.ldw = ldw; p.ldo = ldo;
    p.diag_const = (float)diag_const;
    p.same = same;
    int rc = 1;
    if (m > 128) {
#define CALL(NCV, DPCV) rc = launch_gm<NCV, DPCV, 256>(sp, mapW, p, st);
        DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    } else {
#define CALL(NCV, DPCV) rc = launch_gm<NCV, DPCV, 128>(sp, mapW, p, st);
        DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    }
    return rc;
}
