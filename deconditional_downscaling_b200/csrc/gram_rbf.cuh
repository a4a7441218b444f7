// Single-component RBF fast path of the dense Gram build and of its hyper-parameter adjoint (included by gram.cu).
//
// The generic kernels (gram.cu) evaluate an additive kernel whose component kinds are run-time data.  ncu on B200
// (profiles/r02_gram_kernels_ncu.txt) showed them bound by instruction issue (FP32: 80 % issue-active, ~43 issued
// instructions per entry) or by the FP64 pipe (~27 FP64 instructions per entry, 17 of them the library exp), at 0.26 -
// 0.35 of the HBM rate.  k = os * exp(-|x - y|_l^2 / 2) - the individuals kernel of every model of the reference
// (src/models/cmp.py:60,65 K = self.individuals_kernel(x); experiments/*/models/*.py build it as Scale(RBF)) - needs far
// less:
//   * coordinates are pre-scaled by 1 / (l sqrt 2) (times sqrt(log2 e) in FP32) and the accumulator starts at
//     log(os) (log2 in FP32), so one entry is D subtractions, D fused multiply-adds and one exponential;
//   * FP32: ex2.approx (MUFU) - the accuracy class of __expf, which the generic kernel uses;
//   * FP64: exp(x) = 2^(k / 64) * p(r), k = round(64 x / ln 2) by the magic-number add (no FP64 <-> integer
//     conversion), r = x - k ln 2 / 64 in two fused steps, 2^(j / 64) from a 64-entry shared-memory table, p of degree 5
//     (|r| <= ln 2 / 128, truncation 3.5e-17), the power of two added into the exponent field: 10 FP64 instructions,
//     <= 1 ulp of the table entry plus the rounding of p.  Round 2's first attempt (4-entry table, degree 9) did not
//     beat the library; this one removes 7 more DFMAs.
#pragma once

namespace rbf {

__constant__ double kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,
};

// exponential of the scaled accumulator: FP32 takes a base-2 argument, FP64 a natural one (see Scale)
template <typename T> struct Scale;
template <> struct Scale<float> {
    static __device__ __forceinline__ float coord() { return 0.84932180028801904f; }   // sqrt(log2(e) / 2)
    static __device__ __forceinline__ float base(float os) { return __log2f(os); }
    static __device__ __forceinline__ float inv_sq() { return 1.3862943611198906f; }   // 1 / coord()^2 = 2 ln 2
};
template <> struct Scale<double> {
    static __device__ __forceinline__ double coord() { return 0.70710678118654752; }
    static __device__ __forceinline__ double base(double os) { return log(os); }
    static __device__ __forceinline__ double inv_sq() { return 2.0; }
};

__device__ __forceinline__ float exp_acc(float x, const double*) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ double exp_acc(double x, const double* tab) {
    constexpr double kMagic = 6755399441055744.0;               // 1.5 * 2^52: the low word of x + kMagic is rint(x)
    const double kd = fma(x, 0x1.71547652b82fep+6, kMagic);     // 64 / ln 2
    const int ki = __double2loint(kd);
    const double kf = kd - kMagic;
    double r = fma(kf, -0x1.62e42fee00000p-7, x);               // ln 2 / 64, upper 32 bits
    r = fma(kf, -0x1.a39ef35793c76p-39, r);                     //            remainder
    double p = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = tab[ki & 63] * p;                          // in [1, 2 (1 + 2^-52)]
    const int hi = __double2hiint(v) + ((ki >> 6) << 20);       // times 2^(k div 64)
    // below -708 the result would leave the normal range (and far below, k would leave the low word): flush to zero -
    // a Gram entry under 1e-307 of the outputscale
    return x < -708.0 ? 0.0 : __hiloint2double(hi, __double2loint(v));
}

template <typename T>
__device__ __forceinline__ void load_tab(double* tab) {
    if constexpr (sizeof(T) == 8) {
        if (threadIdx.x < 64) tab[threadIdx.x] = kExp2Tab[threadIdx.x];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: same tiling as gram_fwd_kernel (a CTA owns 32 x VEC columns, their scaled coordinates in registers, and
// walks row tiles of 64 rows whose coordinates pass through a double-buffered shared tile)
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256) gram_rbf_fwd_kernel(const __grid_constant__ DevSpec sp, const T* __restrict__ x1,
                                                           int64_t n1, int64_t ldx1, const T* __restrict__ x2, int64_t n2,
                                                           int64_t ldx2, T* __restrict__ K, int64_t ldk,
                                                           const T* __restrict__ diag_dev, T diag_const, int same,
                                                           int vec_ok, int row_tiles, int rna) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int TCOLS = 32 * VEC;
    constexpr int TROWS = 64;
    __shared__ T sc[D];
    __shared__ T s1[2][TROWS][D];
    __shared__ double tab[sizeof(T) == 8 ? 64 : 1];
    if (threadIdx.x < D) sc[threadIdx.x] = Scale<T>::coord() / ((const T*)sp.ls[0])[threadIdx.x];
    load_tab<T>(tab);
    const T base = Scale<T>::base(((const T*)sp.os[0])[0]);
    __syncthreads();
    const int64_t rowbase = (int64_t)blockIdx.y * TROWS * row_tiles;
    const int64_t col0 = (int64_t)blockIdx.x * TCOLS;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    T b[VEC][D];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int64_t c = col0 + tx * VEC + v;
#pragma unroll
        for (int k = 0; k < D; ++k) b[v][k] = (c < n2) ? x2[c * ldx2 + sp.col[k]] * sc[k] : T(0);
    }
    const T dadd = same ? diag_const + (diag_dev ? diag_dev[0] : T(0)) : T(0);
    for (int t = 0; t < row_tiles; ++t) {
        const int64_t row0 = rowbase + (int64_t)t * TROWS;
        if (row0 >= n1) break;                       // uniform over the CTA
        T (*s1t)[D] = s1[t & 1];
        for (int e = threadIdx.x; e < TROWS * D; e += 256) {
            const int r = e / D, k = e % D;
            s1t[r][k] = (row0 + r < n1) ? x1[(row0 + r) * ldx1 + sp.col[k]] * sc[k] : T(0);
        }
        __syncthreads();
        // the diagonal crosses this tile's rows inside this CTA's columns? (uniform over the CTA)
        const bool on_diag = same && row0 < col0 + TCOLS && row0 + TROWS > col0;
#pragma unroll 2
        for (int rr = 0; rr < TROWS / 8; ++rr) {
            const int r = rr * 8 + ty;
            const int64_t row = row0 + r;
            if (row >= n1) break;
            T a[D];
#pragma unroll
            for (int k = 0; k < D; ++k) a[k] = s1t[r][k];
            T out[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                T acc = base;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const T d = a[k] - b[v][k];
                    acc = fma(-d, d, acc);
                }
                out[v] = exp_acc(acc, tab);
            }
            const int64_t c = col0 + tx * VEC;
            if (on_diag) {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (row == c + v) out[v] += dadd;
            }
            if constexpr (sizeof(T) == 4) {
                if (rna) {   // operand of a single-pass TF32 product: round to nearest (see gram_fwd_kernel)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) out[v] = __uint_as_float(to_tf32(out[v]));
                }
            }
            T* dst = K + row * ldk + c;
            if (vec_ok && c + VEC <= n2) {
                if constexpr (sizeof(T) == 8) {
                    *reinterpret_cast<double2*>(dst) = make_double2(out[0], out[1]);
                } else {
                    *reinterpret_cast<float4*>(dst) = make_float4(out[0], out[1], out[VEC - 2], out[VEC - 1]);
                }
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (c + v < n2) dst[v] = out[v];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// hyper-parameter adjoint for a dense upstream gradient G (n1 x n2):
//   d/d os = sum g k / os,     d/d l_k = sum g k u_k^2 / l_k      (u = (x - y) / l)
// A CTA owns a block of CB columns (scaled coordinates in shared memory, [k][column]) and walks row groups of 8 warps x
// RW rows down it; a lane handles VEC consecutive columns per step - one 16-byte load of G per row, one 16-byte shared
// load per coordinate - and keeps the D + 1 sums in registers for the whole walk.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int D, int RW>
__global__ void __launch_bounds__(256) gram_rbf_bwd_kernel(const __grid_constant__ DevSpec sp, const T* __restrict__ x1,
                                                           int64_t n1, int64_t ldx1, const T* __restrict__ x2, int64_t n2,
                                                           int64_t ldx2, const T* __restrict__ G, int64_t ldg,
                                                           T* __restrict__ dls, T* __restrict__ dos, int cblk, int vec) {
    constexpr int VEC = 16 / sizeof(T);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s2 = reinterpret_cast<T*>(smem_raw);  // [D][cblk]
    __shared__ T sc[D];
    __shared__ double tab[sizeof(T) == 8 ? 64 : 1];
    if (threadIdx.x < D) sc[threadIdx.x] = Scale<T>::coord() / ((const T*)sp.ls[0])[threadIdx.x];
    load_tab<T>(tab);
    const T os = ((const T*)sp.os[0])[0];
    const T base = Scale<T>::base(os);
    __syncthreads();
    const int64_t col0 = (int64_t)blockIdx.x * cblk;
    const int ncols = (int)min((int64_t)cblk, n2 - col0);
    for (int e = threadIdx.x; e < D * cblk; e += 256) {
        const int k = e / cblk, c = e % cblk;
        s2[e] = (c < ncols) ? x2[(col0 + c) * ldx2 + sp.col[k]] * sc[k] : T(0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T gl[D], go = T(0);
#pragma unroll
    for (int k = 0; k < D; ++k) gl[k] = T(0);
    const int64_t rstride = (int64_t)gridDim.y * 8 * RW;
    const int64_t rfirst = ((int64_t)blockIdx.y * 8 + warp) * RW;
    if (vec) {
        // One flat loop over (row group, column step) with the NEXT step's 16-byte loads of G issued before the current
        // step's arithmetic: ncu showed the plain loop waiting on those loads (long-scoreboard stalls 6.3 per issue, 41 %
        // issue-active at 16 resident warps per SM).  ncols % VEC == 0 (host); lanes never meet inside the loop.
        auto load_g = [&](T (&g)[RW][VEC], int64_t row0, int cc) {
#pragma unroll
            for (int r = 0; r < RW; ++r) {
                if (row0 + r < n1) {
                    if constexpr (VEC == 4) {
                        const float4 t = __ldcs(reinterpret_cast<const float4*>(G + (row0 + r) * ldg + col0 + cc));
                        g[r][0] = t.x; g[r][1] = t.y; g[r][VEC - 2] = t.z; g[r][VEC - 1] = t.w;
                    } else {
                        const double2 t = __ldcs(reinterpret_cast<const double2*>(G + (row0 + r) * ldg + col0 + cc));
                        g[r][0] = t.x; g[r][VEC - 1] = t.y;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) g[r][q] = T(0);
                }
            }
        };
        int64_t row0 = rfirst;
        int cc = lane * VEC;
        bool have = row0 < n1 && cc < ncols, fresh = true;
        T a[RW][D], gn[RW][VEC];
        if (have) load_g(gn, row0, cc);
        while (have) {
            if (fresh) {
#pragma unroll
                for (int r = 0; r < RW; ++r)
#pragma unroll
                    for (int k = 0; k < D; ++k)
                        a[r][k] = (row0 + r < n1) ? x1[(row0 + r) * ldx1 + sp.col[k]] * sc[k] : T(0);
            }
            T gv[RW][VEC], bq[D][VEC];
#pragma unroll
            for (int r = 0; r < RW; ++r)
#pragma unroll
                for (int q = 0; q < VEC; ++q) gv[r][q] = gn[r][q];
            int ccn = cc + 32 * VEC;
            int64_t rown = row0;
            fresh = ccn >= ncols;
            if (fresh) {
                ccn = lane * VEC;
                rown = row0 + rstride;
            }
            have = rown < n1;
            if (have) load_g(gn, rown, ccn);
#pragma unroll
            for (int k = 0; k < D; ++k) {
                if constexpr (VEC == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(s2 + k * cblk + cc);
                    bq[k][0] = t.x; bq[k][1] = t.y; bq[k][VEC - 2] = t.z; bq[k][VEC - 1] = t.w;
                } else {
                    const double2 t = *reinterpret_cast<const double2*>(s2 + k * cblk + cc);
                    bq[k][0] = t.x; bq[k][VEC - 1] = t.y;
                }
            }
#pragma unroll
            for (int q = 0; q < VEC; ++q)
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    T d[D], acc = base;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        d[k] = a[r][k] - bq[k][q];
                        acc = fma(-d[k], d[k], acc);
                    }
                    const T gw = gv[r][q] * exp_acc(acc, tab);
                    go += gw;
#pragma unroll
                    for (int k = 0; k < D; ++k) gl[k] = fma(gw * d[k], d[k], gl[k]);
                }
            cc = ccn;
            row0 = rown;
        }
    } else {
        for (int64_t row0 = rfirst; row0 < n1; row0 += rstride) {
            T a[RW][D];
#pragma unroll
            for (int r = 0; r < RW; ++r)
#pragma unroll
                for (int k = 0; k < D; ++k) a[r][k] = (row0 + r < n1) ? x1[(row0 + r) * ldx1 + sp.col[k]] * sc[k] : T(0);
            for (int cc = lane; cc < ncols; cc += 32) {
                T bq[D];
#pragma unroll
                for (int k = 0; k < D; ++k) bq[k] = s2[k * cblk + cc];
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    if (row0 + r >= n1) continue;
                    T d[D], acc = base;
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        d[k] = a[r][k] - bq[k];
                        acc = fma(-d[k], d[k], acc);
                    }
                    const T gw = G[(row0 + r) * ldg + col0 + cc] * exp_acc(acc, tab);
                    go += gw;
#pragma unroll
                    for (int k = 0; k < D; ++k) gl[k] = fma(gw * d[k], d[k], gl[k]);
                }
            }
        }
    }
    // block reduction, one atomic per value and CTA
    __shared__ T red[8][D + 1];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const T v = warp_sum(gl[k]);
        if (lane == 0) red[warp][k] = v;
    }
    {
        const T v = warp_sum(go);
        if (lane == 0) red[warp][D] = v;
    }
    __syncthreads();
    if (threadIdx.x <= D) {
        T v = T(0);
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        if (threadIdx.x < D) {
            // the sums hold g k u'^2 with u' = u * coord(): back to u^2, then d u^2 / d l = -2 u^2 / l and the -1/2 of
            // the exponent leave  + g k u^2 / l
            if (dls) atomicAdd(&dls[threadIdx.x], v * Scale<T>::inv_sq() / ((const T*)sp.ls[0])[threadIdx.x]);
        } else if (dos) {
            atomicAdd(&dos[0], v / os);
        }
    }
}

inline bool enabled() {
    static const bool on = !(getenv("DCGP_GRAM_RBF_FAST") && atoi(getenv("DCGP_GRAM_RBF_FAST")) == 0);
    return on;
}
inline bool applies(const DevSpec& sp) {
    return enabled() && sp.n_comp == 1 && sp.kind[0] == DCGP_KIND_RBF && sp.ndims[0] >= 1 && sp.ndims[0] <= 8;
}

#define DCGP_RBF_DISPATCH_D(d, CALL) \
    switch (d) {                     \
        case 1: CALL(1); break;      \
        case 2: CALL(2); break;      \
        case 3: CALL(3); break;      \
        case 4: CALL(4); break;      \
        case 5: CALL(5); break;      \
        case 6: CALL(6); break;      \
        case 7: CALL(7); break;      \
        default: CALL(8); break;     \
    }

template <typename T>
int launch_fwd(const DevSpec& sp, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2, int64_t ldx2,
               void* K, int64_t ldk, const void* diag_dev, double diag_const, int same, dim3 grid, int vec_ok, int rt,
               int rna, cudaStream_t st) {
#define CALL(DV)                                                                                                    \
    gram_rbf_fwd_kernel<T, DV><<<grid, 256, 0, st>>>(sp, (const T*)x1, n1, ldx1, (const T*)x2, n2, ldx2, (T*)K, ldk, \
                                                     (const T*)diag_dev, (T)diag_const, same, vec_ok, rt, rna)
    DCGP_RBF_DISPATCH_D(sp.ndims[0], CALL)
#undef CALL
    return 0;
}

template <typename T>
int launch_bwd(const DevSpec& sp, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2, int64_t ldx2,
               const void* G, int64_t ldg, void* dls, void* dos, cudaStream_t st) {
    constexpr int VEC = 16 / sizeof(T);
    constexpr int RWMAX = 4;   // rows per warp: 4, 2 in FP64 beyond three dimensions (registers: two CTAs per SM)
    // column block: up to 2048 columns (16 steps of a row group between two loads of row coordinates), 40 KB of shared
    // memory at most
    const int d = sp.ndims[0];
    int cblk = 2048;
    while (cblk > 128 && (cblk / 2 >= n2 || (size_t)d * cblk * sizeof(T) > 40 * 1024)) cblk >>= 1;
    const size_t smem = (size_t)d * cblk * sizeof(T);
    const int64_t gx = (n2 + cblk - 1) / cblk, rows_max = (n1 + 8 * RWMAX - 1) / (8 * RWMAX);
    int64_t gy = (4 * (int64_t)num_sms() + gx - 1) / gx;  // column blocks x row walkers: about four CTAs per SM
    if (gy > rows_max) gy = rows_max;
    if (gy < 1) gy = 1;
    dim3 grid((unsigned)gx, (unsigned)gy);
    const int vec = ((ldg % VEC) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0) && (n2 % VEC) == 0;
#define CALL(DV)                                                                                                       \
    gram_rbf_bwd_kernel<T, DV, (sizeof(T) == 8 && DV > 3) ? 2 : 4><<<grid, 256, smem, st>>>(sp, (const T*)x1, n1, ldx1, (const T*)x2, n2, ldx2,        \
                                                            (const T*)G, ldg, (T*)dls, (T*)dos, cblk, vec)
    DCGP_RBF_DISPATCH_D(d, CALL)
#undef CALL
    return 0;
}

}  // namespace rbf
