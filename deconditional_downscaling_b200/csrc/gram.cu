// Gram assembly kernels (RBF / Matern, ARD lengthscales, additive over active_dims).
//
// Forward: one pass, HBM-write bound.  Each CTA produces a 64 x (32*VEC) tile; every thread
// keeps the scaled coordinates of its VEC columns in registers, walks 8 rows whose scaled
// coordinates are broadcast from shared memory, and issues one 16-byte store per row
// (a warp writes 512 contiguous bytes).  The squared distance is formed from coordinate
// differences, not from the |x|^2+|y|^2-2xy expansion: with an inner dimension of d<=8 the
// tensor pipe could not be fed anyway (the kernel is bound by the N1*N2 output bytes and, in
// FP64, by exp), and the difference form has no cancellation, so the diagonal is exact and
// K stays positive definite in FP32/TF32 builds.
//
// Backward (dense upstream gradient): a warp owns RW rows and strides the columns of a
// column chunk; hyper-parameter and input-coordinate adjoints are accumulated in registers,
// reduced with warp shuffles and committed with one atomicAdd per warp and output value.
#include "common.cuh"
#include "mma.cuh"

namespace {

// commit warp-reduced hyper-parameter adjoints: d/dl_k = sum g*os*w*u_k^2 / l_k
template <typename T, int NC, int DPC>
__device__ __forceinline__ void commit_theta(const DevSpec& sp, const SpecVals<T>& sv, T (&gl)[NC * DPC], T (&go)[NC],
                                             T* dls, T* dos, int lane) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int j = 0; j < DPC; ++j) {
            T v = warp_sum(gl[c * DPC + j]);
            if (lane == 0 && dls && c < sp.n_comp && j < sp.ndims[c])
                atomicAdd(&dls[c * DCGP_MAX_DIMS + j], v * sv.inv_ls[c * DPC + j]);
        }
        T v = warp_sum(go[c]);
        if (lane == 0 && dos && c < sp.n_comp) atomicAdd(&dos[c], v);
    }
}

template <typename T, typename TO, int NC, int DPC>
__global__ void __launch_bounds__(256) gram_fwd_kernel(const __grid_constant__ DevSpec sp, const T* __restrict__ x1,
                                                       int64_t n1, int64_t ldx1, const T* __restrict__ x2, int64_t n2,
                                                       int64_t ldx2, TO* __restrict__ K, int64_t ldk,
                                                       const T* __restrict__ diag_dev, T diag_const, int same,
                                                       int vec_ok, int row_tiles, int rna) {
    // One CTA owns a 128-column (FP32; 64 in FP64) block and walks `row_tiles` 64-row tiles down it: the scaled x2
    // coordinates of its columns stay in registers and the per-CTA set-up (kernel spec, strided x2 gather) is paid once
    // per row_tiles x 64 x TCOLS outputs instead of once per 64 x TCOLS.
    constexpr int TD = NC * DPC;
    constexpr int VEC = 16 / sizeof(TO);
    constexpr int TCOLS = 32 * VEC;
    constexpr int TROWS = 64;
    __shared__ SpecVals<T> sv;
    __shared__ T s1[2][TROWS][TD];
    load_spec_vals<T>(sp, &sv);
    __syncthreads();
    const int64_t rowbase = (int64_t)blockIdx.y * TROWS * row_tiles;
    const int64_t col0 = (int64_t)blockIdx.x * TCOLS;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    T b[VEC][TD];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        int64_t c = col0 + tx * VEC + v;
#pragma unroll
        for (int k = 0; k < TD; ++k) b[v][k] = (c < n2) ? x2[c * ldx2 + sp.col[k]] * sv.inv_ls[k] : T(0);
    }
    T dadd = T(0);
    if (same) dadd = diag_const + (diag_dev ? diag_dev[0] : T(0));
    for (int t = 0; t < row_tiles; ++t) {
    const int64_t row0 = rowbase + (int64_t)t * TROWS;
    if (row0 >= n1) break;                       // uniform over the CTA
    // double-buffered row coordinates: one barrier per tile (a thread refills buffer t & 1 only after the barrier of
    // tile t - 1, which every thread passes after it has finished reading that buffer for tile t - 2)
    T (*s1t)[TD] = s1[t & 1];
    for (int e = threadIdx.x; e < TROWS * TD; e += 256) {
        int r = e / TD, k = e % TD;
        s1t[r][k] = (row0 + r < n1) ? x1[(row0 + r) * ldx1 + sp.col[k]] * sv.inv_ls[k] : T(0);
    }
    __syncthreads();
#pragma unroll 2
    for (int rr = 0; rr < TROWS / 8; ++rr) {
        const int r = rr * 8 + ty;
        const int64_t row = row0 + r;
        if (row >= n1) break;
        T a[TD];
#pragma unroll
        for (int k = 0; k < TD; ++k) a[k] = s1t[r][k];
        TO out[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            T val = eval_pair<T, NC, DPC>(sp, sv.os, a, b[v]);
            if (same && row == col0 + tx * VEC + v) val += dadd;
            if constexpr (sizeof(T) == 4 && sizeof(TO) == 4) {
                // operand of a single-pass TF32 product: round to nearest here, so that the truncating operand read of
                // tcgen05.mma sees an exactly representable value (unbiased instead of -3.5e-4 relative on average)
                if (rna) val = __uint_as_float(to_tf32(val));
            }
            out[v] = (TO)val;
        }
        const int64_t c = col0 + tx * VEC;
        TO* dst = K + row * ldk + c;
        if (vec_ok && c + VEC <= n2) {
            if constexpr (sizeof(TO) == 8) {
                *reinterpret_cast<double2*>(dst) = make_double2((double)out[0], (double)out[1]);
            } else {
                *reinterpret_cast<float4*>(dst) =
                    make_float4((float)out[0], (float)out[1], (float)out[2], (float)out[3]);
            }
        } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                if (c + v < n2) dst[v] = out[v];
        }
    }
    }
}

// ---------------------------------------------------------------------------------------
// backward, dense upstream gradient G (n1 x n2)
// ---------------------------------------------------------------------------------------
template <typename T, int NC, int DPC, int RW, bool DX>
__global__ void __launch_bounds__(256) gram_bwd_kernel(const __grid_constant__ DevSpec sp, const T* __restrict__ x1,
                                                       int64_t n1, int64_t ldx1, const T* __restrict__ x2, int64_t n2,
                                                       int64_t ldx2, const T* __restrict__ G, int64_t ldg,
                                                       T* __restrict__ dls, T* __restrict__ dos, T* __restrict__ dx1,
                                                       int64_t lddx1, int same, int cblk) {
    constexpr int TD = NC * DPC;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* s2 = reinterpret_cast<T*>(smem_raw);  // [TD][cblk]
    __shared__ SpecVals<T> sv;
    load_spec_vals<T>(sp, &sv);
    __syncthreads();
    const int64_t col0 = (int64_t)blockIdx.x * cblk;
    const int ncols = (int)min((int64_t)cblk, n2 - col0);
    for (int e = threadIdx.x; e < TD * cblk; e += 256) {
        int k = e / cblk, c = e % cblk;
        s2[e] = (c < ncols) ? x2[(col0 + c) * ldx2 + sp.col[k]] * sv.inv_ls[k] : T(0);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T gl[TD], go[NC];
#pragma unroll
    for (int k = 0; k < TD; ++k) gl[k] = T(0);
#pragma unroll
    for (int c = 0; c < NC; ++c) go[c] = T(0);
    // a CTA walks its column block down the rows (grid.y is a few waves, not n1 / 8 RW): the hyper-parameter
    // adjoints stay in registers across the walk and reach memory as ONE atomic per value and CTA - with one
    // per warp the ~10 shared addresses serialised hundreds of thousands of atomics
    for (int64_t row0 = ((int64_t)blockIdx.y * 8 + warp) * RW; row0 < n1; row0 += (int64_t)gridDim.y * 8 * RW) {
        T a[RW][TD], gx[RW][TD];
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int k = 0; k < TD; ++k) {
                a[r][k] = (row0 + r < n1) ? x1[(row0 + r) * ldx1 + sp.col[k]] * sv.inv_ls[k] : T(0);
                gx[r][k] = T(0);
            }
        constexpr int VEC = 16 / sizeof(T);  // columns per lane and step: one 16-byte load of G per row
        constexpr bool FAST = sizeof(T) == 4;
        const bool vec = ((ldg % VEC) == 0) && ((reinterpret_cast<uintptr_t>(G) & 15) == 0) && (ncols % VEC) == 0;
        if (vec) {
            for (int cc = lane * VEC; cc < ncols; cc += 32 * VEC) {
                T gv[RW][VEC];
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    if (row0 + r < n1) {
                        if (VEC == 4) {
                            const float4 t = *reinterpret_cast<const float4*>(G + (row0 + r) * ldg + col0 + cc);
                            gv[r][0] = t.x; gv[r][1] = t.y; gv[r][VEC - 2] = t.z; gv[r][VEC - 1] = t.w;
                        } else {
                            const double2 t = *reinterpret_cast<const double2*>(G + (row0 + r) * ldg + col0 + cc);
                            gv[r][0] = t.x; gv[r][VEC - 1] = t.y;
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) gv[r][q] = T(0);
                    }
                }
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    T b[TD];
#pragma unroll
                    for (int k = 0; k < TD; ++k) b[k] = s2[k * cblk + cc + q];
#pragma unroll
                    for (int r = 0; r < RW; ++r) {
                        const T g = gv[r][q];
                        T gs = g;
                        if (DX && same && row0 + r < n1) gs += G[(col0 + cc + q) * ldg + row0 + r];
                        accum_pair<T, NC, DPC, DX, FAST>(sp, sv.os, a[r], b, g, gs, gl, go, gx[r]);
                    }
                }
            }
        } else {
            for (int cc = lane; cc < ncols; cc += 32) {
                const int64_t col = col0 + cc;
                T b[TD];
#pragma unroll
                for (int k = 0; k < TD; ++k) b[k] = s2[k * cblk + cc];
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    const int64_t row = row0 + r;
                    if (row < n1) {
                        const T g = G[row * ldg + col];
                        // d/dx of k(x, x) collects both argument slots: needs G + G^T - a strided (transposed)
                        // read, only paid when coordinate gradients are requested
                        const T gs = (same && DX) ? g + G[col * ldg + row] : g;
                        accum_pair<T, NC, DPC, DX, FAST>(sp, sv.os, a[r], b, g, gs, gl, go, gx[r]);
                    }
                }
            }
        }
        if (DX) {
#pragma unroll
            for (int r = 0; r < RW; ++r)
#pragma unroll
                for (int k = 0; k < TD; ++k) {
                    T v = warp_sum(gx[r][k]);
                    if (lane == 0 && row0 + r < n1 && (k / DPC) < sp.n_comp && (k % DPC) < sp.ndims[k / DPC])
                        atomicAdd(&dx1[(row0 + r) * lddx1 + sp.col[k]], v * sv.inv_ls[k]);
                }
        }
    }
    // block reduction of the hyper-parameter adjoints
    __shared__ T red[8][TD + NC];
#pragma unroll
    for (int k = 0; k < TD; ++k) {
        const T v = warp_sum(gl[k]);
        if (lane == 0) red[warp][k] = v;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const T v = warp_sum(go[c]);
        if (lane == 0) red[warp][TD + c] = v;
    }
    __syncthreads();
    if (threadIdx.x < TD + NC) {
        T v = T(0);
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
        if (threadIdx.x < TD) {
            const int c = threadIdx.x / DPC, j = threadIdx.x % DPC;
            if (dls && c < sp.n_comp && j < sp.ndims[c]) atomicAdd(&dls[c * DCGP_MAX_DIMS + j], v * sv.inv_ls[threadIdx.x]);
        } else {
            const int c = threadIdx.x - TD;
            if (dos && c < sp.n_comp) atomicAdd(&dos[c], v);
        }
    }
}

// ---------------------------------------------------------------------------------------
// bag-summed cross Gram: out[a, i] = sum_{j in bag a} k(z_i, x_j)   (n_bags x M)
// A CTA owns 128 rows of z (one per thread, coordinates in registers) and BCH consecutive
// bags; the bags' scaled coordinates stream through shared memory and are read as warp
// broadcasts.  K_Zx (M x N) is never written.
// ---------------------------------------------------------------------------------------
constexpr int kBagChunk = 16;
constexpr int kXTile = 256;

template <typename T, int NC, int DPC>
__global__ void __launch_bounds__(128) gram_bagsum_kernel(const __grid_constant__ DevSpec sp, const T* __restrict__ z,
                                                          int64_t M, int64_t ldz, const T* __restrict__ x, int64_t ldx,
                                                          const int64_t* __restrict__ offsets, int64_t n_bags,
                                                          T* __restrict__ out, int64_t ldo) {
    constexpr int TD = NC * DPC;
    __shared__ SpecVals<T> sv;
    __shared__ T sx[kXTile][TD];
    load_spec_vals<T>(sp, &sv);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    T a[TD];
#pragma unroll
    for (int k = 0; k < TD; ++k) a[k] = (i < M) ? z[i * ldz + sp.col[k]] * sv.inv_ls[k] : T(0);
    const int64_t bag0 = (int64_t)blockIdx.y * kBagChunk;
    const int64_t bag1 = min(bag0 + kBagChunk, n_bags);
    for (int64_t bag = bag0; bag < bag1; ++bag) {
        const int64_t j0 = offsets[bag], j1 = offsets[bag + 1];
        T sum = T(0);
        for (int64_t jt = j0; jt < j1; jt += kXTile) {
            const int nt = (int)min((int64_t)kXTile, j1 - jt);
            __syncthreads();
            for (int e = threadIdx.x; e < nt * TD; e += 128) {
                int r = e / TD, k = e % TD;
                sx[r][k] = x[(jt + r) * ldx + sp.col[k]] * sv.inv_ls[k];
            }
            __syncthreads();
            for (int r = 0; r < nt; ++r) {
                T b[TD];
#pragma unroll
                for (int k = 0; k < TD; ++k) b[k] = sx[r][k];
                sum += eval_pair<T, NC, DPC>(sp, sv.os, a, b);
            }
        }
        if (i < M) out[bag * ldo + i] = sum;
    }
}

template <typename T, int NC, int DPC>
__global__ void __launch_bounds__(128) gram_bagsum_bwd_kernel(const __grid_constant__ DevSpec sp,
                                                              const T* __restrict__ z, int64_t M, int64_t ldz,
                                                              const T* __restrict__ x, int64_t ldx,
                                                              const int64_t* __restrict__ offsets, int64_t n_bags,
                                                              const T* __restrict__ G, int64_t ldg, T* __restrict__ dls,
                                                              T* __restrict__ dos, T* __restrict__ dz, int64_t lddz) {
    constexpr int TD = NC * DPC;
    __shared__ SpecVals<T> sv;
    __shared__ T sx[kXTile][TD];
    load_spec_vals<T>(sp, &sv);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * 128 + threadIdx.x;
    T a[TD], ga[TD], gl[TD], go[NC];
#pragma unroll
    for (int k = 0; k < TD; ++k) {
        a[k] = (i < M) ? z[i * ldz + sp.col[k]] * sv.inv_ls[k] : T(0);
        ga[k] = T(0);
        gl[k] = T(0);
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) go[c] = T(0);
    const int64_t bag0 = (int64_t)blockIdx.y * kBagChunk;
    const int64_t bag1 = min(bag0 + kBagChunk, n_bags);
    for (int64_t bag = bag0; bag < bag1; ++bag) {
        const int64_t j0 = offsets[bag], j1 = offsets[bag + 1];
        const T g = (i < M) ? G[bag * ldg + i] : T(0);
        for (int64_t jt = j0; jt < j1; jt += kXTile) {
            const int nt = (int)min((int64_t)kXTile, j1 - jt);
            __syncthreads();
            for (int e = threadIdx.x; e < nt * TD; e += 128) {
                int r = e / TD, k = e % TD;
                sx[r][k] = x[(jt + r) * ldx + sp.col[k]] * sv.inv_ls[k];
            }
            __syncthreads();
            for (int r = 0; r < nt; ++r) {
                T b[TD];
#pragma unroll
                for (int k = 0; k < TD; ++k) b[k] = sx[r][k];
                accum_pair<T, NC, DPC>(sp, sv.os, a, b, g, g, gl, go, ga);
            }
        }
    }
    commit_theta<T, NC, DPC>(sp, sv, gl, go, dls, dos, threadIdx.x & 31);
    if (dz && i < M) {
#pragma unroll
        for (int k = 0; k < TD; ++k)
            if ((k / DPC) < sp.n_comp && (k % DPC) < sp.ndims[k / DPC])
                atomicAdd(&dz[i * lddz + sp.col[k]], ga[k] * sv.inv_ls[k]);
    }
}

// within-bag block sums: out[a] = sum_{i,j in bag a} k(x_i, x_j); one CTA per bag.
template <typename T, int NC, int DPC, bool BWD>
__global__ void __launch_bounds__(256) gram_bag_blocksum_kernel(const __grid_constant__ DevSpec sp,
                                                                const T* __restrict__ x, int64_t ldx,
                                                                const int64_t* __restrict__ offsets,
                                                                const T* __restrict__ g_in, T* __restrict__ out,
                                                                T* __restrict__ dls, T* __restrict__ dos) {
    constexpr int TD = NC * DPC;
    __shared__ SpecVals<T> sv;
    __shared__ T red[8];
    load_spec_vals<T>(sp, &sv);
    __syncthreads();
    const int64_t bag = blockIdx.x;
    const int64_t j0 = offsets[bag];
    const int64_t na = offsets[bag + 1] - j0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T sum = T(0);
    T gl[TD], go[NC], ga[TD];
#pragma unroll
    for (int k = 0; k < TD; ++k) gl[k] = ga[k] = T(0);
#pragma unroll
    for (int c = 0; c < NC; ++c) go[c] = T(0);
    const T g = BWD ? g_in[bag] : T(0);
    // a warp takes rows i = warp, warp+8, ...; lanes stride the columns
    for (int64_t i = warp; i < na; i += 8) {
        T a[TD];
#pragma unroll
        for (int k = 0; k < TD; ++k) a[k] = x[(j0 + i) * ldx + sp.col[k]] * sv.inv_ls[k];
        for (int64_t j = lane; j < na; j += 32) {
            T b[TD];
#pragma unroll
            for (int k = 0; k < TD; ++k) b[k] = x[(j0 + j) * ldx + sp.col[k]] * sv.inv_ls[k];
            if constexpr (BWD) accum_pair<T, NC, DPC>(sp, sv.os, a, b, g, g, gl, go, ga);
            else sum += eval_pair<T, NC, DPC>(sp, sv.os, a, b);
        }
    }
    if constexpr (BWD) {
        commit_theta<T, NC, DPC>(sp, sv, gl, go, dls, dos, lane);
    } else {
        sum = warp_sum(sum);
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        if (threadIdx.x == 0) {
            T t = T(0);
            for (int w = 0; w < 8; ++w) t += red[w];
            out[bag] = t;
        }
    }
}

template <typename T>
__global__ void bag_sum_kernel(const T* __restrict__ X, int64_t ldx, int64_t m, const int64_t* __restrict__ offsets,
                               T* __restrict__ out, int64_t ldo) {
    const int64_t bag = blockIdx.x;
    const int64_t j0 = offsets[bag], j1 = offsets[bag + 1];
    for (int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; c < m; c += (int64_t)gridDim.y * blockDim.x) {
        T s = T(0);
        for (int64_t j = j0; j < j1; ++j) s += X[j * ldx + c];
        out[bag * ldo + c] = s;
    }
}

// whitened KL, one pass over Ls
template <typename T>
__global__ void __launch_bounds__(256) whitened_kl_kernel(const T* __restrict__ m, const T* __restrict__ Ls, int64_t M,
                                                          int64_t ldls, T* __restrict__ out, T* __restrict__ dm,
                                                          T* __restrict__ dLs, int64_t lddls) {
    __shared__ T red[8];
    T acc = T(0);
    const int64_t total = M * M;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t i = e / M, j = e % M;
        T g = T(0);
        if (j <= i) {
            const T v = Ls[i * ldls + j];
            acc = fma(v, v, acc);
            g = v;
            if (i == j) {
                acc -= t_log<T>(v * v) + T(1);
                g -= T(1) / v;
                const T mi = m[i];
                acc = fma(mi, mi, acc);
                if (dm) dm[i] = mi;
            }
        }
        if (dLs) dLs[i * lddls + j] = g;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = T(0);
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out, T(0.5) * t);
    }
}

#include "gram_rbf.cuh"   // single-component RBF fast path (namespace rbf)

// ------------------------------------------------------------------------- launchers
template <typename T, typename TO>
int launch_gram_fwd(const DevSpec& sp, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2,
                    int64_t ldx2, void* K, int64_t ldk, const void* diag_dev, double diag_const, int same,
                    cudaStream_t st, int rna = 0) {
    constexpr int VEC = 16 / sizeof(TO);
    // row tiles per CTA (DCGP_GRAM_RT, default 8), fewer when the grid would drop under ~8 CTAs per SM
    static const int rt_max = getenv("DCGP_GRAM_RT") ? atoi(getenv("DCGP_GRAM_RT")) : 8;
    const int64_t gx = (n2 + 32 * VEC - 1) / (32 * VEC), tiles_y = (n1 + 63) / 64;
    int64_t rt = rt_max < 1 ? 1 : rt_max;
    while (rt > 1 && gx * ((tiles_y + rt - 1) / rt) < (int64_t)num_sms() * 8) rt >>= 1;
    dim3 grid((unsigned)gx, (unsigned)((tiles_y + rt - 1) / rt));
    int vec_ok = ((reinterpret_cast<uintptr_t>(K) % 16) == 0) && ((ldk * sizeof(TO)) % 16 == 0);
    if constexpr (sizeof(T) == sizeof(TO)) {
        if (rbf::applies(sp)) {
            rbf::launch_fwd<T>(sp, x1, n1, ldx1, x2, n2, ldx2, K, ldk, diag_dev, diag_const, same, grid, vec_ok, (int)rt, rna, st);
            DCGP_CHECK_LAUNCH();
            return 0;
        }
    }
#define CALL(NCV, DPCV)                                                                                      \
    gram_fwd_kernel<T, TO, NCV, DPCV><<<grid, 256, 0, st>>>(sp, (const T*)x1, n1, ldx1, (const T*)x2, n2, ldx2, \
                                                            (TO*)K, ldk, (const T*)diag_dev, (T)diag_const, same, \
                                                            vec_ok, (int)rt, rna);
    DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    DCGP_CHECK_LAUNCH();
    return 0;
}

template <typename T, int NC, int DPC, int RW>
void launch_gram_bwd_one(const DevSpec& sp, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2,
                         int64_t ldx2, const void* G, int64_t ldg, void* dls, void* dos, void* dx1, int64_t lddx1,
                         int same, cudaStream_t st) {
    const int cblk = 512;
    size_t smem = (size_t)NC * DPC * cblk * sizeof(T);
    // column blocks x row walkers: about four CTAs per SM in total
    const int64_t gx = (n2 + cblk - 1) / cblk, rows_max = (n1 + 8 * RW - 1) / (8 * RW);
    int64_t gy = (4 * (int64_t)num_sms() + gx - 1) / gx;
    if (gy > rows_max) gy = rows_max;
    if (gy < 1) gy = 1;
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (dx1) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(gram_bwd_kernel<T, NC, DPC, RW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gram_bwd_kernel<T, NC, DPC, RW, true><<<grid, 256, smem, st>>>(sp, (const T*)x1, n1, ldx1, (const T*)x2, n2, ldx2,
                                                                       (const T*)G, ldg, (T*)dls, (T*)dos, (T*)dx1, lddx1,
                                                                       same, cblk);
    } else {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(gram_bwd_kernel<T, NC, DPC, RW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        gram_bwd_kernel<T, NC, DPC, RW, false><<<grid, 256, smem, st>>>(sp, (const T*)x1, n1, ldx1, (const T*)x2, n2, ldx2,
                                                                        (const T*)G, ldg, (T*)dls, (T*)dos, (T*)dx1, lddx1,
                                                                        same, cblk);
    }
}

template <typename T>
int launch_gram_bwd(const DevSpec& sp, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2,
                    int64_t ldx2, const void* G, int64_t ldg, void* dls, void* dos, void* dx1, int64_t lddx1,
                    int same, cudaStream_t st) {
    if (!dx1 && rbf::applies(sp)) {   // hyper-parameter adjoints only (coordinate adjoints: the generic kernel)
        rbf::launch_bwd<T>(sp, x1, n1, ldx1, x2, n2, ldx2, G, ldg, dls, dos, st);
        DCGP_CHECK_LAUNCH();
        return 0;
    }
#define CALL(NCV, DPCV)                                                                                             \
    launch_gram_bwd_one<T, NCV, DPCV, (NCV * DPCV <= 4 ? 4 : (NCV * DPCV <= 8 ? 2 : 1))>(sp, x1, n1, ldx1, x2, n2,  \
                                                                                         ldx2, G, ldg, dls, dos,   \
                                                                                         dx1, lddx1, same, st);
    DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    DCGP_CHECK_LAUNCH();
    return 0;
}

}  // namespace

extern "C" int dcgp_gram(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2,
                         int64_t n2, int64_t ldx2, void* K, int64_t ldk, const void* diag_dev, double diag_const,
                         int dtype, int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!x1 || n1 < 0) return -2;
    if (!x2 || n2 < 0) return -5;
    if (!K || ldk < n2) return -8;
    if (n1 == 0 || n2 == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int same = (flags & DCGP_GRAM_SAME) ? 1 : 0;
    if (dtype == DCGP_F64)
        return launch_gram_fwd<double, double>(sp, x1, n1, ldx1, x2, n2, ldx2, K, ldk, diag_dev, diag_const, same, st);
    if (dtype == DCGP_F32) {
        if (flags & DCGP_GRAM_OUT_F64)
            return launch_gram_fwd<float, double>(sp, x1, n1, ldx1, x2, n2, ldx2, K, ldk, diag_dev, diag_const, same,
                                                  st);
        return launch_gram_fwd<float, float>(sp, x1, n1, ldx1, x2, n2, ldx2, K, ldk, diag_dev, diag_const, same, st);
    }
    return -12;
}

int dcgp_gemm_impl(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                   const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype, int flags,
                   cudaStream_t st, const void* Alo, const void* Blo, void* Clo);
int dcgp_gram_bwd_lowrank_tc(const DevSpec& sp, const float* x1, int64_t n1, int64_t ldx1, const float* x2, int64_t n2,
                             int64_t ldx2, const float* P, int64_t ldp, const float* Q, int64_t ldq, int64_t k, double alpha,
                             float* dls, float* dos, cudaStream_t st);

namespace {
// out[r, :] += (diag_const + diag_dev[0]) * W[r, :]: the nugget term of (K + c I) W for a row chunk
template <typename T>
__global__ void nugget_rows_kernel(T* __restrict__ out, int64_t ldo, const T* __restrict__ W, int64_t ldw, int64_t rows,
                                   int64_t m, const T* __restrict__ diag_dev, T diag_const) {
    const T c = diag_const + (diag_dev ? diag_dev[0] : T(0));
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < rows * m; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / m, j = e - r * m;
        out[r * ldo + j] += c * W[r * ldw + j];
    }
}
// Gram row panels: one panel when the whole matrix is small (<= 1 GiB: a single wide GEMM fills the machine
// best), else 2 GiB panels - the full matrix (34 GB at N = 65536 in FP64) is never resident.  (512 MiB panels, the
// round-1 size, are 2048 rows at N = 32768 in FP64: with n = 327 right-hand sides that is 48 tiles of 128 x 128 on 148
// SMs per panel, and K W took 57 ms against 33 ms for the materialised Gram + one GEMM; profiles/r02_bench_c6.json.)
int64_t gm_rows(int64_t n1, int64_t n2, int dtype) {
    const int64_t es = dtype == DCGP_F64 ? 8 : 4;
    if (n1 * n2 * es <= (1ll << 30)) return n1;
    int64_t rows = (2048ll << 20) / (n2 * es);
    rows = (rows / 128) * 128;
    if (rows < 128) rows = 128;
    return rows < n1 ? rows : n1;
}
// ... trimmed so that a panel's GEMM is a whole number of waves of 128 x 128 tiles (m right-hand sides): 8192 rows x 3 tile
// columns are 192 tiles = 1.3 waves of 148 SMs, 6272 rows are 147 tiles (FP64, N = 32768, n = 327: K W 47.5 ms with 2 GiB
// panels against 33 ms for one GEMM on the materialised matrix)
int64_t gm_rows_for(int64_t n1, int64_t n2, int64_t m, int dtype) {
    const int64_t cap = gm_rows(n1, n2, dtype);
    if (cap >= n1) return n1;
    const int64_t ct = (m + 127) / 128, sms = num_sms();
    const int64_t waves = (cap / 128) * ct / sms;
    if (waves < 1) return cap;
    const int64_t rows = (waves * sms / ct) * 128;
    return rows >= 128 ? rows : cap;
}
}  // namespace

int dcgp_gram_matmul_fused_f32(const DevSpec& sp, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2,
                               int64_t ldx2, const void* diag_dev, double diag_const, const void* W, int64_t m, int64_t ldw,
                               void* out, int64_t ldo, int same, cudaStream_t st);

// out = (k(x1, x2) [+ (diag_const + *diag_dev) I when SAME]) W without ever holding the Gram matrix in HBM as a
// whole once it is large: row panels are generated into the workspace and consumed by the tensor-core GEMM.
extern "C" size_t dcgp_gram_matmul_workspace(int64_t n1, int64_t n2, int dtype) {
    if (n1 <= 0 || n2 <= 0) return 0;
    return (size_t)gm_rows(n1, n2, dtype) * n2 * (dtype == DCGP_F64 ? 8 : 4);
}

extern "C" int dcgp_gram_matmul(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2,
                                int64_t n2, int64_t ldx2, const void* diag_dev, double diag_const, const void* W, int64_t m,
                                int64_t ldw, void* out, int64_t ldo, void* workspace, size_t workspace_bytes, int dtype,
                                int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!x1 || n1 < 0) return -2;
    if (!x2 || n2 < 0) return -5;
    if (!W || m < 0 || ldw < m) return -10;
    if (!out || ldo < m) return -13;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -17;
    if (n1 == 0 || m == 0) return 0;
    const int same = (flags & DCGP_GRAM_SAME) ? 1 : 0;
    if (same && n1 != n2) return -6;
    if (n2 == 0) return -6;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t es = dtype == DCGP_F64 ? 8 : 4, rows = gm_rows_for(n1, n2, m, dtype);
    const int gflags = flags & (DCGP_GEMM_TF32X3 | DCGP_GEMM_NO_TCGEN05);
    // the generated panels feed nothing but this product: single-pass TF32 gets them rounded to nearest at the source
    const int rna = (dtype == DCGP_F32 && !(flags & DCGP_GEMM_TF32X3)) ? 1 : 0;
    // FP32 single-pass: the fused tcgen05 kernel (gram_mm.cu) generates the Gram tiles in shared memory - no panels, no
    // workspace.  DCGP_GEMM_NO_TCGEN05 keeps the panel path (A/B runs); shapes the kernel declines fall through.
    if (rna && !(flags & (DCGP_GEMM_NO_TCGEN05 | DCGP_GRAM_NO_FUSE))) {
        const int rc = dcgp_gram_matmul_fused_f32(sp, x1, n1, ldx1, x2, n2, ldx2, diag_dev, diag_const, W, m, ldw, out, ldo, same, st);
        if (rc <= 0) return rc;
    }
    if (!workspace || workspace_bytes < dcgp_gram_matmul_workspace(n1, n2, dtype)) return -15;
    if (rows == n1) {
        // single panel: the symmetric Gram path (nugget applied on the diagonal by the generator itself)
        int rc = dtype == DCGP_F64 ? launch_gram_fwd<double, double>(sp, x1, n1, ldx1, x2, n2, ldx2, workspace, n2, diag_dev,
                                                                     diag_const, same, st)
                                   : launch_gram_fwd<float, float>(sp, x1, n1, ldx1, x2, n2, ldx2, workspace, n2, diag_dev,
                                                                   diag_const, same, st, rna);
        if (rc) return rc;
        return dcgp_gemm_impl(0, 0, n1, m, n2, 1.0, workspace, n2, W, ldw, 0.0, out, ldo, dtype, gflags, st, nullptr, nullptr,
                              nullptr);
    }
    for (int64_t r0 = 0; r0 < n1; r0 += rows) {
        const int64_t nr = (n1 - r0 < rows) ? n1 - r0 : rows;
        const char* xp = (const char*)x1 + r0 * ldx1 * es;
        int rc = dtype == DCGP_F64
                     ? launch_gram_fwd<double, double>(sp, xp, nr, ldx1, x2, n2, ldx2, workspace, n2, nullptr, 0.0, 0, st)
                     : launch_gram_fwd<float, float>(sp, xp, nr, ldx1, x2, n2, ldx2, workspace, n2, nullptr, 0.0, 0, st, rna);
        if (rc) return rc;
        char* op = (char*)out + r0 * ldo * es;
        rc = dcgp_gemm_impl(0, 0, nr, m, n2, 1.0, workspace, n2, W, ldw, 0.0, op, ldo, dtype, gflags, st, nullptr, nullptr, nullptr);
        if (rc) return rc;
        if (same && (diag_dev || diag_const != 0.0)) {
            const char* wp = (const char*)W + r0 * ldw * es;
            const unsigned grid = (unsigned)((nr * m + 255) / 256 < 4096 ? (nr * m + 255) / 256 : 4096);
            if (dtype == DCGP_F64)
                nugget_rows_kernel<double><<<grid, 256, 0, st>>>((double*)op, ldo, (const double*)wp, ldw, nr, m,
                                                                 (const double*)diag_dev, diag_const);
            else
                nugget_rows_kernel<float><<<grid, 256, 0, st>>>((float*)op, ldo, (const float*)wp, ldw, nr, m,
                                                                (const float*)diag_dev, (float)diag_const);
            DCGP_CHECK_LAUNCH();
        }
    }
    return 0;
}

extern "C" int dcgp_gram_bwd(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2,
                             int64_t n2, int64_t ldx2, const void* G, int64_t ldg, void* dls, void* dos, void* dx1,
                             int64_t lddx1, int dtype, int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!x1 || n1 < 0) return -2;
    if (!x2 || n2 < 0) return -5;
    if (!G || ldg < n2) return -8;
    const int same = (flags & DCGP_GRAM_SAME) ? 1 : 0;
    if (same && n1 != n2) return -15;
    if (n1 == 0 || n2 == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DCGP_F64)
        return launch_gram_bwd<double>(sp, x1, n1, ldx1, x2, n2, ldx2, G, ldg, dls, dos, dx1, lddx1, same, st);
    if (dtype == DCGP_F32)
        return launch_gram_bwd<float>(sp, x1, n1, ldx1, x2, n2, ldx2, G, ldg, dls, dos, dx1, lddx1, same, st);
    return -14;
}

extern "C" int dcgp_bag_sum(const void* X, int64_t ldx, int64_t m, const int64_t* offsets, int64_t n_bags, void* out,
                            int64_t ldo, int dtype, void* stream) {
    if (!X || !offsets || !out) return -1;
    if (n_bags <= 0 || m <= 0) return 0;
    dim3 grid((unsigned)n_bags, (unsigned)((m + 127) / 128));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DCGP_F64)
        bag_sum_kernel<double><<<grid, 128, 0, st>>>((const double*)X, ldx, m, offsets, (double*)out, ldo);
    else if (dtype == DCGP_F32)
        bag_sum_kernel<float><<<grid, 128, 0, st>>>((const float*)X, ldx, m, offsets, (float*)out, ldo);
    else
        return -8;
    DCGP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dcgp_gram_bagsum(const dcgp_kernel_spec* spec, const void* z, int64_t M, int64_t ldz, const void* x,
                                int64_t ldx, const int64_t* offsets, int64_t n_bags, void* out, int64_t ldo, int dtype,
                                int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!z || !x || !offsets || !out) return -2;
    if (M <= 0 || n_bags <= 0) return 0;
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((n_bags + kBagChunk - 1) / kBagChunk));
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(NCV, DPCV)                                                                                              \
    if (dtype == DCGP_F64)                                                                                           \
        gram_bagsum_kernel<double, NCV, DPCV><<<grid, 128, 0, st>>>(sp, (const double*)z, M, ldz, (const double*)x,  \
                                                                    ldx, offsets, n_bags, (double*)out, ldo);        \
    else                                                                                                             \
        gram_bagsum_kernel<float, NCV, DPCV><<<grid, 128, 0, st>>>(sp, (const float*)z, M, ldz, (const float*)x, ldx, \
                                                                   offsets, n_bags, (float*)out, ldo);
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -11;
    DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    DCGP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dcgp_gram_bagsum_bwd(const dcgp_kernel_spec* spec, const void* z, int64_t M, int64_t ldz, const void* x,
                                    int64_t ldx, const int64_t* offsets, int64_t n_bags, const void* G, int64_t ldg,
                                    void* dls, void* dos, void* dz, int64_t lddz, int dtype, int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!z || !x || !offsets || !G) return -2;
    if (M <= 0 || n_bags <= 0) return 0;
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((n_bags + kBagChunk - 1) / kBagChunk));
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(NCV, DPCV)                                                                                        \
    if (dtype == DCGP_F64)                                                                                     \
        gram_bagsum_bwd_kernel<double, NCV, DPCV><<<grid, 128, 0, st>>>(                                       \
            sp, (const double*)z, M, ldz, (const double*)x, ldx, offsets, n_bags, (const double*)G, ldg,       \
            (double*)dls, (double*)dos, (double*)dz, lddz);                                                    \
    else                                                                                                       \
        gram_bagsum_bwd_kernel<float, NCV, DPCV><<<grid, 128, 0, st>>>(                                        \
            sp, (const float*)z, M, ldz, (const float*)x, ldx, offsets, n_bags, (const float*)G, ldg,          \
            (float*)dls, (float*)dos, (float*)dz, lddz);
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -15;
    DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    DCGP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dcgp_gram_bag_blocksum(const dcgp_kernel_spec* spec, const void* x, int64_t ldx, const int64_t* offsets,
                                      int64_t n_bags, void* out, int dtype, int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!x || !offsets || !out) return -2;
    if (n_bags <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(NCV, DPCV)                                                                                            \
    if (dtype == DCGP_F64)                                                                                         \
        gram_bag_blocksum_kernel<double, NCV, DPCV, false><<<(unsigned)n_bags, 256, 0, st>>>(                      \
            sp, (const double*)x, ldx, offsets, nullptr, (double*)out, nullptr, nullptr);                          \
    else                                                                                                           \
        gram_bag_blocksum_kernel<float, NCV, DPCV, false><<<(unsigned)n_bags, 256, 0, st>>>(                       \
            sp, (const float*)x, ldx, offsets, nullptr, (float*)out, nullptr, nullptr);
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -7;
    DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    DCGP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dcgp_gram_bag_blocksum_bwd(const dcgp_kernel_spec* spec, const void* x, int64_t ldx,
                                          const int64_t* offsets, int64_t n_bags, const void* g, void* dls, void* dos,
                                          int dtype, int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!x || !offsets || !g) return -2;
    if (n_bags <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(NCV, DPCV)                                                                                            \
    if (dtype == DCGP_F64)                                                                                         \
        gram_bag_blocksum_kernel<double, NCV, DPCV, true><<<(unsigned)n_bags, 256, 0, st>>>(                       \
            sp, (const double*)x, ldx, offsets, (const double*)g, nullptr, (double*)dls, (double*)dos);            \
    else                                                                                                           \
        gram_bag_blocksum_kernel<float, NCV, DPCV, true><<<(unsigned)n_bags, 256, 0, st>>>(                        \
            sp, (const float*)x, ldx, offsets, (const float*)g, nullptr, (float*)dls, (float*)dos);
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -9;
    DCGP_DISPATCH_LAYOUT(sp, CALL);
#undef CALL
    DCGP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dcgp_whitened_kl(const void* m, const void* Ls, int64_t M, int64_t ldls, void* out, void* dm, void* dLs,
                                int64_t lddls, int dtype, void* stream) {
    if (!m || !Ls || !out) return -1;
    if (M <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t total = M * M;
    unsigned grid = (unsigned)min((int64_t)(num_sms() * 4), (total + 255) / 256);
    if (dtype == DCGP_F64)
        whitened_kl_kernel<double><<<grid, 256, 0, st>>>((const double*)m, (const double*)Ls, M, ldls, (double*)out,
                                                         (double*)dm, (double*)dLs, lddls);
    else if (dtype == DCGP_F32)
        whitened_kl_kernel<float><<<grid, 256, 0, st>>>((const float*)m, (const float*)Ls, M, ldls, (float*)out,
                                                        (float*)dm, (float*)dLs, lddls);
    else
        return -9;
    DCGP_CHECK_LAUNCH();
    return 0;
}

namespace {
// Bag-level expected log-likelihood of the aggregate-observation model and its adjoints in one pass:
//   ell = -1/2 sum_a [ (z_a^2 - 2 z_a S_a / n_a + (T_a + S_a^2) / n_a^2) n_a / s2 + log(2 pi s2 / n_a) ]
//   d ell / d S_a = (z_a - S_a / n_a) / s2,   d ell / d T_a = -1 / (2 n_a s2),
//   d ell / d s2  = -1/2 sum_a [ 1 / s2 - q_a n_a / s2^2 ]           (q_a = the bracket's first factor)
template <typename T>
__global__ void __launch_bounds__(256) vbagg_ell_kernel(const T* __restrict__ z, const T* __restrict__ S,
                                                        const T* __restrict__ Tq, const int64_t* __restrict__ offsets,
                                                        int64_t n_bags, const T* __restrict__ noise, T* __restrict__ ell,
                                                        T* __restrict__ dS, T* __restrict__ dT, T* __restrict__ dnoise) {
    __shared__ T red[2][8];
    const T s2 = noise[0];
    T acc = T(0), accn = T(0);
    for (int64_t a = threadIdx.x; a < n_bags; a += 256) {
        const T n = (T)(offsets[a + 1] - offsets[a]);
        const T q = z[a] * z[a] - T(2) * z[a] * S[a] / n + (Tq[a] + S[a] * S[a]) / (n * n);
        acc += q * n / s2 + t_log<T>(T(6.283185307179586476925286766559) * s2 / n);
        accn += T(1) / s2 - q * n / (s2 * s2);
        if (dS) dS[a] = (z[a] - S[a] / n) / s2;
        if (dT) dT[a] = T(-0.5) / (n * s2);
    }
    acc = warp_sum(acc);
    accn = warp_sum(accn);
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = acc;
        red[1][threadIdx.x >> 5] = accn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = T(0), tn = T(0);
        for (int w = 0; w < 8; ++w) {
            t += red[0][w];
            tn += red[1][w];
        }
        ell[0] = T(-0.5) * t;
        if (dnoise) dnoise[0] = T(-0.5) * tn;
    }
}
}  // namespace

extern "C" int dcgp_vbagg_ell(const void* z, const void* S, const void* T, const int64_t* offsets, int64_t n_bags,
                              const void* noise, void* ell, void* dS, void* dT, void* dnoise, int dtype, void* stream) {
    if (!z || !S || !T || !offsets) return -1;
    if (!noise) return -6;
    if (!ell) return -7;
    if (n_bags < 0) return -5;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DCGP_F64)
        vbagg_ell_kernel<double><<<1, 256, 0, st>>>((const double*)z, (const double*)S, (const double*)T, offsets, n_bags,
                                                    (const double*)noise, (double*)ell, (double*)dS, (double*)dT, (double*)dnoise);
    else if (dtype == DCGP_F32)
        vbagg_ell_kernel<float><<<1, 256, 0, st>>>((const float*)z, (const float*)S, (const float*)T, offsets, n_bags,
                                                   (const float*)noise, (float*)ell, (float*)dS, (float*)dT, (float*)dnoise);
    else
        return -11;
    DCGP_CHECK_LAUNCH();
    return 0;
}


// ---------------------------------------------------------------------------------------------------------------
// Gram adjoint for a low-rank upstream gradient G = alpha P Q^T (see dcgp.h)
// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t dcgp_gram_bwd_lowrank_workspace(int64_t n1, int64_t n2, int dtype) {
    if (n1 < 1 || n2 < 1) return 0;
    return (size_t)n1 * n2 * (dtype == DCGP_F64 ? 8 : 4);   // the fallback forms G; the fused F32 path uses none of it
}

extern "C" int dcgp_gram_bwd_lowrank(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2,
                                     int64_t n2, int64_t ldx2, const void* P, int64_t ldp, const void* Q, int64_t ldq, int64_t k,
                                     double alpha, void* dls, void* dos, void* workspace, size_t workspace_bytes, int dtype,
                                     int flags, void* stream) {
    DevSpec sp;
    if (make_dev_spec(spec, &sp)) return -1;
    if (!x1 || n1 < 0 || ldx1 < 1) return -2;
    if (!x2 || n2 < 0 || ldx2 < 1) return -5;
    if (!P || ldp < k) return -8;
    if (!Q || ldq < k) return -10;
    if (k < 1) return -12;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -19;
    if (n1 == 0 || n2 == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (dtype == DCGP_F32 && (flags & DCGP_GRAM_BWD_FUSE) && !(flags & DCGP_GEMM_TF32X3)) {
        const int rc = dcgp_gram_bwd_lowrank_tc(sp, (const float*)x1, n1, ldx1, (const float*)x2, n2, ldx2, (const float*)P, ldp,
                                                (const float*)Q, ldq, k, alpha, (float*)dls, (float*)dos, st);
        if (rc <= 0) return rc;
    }
    // fallback: G into the workspace, then the dense adjoint kernel
    if (!workspace || workspace_bytes < dcgp_gram_bwd_lowrank_workspace(n1, n2, dtype)) return -16;
    int rc = dcgp_gemm_impl(0, 1, n1, n2, k, alpha, P, ldp, Q, ldq, 0.0, workspace, n2, dtype, flags & DCGP_GEMM_TF32X3, st,
                            nullptr, nullptr, nullptr);
    if (rc) return rc;
    return dcgp_gram_bwd(spec, x1, n1, ldx1, x2, n2, ldx2, workspace, n2, dls, dos, nullptr, 0, dtype, 0, stream);
}
