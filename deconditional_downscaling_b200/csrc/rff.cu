// Random Fourier features: Phi = [cos(x W / l), sin(x W / l)]  (N x 2D), x: N x d (d <= 8), W: d x D.
// The inner dimension is d, so this is an HBM-write-bound map (N*2D outputs), not a GEMM:
// one thread per (row, frequency), the row's d coordinates in registers, W read coalesced
// along the frequency index, sincos evaluated once for both halves of the feature vector.
// Backward: a warp owns a row, strides the frequencies, re-evaluates sincos and reduces
// d(x) with warp shuffles; d(lengthscale) is accumulated per lane and committed once per warp.
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ void t_sincos(T p, T* s, T* c);
template <> __device__ __forceinline__ void t_sincos<double>(double p, double* s, double* c) { sincos(p, s, c); }
template <> __device__ __forceinline__ void t_sincos<float>(float p, float* s, float* c) { sincosf(p, s, c); }

template <typename T>
__global__ void __launch_bounds__(256) rff_fwd_kernel(const T* __restrict__ x, int64_t n, int64_t ldx, int d,
                                                      const T* __restrict__ W, int64_t D, const T* __restrict__ ls,
                                                      T* __restrict__ out, int64_t ldo) {
    __shared__ T inv_ls[DCGP_MAX_DIMS];
    if (threadIdx.x < DCGP_MAX_DIMS) inv_ls[threadIdx.x] = threadIdx.x < d ? T(1) / ls[threadIdx.x] : T(0);
    __syncthreads();
    const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
    T w[DCGP_MAX_DIMS];
#pragma unroll
    for (int k = 0; k < DCGP_MAX_DIMS; ++k) w[k] = (k < d && j < D) ? W[k * D + j] * inv_ls[k] : T(0);
    const int64_t i0 = (int64_t)blockIdx.y * 32;
    for (int64_t i = i0; i < min(i0 + 32, n); ++i) {
        T p = T(0);
#pragma unroll
        for (int k = 0; k < DCGP_MAX_DIMS; ++k)
            if (k < d) p = fma(x[i * ldx + k], w[k], p);
        T s, c;
        t_sincos<T>(p, &s, &c);
        if (j < D) {
            out[i * ldo + j] = c;
            out[i * ldo + D + j] = s;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) rff_bwd_kernel(const T* __restrict__ x, int64_t n, int64_t ldx, int d,
                                                      const T* __restrict__ W, int64_t D, const T* __restrict__ ls,
                                                      const T* __restrict__ dphi, int64_t ldg, T* __restrict__ dx,
                                                      int64_t lddx, T* __restrict__ dls) {
    __shared__ T inv_ls[DCGP_MAX_DIMS];
    if (threadIdx.x < DCGP_MAX_DIMS) inv_ls[threadIdx.x] = threadIdx.x < d ? T(1) / ls[threadIdx.x] : T(0);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T gl[DCGP_MAX_DIMS];
#pragma unroll
    for (int k = 0; k < DCGP_MAX_DIMS; ++k) gl[k] = T(0);
    for (int64_t i = (int64_t)blockIdx.x * 8 + warp; i < n; i += (int64_t)gridDim.x * 8) {
        T xi[DCGP_MAX_DIMS], gx[DCGP_MAX_DIMS];
#pragma unroll
        for (int k = 0; k < DCGP_MAX_DIMS; ++k) {
            xi[k] = k < d ? x[i * ldx + k] : T(0);
            gx[k] = T(0);
        }
        for (int64_t j = lane; j < D; j += 32) {
            T w[DCGP_MAX_DIMS];
            T p = T(0);
#pragma unroll
            for (int k = 0; k < DCGP_MAX_DIMS; ++k) {
                w[k] = k < d ? W[k * D + j] * inv_ls[k] : T(0);
                p = fma(xi[k], w[k], p);
            }
            T s, c;
            t_sincos<T>(p, &s, &c);
            const T dp = dphi[i * ldg + D + j] * c - dphi[i * ldg + j] * s;
#pragma unroll
            for (int k = 0; k < DCGP_MAX_DIMS; ++k) {
                gx[k] = fma(dp, w[k], gx[k]);
                gl[k] = fma(-dp * xi[k], w[k], gl[k]);  // d p / d l_k = -x_k W_kj / l_k^2 = -x_k w_k / l_k
            }
        }
        if (dx) {
#pragma unroll
            for (int k = 0; k < DCGP_MAX_DIMS; ++k) {
                T v = warp_sum(gx[k]);
                if (lane == 0 && k < d) dx[i * lddx + k] = v;
            }
        }
    }
    if (dls) {
#pragma unroll
        for (int k = 0; k < DCGP_MAX_DIMS; ++k) {
            T v = warp_sum(gl[k]);
            if (lane == 0 && k < d) atomicAdd(&dls[k], v * inv_ls[k]);
        }
    }
}

}  // namespace

extern "C" int dcgp_rff_features(const void* x, int64_t n, int64_t ldx, int d, const void* W, int64_t D,
                                 const void* lengthscale, void* out, int64_t ldo, int dtype, void* stream) {
    if (!x || n < 0) return -1;
    if (d < 1 || d > DCGP_MAX_DIMS) return -4;
    if (!W || D < 1) return -5;
    if (!lengthscale) return -7;
    if (!out || ldo < 2 * D) return -8;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)((D + 255) / 256), (unsigned)((n + 31) / 32));
    if (dtype == DCGP_F64)
        rff_fwd_kernel<double><<<grid, 256, 0, st>>>((const double*)x, n, ldx, d, (const double*)W, D,
                                                     (const double*)lengthscale, (double*)out, ldo);
    else if (dtype == DCGP_F32)
        rff_fwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, n, ldx, d, (const float*)W, D,
                                                    (const float*)lengthscale, (float*)out, ldo);
    else
        return -10;
    DCGP_CHECK_LAUNCH();
    return 0;
}

extern "C" int dcgp_rff_features_bwd(const void* x, int64_t n, int64_t ldx, int d, const void* W, int64_t D,
                                     const void* lengthscale, const void* dphi, int64_t ldg, void* dx, int64_t lddx,
                                     void* dls, int dtype, void* stream) {
    if (!x || n < 0) return -1;
    if (d < 1 || d > DCGP_MAX_DIMS) return -4;
    if (!W || D < 1) return -5;
    if (!lengthscale) return -7;
    if (!dphi || ldg < 2 * D) return -8;
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned grid = (unsigned)min((int64_t)(num_sms() * 8), (n + 7) / 8);
    if (dtype == DCGP_F64)
        rff_bwd_kernel<double><<<grid, 256, 0, st>>>((const double*)x, n, ldx, d, (const double*)W, D,
                                                     (const double*)lengthscale, (const double*)dphi, ldg, (double*)dx,
                                                     lddx, (double*)dls);
    else if (dtype == DCGP_F32)
        rff_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)x, n, ldx, d, (const float*)W, D,
                                                    (const float*)lengthscale, (const float*)dphi, ldg, (float*)dx,
                                                    lddx, (float*)dls);
    else
        return -13;
    DCGP_CHECK_LAUNCH();
    return 0;
}
