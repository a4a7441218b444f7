// Shared device/host helpers for libdcgp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/dcgp.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libdcgp is written for sm_100a (B200) only"
#endif

#define DCGP_CUDA_ERR(e) (-(1000 + (int)(e)))
// number of kernel launches issued by this library since load (diagnostic only)
void dcgp_count_launch();
#define DCGP_CHECK_LAUNCH()                                 \
    do {                                                    \
        cudaError_t _e = cudaGetLastError();                \
        if (_e != cudaSuccess) return DCGP_CUDA_ERR(_e);    \
        dcgp_count_launch();                                \
    } while (0)

constexpr int kMaxTotalDims = 16;  // sum of active dims over the additive components

// Flattened kernel description passed by value to kernels (__grid_constant__).
// Slot layout: component c owns slots [c*dpc, (c+1)*dpc); unused slots read input column 0
// with inverse lengthscale 0 (so they contribute exactly 0 to the squared distance) and
// unused components carry outputscale 0.  (ncb, dpc) are template buckets, so every inner
// loop is fully unrolled with no runtime predicates.
struct DevSpec {
    int n_comp;
    int ncb;  // component bucket: 1, 2 or 4
    int dpc;  // dims-per-component bucket: 2, 4 or 8   (ncb * dpc <= kMaxTotalDims)
    int kind[DCGP_MAX_COMPONENTS];
    int ndims[DCGP_MAX_COMPONENTS];
    int col[kMaxTotalDims];
    const void* ls[DCGP_MAX_COMPONENTS];
    const void* os[DCGP_MAX_COMPONENTS];
};

inline int make_dev_spec(const dcgp_kernel_spec* s, DevSpec* d) {
    if (!s || s->n_components < 1 || s->n_components > DCGP_MAX_COMPONENTS) return -1;
    d->n_comp = s->n_components;
    d->ncb = s->n_components == 1 ? 1 : (s->n_components == 2 ? 2 : 4);
    int maxd = 1;
    for (int c = 0; c < s->n_components; ++c) {
        const dcgp_component& k = s->comp[c];
        if (k.kind < 0 || k.kind > DCGP_KIND_MATERN52) return -1;
        if (k.ndims < 1 || k.ndims > DCGP_MAX_DIMS || !k.lengthscale || !k.outputscale) return -1;
        if (k.ndims > maxd) maxd = k.ndims;
    }
    d->dpc = maxd <= 2 ? 2 : (maxd <= 4 ? 4 : 8);
    if (d->ncb * d->dpc > kMaxTotalDims) return -1;
    for (int i = 0; i < kMaxTotalDims; ++i) d->col[i] = 0;
    for (int c = 0; c < DCGP_MAX_COMPONENTS; ++c) {
        d->kind[c] = DCGP_KIND_RBF;
        d->ndims[c] = 0;
        d->ls[c] = nullptr;
        d->os[c] = nullptr;
    }
    for (int c = 0; c < s->n_components; ++c) {
        const dcgp_component& k = s->comp[c];
        d->kind[c] = k.kind;
        d->ndims[c] = k.ndims;
        for (int j = 0; j < k.ndims; ++j) {
            if (k.dims[j] < 0) return -1;
            d->col[c * d->dpc + j] = k.dims[j];
        }
        d->ls[c] = k.lengthscale;
        d->os[c] = k.outputscale;
    }
    return 0;
}

// per-CTA copy of the (device-resident) hyper-parameters, padded to the slot layout
template <typename T>
struct SpecVals {
    T inv_ls[kMaxTotalDims];
    T os[DCGP_MAX_COMPONENTS];
};

template <typename T>
__device__ __forceinline__ void load_spec_vals(const DevSpec& sp, SpecVals<T>* sv) {
    for (int c = threadIdx.x; c < DCGP_MAX_COMPONENTS; c += blockDim.x)
        sv->os[c] = (c < sp.n_comp) ? ((const T*)sp.os[c])[0] : T(0);
    for (int s = threadIdx.x; s < kMaxTotalDims; s += blockDim.x) {
        int c = s / sp.dpc, j = s % sp.dpc;
        sv->inv_ls[s] = (c < sp.n_comp && j < sp.ndims[c]) ? T(1) / ((const T*)sp.ls[c])[j] : T(0);
    }
}

#define DCGP_DISPATCH_LAYOUT(sp, CALL)                 \
    do {                                               \
        if ((sp).ncb == 1 && (sp).dpc == 2) CALL(1, 2) \
        else if ((sp).ncb == 1 && (sp).dpc == 4) CALL(1, 4) \
        else if ((sp).ncb == 1) CALL(1, 8)             \
        else if ((sp).ncb == 2 && (sp).dpc == 2) CALL(2, 2) \
        else if ((sp).ncb == 2 && (sp).dpc == 4) CALL(2, 4) \
        else if ((sp).ncb == 2) CALL(2, 8)             \
        else if ((sp).dpc == 2) CALL(4, 2)             \
        else CALL(4, 4)                                \
    } while (0)

template <typename T> __device__ __forceinline__ T t_exp(T x);
// FP64 exp: the CUDA library function.  Round 2 tried a hand-written exp for non-positive arguments (Cody-Waite reduction
// by ln 2 / 4, four-entry table, degree-9 polynomial, <= 2 ulp): the Gram build got SLOWER on B200, 2473 GB/s with
// FP64 <-> integer conversions and 2651 GB/s with the magic-number shift and two Horner chains, against 3030 GB/s with
// exp() (profiles/r02_gram_fp64_exp.txt) - removed.
template <> __device__ __forceinline__ double t_exp<double>(double x) { return exp(x); }
template <> __device__ __forceinline__ float t_exp<float>(float x) { return __expf(x); }
template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_log(T x);
template <> __device__ __forceinline__ double t_log<double>(double x) { return log(x); }
template <> __device__ __forceinline__ float t_log<float>(float x) { return logf(x); }

// phi(r^2) for one base kernel and w(r^2) with  d phi / d u_k = -w * u_k  (u = scaled difference).
template <typename T>
__device__ __forceinline__ T kernel_phi(int kind, T r2) {
    if (kind == DCGP_KIND_RBF) return t_exp<T>(T(-0.5) * r2);
    T r = t_sqrt<T>(r2 > T(1e-30) ? r2 : T(1e-30));
    if (kind == DCGP_KIND_MATERN32) {
        T s = T(1.7320508075688772) * r;
        return (T(1) + s) * t_exp<T>(-s);
    }
    if (kind == DCGP_KIND_MATERN12) return t_exp<T>(-r);
    T s = T(2.23606797749979) * r;
    return (T(1) + s + T(5.0 / 3.0) * r2) * t_exp<T>(-s);
}

template <typename T>
__device__ __forceinline__ void kernel_phi_w(int kind, T r2, T& phi, T& w) {
    if (kind == DCGP_KIND_RBF) {
        phi = t_exp<T>(T(-0.5) * r2);
        w = phi;
        return;
    }
    T r = t_sqrt<T>(r2 > T(1e-30) ? r2 : T(1e-30));
    if (kind == DCGP_KIND_MATERN32) {
        T s = T(1.7320508075688772) * r;
        T e = t_exp<T>(-s);
        phi = (T(1) + s) * e;
        w = T(3) * e;
    } else if (kind == DCGP_KIND_MATERN12) {
        T e = t_exp<T>(-r);
        phi = e;
        w = (r2 > T(1e-30)) ? e / r : T(0);
    } else {
        T s = T(2.23606797749979) * r;
        T e = t_exp<T>(-s);
        phi = (T(1) + s + T(5.0 / 3.0) * r2) * e;
        w = T(5.0 / 3.0) * (T(1) + s) * e;
    }
}

// value of the additive kernel for one pair; a[] and b[] already scaled by 1/l
template <typename T, int NC, int DPC>
__device__ __forceinline__ T eval_pair(const DevSpec& sp, const T* os, const T (&a)[NC * DPC],
                                       const T (&b)[NC * DPC]) {
    T val = T(0);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < DPC; ++j) {
            T d = a[c * DPC + j] - b[c * DPC + j];
            acc = fma(d, d, acc);
        }
        val = fma(os[c], kernel_phi<T>(sp.kind[c], acc), val);
    }
    return val;
}

// Same as kernel_phi_w with hardware-approximate exp / sqrt (MUFU ex2, rsq: ~2 ulp) for FP32 adjoint kernels,
// whose results are reductions over ~10^7 terms checked at 1e-3; FP64 keeps the accurate versions.
template <typename T>
__device__ __forceinline__ void kernel_phi_w_fast(int kind, T r2, T& phi, T& w) {
    kernel_phi_w<T>(kind, r2, phi, w);
}
template <>
__device__ __forceinline__ void kernel_phi_w_fast<float>(int kind, float r2, float& phi, float& w) {
    if (kind == DCGP_KIND_RBF) {
        phi = __expf(-0.5f * r2);
        w = phi;
        return;
    }
    const float rc = r2 > 1e-30f ? r2 : 1e-30f;
    const float rinv = rsqrtf(rc), r = rc * rinv;
    if (kind == DCGP_KIND_MATERN32) {
        const float s = 1.7320508075688772f * r, e = __expf(-s);
        phi = (1.f + s) * e;
        w = 3.f * e;
    } else if (kind == DCGP_KIND_MATERN12) {
        const float e = __expf(-r);
        phi = e;
        w = (r2 > 1e-30f) ? e * rinv : 0.f;
    } else {
        const float s = 2.23606797749979f * r, e = __expf(-s);
        phi = (1.f + s + (5.f / 3.f) * r2) * e;
        w = (5.f / 3.f) * (1.f + s) * e;
    }
}

// accumulate the adjoints of one pair: g multiplies dK/dtheta, gs multiplies dK/da (only when DX)
template <typename T, int NC, int DPC, bool DX = true, bool FAST = false>
__device__ __forceinline__ void accum_pair(const DevSpec& sp, const T* os, const T (&a)[NC * DPC],
                                           const T (&b)[NC * DPC], T g, T gs, T (&gl)[NC * DPC], T (&go)[NC],
                                           T (&ga)[NC * DPC]) {
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        T acc = T(0);
        T u[DPC];
#pragma unroll
        for (int j = 0; j < DPC; ++j) {
            u[j] = a[c * DPC + j] - b[c * DPC + j];
            acc = fma(u[j], u[j], acc);
        }
        T phi, w;
        if (FAST) kernel_phi_w_fast<T>(sp.kind[c], acc, phi, w);
        else kernel_phi_w<T>(sp.kind[c], acc, phi, w);
        go[c] = fma(g, phi, go[c]);
        const T ow = os[c] * w;
        const T gow = g * ow, gsow = -gs * ow;
#pragma unroll
        for (int j = 0; j < DPC; ++j) {
            gl[c * DPC + j] = fma(gow * u[j], u[j], gl[c * DPC + j]);
            if (DX) ga[c * DPC + j] = fma(gsow, u[j], ga[c * DPC + j]);
        }
    }
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
