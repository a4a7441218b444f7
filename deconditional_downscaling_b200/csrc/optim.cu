// Device-side optimiser update of the ELBO step: one fused multi-tensor Adam launch over the flat parameter /
// gradient buffers (training.FlatGrads), predicated on the device failure words of the step's factorisations so
// that a captured CUDA graph can contain the update without a host read-back deciding it.
// Replaces `torch.optim.Adam.step()` at experiments/downscaling/models/variational_cmp.py:146-148,183 and
// experiments/swiss_roll/models/variational_cmp.py:91-110 (same arithmetic: no weight decay, no amsgrad).
#include "common.cuh"

namespace {

__device__ __forceinline__ bool vetoed(const int* __restrict__ veto, int n_veto) {
    bool bad = false;
    for (int i = 0; i < n_veto; ++i) bad |= (veto[i] != 0);
    return bad;
}

// p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps),  m = b1 m + (1 - b1) g,  v = b2 v + (1 - b2) g^2,
// t = step[0] + 1 (the counter itself is advanced by adam_tick_kernel, launched right after on the same stream).
template <typename T>
__global__ void __launch_bounds__(256) adam_kernel(T* __restrict__ p, T* __restrict__ g, T* __restrict__ m,
                                                   T* __restrict__ v, const double* __restrict__ step, int64_t n,
                                                   double lr, double b1, double b2, double eps, double gscale,
                                                   const int* __restrict__ veto, int n_veto, int zero_grads) {
    if (vetoed(veto, n_veto)) return;
    const double t = step[0] + 1.0;
    const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
    const T step_size = (T)(lr / bc1), rs2 = (T)(1.0 / sqrt(bc2));
    const T tb2 = (T)b2, ob1 = (T)(1.0 - b1), ob2 = (T)(1.0 - b2), teps = (T)eps, gs = (T)gscale;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        const T gi = g[i] * gs;
        const T mi = m[i] + ob1 * (gi - m[i]);           // lerp, as torch
        const T vi = fma(tb2, v[i], ob2 * gi * gi);
        m[i] = mi;
        v[i] = vi;
        const T denom = t_sqrt<T>(vi) * rs2 + teps;
        p[i] -= step_size * (mi / denom);
        if (zero_grads) g[i] = T(0);
    }
}

__global__ void adam_tick_kernel(double* step, const int* __restrict__ veto, int n_veto) {
    if (!vetoed(veto, n_veto)) step[0] += 1.0;
}

}  // namespace

extern "C" int dcgp_adam_step(void* params, void* grads, void* exp_avg, void* exp_avg_sq, void* step, int64_t n,
                              double lr, double beta1, double beta2, double eps, double grad_scale, const int* veto,
                              int n_veto, int zero_grads, int dtype, void* stream) {
    if (n < 0) return -6;
    if (n == 0) return 0;
    if (!params || !grads || !exp_avg || !exp_avg_sq || !step) return -1;
    if (n_veto < 0 || (n_veto > 0 && !veto)) return -12;
    if (dtype != DCGP_F64 && dtype != DCGP_F32) return -15;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < 4 * num_sms() ? want : 4 * num_sms());
    if (dtype == DCGP_F64)
        adam_kernel<double><<<grid, 256, 0, st>>>((double*)params, (double*)grads, (double*)exp_avg,
                                                  (double*)exp_avg_sq, (const double*)step, n, lr, beta1, beta2, eps,
                                                  grad_scale, veto, n_veto, zero_grads);
    else
        adam_kernel<float><<<grid, 256, 0, st>>>((float*)params, (float*)grads, (float*)exp_avg, (float*)exp_avg_sq,
                                                 (const double*)step, n, lr, beta1, beta2, eps, grad_scale, veto, n_veto, zero_grads);
    DCGP_CHECK_LAUNCH();
    adam_tick_kernel<<<1, 1, 0, st>>>((double*)step, veto, n_veto);
    DCGP_CHECK_LAUNCH();
    return 0;
}
