// dcgp_vbagg_elbo_fwd / dcgp_vbagg_elbo_bwd: the whole VBagg minibatch ELBO - value and every gradient - behind two C
// calls (SURVEY.md section 8b lists vbagg_elbo_fwd/bwd in the minimum export set).
//
// Reference path replaced (one optimiser step of experiments/*/models/vbagg.py):
//   q(f) of the whitened SVGP              src/variational/variational_strategy.py:14-68
//   per-bag aggregation of q(f)            src/likelihoods/vbagg_gaussian_likelihood.py:54-76
//   expected log-likelihood                src/likelihoods/vbagg_gaussian_likelihood.py:78-111
//   ELBO = ELL / n_bags - beta KL / N      src/mlls/bag_variational_elbo.py:47-69
//   .backward()                            torch autograd over all of the above
//
// Algebra (K_Zx, Sigma_q and the aggregation matrix are never formed):
//   L_Z = chol(K_ZZ + jitter I),  W = L_Z^{-1}
//   Kb[a, :] = sum_{j in bag a} k(Z, x_j)               (dcgp_gram_bagsum)
//   U = W Kb^T                                           (M x n_bags: whitened, bag-summed cross covariance)
//   S_a = u_a^T m + c n_a                                (1_a^T mu_q)
//   D = L_S L_S^T - I,  T_a = 1_a^T K_xx 1_a + var_jitter n_a + u_a^T D u_a     (1_a^T Sigma_q 1_a)
//   ELL(z, S, T, noise) and its adjoints                 (dcgp_vbagg_ell),   KL(m, L_S)   (dcgp_whitened_kl)
// Backward, all as products with W (same derivation as functional._CholInverseFn):
//   dU = m dS^T + 2 (D U) diag(dT);   dD = U diag(dT) U^T;   dL_S = tril(2 dD L_S) - (beta / N) dKL
//   dKb = dU^T W;   dW = tril(dU Kb);   dL_Z = -tril(W^T dW W^T);   Phi = tril(L_Z^T dL_Z), diagonal halved;
//   dK_ZZ = sym(W^T Phi W)  ->  dcgp_gram_bwd;   dKb -> dcgp_gram_bagsum_bwd;   dT -> dcgp_gram_bag_blocksum_bwd
// Every launch goes to the caller's stream; nothing is allocated; the workspace carries the forward intermediates to the
// backward call.
#include "common.cuh"

#include "composite.cuh"

using namespace composite;

namespace {
// S_a = sum_i U[i,a] m_i + c n_a;  T_a = bsum_a + vj n_a + sum_i DU[i,a] U[i,a]      (one thread per bag, coalesced rows)
template <typename T>
__global__ void bag_stats_kernel(const T* __restrict__ U, const T* __restrict__ DU, const T* __restrict__ m, int64_t M,
                                 int64_t nb, const int64_t* __restrict__ off, const T* __restrict__ bsum,
                                 const T* __restrict__ mean_const, T vj, T* __restrict__ S, T* __restrict__ Tq) {
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nb) return;
    T s = T(0), q = T(0);
    for (int64_t i = 0; i < M; ++i) {
        const T u = U[i * nb + a];
        s = fma(u, m[i], s);
        q = fma(DU[i * nb + a], u, q);
    }
    const T na = (T)(off[a + 1] - off[a]);
    S[a] = s + (mean_const ? mean_const[0] * na : T(0));
    Tq[a] = bsum[a] + vj * na + q;
}

// out[0] = ell / n_bags - kl * beta / num_data
template <typename T>
__global__ void combine_kernel(const T* ell, const T* kl, T inv_nb, T klw, T* out) {
    out[0] = ell[0] * inv_nb - kl[0] * klw;
}

// dU[i,a] = s (m_i dS_a + 2 DU[i,a] dT_a);   Ud[i,a] = s U[i,a] dT_a
template <typename T>
__global__ void du_kernel(const T* __restrict__ U, const T* __restrict__ DU, const T* __restrict__ m,
                          const T* __restrict__ dS, const T* __restrict__ dT, int64_t M, int64_t nb, T s,
                          T* __restrict__ dU, T* __restrict__ Ud) {
    const int64_t i = blockIdx.y, a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= nb) return;
    const T ta = dT[a];
    dU[i * nb + a] = s * (m[i] * dS[a] + T(2) * DU[i * nb + a] * ta);
    Ud[i * nb + a] = s * U[i * nb + a] * ta;
}

// dm = s1 * v + s2 * dmkl;   dnoise = s * dn;   dmean_const = s * sum_a dS_a n_a;   g_bsum = s * dT   (one block)
template <typename T>
__global__ void tail_kernel(const T* __restrict__ v, const T* __restrict__ dmkl, int64_t M, T s1, T s2, T* __restrict__ dm,
                            const T* __restrict__ dn, T* __restrict__ dnoise, const T* __restrict__ dS,
                            const T* __restrict__ dT, const int64_t* __restrict__ off, int64_t nb, T* __restrict__ dmc,
                            T* __restrict__ gb) {
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x)
        if (dm) dm[i] = s1 * v[i] + s2 * dmkl[i];
    T acc = T(0);
    for (int64_t a = threadIdx.x; a < nb; a += blockDim.x) {
        acc += dS[a] * (T)(off[a + 1] - off[a]);
        gb[a] = s1 * dT[a];
    }
    acc = warp_sum(acc);
    __shared__ T red[32];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T t = T(0);
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        if (dmc) dmc[0] = s1 * t;
        if (dnoise) dnoise[0] = s1 * dn[0];
    }
}

struct Layout {   // element offsets into the workspace (all of type T), then the byte offset of the solver scratch
    size_t L, W, D, Ls, Kb, U, DU, S, T, dS, dT, bs, sc, t1, t2, t3, t4, Llo, Wlo, Wl2, Dlo, Lslo, Kblo, Ulo, lt1, lt2, end;
};

inline Layout layout(int64_t M, int64_t nb, bool f32) {
    Layout l;
    size_t o = 0;
    const size_t mm = up((size_t)M * M), mn = up((size_t)M * nb), n1 = up((size_t)nb);
    l.L = o; o += mm;      // K_ZZ + jitter I, then its Cholesky factor (lower)
    l.W = o; o += mm;      // L_Z^{-1}
    l.D = o; o += mm;      // L_S L_S^T - I
    l.Ls = o; o += mm;     // tril(L_S), packed
    l.Kb = o; o += mn;     // n_bags x M
    l.U = o; o += mn;      // M x n_bags
    l.DU = o; o += mn;     // M x n_bags
    l.S = o; o += n1;
    l.T = o; o += n1;
    l.dS = o; o += n1;
    l.dT = o; o += n1;
    l.bs = o; o += n1;     // block sums, later the scaled dT
    l.sc = o; o += 32;     // scalars: ell, kl, dnoise
    l.t1 = o; o += mm > mn ? mm : mn;   // scratch
    l.t2 = o; o += mm > mn ? mm : mn;
    l.t3 = o; o += mm > mn ? mm : mn;
    l.t4 = o; o += up((size_t)M * (M + 1));   // dKL / dm_kl
    // FP32: TF32 remainders (x - trunc_tf32(x)) of the operands that several products share, and two slots for the
    // remainders of scratch operands - with them every product runs 3xTF32 on the tcgen05 kernel (unused in FP64)
    l.Llo = l.Wlo = l.Wl2 = l.Dlo = l.Lslo = l.Kblo = l.Ulo = l.lt1 = l.lt2 = 0;
    if (f32) {
        l.Llo = o; o += mm;
        l.Wlo = o; o += mm;
        l.Wl2 = o; o += mm;    // FP32 low part of the FP64 inverse factor (island)
        l.Dlo = o; o += mm;
        l.Lslo = o; o += mm;
        l.Kblo = o; o += mn;
        l.Ulo = o; o += mn;
        l.lt1 = o; o += mm > mn ? mm : mn;
        l.lt2 = o; o += mm > mn ? mm : mn;
    }
    l.end = o;
    return l;
}

inline size_t solver_bytes(int64_t M, int dtype) {
    size_t a = dcgp_potrf_workspace(M, dtype), b = dcgp_trtri_workspace(M, dtype);
    if (dtype == DCGP_F32) {   // FP64 island of FP32 problems
        const size_t a2 = dcgp_potrf_workspace(M, DCGP_F64), b2 = dcgp_trtri_workspace(M, DCGP_F64);
        if (a2 > a) a = a2;
        if (b2 > b) b = b2;
    }
    return ((a > b ? a : b) + 255) & ~(size_t)255;
}

// DCGP_VBAGG_ISLAND (default 1): F32 problems evaluate, factor and invert K_ZZ in FP64 and take the adjoint of the factor in
// FP64 (src/variational/variational_strategy.py:35-44), as dcgp_cmp_elbo_* does; 0 = K_ZZ in FP32 / 3xTF32
inline bool island_enabled() {
    static const bool on = !(getenv("DCGP_VBAGG_ISLAND") && atoi(getenv("DCGP_VBAGG_ISLAND")) == 0);
    return on;
}
inline size_t island_base(const Layout& l, size_t es, int64_t M, int dtype) {
    return ((l.end * es + 255) & ~(size_t)255) + solver_bytes(M, dtype);
}

template <typename T>
int fwd_impl(const dcgp_vbagg_problem* p, void* elbo, int* info, void* ws, size_t ws_bytes, int dtype, int flags,
             cudaStream_t st) {
    const int64_t M = p->M, nb = p->n_bags;
    const Layout l = layout(M, nb, sizeof(T) == 4);
    const size_t need = sizeof(T) == 4 ? island_layout(island_base(l, 4, M, dtype), M, p->ldz).end
                                       : l.end * sizeof(T) + 256 + solver_bytes(M, dtype);
    if (ws_bytes < need) return -5;
    T* w = reinterpret_cast<T*>(ws);
    const bool f32 = sizeof(T) == 4;
    T* const Llo = f32 ? w + l.Llo : nullptr; T* const Wlo = f32 ? w + l.Wlo : nullptr; T* const Dlo = f32 ? w + l.Dlo : nullptr;
    T* const Lslo = f32 ? w + l.Lslo : nullptr; T* const Kblo = f32 ? w + l.Kblo : nullptr; T* const Ulo = f32 ? w + l.Ulo : nullptr;
    T* const q1 = f32 ? w + l.lt1 : nullptr; T* const q2 = f32 ? w + l.lt2 : nullptr;
    (void)Llo; (void)Dlo; (void)Lslo;
    void* solver = reinterpret_cast<unsigned char*>(ws) + ((l.end * sizeof(T) + 255) & ~(size_t)255);
    const size_t sbytes = solver_bytes(M, dtype);
    const int gf = dtype == DCGP_F32 ? (flags | DCGP_GEMM_TF32X3) : flags;
    const dim3 g2((unsigned)((M + 255) / 256), (unsigned)M);
    DCGP_CU_RC(cudaMemsetAsync(w + l.sc, 0, 32 * sizeof(T), st));   // dcgp_whitened_kl accumulates into its output
    // K_ZZ + jitter I -> L_Z -> W
    bool island = false;
    unsigned char* wb = reinterpret_cast<unsigned char*>(ws);
    if constexpr (sizeof(T) == 4) {
        island = island_enabled();
        if (island) {
            const Island a = island_layout(island_base(l, 4, M, dtype), M, p->ldz);
            double *Z64 = (double*)(wb + a.Z), *hyp = (double*)(wb + a.hyp), *L64 = (double*)(wb + a.L), *W64 = (double*)(wb + a.W),
                   *t64 = (double*)(wb + a.t1);
            const int64_t nz = M * p->ldz;
            f32_to_f64_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, st>>>((const float*)p->Z, Z64, nz);
            DCGP_CHECK_LAUNCH();
            spec_to_f64_kernel<<<DCGP_MAX_COMPONENTS, 32, 0, st>>>(*p->spec, hyp);
            DCGP_CHECK_LAUNCH();
            const dcgp_kernel_spec k64 = spec_f64(p->spec, hyp);
            RC(dcgp_gram(&k64, Z64, M, p->ldz, Z64, M, p->ldz, L64, M, nullptr, p->jitter, DCGP_F64, DCGP_GRAM_SAME, st));
            RC(dcgp_potrf(L64, M, M, info, solver, sbytes, DCGP_F64, flags, st));
            tril_copy_kernel<double><<<g2, 256, 0, st>>>(L64, M, t64, M);
            DCGP_CHECK_LAUNCH();
            DCGP_CU_RC(cudaMemcpyAsync(L64, t64, (size_t)M * M * 8, cudaMemcpyDeviceToDevice, st));
            RC(dcgp_trtri(L64, M, M, W64, M, solver, sbytes, DCGP_F64, flags, st));
            f64_to_hilo_kernel<<<(unsigned)(((int64_t)M * M + 255) / 256), 256, 0, st>>>(W64, (float*)(w + l.W), (float*)(w + l.Wl2),
                                                                                   (int64_t)M * M);
            DCGP_CHECK_LAUNCH();
        }
    }
    if (!island) {
        RC(dcgp_gram(p->spec, p->Z, M, p->ldz, p->Z, M, p->ldz, w + l.L, M, nullptr, p->jitter, dtype, DCGP_GRAM_SAME, st));
        RC(dcgp_potrf(w + l.L, M, M, info, solver, sbytes, dtype, gf, st));
        tril_copy_kernel<T><<<g2, 256, 0, st>>>(w + l.L, M, w + l.t1, M);
        DCGP_CHECK_LAUNCH();
        DCGP_CU_RC(cudaMemcpyAsync(w + l.L, w + l.t1, (size_t)M * M * sizeof(T), cudaMemcpyDeviceToDevice, st));
        RC(split_keep<T>(w + l.L, Llo, M, M, st));
        RC(dcgp_trtri(w + l.L, M, M, w + l.W, M, solver, sbytes, dtype, gf, st));
    }
    // bag-summed cross Gram and its whitening
    RC(dcgp_gram_bagsum(p->spec, p->Z, M, p->ldz, p->x, p->ldx, p->offsets, nb, w + l.Kb, M, dtype, 0, st));
    RC(split_keep<T>(w + l.W, Wlo, M, M, st));
    RC(split_keep<T>(w + l.Kb, Kblo, nb, M, st));
    RC(gemm3<T>(0, 1, M, nb, M, 1.0, w + l.W, M, Wlo, w + l.Kb, M, Kblo, 0.0, w + l.U, nb, q1, q2, dtype, gf, st));
    if (island)   // W = hi + lo of the FP64 inverse factor
        RC(gemm3<T>(0, 1, M, nb, M, 1.0, w + l.Wl2, M, nullptr, w + l.Kb, M, Kblo, 1.0, w + l.U, nb, q1, q2, dtype, gf, st));
    RC(split_keep<T>(w + l.U, Ulo, M, nb, st));
    // D = L_S L_S^T - I on the masked factor
    tril_copy_kernel<T><<<g2, 256, 0, st>>>((const T*)p->chol_var, p->ldc, w + l.Ls, M);
    DCGP_CHECK_LAUNCH();
    RC(split_keep<T>(w + l.Ls, Lslo, M, M, st));
    RC(gemm3<T>(0, 1, M, M, M, 1.0, w + l.Ls, M, Lslo, w + l.Ls, M, Lslo, 0.0, w + l.D, M, q1, q2, dtype, gf, st));
    add_diag_kernel<T><<<(unsigned)((M + 255) / 256), 256, 0, st>>>(w + l.D, M, T(-1));
    DCGP_CHECK_LAUNCH();
    RC(split_keep<T>(w + l.D, Dlo, M, M, st));
    RC(gemm3<T>(0, 0, M, nb, M, 1.0, w + l.D, M, Dlo, w + l.U, nb, Ulo, 0.0, w + l.DU, nb, q1, q2, dtype, gf, st));
    // per-bag statistics
    RC(dcgp_gram_bag_blocksum(p->spec, p->x, p->ldx, p->offsets, nb, w + l.bs, dtype, 0, st));
    bag_stats_kernel<T><<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(w + l.U, w + l.DU, (const T*)p->var_mean, M, nb,
                                                                     p->offsets, w + l.bs, (const T*)p->mean_const,
                                                                     (T)p->var_jitter, w + l.S, w + l.T);
    DCGP_CHECK_LAUNCH();
    // ELL with adjoints, KL with adjoints (dm_kl after the M x M block of t4), combination
    RC(dcgp_vbagg_ell(p->z, w + l.S, w + l.T, p->offsets, nb, p->noise, w + l.sc, w + l.dS, w + l.dT, w + l.sc + 2, dtype,
                      st));
    RC(dcgp_whitened_kl(p->var_mean, w + l.Ls, M, M, w + l.sc + 1, w + l.t4 + (size_t)M * M, w + l.t4, M, dtype, st));
    combine_kernel<T><<<1, 1, 0, st>>>(w + l.sc, w + l.sc + 1, (T)(1.0 / (double)nb), (T)(p->beta / p->num_data), (T*)elbo);
    DCGP_CHECK_LAUNCH();
    return 0;
}

template <typename T>
int bwd_impl(const dcgp_vbagg_problem* p, double scale, const dcgp_vbagg_grads* g, void* ws, size_t ws_bytes, int dtype,
             int flags, cudaStream_t st) {
    const int64_t M = p->M, nb = p->n_bags;
    const Layout l = layout(M, nb, sizeof(T) == 4);
    if (ws_bytes < (sizeof(T) == 4 ? island_layout(island_base(l, 4, M, dtype), M, p->ldz).end
                                   : l.end * sizeof(T) + 256 + solver_bytes(M, dtype)))
        return -5;
    T* w = reinterpret_cast<T*>(ws);
    const bool f32 = sizeof(T) == 4;
    T* const Llo = f32 ? w + l.Llo : nullptr; T* const Wlo = f32 ? w + l.Wlo : nullptr; T* const Dlo = f32 ? w + l.Dlo : nullptr;
    T* const Lslo = f32 ? w + l.Lslo : nullptr; T* const Kblo = f32 ? w + l.Kblo : nullptr; T* const Ulo = f32 ? w + l.Ulo : nullptr;
    T* const q1 = f32 ? w + l.lt1 : nullptr; T* const q2 = f32 ? w + l.lt2 : nullptr;
    (void)Llo; (void)Dlo; (void)Lslo;
    const int gf = dtype == DCGP_F32 ? (flags | DCGP_GEMM_TF32X3) : flags;
    const int nc = p->spec->n_components;
    const T s1 = (T)(scale / (double)nb), s2 = (T)(-scale * p->beta / p->num_data);
    const dim3 g2((unsigned)((M + 255) / 256), (unsigned)M), gu((unsigned)((nb + 255) / 256), (unsigned)M);
    if (g->dls) DCGP_CU_RC(cudaMemsetAsync(g->dls, 0, (size_t)nc * DCGP_MAX_DIMS * sizeof(T), st));
    if (g->dos) DCGP_CU_RC(cudaMemsetAsync(g->dos, 0, (size_t)nc * sizeof(T), st));
    if (g->dZ) {
        if (g->lddz < 1) return -3;
        DCGP_CU_RC(cudaMemsetAsync(g->dZ, 0, (size_t)M * g->lddz * sizeof(T), st));
    }
    // dU (t1), U diag(dT) (t2)
    du_kernel<T><<<gu, 256, 0, st>>>(w + l.U, w + l.DU, (const T*)p->var_mean, w + l.dS, w + l.dT, M, nb, s1, w + l.t1,
                                     w + l.t2);
    DCGP_CHECK_LAUNCH();
    // variational mean: dm = s1 U dS + s2 m;  noise, mean constant, scaled dT (into bs)
    RC(dcgp_gemm(0, 0, M, 1, nb, 1.0, w + l.U, nb, w + l.dS, 1, 0.0, w + l.t3, 1, dtype, gf, st));
    tail_kernel<T><<<1, 1024, 0, st>>>(w + l.t3, w + l.t4 + (size_t)M * M, M, s1, s2, (T*)g->dvar_mean, w + l.sc + 2,
                                       (T*)g->dnoise, w + l.dS, w + l.dT, p->offsets, nb, (T*)g->dmean_const, w + l.bs);
    DCGP_CHECK_LAUNCH();
    // variational factor: dL_S = tril(2 dD L_S) + s2 dKL,  dD = (U diag(dT)) U^T (already scaled by s1)
    if (g->dchol_var) {
        RC(gemm3<T>(0, 1, M, M, nb, 1.0, w + l.t2, nb, nullptr, w + l.U, nb, Ulo, 0.0, w + l.t3, M, q1, q2, dtype, gf, st));
        RC(gemm3<T>(0, 0, M, M, M, 1.0, w + l.t3, M, nullptr, w + l.Ls, M, Lslo, 0.0, w + l.t2, M, q1, q2, dtype, gf, st));
        tril_axpby_kernel<T><<<g2, 256, 0, st>>>(w + l.t2, M, T(2), w + l.t4, M, s2, 0, (T*)g->dchol_var, g->lddc);
        DCGP_CHECK_LAUNCH();
    }
    // within-bag block sums
    if (g->dls || g->dos)
        RC(dcgp_gram_bag_blocksum_bwd(p->spec, p->x, p->ldx, p->offsets, nb, w + l.bs, g->dls, g->dos, dtype, 0, st));
    // dKb = dU^T W (n_bags x M) -> bag-summed cross Gram adjoint
    RC(split_keep<T>(w + l.t1, q1, M, nb, st));   // dU is an operand twice
    bool island = false;
    if constexpr (sizeof(T) == 4) island = island_enabled();
    RC(gemm3<T>(1, 0, nb, M, M, 1.0, w + l.t1, nb, q1, w + l.W, M, Wlo, 0.0, w + l.t2, M, q1, q2, dtype, gf, st));
    if (island)
        RC(gemm3<T>(1, 0, nb, M, M, 1.0, w + l.t1, nb, q1, w + l.Wl2, M, nullptr, 1.0, w + l.t2, M, q1, q2, dtype, gf, st));
    RC(dcgp_gram_bagsum_bwd(p->spec, p->Z, M, p->ldz, p->x, p->ldx, p->offsets, nb, w + l.t2, M, g->dls, g->dos, g->dZ,
                            g->lddz, dtype, 0, st));
    // dW = tril(dU Kb);  dL_Z = -tril(W^T dW W^T);  Phi;  dK_ZZ = sym(W^T Phi W)
    RC(gemm3<T>(0, 0, M, M, nb, 1.0, w + l.t1, nb, q1, w + l.Kb, M, Kblo, 0.0, w + l.t3, M, q1, q2, dtype, gf, st));
    tril_axpby_kernel<T><<<g2, 256, 0, st>>>(w + l.t3, M, T(1), nullptr, 0, T(0), 0, w + l.t2, M);
    DCGP_CHECK_LAUNCH();
    if constexpr (sizeof(T) == 4) {
        if (island) {
            // the adjoint of W = chol(K_ZZ)^{-1} in FP64 (as in cmp.cu)
            unsigned char* wb = reinterpret_cast<unsigned char*>(ws);
            const Island a = island_layout(island_base(l, 4, M, dtype), M, p->ldz);
            double *Z64 = (double*)(wb + a.Z), *hyp = (double*)(wb + a.hyp), *L64 = (double*)(wb + a.L), *W64 = (double*)(wb + a.W),
                   *u1 = (double*)(wb + a.t1), *u2 = (double*)(wb + a.t2), *u3 = (double*)(wb + a.t3), *dZ64 = (double*)(wb + a.dZ),
                   *dth = (double*)(wb + a.dth);
            const int64_t mm = (int64_t)M * M;
            f32_to_f64_kernel<<<(unsigned)((mm + 255) / 256), 256, 0, st>>>((const float*)(w + l.t2), u2, mm);
            DCGP_CHECK_LAUNCH();
            RC(dcgp_gemm(1, 0, M, M, M, 1.0, W64, M, u2, M, 0.0, u1, M, DCGP_F64, flags, st));
            RC(dcgp_gemm(0, 1, M, M, M, -1.0, u1, M, W64, M, 0.0, u3, M, DCGP_F64, flags | DCGP_GEMM_LOWER, st));   // only its lower triangle is read
            tril_axpby_kernel<double><<<g2, 256, 0, st>>>(u3, M, 1.0, nullptr, 0, 0.0, 0, u2, M);
            DCGP_CHECK_LAUNCH();
            RC(dcgp_gemm(1, 0, M, M, M, 1.0, L64, M, u2, M, 0.0, u1, M, DCGP_F64, flags | DCGP_GEMM_LOWER, st));    // likewise
            tril_axpby_kernel<double><<<g2, 256, 0, st>>>(u1, M, 1.0, nullptr, 0, 0.0, 1, u2, M);
            DCGP_CHECK_LAUNCH();
            RC(dcgp_gemm(1, 0, M, M, M, 1.0, W64, M, u2, M, 0.0, u1, M, DCGP_F64, flags, st));
            RC(dcgp_gemm(0, 0, M, M, M, 1.0, u1, M, W64, M, 0.0, u3, M, DCGP_F64, flags, st));
            symmetrize_kernel<double><<<g2, 256, 0, st>>>(u3, M, u2);
            DCGP_CHECK_LAUNCH();
            const dcgp_kernel_spec k64 = spec_f64(p->spec, hyp);
            DCGP_CU_RC(cudaMemsetAsync(dth, 0, 512, st));
            DCGP_CU_RC(cudaMemsetAsync(dZ64, 0, (size_t)M * p->ldz * 8, st));
            RC(dcgp_gram_bwd(&k64, Z64, M, p->ldz, Z64, M, p->ldz, u2, M, g->dls ? dth : nullptr, g->dos ? dth + 32 : nullptr,
                             g->dZ ? dZ64 : nullptr, p->ldz, DCGP_F64, DCGP_GRAM_SAME, st));
            acc_f64_kernel<<<dim3(1, (unsigned)nc), 32, 0, st>>>(dth, DCGP_MAX_DIMS, (float*)g->dls, DCGP_MAX_DIMS, nc, DCGP_MAX_DIMS);
            DCGP_CHECK_LAUNCH();
            acc_f64_kernel<<<dim3(1, 1), 32, 0, st>>>(dth + 32, nc, (float*)g->dos, nc, 1, nc);
            DCGP_CHECK_LAUNCH();
            if (g->dZ) {
                acc_f64_kernel<<<dim3((unsigned)((p->ldz + 31) / 32), (unsigned)M), 32, 0, st>>>(
                    dZ64, p->ldz, (float*)g->dZ, g->lddz, M, p->ldz < g->lddz ? p->ldz : g->lddz);
                DCGP_CHECK_LAUNCH();
            }
            return 0;
        }
    }
    RC(gemm3<T>(1, 0, M, M, M, 1.0, w + l.W, M, Wlo, w + l.t2, M, nullptr, 0.0, w + l.t1, M, q1, q2, dtype, gf, st));
    RC(gemm3<T>(0, 1, M, M, M, -1.0, w + l.t1, M, nullptr, w + l.W, M, Wlo, 0.0, w + l.t3, M, q1, q2, dtype, gf, st));
    tril_axpby_kernel<T><<<g2, 256, 0, st>>>(w + l.t3, M, T(1), nullptr, 0, T(0), 0, w + l.t2, M);
    DCGP_CHECK_LAUNCH();
    RC(gemm3<T>(1, 0, M, M, M, 1.0, w + l.L, M, Llo, w + l.t2, M, nullptr, 0.0, w + l.t1, M, q1, q2, dtype, gf, st));
    tril_axpby_kernel<T><<<g2, 256, 0, st>>>(w + l.t1, M, T(1), nullptr, 0, T(0), 1, w + l.t2, M);
    DCGP_CHECK_LAUNCH();
    RC(gemm3<T>(1, 0, M, M, M, 1.0, w + l.W, M, Wlo, w + l.t2, M, nullptr, 0.0, w + l.t1, M, q1, q2, dtype, gf, st));
    RC(gemm3<T>(0, 0, M, M, M, 1.0, w + l.t1, M, nullptr, w + l.W, M, Wlo, 0.0, w + l.t3, M, q1, q2, dtype, gf, st));
    symmetrize_kernel<T><<<g2, 256, 0, st>>>(w + l.t3, M, w + l.t2);
    DCGP_CHECK_LAUNCH();
    RC(dcgp_gram_bwd(p->spec, p->Z, M, p->ldz, p->Z, M, p->ldz, w + l.t2, M, g->dls, g->dos, g->dZ, g->lddz, dtype,
                     DCGP_GRAM_SAME, st));
    return 0;
}

inline int check(const dcgp_vbagg_problem* p) {
    if (!p || !p->spec || !p->Z || !p->x || !p->offsets || !p->z || !p->var_mean || !p->chol_var || !p->noise) return -1;
    if (p->M < 1 || p->n_bags < 1 || p->ldz < 1 || p->ldx < 1 || p->ldc < p->M) return -1;
    if (!(p->num_data > 0.0)) return -1;
    return 0;
}

}  // namespace

extern "C" size_t dcgp_vbagg_elbo_workspace(int64_t M, int64_t n_bags, int64_t ldz, int dtype) {
    if (M < 1 || n_bags < 1 || ldz < 1) return 0;
    const Layout l = layout(M, n_bags, dtype == DCGP_F32);
    if (dtype == DCGP_F32) return island_layout(island_base(l, 4, M, dtype), M, ldz).end;
    return l.end * 8 + 256 + solver_bytes(M, dtype);
}

extern "C" int dcgp_vbagg_elbo_fwd(const dcgp_vbagg_problem* p, void* elbo, int* info, void* workspace,
                                   size_t workspace_bytes, int dtype, int flags, void* stream) {
    if (check(p)) return -1;
    if (!elbo) return -2;
    if (!info) return -3;
    if (!workspace) return -4;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dtype == DCGP_F64 ? fwd_impl<double>(p, elbo, info, workspace, workspace_bytes, dtype, flags, st)
                             : fwd_impl<float>(p, elbo, info, workspace, workspace_bytes, dtype, flags, st);
}

extern "C" int dcgp_vbagg_elbo_bwd(const dcgp_vbagg_problem* p, double scale, const dcgp_vbagg_grads* grads, void* workspace,
                                   size_t workspace_bytes, int dtype, int flags, void* stream) {
    if (check(p)) return -1;
    if (!grads) return -3;
    if (!workspace) return -4;
    if (grads->dchol_var && grads->lddc < p->M) return -3;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dtype == DCGP_F64 ? bwd_impl<double>(p, scale, grads, workspace, workspace_bytes, dtype, flags, st)
                             : bwd_impl<float>(p, scale, grads, workspace, workspace_bytes, dtype, flags, st);
}
