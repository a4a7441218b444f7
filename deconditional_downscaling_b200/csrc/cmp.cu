// dcgp_cmp_elbo_fwd / dcgp_cmp_elbo_bwd: the whole VariationalCMP minibatch ELBO (with individuals noise) - value and
// every gradient - behind two C calls (SURVEY.md section 8b lists cmp_elbo_fwd/bwd in the minimum export set).
//
// Reference path replaced (one optimiser step of experiments/downscaling/models/variational_cmp.py:163-186):
//   q(f) of the whitened SVGP                          src/variational/variational_strategy.py:14-68
//   Lambda^{-1}, l(y, y~)                               src/models/variational_cmp.py:86-113
//   expected log-likelihood with individuals noise     src/likelihoods/cmp_likelihood.py:41-78
//   ELBO = ELL / n - beta KL / num_data                 src/mlls/bag_variational_elbo.py:47-69
//   .backward()                                         torch autograd over all of the above
//
// Algebra (minimal form: no root of Lambda^{-1}, no Sigma_q, no chol(Sigma_q)):
//   L = chol(l(y~, y~) + lbda N I),  A = Lambda^{-1} l(y~, y)                         (N x n deconditioning operator)
//   W = chol(K_ZZ + jitter I)^{-1},  A~ = W k(Z, x),  mu_q = A~^T m + c,  D = L_S L_S^T - I
//   a = A^T mu_q,  H = A^T A,  U = A~ A,  G = A^T (K_xx + var_jitter I) A + U^T D U   (= A^T Sigma_q A)
//   C = s_ind H + s2 I,  e = z - a,  ELL = -1/2 (n log 2 pi + log|C| + e^T C^{-1} e + tr(C^{-1} G))
// Backward: dC = c0 (C^{-1} - u u^T - C^{-1} G C^{-1}), u = C^{-1} e, c0 = -scale / (2 n);  dG = c0 C^{-1};  da = -2 c0 u;
//   dA = 2 s_ind A dC + 2 (K_xx + vj I) A dG + mu_q da^T + A~^T dU,  dU = 2 D U dG,  dD = U dG U^T,  dK_xx = A dG A^T,
//   dmu = A da,  dA~ = dU A^T + m dmu^T,  dK_Zx = W^T dA~,  dW = tril(dA~ K_Zx^T) -> dK_ZZ (as in vbagg.cu),
//   X = Lambda^{-1} dA:  d l(y~, y) = X,  dLambda = -X A^T.
// Work is ordered on the caller's stream (the inducing-point chain forks onto a side lane and joins back before either call
// returns, see SideLane); no device memory is allocated; the workspace carries the forward intermediates to the backward
// call.  F32: factorisations, solves, the whitening products and everything M / n sized run 3xTF32; the
// individuals-sized products run single-pass TF32 on operands rounded to nearest (the precision policy of the Python layer,
// DESIGN.md section 3), K_xx A on panels rounded in the Gram epilogue.
#include <map>

#include "common.cuh"

#include "composite.cuh"

using namespace composite;

namespace {

// C = s_ind H + s2 I,  e = z - a
template <typename T>
__global__ void build_c_kernel(const T* __restrict__ H, int64_t n, const T* __restrict__ s_ind, const T* __restrict__ s2,
                               T* __restrict__ C, const T* __restrict__ z, const T* __restrict__ a, T* __restrict__ e) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    C[i * n + j] = s_ind[0] * H[i * n + j] + (i == j ? s2[0] : T(0));
    if (i == 0) e[j] = z[j] - a[j];
}

template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    T t = T(0);
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    return t;
}

// sc[3] = e^T C^{-1} e and sc[4] = tr(C^{-1} G) come from dot();  sc[0] = ell;  out = ell / n - kl * klw
template <typename T>
__global__ void ell_kernel(int64_t n, const T* __restrict__ logdet, const T* __restrict__ kl, T klw, T* __restrict__ sc,
                           T* __restrict__ out) {
    const T ell = T(-0.5) * ((T)n * T(1.8378770664093453) + logdet[0] + sc[3] + sc[4]);
    sc[0] = ell;
    out[0] = ell / (T)n - kl[0] * klw;
}

// dC = c0 (Cinv - u u^T - T1);  S2 = 2 s_ind dC;  Gd = c0 Cinv (= dG);  Gd2 = 2 Gd
template <typename T>
__global__ void dc_kernel(const T* __restrict__ Cinv, const T* __restrict__ T1, const T* __restrict__ u, int64_t n, T c0,
                          const T* __restrict__ s_ind, T* __restrict__ dC, T* __restrict__ S2, T* __restrict__ Gd,
                          T* __restrict__ Gd2) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const T ci = Cinv[i * n + j];
    const T d = c0 * (ci - u[i] * u[j] - T1[i * n + j]);
    dC[i * n + j] = d;
    S2[i * n + j] = T(2) * s_ind[0] * d;
    Gd[i * n + j] = c0 * ci;
    Gd2[i * n + j] = T(2) * c0 * ci;
}

// da = -2 c0 u;  sc[8] += tr(dC)   (<dC, H> comes from dot() into sc[9])
template <typename T>
__global__ void dscal_kernel(const T* __restrict__ dC, const T* __restrict__ u, int64_t n, T c0, T* __restrict__ da,
                             T* __restrict__ sc) {
    __shared__ T red[32];
    T t = T(0);
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        t += dC[i * n + i];
        da[i] = T(-2) * c0 * u[i];
    }
    t = block_sum(t, red);
    if (threadIdx.x == 0) sc[8] = t;
}

template <typename T>
__global__ void dscal_out_kernel(const T* __restrict__ sc, T* __restrict__ dnoise, T* __restrict__ dind) {
    if (dnoise) dnoise[0] = sc[8];
    if (dind) dind[0] = sc[9];
}

// Mx[i, j] += a_i b_j
template <typename T>
__global__ void rank1_kernel(T* __restrict__ Mx, int64_t rows, int64_t cols, const T* __restrict__ a,
                             const T* __restrict__ b) {
    const int64_t i = blockIdx.y, j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < cols) Mx[i * cols + j] = fma(a[i], b[j], Mx[i * cols + j]);
}

template <typename T>
__global__ void add_scalar_kernel(T* __restrict__ v, int64_t n, const T* __restrict__ c) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += c[0];
}

// dm = v + s2 dmkl;  dmc = sum(dmu)       (one block)
template <typename T>
__global__ void tail_kernel(const T* __restrict__ v, const T* __restrict__ dmkl, int64_t M, T s2, T* __restrict__ dm,
                            const T* __restrict__ dmu, int64_t N, T* __restrict__ dmc) {
    __shared__ T red[32];
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x)
        if (dm) dm[i] = v[i] + s2 * dmkl[i];
    T t = T(0);
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) t += dmu[i];
    t = block_sum(t, red);
    if (threadIdx.x == 0 && dmc) dmc[0] = t;
}

// A second stream for the inducing-point chain (K_ZZ factor, whitening of K_Zx and their adjoints): ~1 ms of short serial
// kernels per step that fit beside the Lambda factorisation / the adjoint solve, which leave most SMs idle (the Python layer
// does the same, variational.side_stream_scope).  One lane per calling stream, created on first use and kept (host-side
// state like the factorisation lanes, see the top of include/dcgp.h); fork / join with events, so the two calls stay
// capturable in a CUDA graph.  DCGP_CMP_SIDE=0 keeps everything on the caller's stream.
struct SideLane {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    int init() {
        if (s) return 0;
        DCGP_CU_RC(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        DCGP_CU_RC(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        DCGP_CU_RC(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        return 0;
    }
};
thread_local std::map<cudaStream_t, SideLane> g_side;
inline bool side_enabled() {
    static const bool on = !(getenv("DCGP_CMP_SIDE") && atoi(getenv("DCGP_CMP_SIDE")) == 0);
    return on;
}

// DCGP_CMP_ISLAND: 1 (default) = forward and adjoint of the K_ZZ factor in FP64; 2 = forward in FP64, adjoint chain in
// FP32 / 3xTF32 on the rounded factor; 0 = no island
inline int island_mode() {
    static const int m = getenv("DCGP_CMP_ISLAND") ? atoi(getenv("DCGP_CMP_ISLAND")) : 1;
    return m;
}
inline bool island_enabled() { return island_mode() != 0; }

struct Layout {
    size_t Lam, NN, NN2, A, Ar, KA, dA, P, Kzx, Kzxlo, At, dAt, dAtlo, dKzx, U, DU, dU, tU, H, G, C, Wc, Cinv, T1, dC, S2, Gd, Gd2, Lz, W, Wlo, D, Ls, t1, t2,
        t3, t4, mu, dmu, a, e, u, da, v, sc, q1, q2, q3, q4, end;
};

inline Layout layout(int64_t N, int64_t n, int64_t M) {
    Layout l;
    size_t o = 0;
    const size_t NNs = up((size_t)N * N), Nn = up((size_t)N * n), MN = up((size_t)M * N), Mn = up((size_t)M * n),
                 nn = up((size_t)n * n), mm = up((size_t)M * M);
    const size_t sq = mm > nn ? mm : nn;
    size_t big = Nn > MN ? Nn : MN;
    if (sq > big) big = sq;
#define TAKE(f, s) l.f = o; o += (s)
    TAKE(Lam, NNs); TAKE(NN, NNs); TAKE(NN2, NNs);
    TAKE(A, Nn); TAKE(Ar, Nn); TAKE(KA, Nn); TAKE(dA, Nn); TAKE(P, Nn);
    TAKE(Kzx, MN); TAKE(Kzxlo, MN); TAKE(At, MN); TAKE(dAt, MN); TAKE(dAtlo, MN); TAKE(dKzx, MN);
    TAKE(U, Mn); TAKE(DU, Mn); TAKE(dU, Mn); TAKE(tU, Mn);
    TAKE(H, nn); TAKE(G, nn); TAKE(C, nn); TAKE(Wc, nn); TAKE(Cinv, nn); TAKE(T1, nn); TAKE(dC, nn); TAKE(S2, nn); TAKE(Gd, nn);
    TAKE(Gd2, nn);
    TAKE(Lz, mm); TAKE(W, mm); TAKE(Wlo, mm); TAKE(D, mm); TAKE(Ls, mm);
    TAKE(t1, sq); TAKE(t2, sq); TAKE(t3, sq); TAKE(t4, up((size_t)M * (M + 1)));
    TAKE(mu, up((size_t)N)); TAKE(dmu, up((size_t)N));
    TAKE(a, up((size_t)n)); TAKE(e, up((size_t)n)); TAKE(u, up((size_t)n)); TAKE(da, up((size_t)n));
    TAKE(v, up((size_t)(M > n ? M : n)));
    TAKE(sc, 32);
    TAKE(q1, big); TAKE(q2, big);   // TF32 remainders / rounded copies of the operands of the current product (F32)
    TAKE(q3, big); TAKE(q4, big);   // the same for the side lane
#undef TAKE
    l.end = o;
    return l;
}

struct Scratch {   // byte offsets behind the typed part
    size_t plan, plan_bytes, solver, solver_bytes, solver2, solver2_bytes, gm, gm_bytes, island, end;
};

inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

inline Scratch scratch(const Layout& l, size_t es, int64_t N, int64_t n, int64_t M, int64_t ldz, int dtype) {
    Scratch s;
    size_t o = up256(l.end * es);
    s.plan = o;
    s.plan_bytes = dcgp_solve_plan_bytes(N, N, dtype);
    o += up256(s.plan_bytes);
    size_t sb = dcgp_potrf_workspace(N, dtype);
    const size_t cands[] = {dcgp_trsm_workspace(N, n, dtype), dcgp_potrf_workspace(M, dtype), dcgp_trtri_workspace(M, dtype),
                            dcgp_potrf_workspace(n, dtype), dcgp_trtri_workspace(n, dtype)};
    for (size_t c : cands)
        if (c > sb) sb = c;
    s.solver = o;
    s.solver_bytes = up256(sb);
    o += s.solver_bytes;
    s.solver2 = o;   // factorisation / inversion scratch of the side lane (K_ZZ)
    {
        size_t a2 = dcgp_potrf_workspace(M, dtype), b2 = dcgp_trtri_workspace(M, dtype);
        if (dtype == DCGP_F32) {   // FP64 island
            const size_t a3 = dcgp_potrf_workspace(M, DCGP_F64), b3 = dcgp_trtri_workspace(M, DCGP_F64);
            if (a3 > a2) a2 = a3;
            if (b3 > b2) b2 = b3;
        }
        s.solver2_bytes = up256(a2 > b2 ? a2 : b2);
    }
    o += s.solver2_bytes;
    s.gm = o;
    s.gm_bytes = up256(dcgp_gram_matmul_workspace(N, N, dtype));
    o += s.gm_bytes;
    s.island = o;
    if (dtype == DCGP_F32) o = island_layout(o, M, ldz).end;
    s.end = o;
    return s;
}

template <typename T>
int fwd_impl(const dcgp_cmp_problem* p, void* elbo, int* info, void* ws, size_t ws_bytes, int dtype, int flags,
             cudaStream_t st) {
    const int64_t N = p->N, n = p->n, M = p->M;
    const Layout l = layout(N, n, M);
    const Scratch s = scratch(l, sizeof(T), N, n, M, p->ldz, dtype);
    if (ws_bytes < s.end) return -5;
    T* w = reinterpret_cast<T*>(ws);
    unsigned char* wb = reinterpret_cast<unsigned char*>(ws);
    const bool f32 = sizeof(T) == 4;
    T* const q1 = f32 ? w + l.q1 : nullptr;
    T* const q2 = f32 ? w + l.q2 : nullptr;
    T* const q3 = f32 ? w + l.q3 : nullptr;
    T* const q4 = f32 ? w + l.q4 : nullptr;
    const T* const Ar = f32 ? w + l.Ar : nullptr;
    const int gf = f32 ? (flags | DCGP_GEMM_TF32X3) : flags;
    const dim3 gM((unsigned)((M + 255) / 256), (unsigned)M), gn((unsigned)((n + 255) / 256), (unsigned)n);
    DCGP_CU_RC(cudaMemsetAsync(w + l.sc, 0, 32 * sizeof(T), st));
    DCGP_CU_RC(cudaMemsetAsync(info, 0, 3 * sizeof(int), st));
    // --- fork: the inducing-point chain runs on the side lane beside the Lambda factorisation
    cudaStream_t ss = st;
    SideLane* lane = nullptr;
    if (side_enabled()) {
        lane = &g_side[st];
        RC(lane->init());
        ss = lane->s;
        DCGP_CU_RC(cudaEventRecord(lane->fork, st));
        DCGP_CU_RC(cudaStreamWaitEvent(ss, lane->fork, 0));
    }
    // --- K_ZZ + jitter I -> L_Z -> W;  A~ = W k(Z, x);  D
    RC(dcgp_gram(p->k_spec, p->Z, M, p->ldz, p->x, N, p->ldx, w + l.Kzx, N, nullptr, 0.0, dtype, 0, ss));
    T* const Kzxlo = f32 ? w + l.Kzxlo : nullptr;      // TF32 remainders of K_Zx: operand of three 3xTF32 products
    RC(split_keep<T>(w + l.Kzx, Kzxlo, M, N, ss));
    bool island = false;
    if constexpr (sizeof(T) == 4) {
        island = island_enabled();
        if (island) {
            const Island a = island_layout(s.island, M, p->ldz);
            double *Z64 = (double*)(wb + a.Z), *hyp = (double*)(wb + a.hyp), *L64 = (double*)(wb + a.L), *W64 = (double*)(wb + a.W),
                   *t64 = (double*)(wb + a.t1);
            const int64_t nz = M * p->ldz;
            f32_to_f64_kernel<<<(unsigned)((nz + 255) / 256), 256, 0, ss>>>((const float*)p->Z, Z64, nz);
            DCGP_CHECK_LAUNCH();
            spec_to_f64_kernel<<<DCGP_MAX_COMPONENTS, 32, 0, ss>>>(*p->k_spec, hyp);
            DCGP_CHECK_LAUNCH();
            const dcgp_kernel_spec k64 = spec_f64(p->k_spec, hyp);
            RC(dcgp_gram(&k64, Z64, M, p->ldz, Z64, M, p->ldz, L64, M, nullptr, p->jitter, DCGP_F64, DCGP_GRAM_SAME, ss));
            RC(dcgp_potrf(L64, M, M, info + 1, wb + s.solver2, s.solver2_bytes, DCGP_F64, flags, ss));
            tril_copy_kernel<double><<<gM, 256, 0, ss>>>(L64, M, t64, M);
            DCGP_CHECK_LAUNCH();
            DCGP_CU_RC(cudaMemcpyAsync(L64, t64, (size_t)M * M * 8, cudaMemcpyDeviceToDevice, ss));
            RC(dcgp_trtri(L64, M, M, W64, M, wb + s.solver2, s.solver2_bytes, DCGP_F64, flags, ss));
            f64_to_hilo_kernel<<<(unsigned)(((int64_t)M * M + 255) / 256), 256, 0, ss>>>(W64, (float*)(w + l.W), (float*)(w + l.Wlo),
                                                                                   (int64_t)M * M);
            DCGP_CHECK_LAUNCH();
            if (island_mode() == 2) {
                f64_to_hilo_kernel<<<(unsigned)(((int64_t)M * M + 255) / 256), 256, 0, ss>>>(L64, (float*)(w + l.Lz), (float*)(w + l.t1),
                                                                                       (int64_t)M * M);
                DCGP_CHECK_LAUNCH();
            }
            RC(gemm3<T>(0, 0, M, N, M, 1.0, w + l.W, M, nullptr, w + l.Kzx, N, Kzxlo, 0.0, w + l.At, N, q3, q4, dtype, gf, ss));
            RC(gemm3<T>(0, 0, M, N, M, 1.0, w + l.Wlo, M, nullptr, w + l.Kzx, N, Kzxlo, 1.0, w + l.At, N, q3, q4, dtype, gf, ss));
        }
    }
    if (!island) {
        RC(dcgp_gram(p->k_spec, p->Z, M, p->ldz, p->Z, M, p->ldz, w + l.Lz, M, nullptr, p->jitter, dtype, DCGP_GRAM_SAME, ss));
        RC(dcgp_potrf(w + l.Lz, M, M, info + 1, wb + s.solver2, s.solver2_bytes, dtype, gf, ss));
        tril_copy_kernel<T><<<gM, 256, 0, ss>>>(w + l.Lz, M, w + l.t1, M);
        DCGP_CHECK_LAUNCH();
        DCGP_CU_RC(cudaMemcpyAsync(w + l.Lz, w + l.t1, (size_t)M * M * sizeof(T), cudaMemcpyDeviceToDevice, ss));
        RC(dcgp_trtri(w + l.Lz, M, M, w + l.W, M, wb + s.solver2, s.solver2_bytes, dtype, gf, ss));
        RC(gemm3<T>(0, 0, M, N, M, 1.0, w + l.W, M, nullptr, w + l.Kzx, N, Kzxlo, 0.0, w + l.At, N, q3, q4, dtype, gf, ss));
    }
    tril_copy_kernel<T><<<gM, 256, 0, ss>>>((const T*)p->chol_var, p->ldc, w + l.Ls, M);
    DCGP_CHECK_LAUNCH();
    RC(gemm3<T>(0, 1, M, M, M, 1.0, w + l.Ls, M, nullptr, w + l.Ls, M, nullptr, 0.0, w + l.D, M, q3, q4, dtype, gf, ss));
    add_diag_kernel<T><<<(unsigned)((M + 255) / 256), 256, 0, ss>>>(w + l.D, M, T(-1));
    DCGP_CHECK_LAUNCH();
    // --- Lambda = l(y~, y~) + lbda N I, its factor, the solve plan, A = Lambda^{-1} l(y~, y)
    RC(dcgp_gram(p->l_spec, p->ext, N, p->ldext, p->ext, N, p->ldext, w + l.Lam, N, nullptr, p->lbda * (double)N, dtype,
                 DCGP_GRAM_SAME, st));
    RC(dcgp_potrf(w + l.Lam, N, N, info, wb + s.solver, s.solver_bytes, dtype, gf, st));
    RC(dcgp_solve_plan(w + l.Lam, N, N, wb + s.plan, s.plan_bytes, dtype, gf, st));
    RC(dcgp_gram(p->l_spec, p->ext, N, p->ldext, p->y, n, p->ldy, w + l.A, n, nullptr, 0.0, dtype, 0, st));
    RC(dcgp_trsm_planned(2, w + l.Lam, N, N, wb + s.plan, s.plan_bytes, w + l.A, n, n, wb + s.solver, s.solver_bytes, dtype, gf,
                         st));
    if (lane) {
        DCGP_CU_RC(cudaEventRecord(lane->join, ss));
        DCGP_CU_RC(cudaStreamWaitEvent(st, lane->join, 0));
    }
    // --- mu_q, a, H, K A, U, G
    RC(gemv<T>(1, N, M, w + l.At, N, (const T*)p->var_mean, w + l.mu, st));
    if (p->mean_const) {
        add_scalar_kernel<T><<<(unsigned)((N + 255) / 256), 256, 0, st>>>(w + l.mu, N, (const T*)p->mean_const);
        DCGP_CHECK_LAUNCH();
    }
    RC(gemv<T>(1, n, N, w + l.A, n, w + l.mu, w + l.a, st));
    if (f32) RC(dcgp_round_tf32(w + l.A, w + l.Ar, N, n, n, n, st));   // A is an operand of ten products
    // Second fork: G = A^T (K_xx + vj I) A + U^T D U (needed only by the trace term and by the backward pass) is built on the
    // side lane - U = A~ A, D U, K_xx A, two tall-K products - while the main lane goes from H straight to C and its factor,
    // a chain of single-CTA links that leaves the SMs to those products (serial: ~1.1 ms, overlapped: ~0.7 ms)
    if (lane) {
        DCGP_CU_RC(cudaEventRecord(lane->fork, st));      // A (rounded), A~, D are ready
        DCGP_CU_RC(cudaStreamWaitEvent(ss, lane->fork, 0));
    }
    RC(gemm1<T>(0, 0, M, n, N, 1.0, w + l.At, N, nullptr, w + l.A, n, Ar, 0.0, w + l.U, n, q3, q4, dtype, gf, ss));
    RC(gemm3<T>(0, 0, M, n, M, 1.0, w + l.D, M, nullptr, w + l.U, n, nullptr, 0.0, w + l.DU, n, q3, q4, dtype, gf, ss));
    RC(dcgp_gram_matmul(p->k_spec, p->x, N, p->ldx, p->x, N, p->ldx, nullptr, p->var_jitter, w + l.A, n, n, w + l.KA, n,
                        wb + s.gm, s.gm_bytes, dtype, DCGP_GRAM_SAME | flags, ss));
    RC(gemm1<T>(1, 0, n, n, N, 1.0, w + l.A, n, Ar, w + l.KA, n, nullptr, 0.0, w + l.G, n, q3, q4, dtype, gf, ss));
    RC(gemm3<T>(1, 0, n, n, M, 1.0, w + l.U, n, nullptr, w + l.DU, n, nullptr, 1.0, w + l.G, n, q3, q4, dtype, gf, ss));
    if (lane) DCGP_CU_RC(cudaEventRecord(lane->join, ss));
    RC(gemm1<T>(1, 0, n, n, N, 1.0, w + l.A, n, Ar, w + l.A, n, Ar, 0.0, w + l.H, n, q1, q2, dtype, gf, st));
    // --- C, e, the Gaussian terms through C^{-1} = Wc^T Wc
    build_c_kernel<T><<<gn, 256, 0, st>>>(w + l.H, n, (const T*)p->ind_noise, (const T*)p->noise, w + l.C, (const T*)p->z,
                                          w + l.a, w + l.e);
    DCGP_CHECK_LAUNCH();
    RC(dcgp_potrf(w + l.C, n, n, info + 2, wb + s.solver, s.solver_bytes, dtype, gf, st));
    tril_copy_kernel<T><<<gn, 256, 0, st>>>(w + l.C, n, w + l.t1, n);
    DCGP_CHECK_LAUNCH();
    RC(dcgp_logdet_chol(w + l.t1, n, n, w + l.sc + 2, dtype, st));
    RC(dcgp_trtri(w + l.t1, n, n, w + l.Wc, n, wb + s.solver, s.solver_bytes, dtype, gf, st));
    RC(gemm3<T>(1, 0, n, n, n, 1.0, w + l.Wc, n, nullptr, w + l.Wc, n, nullptr, 0.0, w + l.Cinv, n, q1, q2, dtype, gf, st));
    RC(gemv<T>(0, n, n, w + l.Cinv, n, w + l.e, w + l.u, st));
    RC(dcgp_whitened_kl(p->var_mean, w + l.Ls, M, M, w + l.sc + 1, w + l.t4 + (size_t)M * M, w + l.t4, M, dtype, st));
    if (lane) DCGP_CU_RC(cudaStreamWaitEvent(st, lane->join, 0));
    RC(dot<T>(w + l.u, w + l.e, n, w + l.sc + 3, st));
    RC(dot<T>(w + l.Cinv, w + l.G, n * n, w + l.sc + 4, st));
    ell_kernel<T><<<1, 1, 0, st>>>(n, w + l.sc + 2, w + l.sc + 1, (T)(p->beta / p->num_data), w + l.sc, (T*)elbo);
    DCGP_CHECK_LAUNCH();
    return 0;
}

template <typename T>
int bwd_impl(const dcgp_cmp_problem* p, double scale, const dcgp_cmp_grads* g, void* ws, size_t ws_bytes, int dtype, int flags,
             cudaStream_t st) {
    const int64_t N = p->N, n = p->n, M = p->M;
    const Layout l = layout(N, n, M);
    const Scratch s = scratch(l, sizeof(T), N, n, M, p->ldz, dtype);
    if (ws_bytes < s.end) return -5;
    T* w = reinterpret_cast<T*>(ws);
    unsigned char* wb = reinterpret_cast<unsigned char*>(ws);
    const bool f32 = sizeof(T) == 4;
    T* const q1 = f32 ? w + l.q1 : nullptr;
    T* const q2 = f32 ? w + l.q2 : nullptr;
    T* const q3 = f32 ? w + l.q3 : nullptr;
    T* const q4 = f32 ? w + l.q4 : nullptr;
    const T* const Ar = f32 ? w + l.Ar : nullptr;
    const int gf = f32 ? (flags | DCGP_GEMM_TF32X3) : flags;
    const T c0 = (T)(-0.5 * scale / (double)n), s2 = (T)(-scale * p->beta / p->num_data);
    const dim3 gM((unsigned)((M + 255) / 256), (unsigned)M), gn((unsigned)((n + 255) / 256), (unsigned)n);
    const int nk = p->k_spec->n_components, nl = p->l_spec->n_components;
    if (g->dk_ls) DCGP_CU_RC(cudaMemsetAsync(g->dk_ls, 0, (size_t)nk * DCGP_MAX_DIMS * sizeof(T), st));
    if (g->dk_os) DCGP_CU_RC(cudaMemsetAsync(g->dk_os, 0, (size_t)nk * sizeof(T), st));
    if (g->dl_ls) DCGP_CU_RC(cudaMemsetAsync(g->dl_ls, 0, (size_t)nl * DCGP_MAX_DIMS * sizeof(T), st));
    if (g->dl_os) DCGP_CU_RC(cudaMemsetAsync(g->dl_os, 0, (size_t)nl * sizeof(T), st));
    if (g->dZ) {
        if (g->lddz < 1) return -3;
        DCGP_CU_RC(cudaMemsetAsync(g->dZ, 0, (size_t)M * g->lddz * sizeof(T), st));
    }
    // --- bag-level block: dC, dG, da, noise adjoints
    RC(gemm3<T>(0, 0, n, n, n, 1.0, w + l.Cinv, n, nullptr, w + l.G, n, nullptr, 0.0, w + l.t1, n, q1, q2, dtype, gf, st));
    RC(gemm3<T>(0, 0, n, n, n, 1.0, w + l.t1, n, nullptr, w + l.Cinv, n, nullptr, 0.0, w + l.T1, n, q1, q2, dtype, gf, st));
    dc_kernel<T><<<gn, 256, 0, st>>>(w + l.Cinv, w + l.T1, w + l.u, n, c0, (const T*)p->ind_noise, w + l.dC, w + l.S2,
                                     w + l.Gd, w + l.Gd2);
    DCGP_CHECK_LAUNCH();
    DCGP_CU_RC(cudaMemsetAsync(w + l.sc + 8, 0, 2 * sizeof(T), st));
    dscal_kernel<T><<<1, 1024, 0, st>>>(w + l.dC, w + l.u, n, c0, w + l.da, w + l.sc);
    DCGP_CHECK_LAUNCH();
    RC(dot<T>(w + l.dC, w + l.H, n * n, w + l.sc + 9, st));
    dscal_out_kernel<T><<<1, 1, 0, st>>>(w + l.sc, (T*)g->dnoise, (T*)g->dind_noise);
    DCGP_CHECK_LAUNCH();
    // --- early fork: the adjoints of L_S and of K_xx need only dG - they go to the side lane at once and run beside the
    // critical path dU -> dA -> Lambda^{-1} dA (the sweeps of that solve leave SMs idle between their dependent products)
    cudaStream_t ss = st;
    SideLane* lane = nullptr;
    if (side_enabled()) {
        lane = &g_side[st];
        RC(lane->init());
        ss = lane->s;
        DCGP_CU_RC(cudaEventRecord(lane->fork, st));
        DCGP_CU_RC(cudaStreamWaitEvent(ss, lane->fork, 0));
    }
    // dD = U dG U^T -> dL_S
    if (g->dchol_var) {
        RC(gemm3<T>(0, 0, M, n, n, 1.0, w + l.U, n, nullptr, w + l.Gd, n, nullptr, 0.0, w + l.tU, n, q3, q4, dtype, gf, ss));
        RC(gemm3<T>(0, 1, M, M, n, 1.0, w + l.tU, n, nullptr, w + l.U, n, nullptr, 0.0, w + l.t1, M, q3, q4, dtype, gf, ss));
        RC(gemm3<T>(0, 0, M, M, M, 1.0, w + l.t1, M, nullptr, w + l.Ls, M, nullptr, 0.0, w + l.t2, M, q3, q4, dtype, gf, ss));
        tril_axpby_kernel<T><<<gM, 256, 0, ss>>>(w + l.t2, M, T(2), w + l.t4, M, s2, 0, (T*)g->dchol_var, g->lddc);
        DCGP_CHECK_LAUNCH();
    }
    // dK_xx = (A dG) A^T -> individuals-kernel hyper-parameters (its own N x N buffer: the main lane forms dLambda meanwhile)
    if (g->dk_ls || g->dk_os) {
        RC(gemm1<T>(0, 0, N, n, n, 1.0, w + l.A, n, Ar, w + l.Gd, n, nullptr, 0.0, w + l.P, n, q3, q4, dtype, gf, ss));
        RC(gemm1<T>(0, 1, N, N, n, 1.0, w + l.P, n, nullptr, w + l.A, n, Ar, 0.0, w + l.NN2, N, q3, q4, dtype, gf, ss));
        RC(dcgp_gram_bwd(p->k_spec, p->x, N, p->ldx, p->x, N, p->ldx, w + l.NN2, N, g->dk_ls, g->dk_os, nullptr, 0, dtype,
                         DCGP_GRAM_SAME, ss));
    }
    // --- main lane: dU = D U (2 dG)
    RC(gemm3<T>(0, 0, M, n, n, 1.0, w + l.DU, n, nullptr, w + l.Gd2, n, nullptr, 0.0, w + l.dU, n, q1, q2, dtype, gf, st));
    // --- second fork point: dU is complete.  Everything that feeds the inducing-point side of the graph (dmu, dA~, dm, the
    // mean constant, then the adjoint of the K_ZZ chain) follows on the side lane; the main lane keeps only dA and the
    // adjoint solve with Lambda
    if (lane) {
        DCGP_CU_RC(cudaEventRecord(lane->fork, st));
        DCGP_CU_RC(cudaStreamWaitEvent(ss, lane->fork, 0));
    }
    // --- main lane: dA = A (2 s_ind dC) + K A (2 dG) + mu da^T + A~^T dU
    RC(gemm1<T>(0, 0, N, n, n, 1.0, w + l.A, n, Ar, w + l.S2, n, nullptr, 0.0, w + l.dA, n, q1, q2, dtype, gf, st));
    RC(gemm1<T>(0, 0, N, n, n, 1.0, w + l.KA, n, nullptr, w + l.Gd2, n, nullptr, 1.0, w + l.dA, n, q1, q2, dtype, gf, st));
    RC(gemm1<T>(1, 0, N, n, M, 1.0, w + l.At, N, nullptr, w + l.dU, n, nullptr, 1.0, w + l.dA, n, q1, q2, dtype, gf, st));
    {
        const dim3 gr((unsigned)((n + 255) / 256), (unsigned)N);
        rank1_kernel<T><<<gr, 256, 0, st>>>(w + l.dA, N, n, w + l.mu, w + l.da);
        DCGP_CHECK_LAUNCH();
    }
    // --- side lane: dmu = A da;  dA~ = dU A^T + m dmu^T;  dm;  mean constant
    RC(gemv<T>(0, N, n, w + l.A, n, w + l.da, w + l.dmu, ss));
    RC(gemm1<T>(0, 1, M, N, n, 1.0, w + l.dU, n, nullptr, w + l.A, n, Ar, 0.0, w + l.dAt, N, q3, q4, dtype, gf, ss));
    {
        const dim3 gr((unsigned)((N + 255) / 256), (unsigned)M);
        rank1_kernel<T><<<gr, 256, 0, ss>>>(w + l.dAt, M, N, (const T*)p->var_mean, w + l.dmu);
        DCGP_CHECK_LAUNCH();
    }
    RC(gemv<T>(0, M, N, w + l.At, N, w + l.dmu, w + l.v, ss));
    tail_kernel<T><<<1, 1024, 0, ss>>>(w + l.v, w + l.t4 + (size_t)M * M, M, s2, (T*)g->dvar_mean, w + l.dmu, N,
                                       (T*)g->dmean_const);
    DCGP_CHECK_LAUNCH();
    // --- dK_Zx = W^T dA~;  dW = tril(dA~ K_Zx^T) -> dK_ZZ
    bool island = false;
    if constexpr (sizeof(T) == 4) island = island_enabled();
    T* const dAtlo = f32 ? w + l.dAtlo : nullptr;      // remainders of dA~: operand of three 3xTF32 products
    const T* const Kzxlo = f32 ? w + l.Kzxlo : nullptr;
    RC(split_keep<T>(w + l.dAt, dAtlo, M, N, ss));
    RC(gemm3<T>(1, 0, M, N, M, 1.0, w + l.W, M, nullptr, w + l.dAt, N, dAtlo, 0.0, w + l.dKzx, N, q3, q4, dtype, gf, ss));
    if (island)   // (still true in mode 2 here: W is a hi + lo pair whenever the forward island ran)
        RC(gemm3<T>(1, 0, M, N, M, 1.0, w + l.Wlo, M, nullptr, w + l.dAt, N, dAtlo, 1.0, w + l.dKzx, N, q3, q4, dtype, gf, ss));
    RC(dcgp_gram_bwd(p->k_spec, p->Z, M, p->ldz, p->x, N, p->ldx, w + l.dKzx, N, g->dk_ls, g->dk_os, g->dZ, g->lddz, dtype, 0,
                     ss));
    RC(gemm3<T>(0, 1, M, M, N, 1.0, w + l.dAt, N, dAtlo, w + l.Kzx, N, Kzxlo, 0.0, w + l.t3, M, q3, q4, dtype, gf, ss));
    tril_axpby_kernel<T><<<gM, 256, 0, ss>>>(w + l.t3, M, T(1), nullptr, 0, T(0), 0, w + l.t2, M);
    DCGP_CHECK_LAUNCH();
    if constexpr (sizeof(T) == 4) {
        if (island_mode() == 2) island = false;    // adjoint chain below in FP32 / 3xTF32 on the rounded factor
    }
    if constexpr (sizeof(T) == 4) {
        if (island) {
            // the adjoint of W = chol(K_ZZ)^{-1} in FP64, as products with W (functional._CholInverseFn)
            const Island a = island_layout(s.island, M, p->ldz);
            double *Z64 = (double*)(wb + a.Z), *hyp = (double*)(wb + a.hyp), *L64 = (double*)(wb + a.L), *W64 = (double*)(wb + a.W),
                   *u1 = (double*)(wb + a.t1), *u2 = (double*)(wb + a.t2), *u3 = (double*)(wb + a.t3), *dZ64 = (double*)(wb + a.dZ),
                   *dth = (double*)(wb + a.dth);
            const int64_t mm = (int64_t)M * M;
            f32_to_f64_kernel<<<(unsigned)((mm + 255) / 256), 256, 0, ss>>>((const float*)(w + l.t2), u2, mm);
            DCGP_CHECK_LAUNCH();
            RC(dcgp_gemm(1, 0, M, M, M, 1.0, W64, M, u2, M, 0.0, u1, M, DCGP_F64, flags, ss));
            RC(dcgp_gemm(0, 1, M, M, M, -1.0, u1, M, W64, M, 0.0, u3, M, DCGP_F64, flags | DCGP_GEMM_LOWER, ss));   // only its lower triangle is read
            tril_axpby_kernel<double><<<gM, 256, 0, ss>>>(u3, M, 1.0, nullptr, 0, 0.0, 0, u2, M);
            DCGP_CHECK_LAUNCH();
            RC(dcgp_gemm(1, 0, M, M, M, 1.0, L64, M, u2, M, 0.0, u1, M, DCGP_F64, flags | DCGP_GEMM_LOWER, ss));    // likewise
            tril_axpby_kernel<double><<<gM, 256, 0, ss>>>(u1, M, 1.0, nullptr, 0, 0.0, 1, u2, M);
            DCGP_CHECK_LAUNCH();
            RC(dcgp_gemm(1, 0, M, M, M, 1.0, W64, M, u2, M, 0.0, u1, M, DCGP_F64, flags, ss));
            RC(dcgp_gemm(0, 0, M, M, M, 1.0, u1, M, W64, M, 0.0, u3, M, DCGP_F64, flags, ss));
            symmetrize_kernel<double><<<gM, 256, 0, ss>>>(u3, M, u2);
            DCGP_CHECK_LAUNCH();
            const dcgp_kernel_spec k64 = spec_f64(p->k_spec, hyp);
            DCGP_CU_RC(cudaMemsetAsync(dth, 0, 512, ss));
            DCGP_CU_RC(cudaMemsetAsync(dZ64, 0, (size_t)M * p->ldz * 8, ss));
            RC(dcgp_gram_bwd(&k64, Z64, M, p->ldz, Z64, M, p->ldz, u2, M, g->dk_ls ? dth : nullptr, g->dk_os ? dth + 32 : nullptr,
                             g->dZ ? dZ64 : nullptr, p->ldz, DCGP_F64, DCGP_GRAM_SAME, ss));
            const int nkc = p->k_spec->n_components;
            acc_f64_kernel<<<dim3(1, (unsigned)nkc), 32, 0, ss>>>(dth, DCGP_MAX_DIMS, (float*)g->dk_ls, DCGP_MAX_DIMS, nkc,
                                                                 DCGP_MAX_DIMS);
            DCGP_CHECK_LAUNCH();
            acc_f64_kernel<<<dim3(1, 1), 32, 0, ss>>>(dth + 32, nkc, (float*)g->dk_os, nkc, 1, nkc);
            DCGP_CHECK_LAUNCH();
            if (g->dZ) {
                // columns of Z that the kernel does not address have zero gradient in both buffers
                acc_f64_kernel<<<dim3((unsigned)((p->ldz + 31) / 32), (unsigned)M), 32, 0, ss>>>(
                    dZ64, p->ldz, (float*)g->dZ, g->lddz, M, p->ldz < g->lddz ? p->ldz : g->lddz);
                DCGP_CHECK_LAUNCH();
            }
        }
    }
    if (!island) {
        RC(gemm3<T>(1, 0, M, M, M, 1.0, w + l.W, M, nullptr, w + l.t2, M, nullptr, 0.0, w + l.t1, M, q3, q4, dtype, gf, ss));
        RC(gemm3<T>(0, 1, M, M, M, -1.0, w + l.t1, M, nullptr, w + l.W, M, nullptr, 0.0, w + l.t3, M, q3, q4, dtype, gf, ss));
        tril_axpby_kernel<T><<<gM, 256, 0, ss>>>(w + l.t3, M, T(1), nullptr, 0, T(0), 0, w + l.t2, M);
        DCGP_CHECK_LAUNCH();
        RC(gemm3<T>(1, 0, M, M, M, 1.0, w + l.Lz, M, nullptr, w + l.t2, M, nullptr, 0.0, w + l.t1, M, q3, q4, dtype, gf, ss));
        tril_axpby_kernel<T><<<gM, 256, 0, ss>>>(w + l.t1, M, T(1), nullptr, 0, T(0), 1, w + l.t2, M);
        DCGP_CHECK_LAUNCH();
        RC(gemm3<T>(1, 0, M, M, M, 1.0, w + l.W, M, nullptr, w + l.t2, M, nullptr, 0.0, w + l.t1, M, q3, q4, dtype, gf, ss));
        RC(gemm3<T>(0, 0, M, M, M, 1.0, w + l.t1, M, nullptr, w + l.W, M, nullptr, 0.0, w + l.t3, M, q3, q4, dtype, gf, ss));
        symmetrize_kernel<T><<<gM, 256, 0, ss>>>(w + l.t3, M, w + l.t2);
        DCGP_CHECK_LAUNCH();
        RC(dcgp_gram_bwd(p->k_spec, p->Z, M, p->ldz, p->Z, M, p->ldz, w + l.t2, M, g->dk_ls, g->dk_os, g->dZ, g->lddz, dtype,
                         DCGP_GRAM_SAME, ss));
    }
    // --- X = Lambda^{-1} dA (in place);  d l(y~, y) = X;  dLambda = -X A^T
    RC(dcgp_trsm_planned(2, w + l.Lam, N, N, wb + s.plan, s.plan_bytes, w + l.dA, n, n, wb + s.solver, s.solver_bytes, dtype, gf,
                         st));
    if (g->dl_ls || g->dl_os) {
        RC(dcgp_gram_bwd(p->l_spec, p->ext, N, p->ldext, p->y, n, p->ldy, w + l.dA, n, g->dl_ls, g->dl_os, nullptr, 0, dtype, 0,
                         st));
        RC(gemm1<T>(0, 1, N, N, n, -1.0, w + l.dA, n, nullptr, w + l.A, n, Ar, 0.0, w + l.NN, N, q1, q2, dtype, gf, st));
        RC(dcgp_gram_bwd(p->l_spec, p->ext, N, p->ldext, p->ext, N, p->ldext, w + l.NN, N, g->dl_ls, g->dl_os, nullptr, 0, dtype,
                         DCGP_GRAM_SAME, st));
    }
    if (lane) {
        DCGP_CU_RC(cudaEventRecord(lane->join, ss));
        DCGP_CU_RC(cudaStreamWaitEvent(st, lane->join, 0));
    }
    return 0;
}

inline int check(const dcgp_cmp_problem* p) {
    if (!p || !p->k_spec || !p->l_spec || !p->Z || !p->x || !p->ext || !p->y || !p->z || !p->var_mean || !p->chol_var ||
        !p->noise || !p->ind_noise)
        return -1;
    if (p->M < 1 || p->N < 1 || p->n < 1 || p->ldz < 1 || p->ldx < 1 || p->ldext < 1 || p->ldy < 1 || p->ldc < p->M) return -1;
    if (!(p->num_data > 0.0)) return -1;
    return 0;
}

}  // namespace

extern "C" size_t dcgp_cmp_elbo_workspace(int64_t N, int64_t n, int64_t M, int64_t ldz, int dtype) {
    if (N < 1 || n < 1 || M < 1 || ldz < 1) return 0;
    const size_t es = dtype == DCGP_F64 ? 8 : 4;
    return scratch(layout(N, n, M), es, N, n, M, ldz, dtype).end;
}

extern "C" int dcgp_cmp_elbo_fwd(const dcgp_cmp_problem* p, void* elbo, int* info, void* workspace, size_t workspace_bytes,
                                 int dtype, int flags, void* stream) {
    if (check(p)) return -1;
    if (!elbo) return -2;
    if (!info) return -3;
    if (!workspace) return -4;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dtype == DCGP_F64 ? fwd_impl<double>(p, elbo, info, workspace, workspace_bytes, dtype, flags, st)
                             : fwd_impl<float>(p, elbo, info, workspace, workspace_bytes, dtype, flags, st);
}

extern "C" int dcgp_cmp_elbo_bwd(const dcgp_cmp_problem* p, double scale, const dcgp_cmp_grads* grads, void* workspace,
                                 size_t workspace_bytes, int dtype, int flags, void* stream) {
    if (check(p)) return -1;
    if (!grads) return -3;
    if (!workspace) return -4;
    if (grads->dchol_var && grads->lddc < p->M) return -3;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    return dtype == DCGP_F64 ? bwd_impl<double>(p, scale, grads, workspace, workspace_bytes, dtype, flags, st)
                             : bwd_impl<float>(p, scale, grads, workspace, workspace_bytes, dtype, flags, st);
}
