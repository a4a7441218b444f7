/*
 * dcgp.h — C ABI of libdcgp.so: the B200 (sm_100a) hot path of the
 * deconditional-GP framework (Gram assembly -> Cholesky / solves -> bag ELBO).
 *
 * The reference (shahineb/deconditional-downscaling) has no FFI of its own:
 * every dense operation on this path is issued through gpytorch 1.4.0
 * LazyTensor algebra into ATen/cuBLAS/cuSOLVER.  Each entry point below
 * replaces the library call sequence triggered at the cited reference line;
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the
 * reference side.
 *
 * Conventions
 *   - plain C: pointers + sizes, no C++/torch types.  All data pointers are
 *     DEVICE pointers unless stated; the caller owns every buffer (inputs,
 *     outputs, workspace): the library never allocates device memory and
 *     never synchronises the stream.
 *   - host-side state the library does keep (all of it created lazily, none
 *     of it visible in results): per calling thread and per caller stream a
 *     set of three internal CUDA streams (high / middle / low priority) and
 *     their events - the "lanes" of the look-ahead factorisation and solves -
 *     plus one side lane (a stream and two events) of dcgp_cmp_elbo_fwd/bwd,
 *     all forked from and joined back into the caller's stream with events
 *     only, so calls stay capturable in a CUDA graph; `cudaFuncSetAttribute` latches per kernel;
 *     the SM count; tuning switches read once from the environment (DCGP_*,
 *     DESIGN.md section 7); a relaxed-atomic diagnostic launch counter.
 *     Calls are safe from several host threads (thread-local lanes) and from
 *     several streams of one thread (one lane set per stream); one process
 *     drives one device (the state is keyed by stream, not by device).
 *   - matrices are ROW-MAJOR with an explicit leading dimension (elements
 *     between consecutive rows).
 *   - dtype: DCGP_F64 computes and stores in double (FP64 tensor/vector
 *     pipes); DCGP_F32 stores float and contracts on TF32 tensor cores.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - return value: 0 = success; -k = argument k (1-based) invalid;
 *     <= -1000 = -(1000 + cudaError_t).  Factorisation failures are reported
 *     LAPACK-style through a device `int* info` (0 ok, i>0: leading minor i
 *     not positive definite) so no host synchronisation is forced.
 */
#ifndef DCGP_H_
#define DCGP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DCGP_API __attribute__((visibility("default")))
#else
#define DCGP_API
#endif

#define DCGP_F64 0
#define DCGP_F32 1

#define DCGP_KIND_RBF 0      /* os * exp(-r^2/2)                       gpytorch RBFKernel     */
#define DCGP_KIND_MATERN32 1 /* os * (1+sqrt3 r) exp(-sqrt3 r)         gpytorch MaternKernel(nu=1.5) */
#define DCGP_KIND_MATERN12 2 /* os * exp(-r)                           MaternKernel(nu=0.5)   */
#define DCGP_KIND_MATERN52 3 /* os * (1+sqrt5 r+5/3 r^2) exp(-sqrt5 r) MaternKernel(nu=2.5)   */

#define DCGP_MAX_COMPONENTS 4
#define DCGP_MAX_DIMS 8 /* active dims per component; total over components <= 16 */

/* flags */
#define DCGP_GRAM_SAME 1        /* x2 is x1 (enables +diag and the symmetric fast path)  */
#define DCGP_GRAM_OUT_F64 2     /* dtype F32 inputs/math, store K as double (fp64 island) */
#define DCGP_GRAM_BWD_FUSE 512  /* dcgp_gram_bwd_lowrank: contract G tile by tile inside the tcgen05 product (G never stored) */
#define DCGP_GRAM_NO_FUSE 256   /* dcgp_gram_matmul: keep the panel path (Gram rows into the workspace + GEMM) even where
                                   the fused tcgen05 kernel would take the product (A/B runs) */
#define DCGP_GEMM_LOWER 1       /* only tiles touching the lower triangle of C are computed/written */
#define DCGP_GEMM_NO_SPLITK 2   /* C aliases an input (in-place): the K loop is never split across CTAs */
#define DCGP_GEMM_INPLACE 16    /* C aliases A (then n <= 128) or B (then m <= 128): single tile along that dimension */
#define DCGP_GEMM_A_LOWER 32  /* op(A) is lower triangular (zero for k > row): k-tiles past the diagonal are skipped */
#define DCGP_GEMM_A_UPPER 64  /* op(A) is upper triangular (zero for k < row) */
#define DCGP_GEMM_NO_TCGEN05 8  /* F32 only: keep the contraction on the mma.sync kernel (A/B testing) */
#define DCGP_GEMM_TF32X3 4      /* F32 only: error-compensated 3xTF32 (~FP32 accuracy) instead of single-pass
                                   TF32; also honoured in the flags of dcgp_potrf / dcgp_trsm / dcgp_potrs */

/* One ScaleKernel(base kernel with ARD lengthscales, active_dims) term of an
 * additive kernel.  lengthscale / outputscale are DEVICE pointers to the
 * already-transformed (softplus) values in the compute dtype. */
typedef struct dcgp_component {
    int32_t kind;
    int32_t ndims;
    int32_t dims[DCGP_MAX_DIMS];
    const void* lengthscale; /* [ndims] */
    const void* outputscale; /* [1]     */
} dcgp_component;

typedef struct dcgp_kernel_spec {
    int32_t n_components;
    int32_t reserved;
    dcgp_component comp[DCGP_MAX_COMPONENTS];
} dcgp_kernel_spec;

DCGP_API int dcgp_version(void);
/* compiled SM architecture * 10 (1000 for sm_100a) */
DCGP_API int dcgp_arch(void);
/* kernels launched by the library since it was loaded (diagnostic counter, relaxed atomic) */
DCGP_API unsigned long long dcgp_launch_count(void);

/* ---- Gram assembly ---------------------------------------------------
 * K[i,j] = sum_c os_c * phi_c(|| (x1_i - x2_j)[dims_c] / l_c ||)  (+ diag on i==j if SAME)
 * Replaces gpytorch Kernel.__call__ -> covar_dist -> exp/matern post-processing
 * (reference call sites src/models/cmp.py:60,65,98,99; src/models/variational_cmp.py:79,99,103;
 *  src/kernels/cme_aggregate_kernel.py:38-39; src/means/cme_aggregate_mean.py:36).
 * diag_dev: optional device scalar added to the diagonal (e.g. sigma_ind^2),
 * diag_const: host constant added to the diagonal (lbda*N, jitter). */
DCGP_API int dcgp_gram(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2,
              int64_t ldx2, void* K, int64_t ldk, const void* diag_dev, double diag_const, int dtype, int flags,
              void* stream);

/* Gram adjoint for a LOW-RANK upstream gradient: dls / dos += sum_ij G_ij d k(x1_i, x2_j) / d theta with
 * G = alpha P Q^T (P: n1 x k, Q: n2 x k).  In the ELBO step the adjoints of Lambda = l(y~, y~) + lbda N I and of K_xx
 * arrive in exactly this form (torch autograd forms them densely: the backward of src/models/variational_cmp.py:99-103 and
 * of src/likelihoods/cmp_likelihood.py:58-61).  Default: G goes into the workspace (dcgp_gram_bwd_lowrank_workspace bytes)
 * by one product and dcgp_gram_bwd contracts it.  With DCGP_GRAM_BWD_FUSE (F32, single-pass TF32 on operands rounded to
 * nearest, at most 8 coordinate slots, n1, n2 >= 128): one tcgen05 launch whose epilogue contracts each accumulator tile
 * with the kernel derivatives - G is never stored and the workspace is not touched (536 MB less at N = 8192 x 2), but
 * the eight epilogue warps make it slower than the two-call route on B200 (8192^2, k = 1024: 368 vs 328 us for an RBF
 * kernel, 462 vs 431 us for Matern + RBF), so it is a memory option, not the default. */
DCGP_API size_t dcgp_gram_bwd_lowrank_workspace(int64_t n1, int64_t n2, int dtype);
DCGP_API int dcgp_gram_bwd_lowrank(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2,
                                   int64_t n2, int64_t ldx2, const void* P, int64_t ldp, const void* Q, int64_t ldq,
                                   int64_t k, double alpha, void* dls, void* dos, void* workspace, size_t workspace_bytes,
                                   int dtype, int flags, void* stream);

/* out = (k(x1, x2) [+ (diag_const + *diag_dev) I with DCGP_GRAM_SAME]) W, W is n2 x m: the fused K W of
 * SURVEY.md section 8a row 6 (src/kernels/cme_aggregate_kernel.py:58-70 forms K R by materialising K).  The
 * Gram matrix exists only as row panels in the workspace (a single panel up to 1 GiB, 512 MiB panels beyond). */
DCGP_API size_t dcgp_gram_matmul_workspace(int64_t n1, int64_t n2, int dtype);
DCGP_API int dcgp_gram_matmul(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2,
                     int64_t n2, int64_t ldx2, const void* diag_dev, double diag_const, const void* W, int64_t m,
                     int64_t ldw, void* out, int64_t ldo, void* workspace, size_t workspace_bytes, int dtype, int flags,
                     void* stream);

/* Backward of dcgp_gram for a dense upstream gradient G (n1 x n2):
 *   dls[c*DCGP_MAX_DIMS + k] += sum_ij G_ij dK_ij/dl_{c,k}
 *   dos[c]                   += sum_ij G_ij dK_ij/dos_c
 *   dx1[i, col]              += sum_j  G_ij dK_ij/dx1_{i,col}      (optional, may be NULL)
 * With DCGP_GRAM_SAME, x2 is x1 and dx1 receives both argument slots
 * (G_ij + G_ji), i.e. the gradient w.r.t. the shared input.
 * Outputs are ACCUMULATED into (caller zeroes them).  Replaces autograd through the
 * same gpytorch ops (hyper-parameter leaves listed in SURVEY.md §8a). */
DCGP_API int dcgp_gram_bwd(const dcgp_kernel_spec* spec, const void* x1, int64_t n1, int64_t ldx1, const void* x2, int64_t n2,
                  int64_t ldx2, const void* G, int64_t ldg, void* dls, void* dos, void* dx1, int64_t lddx1, int dtype,
                  int flags, void* stream);

/* ---- dense contraction --------------------------------------------------
 * C = alpha * op(A) op(B) + beta * C ; op(X) = X or X^T ; C is m x n, inner dim k.
 * F64: DMMA (mma.sync m8n8k4 f64); F32: TF32 tensor cores.  With DCGP_GEMM_LOWER only
 * tiles intersecting the lower triangle are touched (SYRK trailing update).
 * Replaces the MatmulLazyTensor / torch.matmul products of
 * src/kernels/cme_aggregate_kernel.py:63-70, src/likelihoods/cmp_likelihood.py:61-73,
 * src/variational/variational_strategy.py:44-65. */
DCGP_API int dcgp_gemm(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
              const void* B, int64_t ldb, double beta, void* C, int64_t ldc, int dtype, int flags, void* stream);

/* C (lower triangle only) = alpha * op(A) op(A)^T + beta * C; trans = 0: A is n x k, trans = 1: A is k x n.
 * The trailing updates of the factorisation and the A^T A / U^T U terms of the ELBO
 * (src/likelihoods/cmp_likelihood.py:61-64) are of this form. */
DCGP_API int dcgp_syrk(int trans, int64_t n, int64_t k, double alpha, const void* A, int64_t lda, double beta, void* C,
              int64_t ldc, int dtype, int flags, void* stream);

/* 3xTF32 with the tensor core's own operand rounding: tcgen05.mma kind::tf32 reads an FP32 operand
 * truncated to TF32 (x_hi); dcgp_split_lo writes what it drops, x_lo = x - x_hi.  dcgp_gemm_ex takes
 * the remainder matrices of A and B (same shapes and leading dimensions, FP32 only, may be NULL) and
 * then evaluates A_lo B_hi + A_hi B_lo + A_hi B_hi in one tcgen05 launch (~FP32 accuracy at a third
 * of the TF32 rate); Clo (may be NULL) receives the remainders of the result for a later product.
 * Used for the whitening products L_Z^{-1} K_Zx of the SVGP (gpytorch variational_strategy.py:147-163
 * keeps that solve in double precision; here the factor and its inverse are FP64 and the product runs
 * on an FP32 hi/lo pair of the inverse). */
DCGP_API int dcgp_split_lo(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, void* stream);
/* dst = src rounded to the nearest TF32 value (cvt.rna.tf32.f32), rows x cols FP32, row-major.  tcgen05.mma reads FP32
 * operands by truncation (round toward zero): on raw data a single-pass TF32 product is biased by about -7e-4 relative
 * (the mean of the two truncation errors) - three chained products already leave the 1e-3 tolerance.  Operands rounded
 * to nearest first make the product unbiased, which is what cuBLAS TF32 - the reference's GPU path
 * (torch.backends.cuda.matmul.allow_tf32) - does inside its kernels. */
DCGP_API int dcgp_round_tf32(const void* src, void* dst, int64_t rows, int64_t cols, int64_t lds, int64_t ldd, void* stream);
DCGP_API int dcgp_gemm_ex(int transa, int transb, int64_t m, int64_t n, int64_t k, double alpha, const void* A, int64_t lda,
                 const void* Alo, const void* B, int64_t ldb, const void* Blo, double beta, void* C, int64_t ldc,
                 void* Clo, int dtype, int flags, void* stream);

/* ---- Cholesky and triangular solves -----------------------------------------
 * dcgp_potrf: A (n x n, lower triangle read) is overwritten by L (A = L L^T) in its lower
 * triangle; blocked right-looking, panel on CUDA cores, TRSM + SYRK trailing update on
 * tensor cores.  *info (device) = 0 or the 1-based index of the first non-positive pivot.
 * Replaces psd_safe_cholesky / root_inv_decomposition / inv_quad_logdet Cholesky
 * (src/models/cmp.py:66-67, src/models/variational_cmp.py:99-100,
 *  src/variational/variational_strategy.py:35, src/likelihoods/cmp_likelihood.py:58,69,73). */
DCGP_API size_t dcgp_potrf_workspace(int64_t n, int dtype);
DCGP_API int dcgp_potrf(void* A, int64_t n, int64_t lda, int* info, void* workspace, size_t workspace_bytes, int dtype,
               int flags, void* stream);

/* dcgp_trsm: solves op(L) X = B in place (B is n x m row-major), L lower triangular,
 * trans = 0: L X = B, trans = 1: L^T X = B.  Replaces TriangularLazyTensor.inv_matmul /
 * torch.triangular_solve (src/variational/variational_strategy.py:44) and, chained,
 * cholesky_solve behind LazyTensor.inv_matmul (cmp_prediction_strategy.py:60,164).
 * Workspace: pass max(m, ldb) as m when B is a strided view (the blocked sweeps keep B's row stride; a workspace sized for
 * a narrower B is still accepted but takes the slower 128-block path). */
DCGP_API size_t dcgp_trsm_workspace(int64_t n, int64_t m, int dtype);
DCGP_API int dcgp_trsm(int trans, const void* L, int64_t n, int64_t ldl, void* B, int64_t m, int64_t ldb, void* workspace,
              size_t workspace_bytes, int dtype, int flags, void* stream);
/* dcgp_potrs: X = (L L^T)^{-1} B in place = trsm(0) then trsm(1). */
DCGP_API int dcgp_potrs(const void* L, int64_t n, int64_t ldl, void* B, int64_t m, int64_t ldb, void* workspace,
               size_t workspace_bytes, int dtype, int flags, void* stream);
/* out[0] = 2 * sum_i log L_ii   (log det of L L^T) */
/* Solve plans: everything about a factor L that a blocked substitution sweep needs beyond L itself (the
 * inverted 512-wide diagonal blocks and, for 3xTF32, the TF32 remainders of L and of those blocks).  The
 * forward solve and its adjoint in one ELBO step use the same Lambda^{-1}
 * (src/models/variational_cmp.py:69-83 and its autograd backward): build the plan once, then
 * dcgp_trsm_planned(trans = 0 | 1 | 2 (both sweeps, = potrs)).  plan_bytes == 0 (small n) or plan == NULL
 * falls back to dcgp_trsm with the same workspace contract. */
DCGP_API size_t dcgp_solve_plan_bytes(int64_t n, int64_t ldl, int dtype);
DCGP_API int dcgp_solve_plan(const void* L, int64_t n, int64_t ldl, void* plan, size_t plan_bytes, int dtype, int flags,
                    void* stream);
DCGP_API int dcgp_trsm_planned(int trans, const void* L, int64_t n, int64_t ldl, const void* plan, size_t plan_bytes,
                      void* B, int64_t m, int64_t ldb, void* workspace, size_t workspace_bytes, int dtype, int flags,
                      void* stream);

DCGP_API int dcgp_logdet_chol(const void* L, int64_t n, int64_t ldl, void* out, int dtype, void* stream);

/* W = L^{-1} for lower-triangular L (n x n, row-major); W is packed (ldw == n), lower, zero above the
 * diagonal.  Blocked recursive doubling on the tensor cores.  The bag-level Gaussian terms of
 * src/likelihoods/cmp_likelihood.py:60-78 (log|C|, e^T C^{-1} e, tr(C^{-1} G) and their adjoints) are all
 * products with C^{-1} = W^T W, so one inversion replaces eight triangular solves per ELBO step. */
DCGP_API size_t dcgp_trtri_workspace(int64_t n, int dtype);
DCGP_API int dcgp_trtri(const void* L, int64_t n, int64_t ldl, void* W, int64_t ldw, void* workspace,
               size_t workspace_bytes, int dtype, int flags, void* stream);

/* ---- bag aggregation ------------------------------------------------------------
 * Bags are contiguous row ranges: bag a = rows [offsets[a], offsets[a+1]) (int64 device array
 * of n_bags+1 entries).  Replaces the python loops over bags of
 * src/likelihoods/vbagg_gaussian_likelihood.py:75,99-101. */
/* out[a, :] = sum_{i in bag a} X[i, :]            (X is N x m) */
DCGP_API int dcgp_bag_sum(const void* X, int64_t ldx, int64_t m, const int64_t* offsets, int64_t n_bags, void* out,
                 int64_t ldo, int dtype, void* stream);
/* out[a, i] = sum_{j in bag a} k(z_i, x_j)        (n_bags x M) — bag-summed cross Gram, K_Zx never stored */
DCGP_API int dcgp_gram_bagsum(const dcgp_kernel_spec* spec, const void* z, int64_t M, int64_t ldz, const void* x, int64_t ldx,
                     const int64_t* offsets, int64_t n_bags, void* out, int64_t ldo, int dtype, int flags,
                     void* stream);
/* backward of dcgp_gram_bagsum for upstream G (n_bags x M): accumulates dls, dos, dz */
DCGP_API int dcgp_gram_bagsum_bwd(const dcgp_kernel_spec* spec, const void* z, int64_t M, int64_t ldz, const void* x,
                         int64_t ldx, const int64_t* offsets, int64_t n_bags, const void* G, int64_t ldg, void* dls,
                         void* dos, void* dz, int64_t lddz, int dtype, int flags, void* stream);
/* out[a] = sum_{i,j in bag a} k(x_i, x_j)         within-bag block sums 1_a^T K_xx 1_a */
DCGP_API int dcgp_gram_bag_blocksum(const dcgp_kernel_spec* spec, const void* x, int64_t ldx, const int64_t* offsets,
                           int64_t n_bags, void* out, int dtype, int flags, void* stream);
/* backward of dcgp_gram_bag_blocksum for upstream g (n_bags): accumulates dls, dos */
DCGP_API int dcgp_gram_bag_blocksum_bwd(const dcgp_kernel_spec* spec, const void* x, int64_t ldx, const int64_t* offsets,
                               int64_t n_bags, const void* g, void* dls, void* dos, int dtype, int flags,
                               void* stream);

/* ---- random Fourier features ------------------------------------------------------
 * out[i, j] = cos(p_ij), out[i, D + j] = sin(p_ij), p_ij = sum_k x[i,k] W[k,j] / l_k   (N x 2D)
 * x: N x d (d <= DCGP_MAX_DIMS, already restricted to the active dims), W: d x D row-major.
 * Replaces RFFKernel._featurize (src/kernels/rff_kernel.py:82-89; gpytorch RFFKernel).
 * _bwd: dx (N x d, overwritten, optional) and dls[d] (accumulated, optional) for upstream dphi. */
DCGP_API int dcgp_rff_features(const void* x, int64_t n, int64_t ldx, int d, const void* W, int64_t D,
                               const void* lengthscale, void* out, int64_t ldo, int dtype, void* stream);
DCGP_API int dcgp_rff_features_bwd(const void* x, int64_t n, int64_t ldx, int d, const void* W, int64_t D,
                                   const void* lengthscale, const void* dphi, int64_t ldg, void* dx, int64_t lddx,
                                   void* dls, int dtype, void* stream);

/* VBagg expected log-likelihood from the per-bag statistics S_a = 1_a^T mu_q, T_a = 1_a^T Sigma_q 1_a (produced by
 * dcgp_gram_bagsum / dcgp_gram_bag_blocksum and the whitening products) and its adjoints dS, dT, dnoise (any of
 * them may be NULL) in one launch: src/likelihoods/vbagg_gaussian_likelihood.py:78-111 without the dense
 * covariance of :74 and the per-bag python loops of :75, :99-101.  noise / ell / dnoise are device scalars. */
DCGP_API int dcgp_vbagg_ell(const void* z, const void* S, const void* T, const int64_t* offsets, int64_t n_bags,
                   const void* noise, void* ell, void* dS, void* dT, void* dnoise, int dtype, void* stream);

/* ---- whitened KL ---------------------------------------------------------------
 * out[0] += 0.5 * (||tril(Ls)||_F^2 + ||m||^2 - M - sum_i log Ls_ii^2) in one pass over Ls (out is ACCUMULATED
 * with atomics: the caller zeroes it)
 * (src/mlls/bag_variational_elbo.py:49 -> variational_strategy.kl_divergence()).
 * dLs (M x M, optional) = tril(Ls) - diag(1/Ls_ii), dm (optional) = m. */
DCGP_API int dcgp_whitened_kl(const void* m, const void* Ls, int64_t M, int64_t ldls, void* out, void* dm, void* dLs,
                     int64_t lddls, int dtype, void* stream);

/* ---- optimiser update ------------------------------------------------------------
 * One fused multi-tensor Adam update over flat buffers of n values (torch.optim.Adam arithmetic, no weight decay,
 * no amsgrad): the `optimizer.step()` of experiments/downscaling/models/variational_cmp.py:183 and
 * experiments/swiss_roll/models/variational_cmp.py:110.  `step` is a DEVICE double holding the number of updates
 * applied so far (advanced here).  `veto` (optional): n_veto device ints - if any is non-zero (a failed
 * factorisation of the step, see `info` of dcgp_potrf) nothing is changed, so the update can sit inside a captured
 * CUDA graph with no host decision.  grads are scaled by grad_scale first (1 / world size after a SUM all-reduce)
 * and zeroed afterwards when zero_grads != 0. */
DCGP_API int dcgp_adam_step(void* params, void* grads, void* exp_avg, void* exp_avg_sq, void* step, int64_t n, double lr,
                            double beta1, double beta2, double eps, double grad_scale, const int* veto, int n_veto,
                            int zero_grads, int dtype, void* stream);

/* ---- whole VBagg minibatch ELBO: value and every gradient behind two calls ---------------------------
 * SURVEY.md section 8b `vbagg_elbo_fwd/bwd`.  Replaces, for one optimiser step of the VBagg model
 * (experiments/swiss_roll/models/vbagg.py, experiments/downscaling/models/vbagg.py): q(f) of the whitened SVGP
 * (src/variational/variational_strategy.py:14-68), its per-bag aggregation and the expected log-likelihood
 * (src/likelihoods/vbagg_gaussian_likelihood.py:54-111), the ELBO combination ELL / n_bags - beta KL / num_data
 * (src/mlls/bag_variational_elbo.py:47-69) and torch's .backward() over all of them.  K_Zx, Sigma_q and the
 * aggregation matrix are never formed; the algebra is written out at the top of csrc/vbagg.cu.
 *
 * All pointers are device pointers of the compute dtype unless noted; hyper-parameters are VALUES (lengthscale,
 * outputscale, noise after their constraint), as in dcgp_kernel_spec; gradients are with respect to those values.
 * dcgp_vbagg_elbo_fwd writes the ELBO (device scalar) and leaves its intermediates in the workspace;
 * dcgp_vbagg_elbo_bwd must follow on the same stream with the same problem and workspace and writes
 * scale * d ELBO / d(.) into every non-NULL field of `grads` (scale = -1 for the loss the trainers minimise).
 * `info`: device int, LAPACK-style status of the K_ZZ factorisation (0 = ok), as for dcgp_potrf.
 * F32 runs every product error-compensated (3xTF32) and keeps the reference's FP64 island (variational_strategy.py:35-44):
 * K_ZZ is evaluated, factored and inverted in double, its inverse factor applied as an FP32 hi + lo pair, its adjoint
 * taken in double again (DCGP_VBAGG_ISLAND=0: K_ZZ in FP32 / 3xTF32). */
typedef struct dcgp_vbagg_problem {
    const dcgp_kernel_spec* spec; /* individuals kernel k                                              */
    const void* Z;                /* inducing points, M x d row-major (d = columns the spec addresses)  */
    int64_t M, ldz;
    const void* x;                /* individuals of the minibatch, bag after bag, N x d                 */
    int64_t ldx;
    const int64_t* offsets;       /* n_bags + 1 row offsets (device int64)                              */
    int64_t n_bags;
    const void* z;                /* bag observations, n_bags                                           */
    const void* var_mean;         /* variational mean m, M                                              */
    const void* chol_var;         /* variational factor L_S, M x M; only its lower triangle is read     */
    int64_t ldc;
    const void* noise;            /* device scalar sigma^2                                              */
    const void* mean_const;       /* device scalar of a ConstantMean, NULL for ZeroMean                 */
    double num_data, beta;        /* BagVariationalELBO(num_data, beta)                                 */
    double jitter;                /* added to diag K_ZZ: 1e-3 in the reference                          */
    double var_jitter;            /* added to diag K_xx: 1e-4 in the reference                          */
} dcgp_vbagg_problem;

typedef struct dcgp_vbagg_grads {   /* any field may be NULL */
    void* dZ;                       /* M x d, row stride lddz (overwritten)               */
    int64_t lddz;
    void* dls;                      /* [n_components][DCGP_MAX_DIMS] (overwritten)        */
    void* dos;                      /* [n_components] (overwritten)                       */
    void* dvar_mean;                /* M                                                  */
    void* dchol_var;                /* M x M, row stride lddc, zero above the diagonal    */
    int64_t lddc;
    void* dnoise;                   /* scalar                                             */
    void* dmean_const;              /* scalar                                             */
} dcgp_vbagg_grads;

/* ldz: the row stride of Z that will be passed (F32 problems keep double copies of Z and dZ for the FP64 island) */
DCGP_API size_t dcgp_vbagg_elbo_workspace(int64_t M, int64_t n_bags, int64_t ldz, int dtype);
DCGP_API int dcgp_vbagg_elbo_fwd(const dcgp_vbagg_problem* problem, void* elbo, int* info, void* workspace,
                                 size_t workspace_bytes, int dtype, int flags, void* stream);
DCGP_API int dcgp_vbagg_elbo_bwd(const dcgp_vbagg_problem* problem, double scale, const dcgp_vbagg_grads* grads,
                                 void* workspace, size_t workspace_bytes, int dtype, int flags, void* stream);

/* ---- whole VariationalCMP minibatch ELBO (with individuals noise): value and every gradient behind two calls --------
 * SURVEY.md section 8b `cmp_elbo_fwd/bwd`.  Replaces, for one optimiser step of
 * experiments/downscaling/models/variational_cmp.py:163-186 (and experiments/swiss_roll/models/variational_cmp.py):
 * model(x) (src/variational/variational_strategy.py:14-68), get_elbo_computation_parameters
 * (src/models/variational_cmp.py:86-113), CMPLikelihood._compute_expected_log_prob_with_individuals_noise
 * (src/likelihoods/cmp_likelihood.py:41-78), BagVariationalELBO (src/mlls/bag_variational_elbo.py:47-69) and
 * .backward().  The algebra (no root of Lambda^{-1}, no Sigma_q) is written out at the top of csrc/cmp.cu.
 * Conventions as for dcgp_vbagg_elbo_*: device pointers of the compute dtype, hyper-parameter VALUES, gradients with
 * respect to those values scaled by `scale`, the workspace carries the forward intermediates to the backward call.
 * `info`: THREE device ints - status of the factorisations of Lambda, K_ZZ and the bag-level matrix C (0 = ok).
 * F32 problems keep the reference's FP64 island (variational_strategy.py:35-44): K_ZZ is evaluated, factored and inverted
 * in double, its inverse factor applied as an FP32 hi + lo pair, its adjoint taken in double again - on a side lane beside
 * the Lambda factorisation, so it costs no wall-clock (DCGP_CMP_ISLAND=0 keeps K_ZZ in FP32 / 3xTF32). */
typedef struct dcgp_cmp_problem {
    const dcgp_kernel_spec* k_spec; /* individuals kernel k                                                  */
    const dcgp_kernel_spec* l_spec; /* bag kernel l                                                          */
    const void* Z;                  /* inducing points, M x d                                                */
    int64_t M, ldz;
    const void* x;                  /* individuals of the CME sample, N x d                                  */
    int64_t N, ldx;
    const void* ext;                /* their bag values y~ (extended_bags_values), N x r                     */
    int64_t ldext;
    const void* y;                  /* bag values of the minibatch, n x r                                    */
    int64_t n, ldy;
    const void* z;                  /* aggregate observations, n                                             */
    const void* var_mean;           /* variational mean m, M                                                 */
    const void* chol_var;           /* variational factor L_S, M x M; only its lower triangle is read        */
    int64_t ldc;
    const void* noise;              /* device scalar: aggregate noise sigma^2 (likelihood.noise)             */
    const void* ind_noise;          /* device scalar: individuals noise (noise_kernel.outputscale)           */
    const void* mean_const;         /* device scalar of a ConstantMean, NULL for ZeroMean                    */
    double lbda;                    /* CME regulariser: Lambda = l(y~, y~) + lbda N I                         */
    double num_data, beta;
    double jitter, var_jitter;      /* 1e-3 on K_ZZ, 1e-4 on K_xx in the reference                           */
} dcgp_cmp_problem;

typedef struct dcgp_cmp_grads {     /* any field may be NULL; all are overwritten */
    void* dZ;                       /* M x d, row stride lddz                              */
    int64_t lddz;
    void* dk_ls;                    /* [k components][DCGP_MAX_DIMS]                       */
    void* dk_os;                    /* [k components]                                      */
    void* dl_ls;                    /* [l components][DCGP_MAX_DIMS]                       */
    void* dl_os;                    /* [l components]                                      */
    void* dvar_mean;                /* M                                                   */
    void* dchol_var;                /* M x M, row stride lddc, zero above the diagonal     */
    int64_t lddc;
    void* dnoise;                   /* scalar                                              */
    void* dind_noise;               /* scalar                                              */
    void* dmean_const;              /* scalar                                              */
} dcgp_cmp_grads;

/* ldz: the row stride of Z that will be passed (the FP64 island of F32 problems keeps double copies of Z and dZ) */
DCGP_API size_t dcgp_cmp_elbo_workspace(int64_t N, int64_t n, int64_t M, int64_t ldz, int dtype);
DCGP_API int dcgp_cmp_elbo_fwd(const dcgp_cmp_problem* problem, void* elbo, int* info, void* workspace,
                               size_t workspace_bytes, int dtype, int flags, void* stream);
DCGP_API int dcgp_cmp_elbo_bwd(const dcgp_cmp_problem* problem, double scale, const dcgp_cmp_grads* grads, void* workspace,
                               size_t workspace_bytes, int dtype, int flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCGP_H_ */
