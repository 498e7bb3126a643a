/*
 * b200gp.h -- C-ABI of libb200gp.so: the B200-native exact-GP posterior path.
 *
 * The reference (ziatdinovmax/gpax v0.1.9) is pure Python on JAX and has no FFI; its two seams on
 * this path are Python callables:
 *   kernel seam     gpax/kernels/kernels.py:17      kernel(X, Z, params, noise, jitter) -> K
 *   posterior seam  gpax/models/gp.py:253-255       get_mvn_posterior(X_new, params, noiseless, **kw)
 * The entry points below are what a ctypes binding placed behind those two callables calls
 * (INTEGRATION.md shows the stub).  Every entry point names the reference lines it replaces.
 *
 * Conventions
 *   - plain C symbols, plain pointers and sizes; no torch / numpy types
 *   - all matrices are fp64, ROW-MAJOR with an explicit leading dimension (elements)
 *   - pointers are HOST pointers unless B2GP_FLAG_DEVICE_PTRS is set in `flags`, in which case every
 *     array argument (not `info`, not `timing`) is a device pointer obtained from b2gp_dev_alloc
 *   - the caller owns every buffer it passes; the library owns only the ctx and its workspaces
 *   - return value: 0 ok, <0 argument / CUDA failure (text via b2gp_last_error).  A numerical failure
 *     is NOT an error status: `info[s] > 0` is the 1-based index of the first non-positive pivot of
 *     draw s and that draw's outputs are NaN (the reference yields NaNs, never an exception, and
 *     post-filters them: gpax/models/gp.py:396-398)
 *   - a ctx is not thread-safe; calls are synchronous from the caller's point of view
 */
#ifndef B200GP_H
#define B200GP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2GP_VERSION 100

typedef struct b2gp_ctx b2gp_ctx;

/* kernel families: gpax/kernels/kernels.py:44-65 (RBF), 68-91 (Matern-5/2), 94-117 (Periodic);
 * name table gpax/kernels/kernels.py:227-241 */
enum { B2GP_KERNEL_RBF = 0, B2GP_KERNEL_MATERN52 = 1, B2GP_KERNEL_PERIODIC = 2,
       /* b2gp_gram only -- NNGP kernels, gpax/kernels/kernels.py:120-224 (erf / ReLU activation): `scale` carries var_w,
        * `period` carries var_b, lengthscale[0] carries the depth; non-stationary, so the posterior entry points (which
        * need k(x, x) in closed form) do not take them -- the shell routes them through its callable-kernel path */
       B2GP_KERNEL_NNGP_ERF = 3, B2GP_KERNEL_NNGP_RELU = 4 };

enum {
    B2GP_OK = 0,
    B2GP_ERR_ARG = -1,
    B2GP_ERR_CUDA = -2,
    B2GP_ERR_NOMEM = -3,
    B2GP_ERR_UNSUPPORTED = -4
};

enum {
    B2GP_FLAG_DEVICE_PTRS = 1u << 0, /* array arguments are device pointers                                  */
    B2GP_FLAG_LOWER_ONLY  = 1u << 1, /* b2gp_gram with same_xz: write only the lower triangle (j <= i)       */
    B2GP_FLAG_F32         = 1u << 2, /* b2gp_gram / b2gp_posterior(_batch) / b2gp_sparse_posterior: the DATA arrays (X, Z, y, X_new, Xu,
                                        noise_vec, eps in; K, mean, var, cov, y_sampled out) are float, the reference's default
                                        precision (gpax/utils/utils.py:19-21); theta stays double.  Widened / narrowed on the
                                        device, everything in between is fp64                                  */
    B2GP_OUT_MEAN         = 1u << 4, /* b2gp_posterior: produce mean[S,P]                                    */
    B2GP_OUT_VAR          = 1u << 5, /* ... var[S,P] = diag(cov)     (viGP.predict, vigp.py:184-185)         */
    B2GP_OUT_COV          = 1u << 6, /* ... cov[S,P,P]               (get_mvn_posterior, gp.py:272)          */
    B2GP_OUT_SAMPLE       = 1u << 7  /* ... y_sampled[S,n,P] = mean + chol(cov) eps   (gp.py:292)            */
};

/* per-call device timing (CUDA events on the library's own streams), filled when non-NULL.
 * The *_ms stage fields are sums of stream time over the draws (they can exceed total_ms when
 * several draws are in flight); total_ms is first-launch to last-completion on the device,
 * h2d/d2h are the host<->device copies of the host-pointer entry points. */
typedef struct b2gp_timing {
    double total_ms;
    double gram_ms;
    double potrf_ms;
    double trsm_ms;
    double epilogue_ms;
    double h2d_ms;
    double d2h_ms;
    double flops;        /* algorithmic flops of the call: S * (N^3/3 + N^2 (P+1) + ...)   */
    double gram_bytes;   /* algorithmic bytes written by the Gram builds                                 */
    int64_t launches;    /* kernels launched by the call                                                 */
    double host_enqueue_ms; /* host wall time spent issuing the call's work (before waiting for the device)  */
} b2gp_timing;

int  b2gp_version(void);

/* lifecycle --------------------------------------------------------------------------------------
 * One context = one CUDA device = one process per GPU.  (SURVEY.md 8b sketched `b2gp_ctx_create(n_dev, dev_ids, out)` with
 * ncclCommInitAll inside one process; the multi-GPU forms of the path run one process per GPU instead and join their
 * contexts with b2gp_dist_init below.) */
int  b2gp_ctx_create(int device, b2gp_ctx** out);
int  b2gp_ctx_destroy(b2gp_ctx* ctx);
const char* b2gp_last_error(const b2gp_ctx* ctx);
/* options: "streams" (draws in flight, 1..16, default 2); "ozaki" (0: fp64 DMMA only, 6 / 7: int8 tcgen05 base-256 digit
 * planes, -1 (default): 6 or 7 chosen per call from a bound on cond(K)); the full table is in INTEGRATION.md */
int  b2gp_set_option(b2gp_ctx* ctx, const char* key, int64_t value);
int  b2gp_device_info(b2gp_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* mem_bytes);
/* device timing of the most recent entry-point call on this ctx (every call records total_ms) */
int  b2gp_last_timing(b2gp_ctx* ctx, b2gp_timing* out);

/* device memory, for callers that keep inputs resident in HBM (replaces jax.device_put,
 * gpax/models/gp.py:388-391,416-428) ----------------------------------------------------------------*/
int  b2gp_dev_alloc(b2gp_ctx* ctx, size_t bytes, void** dptr);
int  b2gp_dev_free(b2gp_ctx* ctx, void* dptr);
int  b2gp_host_alloc(b2gp_ctx* ctx, size_t bytes, void** hptr);   /* pinned host memory */
int  b2gp_host_free(b2gp_ctx* ctx, void* hptr);
int  b2gp_h2d(b2gp_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int  b2gp_d2h(b2gp_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
int  b2gp_sync(b2gp_ctx* ctx);

/* Gram build -- replaces square_scaled_distance + RBFKernel / MaternKernel / PeriodicKernel
 * (gpax/kernels/kernels.py:28-41, 44-65, 68-91, 94-117).
 *   X[n,d], Z[m,d] row-major; lengthscale[d] (a scalar lengthscale is broadcast by the caller);
 *   K[n,m] with leading dimension ldk.  `diag_add` = noise + jitter is added on i == j iff
 *   same_xz != 0; the caller sets same_xz from the reference's rule `X.shape == Z.shape`
 *   (kernels.py:63, 89, 115).  `period` is read for B2GP_KERNEL_PERIODIC only.                     */
int  b2gp_gram(b2gp_ctx* ctx, int kind,
               const double* X, int64_t n, const double* Z, int64_t m, int d,
               const double* lengthscale, double scale, double period,
               double diag_add, int same_xz,
               double* K, int64_t ldk, unsigned flags);

/* Multi-task Gram matrix -- gpax/kernels/mtkernels.py:19-58 (index_kernel) and 61-125 (MultitaskKernel):
 * K[i,j] = (k_data(x_i, z_j) + jitter if same point) * B[taskX[i], taskZ[j]], plus noise_task[taskX[i]] + jitter on i == j,
 * both iff same_xz (the reference's shape rule).  B[T,T] = W W^T + diag(v) is formed by the caller.  Host fp64 / int32
 * arrays.  The Kronecker form of MultivariateKernel (mtkernels.py:128-192) is this call on inputs repeated once per
 * task with group = T (`group` consecutive rows are one data point; 1 otherwise).                                     */
int  b2gp_gram_multitask(b2gp_ctx* ctx, int kind, const double* X, const int* taskX, int64_t n,
                         const double* Z, const int* taskZ, int64_t m, int d,
                         const double* lengthscale, double scale, double period,
                         const double* B, int T, const double* noise_task, double jitter, int same_xz, int group,
                         double* K, int64_t ldk, unsigned flags);

/* Cholesky factorisation A = L L^T of the lower triangle, in place (row-major, lower); the strict
 * upper triangle is not referenced and not modified.  Stands where the reference inverts k_XX
 * (jnp.linalg.inv, gpax/models/gp.py:271) and where viSparseGP calls jax.scipy.linalg.cholesky
 * (gpax/models/sparse_gp.py:194,201).  *info = 0, or the 1-based index of the first bad pivot.      */
int  b2gp_potrf(b2gp_ctx* ctx, int64_t n, double* A, int64_t lda, int* info, unsigned flags);

/* Triangular solve with the factor: overwrites the nrhs right-hand sides with L^{-1} b.
 * B holds one right-hand side per ROW: B[r, 0..n) is b_r, leading dimension ldb (i.e. the n x nrhs
 * matrix of right-hand sides in column-major order).  Replaces solve_triangular(L, ., lower=True)
 * (gpax/models/sparse_gp.py:197,207,209) and the K^{-1} products of gp.py:272-273.
 * L must be the output of b2gp_potrf made through the same ctx immediately before (the solve
 * reuses the inverted diagonal blocks that factorisation left in the ctx).                          */
int  b2gp_trsm_lower(b2gp_ctx* ctx, int64_t n, int64_t nrhs,
                     const double* L, int64_t ldl, double* B, int64_t ldb, unsigned flags);

/* C[m,n] = beta C + alpha A[m,k] B[n,k]^T (fp64, DMMA tensor pipe); lower_only != 0 updates only
 * j <= i (SYRK when A == B).  The trailing-update kernel of the factorisation, exported for the
 * roofline measurement and the parity tests; replaces the jnp.matmul calls of gp.py:272-273.       */
int  b2gp_gemm_nt(b2gp_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha,
                  const double* A, int64_t lda, const double* B, int64_t ldb,
                  double beta, double* C, int64_t ldc, int lower_only, unsigned flags);

/* The posterior, batched over S hyper-parameter draws -- replaces ExactGP.get_mvn_posterior
 * (gpax/models/gp.py:253-277), _predict's sampling (gp.py:279-293), the vmap over draws in predict
 * (gp.py:393-395) and viGP.predict (gpax/models/vigp.py:178-185).
 *   Xtr[N,d], Xnew[P,d]                    row-major
 *   yres[N] (yres_stride == 0) or yres[S, yres_stride]   y_train minus the mean function (gp.py:262-265)
 *   theta[S, d+3]                          per draw: lengthscale[0..d), k_scale, noise, period
 *   noiseless                              gp.py:260-261: noise_p = noise * (1 - noiseless)
 *   jitter                                 the **kwargs jitter of gp.py:267,269 (default 1e-6)
 *   flags                                  B2GP_OUT_* (+ B2GP_FLAG_DEVICE_PTRS)
 *   mean[S,P], var[S,P], cov[S,P,P]        outputs selected by flags (others may be NULL)
 *   eps[S,n_samp,P], y_sampled[S,n_samp,P] standard-normal draws in, posterior samples out
 *   info[S]                                0 or first bad pivot of k_XX (>0) / of cov (<0, sampling)  */
int  b2gp_posterior(b2gp_ctx* ctx, int kind,
                    const double* Xtr, int64_t N, const double* yres, int64_t yres_stride,
                    const double* Xnew, int64_t P, int d, int64_t S,
                    const double* theta, int noiseless, double jitter, unsigned flags,
                    double* mean, double* var, double* cov,
                    const double* eps, int64_t n_samp, double* y_sampled,
                    int* info, b2gp_timing* timing);

/* The same posterior with everything that may differ between the S members of the batch:
 *   Xtr[S, xtr_stride] / Xnew[S, xnew_stride]   per-member training / test inputs (stride in doubles; 0 = shared [N,d] / [P,d]):
 *       the outer task axis of vExactGP (gpax/models/vgp.py:125-172) and the per-draw perturbed inputs X_prime of
 *       UIGP (gpax/models/uigp.py:131-150)
 *   noise_vec[N] or [S, noise_vec_stride]       per-point noise variances added to the diagonal of k_XX on top of
 *       theta's scalar noise: MeasuredNoiseGP (k + diag(measured_noise), gpax/models/mngp.py:92-97) and VarNoiseGP
 *       (k + diag(exp(log_var)), gpax/models/hskgp.py:143-148); NULL = none.
 * All other arguments as b2gp_posterior.                                                                          */
int  b2gp_posterior_batch(b2gp_ctx* ctx, int kind,
                          const double* Xtr, int64_t xtr_stride, int64_t N, const double* yres, int64_t yres_stride,
                          const double* Xnew, int64_t xnew_stride, int64_t P, int d, int64_t S,
                          const double* theta, const double* noise_vec, int64_t noise_vec_stride,
                          int noiseless, double jitter, unsigned flags,
                          double* mean, double* var, double* cov,
                          const double* eps, int64_t n_samp, double* y_sampled,
                          int* info, b2gp_timing* timing);

/* Nystrom / VFE sparse posterior for one theta -- replaces viSparseGP.get_mvn_posterior
 * (gpax/models/sparse_gp.py:173-223).  Xu[M,d] inducing points; theta[d+3] as above;
 * outputs mean[P] and var[P] (B2GP_OUT_VAR) and/or cov[P,P] (B2GP_OUT_COV).                        */
int  b2gp_sparse_posterior(b2gp_ctx* ctx, int kind,
                           const double* Xu, int64_t M, const double* Xtr, int64_t N, const double* yres,
                           const double* Xnew, int64_t P, int d,
                           const double* theta, int noiseless, double jitter, unsigned flags,
                           double* mean, double* var, double* cov,
                           int* info, b2gp_timing* timing);

/* Fit side: value and gradient of the exact-GP log marginal likelihood
 *   log N(yres; 0, K_theta),  K_theta = kernel(X, X, theta, noise, jitter)
 * i.e. the numpyro.sample("y", MultivariateNormal(f_loc, covariance_matrix=k), obs=y) term of
 * gpax/models/gp.py:158-164 and its reverse-mode derivative.  grad[d+3] is w.r.t. (log lengthscale[0..d),
 * log k_scale, log noise, log period); theta (d+3) is a HOST pointer; value, grad, alpha_out[N] = K^{-1} yres
 * (optional) are HOST outputs; X, yres follow `flags`.  d <= 16.                                          */
int  b2gp_mll(b2gp_ctx* ctx, int kind, const double* X, int64_t N, const double* yres, int d,
              const double* theta, double jitter, unsigned flags,
              double* value, double* grad, double* alpha_out, int* info);

/* b2gp_mll with a vector of per-point noise variances on the diagonal, K = kernel(X, X, theta, noise, jitter) + diag(noise_vec)
 * (the likelihoods of gpax/models/mngp.py:92-97 and gpax/models/hskgp.py:143-148), and grad_noise_vec[N] (HOST, optional,
 * needs grad) = d value / d noise_vec[i] = 1/2 (alpha_i^2 - K^{-1}_ii).                                              */
int  b2gp_mll_v(b2gp_ctx* ctx, int kind, const double* X, int64_t N, const double* yres, int d,
                const double* theta, const double* noise_vec, double jitter, unsigned flags,
                double* value, double* grad, double* alpha_out, double* grad_noise_vec, int* info);

/* Fit side of the sparse GP: the VFE bound of viSparseGP.model (gpax/models/sparse_gp.py:62-114)
 *   log LowRankMVN(yres; 0, W^T W + noise I) - 1/2 clip(sum_n (Kff_nn - Qff_nn) / noise, 0),  W = Luu^{-1} K(Xu, X)
 * and its gradient w.r.t. (log lengthscale[d], log k_scale, log noise, log period) in grad_theta[d+3] and w.r.t. the
 * inducing inputs in grad_Xu[M,d] (the reference differentiates the same expression with JAX; Xu is a numpyro.param,
 * sparse_gp.py:69-70).  theta is a HOST pointer; value / grads are HOST outputs; Xu, X, yres follow `flags`. d <= 16.  */
int  b2gp_sparse_elbo(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, const double* X, int64_t N,
                      const double* yres, int d, const double* theta, double jitter, unsigned flags,
                      double* value, double* grad_theta, double* grad_Xu, int* info);

/* Samples of S multivariate normals: y[s,i,:] = mean[s,:] + chol(cov[s]) eps[s,i,:], i < n -- replaces
 * numpyro.distributions.MultivariateNormal(mean, cov).sample (gpax/models/gp.py:292, gpax/acquisition/base_acq.py:221)
 * where the caller changed cov after the posterior call (gpax/models/hskgp.py:201-204 adds the predicted noise variance).
 * info[s] = 0 or the first bad pivot of cov[s] (that member's samples are NaN).                                    */
int  b2gp_mvn_sample(b2gp_ctx* ctx, const double* mean, const double* cov, int64_t S, int64_t P,
                     const double* eps, int64_t n, double* y, int* info, unsigned flags);

/* ---- acquisition epilogues on the posterior's outputs, host or device pointers ----------
 * kind: 0 EI, 1 UCB, 2 UE, 3 POI -- gpax/acquisition/base_acq.py:20-71 (ei), 74-104 (ucb), 107-130 (ue), 133-155 (poi).
 *   mean[R,P], var[R,P] -> out[R,P]; row r is one posterior (R = 1 for viGP / the pooled moments of an MCMC model,
 *   R = number of sub-sampled draws for the q-batch functions, gpax/acquisition/batch_acquisition.py:110-116).
 *   have_best = 0: best_f is derived from each row's mean (max when maximize, else min: base_acq.py:59-60);
 *   param = beta (UCB) or xi (POI).                                                                             */
int  b2gp_acq_moments(b2gp_ctx* ctx, int kind, const double* mean, const double* var, int64_t R, int64_t P,
                      int have_best, double best_f, double param, int maximize, double* out, unsigned flags);
/* Same, from posterior samples y[R,P] (R = S*n rows of y_sampled): column mean and population variance
 * (gpax/acquisition/acquisition.py:31-34), then the acquisition function.  mean_out / var_out [P] optional.      */
int  b2gp_acq_samples(b2gp_ctx* ctx, int kind, const double* y, int64_t R, int64_t P,
                      int have_best, double best_f, double param, int maximize,
                      double* out, double* mean_out, double* var_out, unsigned flags);
/* Knowledge gradient (gpax/acquisition/base_acq.py:158-232) from one posterior: mean[P], cov[P,P] of the candidates,
 * ysim[n,P] simulated observations (base_acq.py:223).  The reference re-inverts the (N+1)x(N+1) training covariance
 * for every candidate and simulation; here the rank-1 (block-inverse) update of the posterior mean is evaluated in
 * closed form.  diag_sub = noise_p + jitter (what `cov` carries on its diagonal beyond the latent covariance),
 * noise_plus_jitter = the diagonal term of the appended training point.  out[P].                                 */
int  b2gp_kg(b2gp_ctx* ctx, const double* mean, const double* cov, int64_t P, const double* ysim, int64_t n,
             double diag_sub, double noise_plus_jitter, int maximize, double* out, unsigned flags);

/* ---- multi-GPU building blocks.  One process per GPU; the exchange steps
 * (panel broadcast, M x M all-reduce) are issued by the host side over NCCL on these same device
 * buffers (gpax_b200/distributed.py).  All array pointers below are DEVICE pointers. ----------------*/

/* ---- in-library multi-GPU (one process per GPU, NCCL loaded at run time; gpax_b200/csrc/dist.cuh) ------------------
 * b2gp_dist_unique_id: rank 0 obtains the 128-byte NCCL id and hands it to the other ranks by any host channel
 *   (gpax_b200/dist.py uses a TCP socket at MASTER_ADDR).  b2gp_dist_init: every rank, same id; builds the world
 *   communicator and the row / column communicators of a grid_rows x grid_cols process grid (rank = row * grid_cols + col).
 * b2gp_dist_posterior (COLLECTIVE, host pointers, inputs replicated on every rank): exact-GP posterior mean and diagonal
 *   variance -- gpax/models/gp.py:253-277 / gpax/models/vigp.py:178-185 -- with k_XX 2-D block-cyclic over the grid in
 *   nb x nb tiles (N a multiple of nb, nb a multiple of 128), generated in place; right-looking Cholesky with the panel
 *   solve spread over the process column, the panel broadcast along process rows and all-gathered down process columns,
 *   look-ahead of one panel, trailing updates on the int8 tcgen05 kernel; the right-hand sides ride below the matrix.
 *   Every rank receives mean[P], var[P] (B2GP_OUT_VAR) and info.
 * b2gp_dist_layout: the block-cyclic index algebra as a pure function (no GPU), see dist.cuh.                         */
int  b2gp_dist_unique_id(void* id128);
int  b2gp_dist_init(b2gp_ctx* ctx, const void* id128, int rank, int nranks, int grid_rows, int grid_cols);
int  b2gp_dist_info(b2gp_ctx* ctx, int* rank, int* nranks, int* grid_rows, int* grid_cols);
int  b2gp_dist_finalize(b2gp_ctx* ctx);
int  b2gp_dist_posterior(b2gp_ctx* ctx, int kind, const double* Xtr, int64_t N, const double* yres,
                         const double* Xnew, int64_t P, int d, const double* theta, int noiseless, double jitter,
                         int64_t nb, unsigned flags, double* mean, double* var, int* info, b2gp_timing* timing);
/* b2gp_dist_sparse_posterior (COLLECTIVE, host pointers): N-sharded Nystrom / VFE posterior --
 *   gpax/models/sparse_gp.py:173-223 -- every rank passes ITS shard of the training set and the same Xu, X_new, theta;
 *   per-rank statistics, one NCCL all-reduce of the M x M matrix + M-vector inside the library, replicated finish.    */
int  b2gp_dist_sparse_posterior(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, const double* Xtr_shard, int64_t N_shard,
                                const double* y_shard, const double* Xnew, int64_t P, int d, const double* theta,
                                int noiseless, double jitter, unsigned flags, double* mean, double* var, int* info,
                                b2gp_timing* timing);
int  b2gp_dist_layout(int64_t T, int64_t R, int64_t nb, int pr, int pc, int row, int col, int64_t k, int64_t* out6);

/* N-sharded sparse posterior: per-shard statistics, then the posterior from their sum.
 *   Kpart[M,M] (lower) = W W^T / noise,  cpart[M] = W y / noise  with W = Luu^{-1} K(Xu, Xtr_shard)
 *   (gpax/models/sparse_gp.py:193-199, 203-204 restricted to a shard; sums over shards give the full terms).
 *   theta is a HOST pointer (d+3 values).                                                           */
int  b2gp_sparse_partial(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, const double* Xtr, int64_t N,
                         const double* yres, int d, const double* theta, double jitter,
                         double* Kpart, int64_t ldk, double* cpart, int* info);
/* Ksum is overwritten (+I, then its Cholesky factor): sparse_gp.py:200-217.                          */
int  b2gp_sparse_finish(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, double* Ksum, int64_t ldk,
                        const double* csum, const double* Xnew, int64_t P, int d, const double* theta,
                        int noiseless, double jitter, unsigned flags,
                        double* mean, double* var, double* cov, int* info);

/* Block-cyclic Cholesky: factor one diagonal block and export the inverted 128x128 sub-blocks
 * (ceil(n/128)*128*128 doubles) so that panel solves can run later / on other blocks.              */
int  b2gp_potrf_inv(b2gp_ctx* ctx, int64_t n, double* A, int64_t lda, double* Linv_out, int* info);
/* B (nrhs rows of length n, leading dimension ldb) <- B L^{-T} using Linv from b2gp_potrf_inv.      */
int  b2gp_trsm_inv(b2gp_ctx* ctx, int64_t n, int64_t nrhs, const double* L, int64_t ldl,
                   const double* Linv, double* B, int64_t ldb);
/* dot[r] (+)= scale * <R[r,0:len), w>,  nrm[r] (+)= |R[r,0:len)|^2 ; either output may be NULL.      */
int  b2gp_rowdot(b2gp_ctx* ctx, int64_t rows, int64_t len, const double* R, int64_t ldr, const double* w,
                 double scale, double* dot, double* nrm, int accumulate);
int  b2gp_copy2d(b2gp_ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols);

#ifdef __cplusplus
}
#endif
#endif /* B200GP_H */
