// potrf.cuh -- blocked Cholesky A = L L^T (row-major, lower) and the triangular solves with L.
//
// Stands where the reference forms jnp.linalg.inv(k_XX) (gpax/models/gp.py:271) and where
// viSparseGP calls cholesky / solve_triangular (gpax/models/sparse_gp.py:194,197,201,207,209).
//
// Structure (everything above the 128x128 leaf is the DMMA GEMM of gemm_dmma.cuh):
//   potrf_rec(A, n):  n <= 128 -> potrf_diag_kernel (one CTA, shared memory): factor the block AND
//                     leave its inverse Linv in the workspace;
//                     else  potrf_rec(A11); A21 <- A21 L11^{-T} (trsm_rec); A22 -= A21 A21^T (SYRK,
//                     lower tiles only); potrf_rec(A22).
//   trsm_rec(B, L):   B <- B L^{-T} for B with one right-hand side per row.  n <= 128 -> one in-place
//                     GEMM with the inverted diagonal block, B <- B Linv^T; else split L, recurse,
//                     with B2 -= B1 L21^T in between.
// All flops above the leaf are K-major x K-major "NT" GEMMs, so one tensor-pipe kernel serves all.
#pragma once
#include "common.cuh"
#include "gemm_dmma.cuh"

constexpr int PD_LD = 129;       // shared-memory stride of the 128x128 block
constexpr int PD_ILD = 65;       // stride of the 64x64 inverse buffers
constexpr int PD_THREADS = 256;
constexpr int PD_SMEM = (128 * PD_LD + 2 * 64 * PD_ILD) * (int)sizeof(double);

// One CTA factors the n x n (n <= 128) lower-triangular block at A in shared memory, writes L back
// and writes inv(L) (lower, zero above the diagonal, leading dimension 128) to Linv.
// A non-positive or NaN pivot records *info = index_base + j + 1 (first one wins) and lets NaNs
// propagate; callers NaN-fill the outputs of that draw.
__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int n, double* __restrict__ Linv, int* info, int index_base) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* I1 = sm + 128 * PD_LD;
    double* I2 = I1 + 64 * PD_ILD;
    const int tid = threadIdx.x;

    for (int idx = tid; idx < 128 * 128; idx += PD_THREADS) {
        const int i = idx >> 7, j = idx & 127;
        double v = 0.0;
        if (i < n && j <= i) v = A[(int64_t)i * lda + j];
        S[i * PD_LD + j] = v;
    }
    __syncthreads();

    // ---- factorisation: 32-wide column panels, unblocked inside a panel, 4x4 register tiles for
    //      the trailing update
    for (int b0 = 0; b0 < n; b0 += 32) {
        const int bw = min(32, n - b0);
        const int bend = b0 + bw;
        for (int j = b0; j < bend; ++j) {
            const double piv = S[j * PD_LD + j];
            if (!(piv > 0.0) && tid == 0) atomicCAS(info, 0, index_base + j + 1);
            const double dj = sqrt(piv);
            __syncthreads();  // everyone has read the pivot
            if (tid == 0) S[j * PD_LD + j] = dj;
            for (int i = j + 1 + tid; i < n; i += PD_THREADS) S[i * PD_LD + j] = S[i * PD_LD + j] / dj;
            __syncthreads();
            const int cols = bend - (j + 1);
            const int rows = n - (j + 1);
            if (cols > 0) {
                for (int idx = tid; idx < rows * cols; idx += PD_THREADS) {
                    const int i = j + 1 + idx / cols;
                    const int k = j + 1 + idx % cols;
                    if (k <= i) S[i * PD_LD + k] -= S[i * PD_LD + j] * S[k * PD_LD + j];
                }
            }
            // (the next iteration's pivot read is ordered by the sync below)
            __syncthreads();
        }
        const int R = n - bend;
        if (R > 0) {
            const int T = (R + 3) >> 2;
            for (int mt = tid; mt < T * T; mt += PD_THREADS) {
                const int tr = mt / T, tc = mt % T;
                if (tc > tr) continue;
                const int r0 = bend + 4 * tr, c0 = bend + 4 * tc;
                double acc[4][4];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
                for (int c = b0; c < bend; ++c) {
                    double a[4], b[4];
#pragma unroll
                    for (int x = 0; x < 4; ++x) a[x] = (r0 + x < n) ? S[(r0 + x) * PD_LD + c] : 0.0;
#pragma unroll
                    for (int y = 0; y < 4; ++y) b[y] = (c0 + y < n) ? S[(c0 + y) * PD_LD + c] : 0.0;
#pragma unroll
                    for (int x = 0; x < 4; ++x)
#pragma unroll
                        for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
                }
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        const int r = r0 + x, c = c0 + y;
                        if (r < n && c <= r) S[r * PD_LD + c] -= acc[x][y];
                    }
            }
            __syncthreads();
        }
    }

    // ---- write L back (lower triangle only)
    for (int idx = tid; idx < 128 * 128; idx += PD_THREADS) {
        const int i = idx >> 7, j = idx & 127;
        if (i < n && j <= i) A[(int64_t)i * lda + j] = S[i * PD_LD + j];
    }

    // ---- inverse of L by the 2x2 block formula  inv = [[I11, 0], [-I22 L21 I11, I22]]
    const int n1 = min(n, 64), n2 = n - n1;
    for (int idx = tid; idx < 2 * 64 * PD_ILD; idx += PD_THREADS) I1[idx] = 0.0;  // I1 and I2 are contiguous
    __syncthreads();
    if (tid < 64) {
        const int j = tid;  // column j of inv(L11) by forward substitution
        if (j < n1) {
            for (int i = j; i < n1; ++i) {
                double s = (i == j) ? 1.0 : 0.0;
                for (int k = j; k < i; ++k) s = fma(-S[i * PD_LD + k], I1[k * PD_ILD + j], s);
                I1[i * PD_ILD + j] = s / S[i * PD_LD + i];
            }
        }
    } else if (tid < 128) {
        const int j = tid - 64;  // column j of inv(L22)
        if (j < n2) {
            for (int i = j; i < n2; ++i) {
                double s = (i == j) ? 1.0 : 0.0;
                for (int k = j; k < i; ++k) s = fma(-S[(64 + i) * PD_LD + 64 + k], I2[k * PD_ILD + j], s);
                I2[i * PD_ILD + j] = s / S[(64 + i) * PD_LD + 64 + i];
            }
        }
    }
    __syncthreads();
    // T = L21 * I11  (n2 x n1), parked in the unused upper-right quadrant S[r][64 + c]
    const int tr = tid >> 4, tc = tid & 15;
    if (n2 > 0) {
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
        for (int k = 4 * tc; k < n1; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) a[x] = S[(64 + 4 * tr + x) * PD_LD + k];  // rows >= n are zero
#pragma unroll
            for (int y = 0; y < 4; ++y) b[y] = I1[k * PD_ILD + 4 * tc + y];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) S[(4 * tr + x) * PD_LD + 64 + 4 * tc + y] = acc[x][y];
    }
    __syncthreads();
    // I21 = -I22 * T  and the final store of the whole inverse (leading dimension 128)
    {
        double acc[4][4];
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y] = 0.0;
        if (n2 > 0) {
            const int kmax = min(4 * tr + 4, n2);
            for (int k = 0; k < kmax; ++k) {
                double a[4], b[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) a[x] = I2[(4 * tr + x) * PD_ILD + k];
#pragma unroll
                for (int y = 0; y < 4; ++y) b[y] = S[k * PD_LD + 64 + 4 * tc + y];
#pragma unroll
                for (int x = 0; x < 4; ++x)
#pragma unroll
                    for (int y = 0; y < 4; ++y) acc[x][y] = fma(a[x], b[y], acc[x][y]);
            }
        }
#pragma unroll
        for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) {
                const int r = 4 * tr + x, c = 4 * tc + y;
                Linv[(int64_t)r * 128 + c] = I1[r * PD_ILD + c];                // I11
                Linv[(int64_t)r * 128 + 64 + c] = 0.0;                           // upper right
                Linv[(int64_t)(64 + r) * 128 + c] = -acc[x][y];                  // I21
                Linv[(int64_t)(64 + r) * 128 + 64 + c] = I2[r * PD_ILD + c];    // I22
            }
    }
}

static int potrf_diag(b2gp_ctx* ctx, cudaStream_t st, double* A, int64_t lda, int n, double* Linv_blk, int* info,
                      int index_base) {
    static bool attr = false;
    if (!attr) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM));
        attr = true;
    }
    potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM, st>>>(A, lda, n, Linv_blk, info, index_base);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B2GP_OK;
}

static inline int64_t split_point(int64_t n) {
    // first half rounded up to a multiple of the leaf, so only the last block can be ragged
    int64_t h = round_up((n + 1) / 2, B2GP_LEAF);
    return h >= n ? n - (n > B2GP_LEAF ? B2GP_LEAF : 0) : h;
}

// B (m x n, one right-hand side per row) <- B L^{-T}; L is n x n lower at `L`, its inverted diagonal
// blocks at `Linv` (block b0 first).
static int trsm_rec(b2gp_ctx* ctx, cudaStream_t st, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl,
                    int64_t n, const double* Linv) {
    if (m <= 0 || n <= 0) return B2GP_OK;
    if (n <= B2GP_LEAF) {
        // in place: C aliases A, one column tile
        return gemm_nt(ctx, st, m, n, n, 1.0, B, ldb, Linv, 128, 0.0, B, ldb, false);
    }
    const int64_t n1 = split_point(n), n2 = n - n1;
    RET_IF(trsm_rec(ctx, st, B, ldb, m, L, ldl, n1, Linv));
    RET_IF(gemm_nt(ctx, st, m, n2, n1, -1.0, B, ldb, L + n1 * ldl, ldl, 1.0, B + n1, ldb, false));
    return trsm_rec(ctx, st, B + n1, ldb, m, L + n1 * ldl + n1, ldl, n2, Linv + (n1 / B2GP_LEAF) * 128 * 128);
}

static int potrf_rec(b2gp_ctx* ctx, cudaStream_t st, double* A, int64_t lda, int64_t n, double* Linv, int* info,
                     int64_t index_base) {
    if (n <= 0) return B2GP_OK;
    if (n <= B2GP_LEAF) return potrf_diag(ctx, st, A, lda, (int)n, Linv, info, (int)index_base);
    const int64_t n1 = split_point(n), n2 = n - n1;
    RET_IF(potrf_rec(ctx, st, A, lda, n1, Linv, info, index_base));
    double* A21 = A + n1 * lda;
    double* A22 = A21 + n1;
    RET_IF(trsm_rec(ctx, st, A21, lda, n2, A, lda, n1, Linv));
    RET_IF(gemm_nt(ctx, st, n2, n2, n1, -1.0, A21, lda, A21, lda, 1.0, A22, lda, true));
    return potrf_rec(ctx, st, A22, lda, n2, Linv + (n1 / B2GP_LEAF) * 128 * 128, info, index_base + n1);
}

static inline int64_t linv_bytes(int64_t n) { return ceil_div(n, B2GP_LEAF) * 128 * 128 * (int64_t)sizeof(double); }
