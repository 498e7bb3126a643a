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

// Short-latency 1/x and 1/sqrt(x) for the pivot chain of the diagonal block.  fp64 arithmetic has a
// dependent-issue latency of ~30 cycles here, so the number of dependent operations is what counts:
// the MUFU fp64 seeds (rcp/rsqrt.approx.ftz.f64, ~2^-20 relative) take one operation, and one cubic
// step (error e -> ~e^3 <= 2^-60, i.e. rounding level) takes three:
//     1/x:       e = 1 - x y;     y <- y + y (e + e^2)
//     1/sqrt(x): e = 1 - x y^2;   y <- y + y e (1/2 + 3/8 e)
// Non-positive / NaN pivots give NaN or Inf, which propagate as include/b200gp.h documents.
__device__ __forceinline__ double pivot_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
}
__device__ __forceinline__ double pivot_rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double xy = x * y;
    const double e = fma(-xy, y, 1.0);
    const double t = fma(0.375, e, 0.5);
    return fma(y * e, t, y);
}

constexpr int PD_LD = 132;       // shared-memory stride: 132 % 32 == 4 -> DMMA fragment loads (8 rows x 32 B) conflict-free
constexpr int PD_THREADS = 256;
constexpr int PD_SMEM = (128 * PD_LD + 32) * (int)sizeof(double);

// One CTA factors the n x n (n <= 128) lower-triangular block at A in shared memory, writes L back
// and writes inv(L) (lower, zero above the diagonal, leading dimension 128) to Linv.
//
//   for each 32-column panel b:
//     (a) warp 0 factors the 32x32 diagonal block in REGISTERS (lane i owns row i; pivots and
//         multipliers travel by warp shuffle), writes it to global, then inverts it in registers and
//         leaves inv(D_b) in the block's place in shared memory;
//     (b) all warps: panel rows below  X = A_panel * inv(D_b)^T      (DMMA, operands in shared memory)
//     (c) all warps: trailing update   S22 -= X X^T on 8x8 tiles      (DMMA)
//   then the inverse is assembled in place by the 2x2 block formula, two levels (32 -> 64 -> 128),
//   four DMMA products with row / column ownership so that every product can overwrite its operand.
//
// A non-positive or NaN pivot records *info = index_base + j + 1 (first one wins) and lets NaNs
// propagate; callers NaN-fill the outputs of that draw.
__global__ void __launch_bounds__(PD_THREADS, 1)
potrf_diag_kernel(double* __restrict__ A, int64_t lda, int n, double* __restrict__ Linv, int* info, int index_base,
                  long long* prof) {
    extern __shared__ __align__(16) double sm[];
    double* S = sm;
    double* RD = sm + 128 * PD_LD;  // reciprocal pivots of the current diagonal block
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const unsigned FULL = 0xffffffffu;
    int pslot = 0;
#define PD_PROF()                                                        \
    do {                                                                 \
        if (prof && tid == 0) {                                          \
            long long gt_;                                               \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));      \
            prof[2 * pslot] = clock64();                                 \
            prof[2 * pslot + 1] = gt_;                                   \
            ++pslot;                                                     \
        }                                                                \
    } while (0)
    PD_PROF();

    const bool gvec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (gvec) {
        // 8192 double2 elements, 32 per thread, 8 loads in flight per thread
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = (it * 8 + u) * PD_THREADS + tid;
                const int i = idx >> 6, j = (idx & 63) * 2;
                v[u] = make_double2(0.0, 0.0);
                if (i < n && j <= i) v[u] = *reinterpret_cast<const double2*>(A + (int64_t)i * lda + j);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = (it * 8 + u) * PD_THREADS + tid;
                const int i = idx >> 6, j = (idx & 63) * 2;
                if (j + 1 > i) v[u].y = 0.0;
                *reinterpret_cast<double2*>(&S[i * PD_LD + j]) = v[u];
            }
        }
    } else {
        for (int idx = tid; idx < 128 * 128; idx += PD_THREADS) {
            const int i = idx >> 7, j = idx & 127;
            double v = 0.0;
            if (i < n && j <= i) v = A[(int64_t)i * lda + j];
            S[i * PD_LD + j] = v;
        }
    }
    __syncthreads();
    PD_PROF();

    const int nb = (n + 31) >> 5;
    for (int b = 0; b < nb; ++b) {
        const int b0 = 32 * b;
        const int bw = min(32, n - b0);
        if (warp == 0) {
            // ---- (a) 32x32 diagonal block in registers
            double a[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) a[k] = S[(b0 + lane) * PD_LD + b0 + k];
            if (lane >= bw) {
#pragma unroll
                for (int k = 0; k < 32; ++k) a[k] = (k == lane) ? 1.0 : 0.0;  // identity padding of a ragged block
            }
            if (b == 0) PD_PROF();
            // Unnormalised right-looking elimination: a[] holds the running Schur complement, the update
            // of column k by column j is a[k] -= (a[j] / piv_j) * a_kj, so the per-column dependency chain is
            //   shfl(piv) -> 1/piv (4 ops) -> a[j]/piv -> fma   and the rsqrt scaling of L is off the chain.
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const double piv = __shfl_sync(FULL, a[j], j);
                if (!(piv > 0.0) && lane == 0 && j < bw) atomicCAS(info, 0, index_base + b0 + j + 1);
                const double r = pivot_rcp(piv);
                const double ajr = a[j] * r;
#pragma unroll
                for (int k = j + 1; k < 32; ++k) {
                    const double akj = __shfl_sync(FULL, a[j], k);
                    a[k] = fma(-ajr, akj, a[k]);
                }
                const double rs = pivot_rsqrt(piv);                 // 1 / L_jj
                double sq = piv * rs;                               // L_jj = sqrt(piv), one Newton correction
                sq = fma(fma(-sq, sq, piv), 0.5 * rs, sq);
                a[j] = (lane == j) ? sq : a[j] * rs;
                if (lane == j) RD[j] = rs;
            }
            if (b == 0) PD_PROF();
            if (lane < bw) {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k <= lane) {
                        A[(int64_t)(b0 + lane) * lda + b0 + k] = a[k];
                        S[(b0 + lane) * PD_LD + b0 + k] = a[k];
                    }
            } else {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                    if (k <= lane) S[(b0 + lane) * PD_LD + b0 + k] = a[k];
            }
            __syncwarp();
            if (b == 0) PD_PROF();
            // inverse of the block: lane c computes column c by forward substitution, L broadcast from smem
            double x[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int k = 0; k < i; ++k) {
                    const double l = S[(b0 + i) * PD_LD + b0 + k];
                    if (k & 1)
                        s1 = fma(l, x[k], s1);
                    else
                        s0 = fma(l, x[k], s0);
                }
                const double rdi = RD[i];
                x[i] = (lane == i) ? rdi : -rdi * (s0 + s1);
                if (i < lane) x[i] = 0.0;
            }
            __syncwarp();
            if (b == 0) PD_PROF();
#pragma unroll
            for (int i = 0; i < 32; ++i) S[(b0 + i) * PD_LD + b0 + lane] = x[i];
        }
        __syncthreads();
        PD_PROF();
        const int r_begin = b0 + 32;
        if (r_begin < n) {
            const int ntile = (n - r_begin + 7) >> 3;
            // ---- (b) X = A_panel * inv(D)^T : X[r][c] = sum_k A[r][k] * Dinv[c][k]
            for (int rt = warp; rt < ntile; rt += PD_THREADS / 32) {
                const int r0 = r_begin + 8 * rt;
                double af[8];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) af[kk] = S[(r0 + g) * PD_LD + b0 + 4 * kk + t];
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    double c0 = 0.0, c1 = 0.0;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const double bb = S[(b0 + 8 * j + g) * PD_LD + b0 + 4 * kk + t];
                        dmma884(c0, c1, af[kk], bb);
                    }
                    *reinterpret_cast<double2*>(&S[(r0 + g) * PD_LD + b0 + 8 * j + 2 * t]) = make_double2(c0, c1);
                }
            }
            __syncthreads();
            // ---- (c) trailing update on 8x8 tiles of the lower triangle
            const int npairs = ntile * (ntile + 1) / 2;
            // three tile pairs in flight per warp: DMMA accumulation chains are ~8 deep and slow to retire
            for (int p0 = 3 * warp; p0 < npairs; p0 += 3 * (PD_THREADS / 32)) {
                int r0[3], c0i[3];
                double acc[3][2];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    const int p = min(p0 + u, npairs - 1);
                    int rt = 0;
                    while ((rt + 1) * (rt + 2) / 2 <= p) ++rt;
                    r0[u] = r_begin + 8 * rt;
                    c0i[u] = r_begin + 8 * (p - rt * (rt + 1) / 2);
                    acc[u][0] = acc[u][1] = 0.0;
                }
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const double aa = S[(r0[u] + g) * PD_LD + b0 + 4 * kk + t];
                        const double bb = S[(c0i[u] + g) * PD_LD + b0 + 4 * kk + t];
                        dmma884(acc[u][0], acc[u][1], aa, bb);
                    }
                }
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    if (p0 + u < npairs) {
                        double2* dst = reinterpret_cast<double2*>(&S[(r0[u] + g) * PD_LD + c0i[u] + 2 * t]);
                        double2 old = *dst;
                        old.x -= acc[u][0];
                        old.y -= acc[u][1];
                        *dst = old;
                    }
                }
            }
            __syncthreads();
        }
        PD_PROF();
    }

    // ---- write the off-diagonal-block part of L back (diagonal blocks went out in (a))
    if (gvec) {
#pragma unroll 1
        for (int it = 0; it < 4; ++it) {
            double2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = (it * 8 + u) * PD_THREADS + tid;
                const int i = idx >> 6, j = (idx & 63) * 2;
                v[u] = *reinterpret_cast<const double2*>(&S[i * PD_LD + j]);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int idx = (it * 8 + u) * PD_THREADS + tid;
                const int i = idx >> 6, j = (idx & 63) * 2;
                if (i < n && (j >> 5) < (i >> 5)) *reinterpret_cast<double2*>(A + (int64_t)i * lda + j) = v[u];
            }
        }
    } else {
        for (int idx = tid; idx < 128 * 128; idx += PD_THREADS) {
            const int i = idx >> 7, j = idx & 127;
            if (i < n && (j >> 5) < (i >> 5)) A[(int64_t)i * lda + j] = S[i * PD_LD + j];
        }
    }
    __syncthreads();
    PD_PROF();

    // ---- inverse assembly, level 1: the two 64-blocks.  M = -P1 * (L10 * P0), in place of L10.
    {
        const int base = 64 * (warp >> 2);
        const int q = warp & 3;
        // step 1: T = L10 * P0   (row ownership: this warp owns rows r0..r0+7 of T)
        {
            const int r0 = base + 32 + 8 * q;
            double af[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) af[kk] = S[(r0 + g) * PD_LD + base + 4 * kk + t];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    if (kk < 2 * j) continue;  // P0 is lower triangular: rows k < 8j of this column tile are zero
                    const double bb = S[(base + 4 * kk + t) * PD_LD + base + 8 * j + g];
                    dmma884(c0, c1, af[kk], bb);
                }
                *reinterpret_cast<double2*>(&S[(r0 + g) * PD_LD + base + 8 * j + 2 * t]) = make_double2(c0, c1);
            }
        }
        __syncthreads();
        // step 2: M10 = -P1 * T   (column ownership: this warp owns columns c0..c0+7)
        {
            const int c0i = base + 8 * q;
            double bf[8];
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) bf[kk] = S[(base + 32 + 4 * kk + t) * PD_LD + c0i + g];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    if (kk > 2 * i + 1) continue;  // P1 lower triangular: k <= r
                    const double aa = S[(base + 32 + 8 * i + g) * PD_LD + base + 32 + 4 * kk + t];
                    dmma884(c0, c1, aa, bf[kk]);
                }
                *reinterpret_cast<double2*>(&S[(base + 32 + 8 * i + g) * PD_LD + c0i + 2 * t]) = make_double2(-c0, -c1);
            }
        }
        __syncthreads();
    }
    // ---- level 2: M_BA = -I_B * (L_BA * I_A) with I_A = S[0:64,0:64], I_B = S[64:128,64:128], L_BA = S[64:128,0:64]
    {
        // step 3: T2 = L_BA * I_A   (row ownership)
        {
            const int r0 = 64 + 8 * warp;
            double af[16];
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) af[kk] = S[(r0 + g) * PD_LD + 4 * kk + t];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    if (kk < 2 * j) continue;
                    const double bb = S[(4 * kk + t) * PD_LD + 8 * j + g];
                    dmma884(c0, c1, af[kk], bb);
                }
                *reinterpret_cast<double2*>(&S[(r0 + g) * PD_LD + 8 * j + 2 * t]) = make_double2(c0, c1);
            }
        }
        __syncthreads();
        // step 4: M_BA = -I_B * T2   (column ownership)
        {
            const int c0i = 8 * warp;
            double bf[16];
#pragma unroll
            for (int kk = 0; kk < 16; ++kk) bf[kk] = S[(64 + 4 * kk + t) * PD_LD + c0i + g];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                double c0 = 0.0, c1 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) {
                    if (kk > 2 * i + 1) continue;
                    const double aa = S[(64 + 8 * i + g) * PD_LD + 64 + 4 * kk + t];
                    dmma884(c0, c1, aa, bf[kk]);
                }
                *reinterpret_cast<double2*>(&S[(64 + 8 * i + g) * PD_LD + c0i + 2 * t]) = make_double2(-c0, -c1);
            }
        }
        __syncthreads();
    }
    PD_PROF();
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        double2 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = (it * 8 + u) * PD_THREADS + tid;
            const int i = idx >> 6, j = (idx & 63) * 2;
            v[u] = *reinterpret_cast<const double2*>(&S[i * PD_LD + j]);
            if (j > i) v[u].x = 0.0;
            if (j + 1 > i) v[u].y = 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = (it * 8 + u) * PD_THREADS + tid;
            *reinterpret_cast<double2*>(Linv + 2 * (int64_t)idx) = v[u];
        }
    }
    __syncthreads();
    PD_PROF();
#undef PD_PROF
}

static int potrf_diag(b2gp_ctx* ctx, cudaStream_t st, double* A, int64_t lda, int n, double* Linv_blk, int* info,
                      int index_base) {
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM));
        attr.done(ctx->device);
    }
    potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM, st>>>(A, lda, n, Linv_blk, info, index_base, nullptr);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B2GP_OK;
}

static inline int64_t split_point(int64_t n) {
    // first half rounded up to a multiple of the leaf, so only the last block can be ragged
    int64_t h = round_up((n + 1) / 2, B2GP_LEAF);
    return h >= n ? n - (n > B2GP_LEAF ? B2GP_LEAF : 0) : h;
}

// ---------------------------------------------------------------------------------------------- strip solve
// B (m x n) <- B L^{-T} for a narrow factor (n <= 512) in ONE launch: the rows of B are independent, so a CTA takes a
// 32-row strip through the whole block forward substitution
//     for each 128-column block j:   T = B_j - sum_{k < 128 j} X[:, k] L[j-block, k]^T ;   X_j = T Linv_j^T
// with the same cp.async / DMMA tile loop as gemm_nt_kernel<32, 128, 1, 8, 3>, reading its own earlier results back
// through L2.  The recursion it replaces issues 2^(log2(n/128)+1) - 1 dependent launches (7 for n = 512) of thin k = 128
// GEMMs whose cost is all prologue and epilogue.
struct TrsmStripArgs {
    double* B;
    int64_t ldb;
    int m, n;
    const double* L;
    int64_t ldl;
    const double* Linv;  // 128 x 128 inverted diagonal blocks, block 0 first
};

constexpr int TS_BM = 32, TS_BN = 128, TS_STAGES = 3, TS_THREADS = 256;
constexpr int TS_STAGE_ELEMS = (TS_BM + TS_BN) * GEMM_LDS;
constexpr int TS_SMEM = TS_STAGES * TS_STAGE_ELEMS * (int)sizeof(double);

// acc (32 x 128 tile, warp w owns columns 16 w .. 16 w + 15) = A[row0.., 0..k) * Bop[0..brows, 0..k)^T
template <bool ALIGNED>
__device__ __forceinline__ void ts_tile_gemm(double (&acc)[4][2][2], double* smem, const double* A, int64_t lda, int a_rows, int row0,
                                             const double* Bop, int64_t ldbop, int b_rows, int k, int tid) {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const int KT = (k + GEMM_BK - 1) / GEMM_BK;
#pragma unroll
    for (int s = 0; s < TS_STAGES - 1; ++s) {
        if (s < KT) {
            double* As = smem + s * TS_STAGE_ELEMS;
            load_slice<TS_BM, TS_THREADS, ALIGNED>(As, A, lda, a_rows, k, row0, s * GEMM_BK, tid);
            load_slice<TS_BN, TS_THREADS, ALIGNED>(As + TS_BM * GEMM_LDS, Bop, ldbop, b_rows, k, 0, s * GEMM_BK, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<TS_STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + TS_STAGES - 1;
            if (nk < KT) {
                double* As = smem + (nk % TS_STAGES) * TS_STAGE_ELEMS;
                load_slice<TS_BM, TS_THREADS, ALIGNED>(As, A, lda, a_rows, k, row0, nk * GEMM_BK, tid);
                load_slice<TS_BN, TS_THREADS, ALIGNED>(As + TS_BM * GEMM_LDS, Bop, ldbop, b_rows, k, 0, nk * GEMM_BK, tid);
            }
            cp_async_commit();
        }
        const double* As = smem + (kt % TS_STAGES) * TS_STAGE_ELEMS + g * GEMM_LDS + t4;
        const double* Bs = smem + (kt % TS_STAGES) * TS_STAGE_ELEMS + TS_BM * GEMM_LDS + (warp * 16 + g) * GEMM_LDS + t4;
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; ++kk) {
            double a[4], b[2];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[i * 8 * GEMM_LDS + kk * 4];
#pragma unroll
            for (int j = 0; j < 2; ++j) b[j] = Bs[j * 8 * GEMM_LDS + kk * 4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();
    __syncthreads();   // every thread is done with the stages (the next call refills them) and with its reads of A
}

template <bool ALIGNED>
__global__ void __launch_bounds__(TS_THREADS, 2) trsm_strip_kernel(const TrsmStripArgs p) {
    extern __shared__ __align__(16) double ts_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
    const int row0 = blockIdx.x * TS_BM;
    const int nblk = (p.n + 127) / 128;
    double acc[4][2][2];
    for (int jb = 0; jb < nblk; ++jb) {
        const int c0 = jb * 128;
        const int cw = p.n - c0 < 128 ? p.n - c0 : 128;
        if (jb > 0) {
            // T = B_j - X[:, 0..c0) L[c0.., 0..c0)^T, written over B_j
            ts_tile_gemm<ALIGNED>(acc, ts_smem, p.B, p.ldb, p.m, row0, p.L + (int64_t)c0 * p.ldl, p.ldl, cw, c0, tid);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = row0 + i * 8 + g;
                if (r >= p.m) continue;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int c = warp * 16 + j * 8 + t4 * 2;
                    double* dst = p.B + (int64_t)r * p.ldb + c0 + c;
                    if (c < cw) dst[0] = -acc[i][j][0] + dst[0];
                    if (c + 1 < cw) dst[1] = -acc[i][j][1] + dst[1];
                }
            }
            __syncthreads();   // T complete (CTA-private rows) before it is read back as the A operand
        }
        // X_j = T Linv_j^T, in place: all of T is in flight / in shared memory before the first store (see ts_tile_gemm's tail)
        ts_tile_gemm<ALIGNED>(acc, ts_smem, p.B + c0, p.ldb, p.m, row0, p.Linv + (int64_t)jb * 128 * 128, 128, cw, cw, tid);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = row0 + i * 8 + g;
            if (r >= p.m) continue;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int c = warp * 16 + j * 8 + t4 * 2;
                double* dst = p.B + (int64_t)r * p.ldb + c0 + c;
                if (c < cw) dst[0] = acc[i][j][0];
                if (c + 1 < cw) dst[1] = acc[i][j][1];
            }
        }
        __syncthreads();       // X_j visible to this CTA's later loads
    }
}

static int launch_trsm_strip(b2gp_ctx* ctx, cudaStream_t st, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n,
                             const double* Linv) {
    TrsmStripArgs a;
    a.B = B;
    a.ldb = ldb;
    a.m = (int)m;
    a.n = (int)n;
    a.L = L;
    a.ldl = ldl;
    a.Linv = Linv;
    const bool aligned = ((ldb & 1) == 0) && ((ldl & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(L) & 15) == 0) && ((reinterpret_cast<uintptr_t>(Linv) & 15) == 0);
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(trsm_strip_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
        CUDA_TRY(ctx, cudaFuncSetAttribute(trsm_strip_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SMEM));
        attr.done(ctx->device);
    }
    const unsigned grid = (unsigned)ceil_div(m, TS_BM);
    if (aligned)
        trsm_strip_kernel<true><<<grid, TS_THREADS, TS_SMEM, st>>>(a);
    else
        trsm_strip_kernel<false><<<grid, TS_THREADS, TS_SMEM, st>>>(a);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B2GP_OK;
}

// B (m x n, one right-hand side per row) <- B L^{-T}; L is n x n lower at `L`, its inverted diagonal
// blocks at `Linv` (block b0 first).
static int panel_solve_all_rows(b2gp_ctx* ctx, cudaStream_t st, Slot& sl, double* rows, int64_t ldr, int64_t r, const double* L,
                                int64_t ldl, int64_t n, const double* Linv128, double* Ukeep = nullptr);

// `panel_route` = false inside panel_solve_all_rows itself: forming U = I L^{-T} must not recurse into the panel route
// (its scratch, Slot::panelU, is the U being formed).
static int trsm_rec(b2gp_ctx* ctx, cudaStream_t st, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl,
                    int64_t n, const double* Linv, bool panel_route = true) {
    if (m <= 0 || n <= 0) return B2GP_OK;
    // many right-hand sides against a factor block of at most `panel` columns: the explicit inverse of the block and ONE
    // int8 tcgen05 GEMM over all rows (see potrf_tall) instead of m/32 strips at a fraction of the DMMA rate
    if (panel_route && n > B2GP_LEAF && n <= ctx->panel && m >= 1024 && ctx->ozaki != 0) {
        Slot* sl = slot_of(ctx, st);
        if (sl) return panel_solve_all_rows(ctx, st, *sl, B, ldb, m, L, ldl, n, Linv);
    }
    if (n <= B2GP_LEAF) {
        // in place: C aliases A, one column tile
        return gemm_nt(ctx, st, m, n, n, 1.0, B, ldb, Linv, 128, 0.0, B, ldb, false);
    }
    if (n <= ctx->trsm_strip) return launch_trsm_strip(ctx, st, B, ldb, m, L, ldl, n, Linv);
    const int64_t n1 = split_point(n), n2 = n - n1;
    RET_IF(trsm_rec(ctx, st, B, ldb, m, L, ldl, n1, Linv, panel_route));
    RET_IF(gemm_nt(ctx, st, m, n2, n1, -1.0, B, ldb, L + n1 * ldl, ldl, 1.0, B + n1, ldb, false));
    return trsm_rec(ctx, st, B + n1, ldb, m, L + n1 * ldl + n1, ldl, n2, Linv + (n1 / B2GP_LEAF) * 128 * 128, panel_route);
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

// ---------------------------------------------------------------------------------------------- tall-panel factorisation
// potrf_tall(A, n, r): factor the n x n diagonal block at A and solve the r rows stored below it,
//     A[n .. n+r, 0 .. n)  <-  A[n .. n+r, 0 .. n) L^{-T},
// where "the rows below" are the rest of the matrix AND any right-hand-side rows appended to it (the posterior stores
// [k_pX; y^T] under k_XX, so the solve V^T = k_pX L^{-T} of gp.py:272-273 rides along with the factorisation's own panel
// solves instead of being a second pass over L).  Recursion:
//     n <= panel:  potrf_rec on the block (128-wide leaves, fp64), U = L^{-T} by solving the identity, then ONE int8
//                  tcgen05 GEMM  rows <- rows U  (k = panel, B operand read transposed and k-triangular, C overwrites A's
//                  storage) for all r rows at once;
//     else:        potrf_tall(A11, n1, n2 + r);   [A22; E2] -= [A21; E1] A21^T  (one int8 GEMM over the lower
//                  trapezoid, k = n1);   potrf_tall(A22, n2, r).
// Against potrf_rec / trsm_rec this replaces the trsm recursion (2 N / 128 strip and thin-GEMM launches at 5-13 % of the
// DMMA peak) by N / panel machine-filling int8 GEMMs at the cost of 2x the flops of the diagonal-block solves
// (N^2 panel flops, 1.4e11 at N = 16384 against 1.5e12 for the factorisation), and every GEMM is as tall as the matrix.
static inline bool use_tall(const b2gp_ctx* ctx, int64_t n) { return ctx->ozaki != 0 && ctx->panel >= 128 && n >= ctx->tall_min; }

// `Ukeep` (n x round_up(n, 8) doubles, caller's storage) receives U instead of the slot's scratch: the factor cache keeps
// the explicit inverses of the diagonal blocks so that later solves against the same factor (trsm_tall) need not redo them.
static int panel_solve_all_rows(b2gp_ctx* ctx, cudaStream_t st, Slot& sl, double* rows, int64_t ldr, int64_t r, const double* L,
                                int64_t ldl, int64_t n, const double* Linv128, double* Ukeep) {
    const int64_t ldu = round_up(n, 8);
    if (!Ukeep) RET_IF(ensure(ctx, sl.panelU, (size_t)n * ldu * 8));
    double* U = Ukeep ? Ukeep : (double*)sl.panelU.p;
    set_identity_kernel<<<grid_for(n * n), 256, 0, st>>>(U, ldu, n);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    RET_IF(trsm_rec(ctx, st, U, ldu, n, L, ldl, n, Linv128, false));   // U = I L^{-T}
    // rows <- rows L^{-T} = rows (L^{-1})^T: NT GEMM whose B operand L^{-1} is U read transposed
    return ozaki_dispatch(ctx, st, r, n, n, 1.0, rows, ldr, U, ldu, rows, ldr, false, true, true, true);
}

// `Ukeep`: optional storage of panel x panel doubles per diagonal block (block b at Ukeep + b panel^2) that receives the
// blocks' explicit inverses U_b = L_bb^{-T} (leading dimension round_up(block size, 8)); see trsm_tall.
static int potrf_tall(b2gp_ctx* ctx, cudaStream_t st, Slot& sl, double* A, int64_t lda, int64_t n, int64_t r, double* Linv128,
                      int* info, int64_t index_base, double* Ukeep = nullptr) {
    if (n <= 0) return B2GP_OK;
    const int64_t NB = ctx->panel;
    if (n <= NB) {
        RET_IF(potrf_rec(ctx, st, A, lda, n, Linv128, info, index_base));
        double* Ub = Ukeep ? Ukeep + (index_base / NB) * NB * NB : nullptr;
        if (r > 0) RET_IF(panel_solve_all_rows(ctx, st, sl, A + n * lda, lda, r, A, lda, n, Linv128, Ub));
        return B2GP_OK;
    }
    const int64_t nblk = ceil_div(n, NB);
    const int64_t n1 = (nblk + 1) / 2 * NB, n2 = n - n1;
    RET_IF(potrf_tall(ctx, st, sl, A, lda, n1, n2 + r, Linv128, info, index_base, Ukeep));
    double* Pn = A + n1 * lda;   // [A21; E1]: n2 + r rows, n1 columns, solved
    RET_IF(gemm_nt(ctx, st, n2 + r, n2, n1, -1.0, Pn, lda, Pn, lda, 1.0, Pn + n1, lda, true));
    return potrf_tall(ctx, st, sl, Pn + n1, lda, n2, r, Linv128 + (n1 / B2GP_LEAF) * 128 * 128, info, index_base + n1, Ukeep);
}

// B (m rows, one right-hand side per row) <- B L^{-T} against a factor whose diagonal blocks' explicit inverses were kept
// by potrf_tall (`Ukeep`, block width NB): per block column  B_b <- B_b U_b  (int8 GEMM, k = NB, k-triangular) and
// B[:, after b] -= B_b L[after b, b]^T.  2 N / NB machine-wide GEMMs instead of trsm_rec's 2 N / 128 strip and thin-GEMM
// launches -- the solve of every posterior call that reuses a cached factor (chunk loops, repeated predictions).
static int trsm_tall(b2gp_ctx* ctx, cudaStream_t st, double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t N,
                     const double* Ukeep, int64_t NB) {
    for (int64_t c0 = 0; c0 < N; c0 += NB) {
        const int64_t n = N - c0 < NB ? N - c0 : NB, rest = N - c0 - n;
        const double* U = Ukeep + (c0 / NB) * NB * NB;
        RET_IF(ozaki_dispatch(ctx, st, m, n, n, 1.0, B + c0, ldb, U, round_up(n, 8), B + c0, ldb, false, true, true, true));
        if (rest > 0) RET_IF(gemm_nt(ctx, st, m, rest, n, -1.0, B + c0, ldb, L + (c0 + n) * ldl + c0, ldl, 1.0, B + c0 + n, ldb, false));
    }
    return B2GP_OK;
}

// factorisation (+ solve of r appended rows) by whichever scheme fits the size
static int potrf_auto(b2gp_ctx* ctx, cudaStream_t st, double* A, int64_t lda, int64_t n, int64_t r, double* Linv128, int* info) {
    Slot* sl = slot_of(ctx, st);
    if (sl && use_tall(ctx, n)) return potrf_tall(ctx, st, *sl, A, lda, n, r, Linv128, info, 0);
    RET_IF(potrf_rec(ctx, st, A, lda, n, Linv128, info, 0));
    if (r > 0) RET_IF(trsm_rec(ctx, st, A + n * lda, lda, r, A, lda, n, Linv128));
    return B2GP_OK;
}

static inline int64_t linv_bytes(int64_t n) { return ceil_div(n, B2GP_LEAF) * 128 * 128 * (int64_t)sizeof(double); }
