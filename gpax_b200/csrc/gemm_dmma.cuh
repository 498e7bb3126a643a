// gemm_dmma.cuh -- C[m,n] = beta*C + alpha * A[m,k] * B[n,k]^T in fp64 on the DMMA tensor pipe.
//
// This is the trailing-update kernel of the blocked Cholesky (SYRK when A == B and lower_only),
// the B <- B * Linv^T "triangular solve by inverted diagonal block" kernel, and the covariance
// epilogue cov = k_pp - V^T V.  It stands where the reference calls jnp.matmul on the outputs of
// jnp.linalg.inv (gpax/models/gp.py:271-273).
//
// sm_100a has no fp64 tcgen05 MMA (kind::f16/tf32/f8f6f4/i8/mx* only); the native fp64 tensor
// instruction is DMMA.8x8x4 (`mma.sync.m8n8k4.f64`; the m16n8k{4,8,16} PTX shapes lower to
// sequences of it, checked with cuobjdump).  Both operands are K-major (rows contiguous in k), so
// one shared-memory layout and one fragment pattern serve A and B:
//     a = As[(row0 + g) * LDS + 4*kk + t]      b = Bs[(col0 + g) * LDS + 4*kk + t]
// with g = lane/4, t = lane%4.  LDS = 20 doubles (160 B) makes the 8 rows x 32 B a half-warp reads
// land in 8 distinct 32-B bank groups -> conflict-free LDS.64.
//
// Pipeline: STAGES-deep cp.async (LDGSTS) ring of [BM+BN] x 16-double k-slices, one
// __syncthreads per k-slice.  Roofline: DMMA-bound; algorithmic flops per launch 2*m*n*k
// (m*n*k for the lower-only SYRK half).
#pragma once
#include "common.cuh"

struct GemmArgs {
    int m, n, k;
    const double* A;
    int64_t lda;
    const double* B;
    int64_t ldb;
    double* C;
    int64_t ldc;
    double alpha, beta;
    int lower_only;
    int tiles_m, tiles_n;
    int tile_base = 0;  // first 128x128 tile index covered by this launch (tail launches, see gemm_tma.cuh)
    int sub = 0;        // 1: blockIdx.x enumerates the four 64x64 quarters of 128x128 tiles tile_base, tile_base+1, ...
};

constexpr int GEMM_BK = 16;
constexpr int GEMM_LDS = 20;

__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem, const void* gmem, int src_bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Load ROWS x 16 doubles (rows [row0, row0+ROWS) of a row-major matrix, columns [k0, k0+16)) into
// shared memory with stride GEMM_LDS; out-of-range rows / columns are zero-filled (cp.async zfill).
template <int ROWS, int NT, bool ALIGNED>
__device__ __forceinline__ void load_slice(double* sm, const double* __restrict__ base, int64_t ld, int rows_total,
                                           int k_total, int row0, int k0, int tid) {
    constexpr int VEC = ALIGNED ? 2 : 1;
    constexpr int CPR = GEMM_BK / VEC;  // chunks per row
#pragma unroll
    for (int c = tid; c < ROWS * CPR; c += NT) {
        const int r = c / CPR;
        const int ch = c % CPR;
        const int gr = row0 + r;
        const int gk = k0 + ch * VEC;
        int bytes = (k_total - gk) * 8;
        bytes = bytes < 0 ? 0 : (bytes > VEC * 8 ? VEC * 8 : bytes);
        if (gr >= rows_total) bytes = 0;
        const double* src = bytes > 0 ? base + (int64_t)gr * ld + gk : base;
        double* dst = sm + r * GEMM_LDS + ch * VEC;
        if (ALIGNED)
            cp_async_16(dst, src, bytes);
        else
            cp_async_8(dst, src, bytes);
    }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, bool ALIGNED, int MINB>
__global__ void __launch_bounds__(WARPS_M* WARPS_N * 32, MINB) gemm_nt_kernel(const GemmArgs p) {
    constexpr int NT = WARPS_M * WARPS_N * 32;
    constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;
    constexpr int MI = WTM / 8, NI = WTN / 8;
    constexpr int STAGE_ELEMS = (BM + BN) * GEMM_LDS;
    extern __shared__ __align__(16) double smem[];

    int ti, tj;
    {
        const int x = p.sub ? p.tile_base + (int)(blockIdx.x >> 2) : p.tile_base + (int)blockIdx.x;
        const int tn = p.sub ? (p.tiles_n + 1) / 2 : p.tiles_n;   // tile columns in units of the indexed (128-wide) tiles
        if (p.lower_only) {
            int t = (int)((sqrt(8.0 * (double)x + 1.0) - 1.0) * 0.5);
            while ((int64_t)(t + 1) * (t + 2) / 2 <= x) ++t;
            while ((int64_t)t * (t + 1) / 2 > x) --t;
            ti = t;
            tj = x - t * (t + 1) / 2;
        } else {
            ti = x / tn;
            tj = x % tn;
        }
        if (p.sub) {
            ti = 2 * ti + (int)((blockIdx.x >> 1) & 1);
            tj = 2 * tj + (int)(blockIdx.x & 1);
            if (p.lower_only && tj > ti) return;
        }
    }
    const int row0 = ti * BM, col0 = tj * BN;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int wm = warp / WARPS_N, wn = warp % WARPS_N;
    const int g = lane >> 2, t4 = lane & 3;

    double acc[MI][NI][2];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const int KT = (p.k + GEMM_BK - 1) / GEMM_BK;

    // prologue
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            double* As = smem + s * STAGE_ELEMS;
            double* Bs = As + BM * GEMM_LDS;
            load_slice<BM, NT, ALIGNED>(As, p.A, p.lda, p.m, p.k, row0, s * GEMM_BK, tid);
            load_slice<BN, NT, ALIGNED>(Bs, p.B, p.ldb, p.n, p.k, col0, s * GEMM_BK, tid);
        }
        cp_async_commit();
    }

    for (int kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nk = kt + STAGES - 1;
            if (nk < KT) {
                double* As = smem + (nk % STAGES) * STAGE_ELEMS;
                double* Bs = As + BM * GEMM_LDS;
                load_slice<BM, NT, ALIGNED>(As, p.A, p.lda, p.m, p.k, row0, nk * GEMM_BK, tid);
                load_slice<BN, NT, ALIGNED>(Bs, p.B, p.ldb, p.n, p.k, col0, nk * GEMM_BK, tid);
            }
            cp_async_commit();
        }
        const double* As = smem + (kt % STAGES) * STAGE_ELEMS + (wm * WTM + g) * GEMM_LDS + t4;
        const double* Bs = smem + (kt % STAGES) * STAGE_ELEMS + BM * GEMM_LDS + (wn * WTN + g) * GEMM_LDS + t4;
#pragma unroll
        for (int kk = 0; kk < GEMM_BK / 4; ++kk) {
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = As[i * 8 * GEMM_LDS + kk * 4];
#pragma unroll
            for (int j = 0; j < NI; ++j) b[j] = Bs[j * 8 * GEMM_LDS + kk * 4];
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: C = beta*C + alpha*acc.  Fragment (i,j): row g, columns 2*t4, 2*t4+1.
    const bool vec_ok = ((p.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = row0 + wm * WTM + i * 8 + g;
        if (r >= p.m) continue;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int c = col0 + wn * WTN + j * 8 + t4 * 2;
            if (c >= p.n) continue;
            if (p.lower_only && c > r) continue;
            double* dst = p.C + (int64_t)r * p.ldc + c;
            const bool two = (c + 1 < p.n) && !(p.lower_only && c + 1 > r);
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            if (two && vec_ok) {
                if (p.beta != 0.0) {
                    const double2 old = *reinterpret_cast<const double2*>(dst);
                    v0 += p.beta * old.x;
                    v1 += p.beta * old.y;
                }
                *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            } else {
                if (p.beta != 0.0) v0 += p.beta * dst[0];
                dst[0] = v0;
                if (two) {
                    if (p.beta != 0.0) v1 += p.beta * dst[1];
                    dst[1] = v1;
                }
            }
        }
    }
}

template <int BM, int BN, int WARPS_M, int WARPS_N, int STAGES, int MINB>
static int launch_gemm_cfg(b2gp_ctx* ctx, cudaStream_t st, GemmArgs& a) {
    constexpr int smem_bytes = STAGES * (BM + BN) * GEMM_LDS * (int)sizeof(double);
    const bool aligned = ((a.lda & 1) == 0) && ((a.ldb & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.A) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(a.B) & 15) == 0);
    a.tiles_m = (a.m + BM - 1) / BM;
    a.tiles_n = (a.n + BN - 1) / BN;
    int64_t grid = a.lower_only ? (int64_t)a.tiles_m * (a.tiles_m + 1) / 2 : (int64_t)a.tiles_m * a.tiles_n;
    if (grid <= 0) return B2GP_OK;
    auto kern = aligned ? gemm_nt_kernel<BM, BN, WARPS_M, WARPS_N, STAGES, true, MINB>
                        : gemm_nt_kernel<BM, BN, WARPS_M, WARPS_N, STAGES, false, MINB>;
    static PerDeviceOnce attr_set[2];
    if (attr_set[aligned ? 1 : 0].need(ctx->device)) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set[aligned ? 1 : 0].done(ctx->device);
    }
    kern<<<(unsigned)grid, WARPS_M * WARPS_N * 32, smem_bytes, st>>>(a);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B2GP_OK;
}

static int gemm_tma_dispatch(b2gp_ctx* ctx, cudaStream_t st, GemmArgs& a);  // gemm_tma.cuh
static int ozaki_dispatch(b2gp_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                          const double* B, int64_t ldb, double* C, int64_t ldc, bool lower_only, bool overwrite, bool transB,
                          bool ktri);  // ozaki.cuh

// C = beta*C + alpha*A*B^T.  lower_only requires a square C (m == n) whose diagonal is the matrix
// diagonal.  `inplace_rows` marks the B <- B*Linv^T use where C aliases A: that is only safe with a
// single column tile (n <= 128), which the 128-wide configuration guarantees.
static int gemm_nt(b2gp_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                   int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, bool lower_only) {
    if (m <= 0 || n <= 0) return B2GP_OK;
    GemmArgs a;
    a.m = (int)m;
    a.n = (int)n;
    a.k = (int)k;
    a.A = A;
    a.lda = lda;
    a.B = B;
    a.ldb = ldb;
    a.C = C;
    a.ldc = ldc;
    a.alpha = alpha;
    a.beta = beta;
    a.lower_only = lower_only ? 1 : 0;
    if (lower_only && m < n) return set_err(ctx, B2GP_ERR_ARG, "gemm_nt", "lower_only needs m >= n", __FILE__, __LINE__);
    // Large rank-k updates C += alpha A B^T (the trailing updates of the factorisation and of the blocked solves) go
    // to the int8 tcgen05 path when it is enabled: measured 64 TFLOP/s-equivalent with 8 digit planes against
    // 35 for DMMA (ozaki.cuh).  It needs beta == 1, k within the int32 accumulation bound, enough 128x64 tiles to
    // fill the machine twice, and operands distinct from C (the in-place solve keeps the DMMA kernel).
    if (ctx->ozaki && beta == 1.0 && k >= 512 && C != A && C != B) {
        const int64_t tm = ceil_div(m, 128), tn = ceil_div(n, 64), sq = ceil_div(n, 128);
        const int64_t toz = lower_only ? sq * (sq + 1) + (tm - sq) * tn : tm * tn;   // lower triangle (+ the rows below it)
        if (toz >= ctx->oz_min_tiles) {
            const int rc = ozaki_dispatch(ctx, st, m, n, k, alpha, A, lda, B, ldb, C, ldc, lower_only, false, false, false);
            if (rc != B2GP_ERR_UNSUPPORTED) return rc;
        }
    }
    if (lower_only && m > n) {
        // trapezoid on the fp64 kernels (their lower-only tile maps are square): the square part, then the rows below it
        RET_IF(gemm_nt(ctx, st, n, n, k, alpha, A, lda, B, ldb, beta, C, ldc, true));
        return gemm_nt(ctx, st, m - n, n, k, alpha, A + n * lda, lda, B, ldb, beta, C + n * ldc, ldc, false);
    }
    // Tile choice.  A 128x128 tile keeps one SM busy for 128*128*k/64 cycles (DMMA: 64 fp64 FMA/clk/SM),
    // i.e. ~17 us per k = 128, however few tiles there are; when the 128x128 grid would leave most of
    // the 148 SMs idle, spend the same flops on more, smaller tiles.  The in-place triangular-solve
    // use (C aliases A, n <= 128) needs a single column tile, which all three configurations give.
    const int64_t tm128 = ceil_div(m, 128), tn128 = ceil_div(n, 128);
    const int64_t t128 = lower_only ? tm128 * (tm128 + 1) / 2 : tm128 * tn128;
    if (t128 >= 112) {
        if (ctx->use_tma) {
            const int rc = gemm_tma_dispatch(ctx, st, a);
            if (rc != B2GP_ERR_UNSUPPORTED) return rc;
        }
        // measured (tools/gemm_cfg.py): 3 stages beat 4 at large k; 16 warps (4 per SM sub-partition) beat 8 at small k
        if (k <= 1024) return launch_gemm_cfg<128, 128, 4, 4, 3, 1>(ctx, st, a);
        return launch_gemm_cfg<128, 128, 2, 4, 3, 1>(ctx, st, a);
    }
    if (lower_only) {
        // square tiles only for the triangular tile map
        return launch_gemm_cfg<64, 64, 2, 4, 4, 2>(ctx, st, a);
    }
    // latency-bound regime: minimise (waves) x (time of one tile), in units of a 32x128 tile; the 128x128 kernel holds
    // one CTA per SM (148 slots), the two smaller ones two (296 slots)
    const int64_t t64 = ceil_div(m, 64) * tn128, t32 = ceil_div(m, 32) * tn128;
    const int64_t c128 = ceil_div(t128, 148) * 4, c64 = ceil_div(t64, 296) * 2, c32 = ceil_div(t32, 296) * 1;
    if (c32 <= c64 && c32 <= c128) return launch_gemm_cfg<32, 128, 1, 8, 3, 2>(ctx, st, a);
    if (c64 <= c128) return launch_gemm_cfg<64, 128, 2, 4, 3, 2>(ctx, st, a);
    return launch_gemm_cfg<128, 128, 4, 4, 3, 1>(ctx, st, a);
}
