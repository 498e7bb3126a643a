// gram.cuh -- fused pairwise-distance + covariance-function Gram build.
//
// Replaces square_scaled_distance + RBFKernel / MaternKernel / PeriodicKernel
// (gpax/kernels/kernels.py:28-41, 44-65, 68-91, 94-117): one kernel reads X[n,d], Z[m,d] and the
// hyper-parameters of one draw and writes K[n,m]; no [n,m] intermediate (XZ, r2, the [n,m,d]
// difference tensor of the periodic kernel) ever reaches HBM.
//
// The arithmetic follows the reference operation for operation so that rounding is of the same
// structure: scaled = X / lengthscale (a division), X2 = sum scaled^2, r2 = (X2 - 2*XZ) + Z2 clipped
// at 0, Matern r = sqrt(r2 + 1e-12), periodic arg = (pi * (x - z)) / period.  X2 and XZ use the
// same fma chain, so r2 of a point with itself is exactly 0.
//
// Roofline: HBM writes.  Algorithmic bytes = 8*n*m written + 8*(n+m)*d read; with
// B2GP_FLAG_LOWER_ONLY only tiles touching j <= i are produced (the factorisation reads nothing else).
// Tile: 64 rows x 128 columns per 256-thread CTA, each thread 16 rows x 2 adjacent columns, 16-byte
// stores (a warp writes 512 contiguous bytes per row).
#pragma once
#include "common.cuh"

constexpr int GRAM_BM = 64;
constexpr int GRAM_BN = 128;
constexpr int GRAM_THREADS = 256;
constexpr int GRAM_MAX_D = 64;

// theta layout of one draw (device memory): lengthscale[0..d), k_scale, noise, period
struct GramArgs {
    const double* X;
    const double* Z;
    int64_t n, m;
    int d;
    int kind;
    const double* theta;  // device pointer
    double noise_mult;    // diagonal term = noise * noise_mult + jitter, added iff same_xz
    double jitter;
    int same_xz;
    int lower_only;
    double* K;
    int64_t ldk;
};

__device__ __forceinline__ double cov_from_r2(int kind, double r2, double scale) {
    // kernels.py:61-62 (RBF) and kernels.py:85-88 (Matern-5/2 with the 1e-12 epsilon of :20-21)
    if (kind == B2GP_KERNEL_RBF) return scale * exp(-0.5 * r2);
    const double r = sqrt(r2 + 1e-12);
    const double s5r = 2.23606797749979 * r;  // 5**0.5 rounded to nearest double, as Python computes it
    return scale * (1.0 + s5r + (5.0 / 3.0) * r2) * exp(-s5r);
}

__global__ void __launch_bounds__(GRAM_THREADS) gram_kernel(const GramArgs p) {
    extern __shared__ __align__(16) double sm[];
    // layout: Xs[GRAM_BM][d] | x2[GRAM_BM] | Zt[d][GRAM_BN] | z2[GRAM_BN] | ell[d]
    const int d = p.d;
    double* Xs = sm;
    double* x2 = Xs + GRAM_BM * d;
    double* Zt = x2 + GRAM_BM;
    double* z2 = Zt + d * GRAM_BN;
    double* ell = z2 + GRAM_BN;

    const int64_t row0 = (int64_t)blockIdx.y * GRAM_BM;
    const int64_t col0 = (int64_t)blockIdx.x * GRAM_BN;
    if (p.lower_only && col0 > row0 + GRAM_BM - 1) return;
    const int tid = threadIdx.x;
    const bool periodic = (p.kind == B2GP_KERNEL_PERIODIC);

    if (tid < d) ell[tid] = p.theta[tid];
    __syncthreads();
    const double scale = p.theta[d];
    const double noise = p.theta[d + 1];
    const double period = p.theta[d + 2];

    // stage the rows: scaled by 1/lengthscale (division, kernels.py:35-36) for RBF/Matern, raw for periodic
    for (int idx = tid; idx < GRAM_BM * d; idx += GRAM_THREADS) {
        const int r = idx / d, k = idx % d;
        const int64_t gr = row0 + r;
        double v = (gr < p.n) ? p.X[gr * d + k] : 0.0;
        Xs[r * d + k] = periodic ? v : v / ell[k];
    }
    for (int idx = tid; idx < GRAM_BN * d; idx += GRAM_THREADS) {
        const int c = idx / d, k = idx % d;
        const int64_t gc = col0 + c;
        double v = (gc < p.m) ? p.Z[gc * d + k] : 0.0;
        Zt[k * GRAM_BN + c] = periodic ? v : v / ell[k];
    }
    __syncthreads();
    if (!periodic) {
        if (tid < GRAM_BM) {
            double s = 0.0;
            for (int k = 0; k < d; ++k) s = fma(Xs[tid * d + k], Xs[tid * d + k], s);
            x2[tid] = s;
        } else if (tid < GRAM_BM + GRAM_BN) {
            const int c = tid - GRAM_BM;
            double s = 0.0;
            for (int k = 0; k < d; ++k) s = fma(Zt[k * GRAM_BN + c], Zt[k * GRAM_BN + c], s);
            z2[c] = s;
        }
    }
    __syncthreads();

    const int cl = (tid & 63) * 2;  // local column pair
    const int rg = tid >> 6;        // 0..3
    const int64_t gc = col0 + cl;
    if (gc >= p.m) return;
    const bool has2 = (gc + 1 < p.m);
    const bool vec_ok = ((p.ldk & 1) == 0) && ((reinterpret_cast<uintptr_t>(p.K) & 15) == 0);
    const double diag_add = noise * p.noise_mult + p.jitter;

#pragma unroll 4
    for (int i = 0; i < GRAM_BM / 4; ++i) {
        const int rl = rg + 4 * i;
        const int64_t gr = row0 + rl;
        if (gr >= p.n) break;
        if (p.lower_only && gc > gr) continue;
        double v0, v1;
        if (!periodic) {
            double xz0 = 0.0, xz1 = 0.0;
            for (int k = 0; k < d; ++k) {
                const double x = Xs[rl * d + k];
                xz0 = fma(x, Zt[k * GRAM_BN + cl], xz0);
                xz1 = fma(x, Zt[k * GRAM_BN + cl + 1], xz1);
            }
            double r20 = (x2[rl] - 2.0 * xz0) + z2[cl];       // kernels.py:40
            double r21 = (x2[rl] - 2.0 * xz1) + z2[cl + 1];
            r20 = r20 < 0.0 ? 0.0 : r20;                      // kernels.py:41 (clip(0); NaN stays NaN)
            r21 = r21 < 0.0 ? 0.0 : r21;
            v0 = cov_from_r2(p.kind, r20, scale);
            v1 = cov_from_r2(p.kind, r21, scale);
        } else {
            double s0 = 0.0, s1 = 0.0;
            for (int k = 0; k < d; ++k) {
                const double x = Xs[rl * d + k];
                // kernels.py:111-113
                const double a0 = sin(3.141592653589793 * (x - Zt[k * GRAM_BN + cl]) / period) / ell[k];
                const double a1 = sin(3.141592653589793 * (x - Zt[k * GRAM_BN + cl + 1]) / period) / ell[k];
                s0 += a0 * a0;
                s1 += a1 * a1;
            }
            v0 = scale * exp(-2.0 * s0);
            v1 = scale * exp(-2.0 * s1);
        }
        if (p.same_xz) {  // kernels.py:63-64
            if (gr == gc) v0 += diag_add;
            if (gr == gc + 1) v1 += diag_add;
        }
        double* dst = p.K + gr * p.ldk + gc;
        const bool w1 = has2 && !(p.lower_only && gc + 1 > gr);
        if (w1 && vec_ok) {
            *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
        } else {
            dst[0] = v0;
            if (w1) dst[1] = v1;
        }
    }
}

// k(x,x) + diagonal term, computed exactly as the Gram kernel computes a diagonal entry
__device__ __forceinline__ double cov_self(int kind, double scale) {
    if (kind == B2GP_KERNEL_PERIODIC) return scale;  // sin(0) = 0, exp(-0) = 1
    return cov_from_r2(kind, 0.0, scale);
}

static int launch_gram(b2gp_ctx* ctx, cudaStream_t st, int kind, const double* X, int64_t n, const double* Z, int64_t m,
                       int d, const double* theta_dev, double noise_mult, double jitter, int same_xz, int lower_only,
                       double* K, int64_t ldk) {
    if (n <= 0 || m <= 0) return B2GP_OK;
    if (d < 1 || d > GRAM_MAX_D) return set_err(ctx, B2GP_ERR_UNSUPPORTED, "gram", "1 <= d <= 64", __FILE__, __LINE__);
    GramArgs a;
    a.X = X;
    a.Z = Z;
    a.n = n;
    a.m = m;
    a.d = d;
    a.kind = kind;
    a.theta = theta_dev;
    a.noise_mult = noise_mult;
    a.jitter = jitter;
    a.same_xz = same_xz;
    a.lower_only = lower_only;
    a.K = K;
    a.ldk = ldk;
    const size_t smem = (size_t)(GRAM_BM * d + GRAM_BM + d * GRAM_BN + GRAM_BN + d) * sizeof(double);
    static bool attr = false;
    if (!attr) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        attr = true;
    }
    dim3 grid((unsigned)ceil_div(m, GRAM_BN), (unsigned)ceil_div(n, GRAM_BM));
    gram_kernel<<<grid, GRAM_THREADS, smem, st>>>(a);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B2GP_OK;
}
