// common.cuh -- context, error handling, device buffers for libb200gp.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200gp.h"

#define B2GP_MAX_STREAMS 16
#define B2GP_LEAF 128  // diagonal-block size of the factorisation (one CTA, shared memory)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

#define OZ_LISTS 192
// scratch of the int8-split (Ozaki) GEMM path, one per stream: digit planes, row scales, tile list
struct OzTileList {
    int tm = -1, tn = -1, lower = -1, cl = -1;
    int64_t count = 0;
    DevBuf dev;
    std::vector<int2> host;         // kept alive for the asynchronous upload
};
struct OzWork {
    DevBuf planesA, planesB, scaleA, scaleB, prof;
    OzTileList lists[OZ_LISTS];     // the (tiles_m, tiles_n, lower) shapes one factorisation cycles through
    int next_list = 0;
};

// One "slot" = the workspace of one posterior draw in flight.
struct Slot {
    cudaStream_t stream = nullptr;
    DevBuf A;      // N x ldA      k_XX, then its factor L (lower)
    DevBuf Vt;     // (P+1) x ldV  rows 0..P-1 = k_pX (gp.py:268), row P = y_res; then V^T, w^T
    DevBuf Linv;   // nblk x 128 x 128 inverted diagonal blocks of L
    DevBuf cov;    // P x ldC      posterior covariance / its factor
    DevBuf LinvC;  // inverted diagonal blocks of chol(cov)
    DevBuf misc;   // small scratch
    DevBuf panelU; // panel x panel scratch of the tall-panel factorisation: L_jj^{-T} of the current diagonal block
    OzWork oz;
    int oz_planes = 7;  // digit planes of the int8 path for the work queued on this slot when ctx->ozaki == -1 (auto)
    cudaEvent_t ev[8];
};

struct b2gp_ctx {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t mem_bytes = 0;
    int n_streams = 2;
    int use_tma = 1;  // large GEMMs through the TMA / mbarrier persistent kernel (gemm_tma.cuh)
    int enqueue_threads = 1;  // queue the draws of a multi-draw posterior from one host thread per slot
    int big_grid = 0;        // CTAs of the persistent kernels (0 = one per SM); fewer leaves SMs for other streams' small kernels
    int oz_min_tiles = 148;  // smallest 128x64-tile count handed to the int8 path
    int trsm_strip = 256;  // widest factor solved by the one-launch strip kernel (0: recurse down to the 128 leaves)
    int oz_cluster = 2;  // 2: CTA pairs share the A digit planes by TMA multicast; 1: independent CTAs
    int panel = 1024;      // diagonal-block width of the tall-panel factorisation (potrf_tall); 0: recursive potrf_rec / trsm_rec only
    int tall_min = 2048;   // smallest N factored by potrf_tall
    int oz_debug = 0;  // see OzArgs::debug (0 in production)
    // 0: fp64 DMMA only; 6 / 7: large rank-k updates through the int8 tcgen05 path with that many base-256 digit planes
    // (46 / 54 bits per operand); -1: 6 or 7 per factorisation from a bound on cond(K), see oz_auto_planes()
    int ozaki = -1;
    Slot slots[B2GP_MAX_STREAMS];
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_a = nullptr, ev_b = nullptr;
    // staging for host-pointer entry points
    DevBuf d_in[8];
    DevBuf d_out[4];
    DevBuf d_info;
    // factor bookkeeping for b2gp_trsm_lower (host-pointer mode keeps the factor resident)
    DevBuf last_linv;
    int64_t last_n = 0;
    std::atomic<int64_t> launches{0};  // kernels queued (draws may be queued from several host threads)
    std::string err;
};

static inline int set_err(b2gp_ctx* ctx, int code, const char* what, const char* detail, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s:%d)", what, detail ? detail : "", file, line);
    if (ctx) {
        static std::mutex mu;  // draws of one call may be queued (and fail) on several host threads
        std::lock_guard<std::mutex> lock(mu);
        ctx->err = buf;
    }
    return code;
}

#define CUDA_TRY(ctx, expr)                                                                        \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return set_err((ctx), B2GP_ERR_CUDA, #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define ARG_CHECK(ctx, cond)                                                                       \
    do {                                                                                           \
        if (!(cond)) return set_err((ctx), B2GP_ERR_ARG, "bad argument", #cond, __FILE__, __LINE__); \
    } while (0)

#define RET_IF(expr)                \
    do {                            \
        int r_ = (expr);            \
        if (r_ != B2GP_OK) return r_; \
    } while (0)

static inline int ensure(b2gp_ctx* ctx, DevBuf& b, size_t bytes) {
    if (b.cap >= bytes && b.p) return B2GP_OK;
    if (b.p) {
        CUDA_TRY(ctx, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    // round up so that slowly growing requests do not re-allocate every call
    size_t want = (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1);
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
        b.p = nullptr;
        return set_err(ctx, B2GP_ERR_NOMEM, "cudaMalloc", cudaGetErrorString(e), __FILE__, __LINE__);
    }
    b.cap = want;
    return B2GP_OK;
}

// cudaFuncSetAttribute is per device: a process may hold contexts on several devices (Context(device=1) next to
// the default one), so the "already opted in to large dynamic shared memory" memo is a bit per device.
struct PerDeviceOnce {
    std::atomic<uint64_t> mask{0};
    bool need(int dev) const { return ((mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0; }
    void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

static inline Slot* slot_of(b2gp_ctx* ctx, cudaStream_t st) {
    for (int i = 0; i < B2GP_MAX_STREAMS; ++i)
        if (ctx->slots[i].stream == st) return &ctx->slots[i];
    return nullptr;
}

// digit planes for the int8 GEMMs queued on stream `st`
static inline int oz_planes_for(b2gp_ctx* ctx, cudaStream_t st) {
    if (ctx->ozaki > 0) return ctx->ozaki;
    Slot* sl = slot_of(ctx, st);
    return sl ? sl->oz_planes : 7;
}

// Accuracy-aware plane count: the error the digit-plane GEMMs add to a posterior grows like
// cond(K) * 1.3e-16 with 46-bit operands (6 planes) and cond(K) * 3e-18 with 54-bit ones (7 planes); against the 1e-9
// parity bar 6 planes are safe while cond(K) <= 1e6.  cond(K) is bounded from the trace: lambda_max <= N k_scale +
// sigma^2 + jitter, lambda_min >= sigma^2 + jitter  (K = k(X, X) + (sigma^2 + jitter) I, k(x, x) = k_scale).
static inline int oz_auto_planes(double n, double k_scale, double noise, double jitter) {
    const double floor_ = noise + jitter;
    if (!(floor_ > 0.0) || !(k_scale > 0.0)) return 7;
    return (n * k_scale + floor_) / floor_ <= 1e6 ? 6 : 7;
}

// CTAs a persistent (one CTA per SM) kernel should launch
static inline int persist_sms(b2gp_ctx* ctx) { return (ctx->big_grid > 0 && ctx->big_grid < ctx->sm_count) ? ctx->big_grid : ctx->sm_count; }

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }
