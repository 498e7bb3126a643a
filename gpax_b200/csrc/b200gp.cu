// b200gp.cu -- C-ABI entry points of libb200gp.so (see include/b200gp.h for the contract and the
// reference lines each entry point replaces).  Host code here only orchestrates: workspace, streams,
// the order of kernel launches.  No torch, no JAX, no CPU fallback: every numerical result is
// produced by the sm_100a kernels in gram.cuh / gemm_dmma.cuh / potrf.cuh / posterior.cuh.
#include <chrono>
#include <thread>

#include "common.cuh"
#include "gemm_dmma.cuh"
#include "gemm_tma.cuh"
#include "ozaki.cuh"
#include "gram.cuh"
#include "mll.cuh"
#include "posterior.cuh"
#include "potrf.cuh"
#include "sparse_elbo.cuh"
#include "acq.cuh"
#include "dist.cuh"

// ------------------------------------------------------------------------------------------ helpers
namespace {

struct EventPool {
    std::vector<cudaEvent_t> ev;
    size_t next = 0;
    cudaEvent_t get() {
        if (next == ev.size()) {
            cudaEvent_t e;
            if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
            ev.push_back(e);
        }
        return ev[next++];
    }
    void reset() { next = 0; }
    void destroy() {
        for (auto e : ev) cudaEventDestroy(e);
        ev.clear();
        next = 0;
    }
};

struct Extra {  // ctx-private state that is not part of the struct the kernels' headers see
    EventPool pool;
    b2gp_timing last{};
    cudaEvent_t slot_done[B2GP_MAX_STREAMS] = {};
    cudaEvent_t inputs_ready = nullptr;
    DevBuf theta1;     // one-draw theta for b2gp_gram
    DevBuf potrf_buf;  // staging for host-pointer b2gp_potrf / trsm / gemm
    DevBuf gemm_buf[3];
    std::vector<void*> user_allocs;
    DevBuf eb[12];     // scratch of b2gp_sparse_elbo
    DevBuf f32_in[8];  // fp32 staging of the inputs / outputs of calls made with B2GP_FLAG_F32
    DevBuf f32_out[4];
    // factor cache of slot 0 (host-pointer, single-draw calls): predict_in_batches / viGP chunk loops call the
    // posterior repeatedly with the same training set and theta; the reference re-inverts k_XX every time
    // (gp.py:319-322 -> gp.py:269-271), here the factor L and its inverted diagonal blocks are kept.
    struct {
        bool valid = false;
        int kind = -1, d = 0;
        int64_t N = 0;
        double jitter = 0.0;
        std::vector<double> theta;
        std::vector<char> X;   // raw bytes of the caller's training inputs (fp64 or fp32)
        int info = 0;
        int64_t U_nb = 0;      // > 0: `Ukeep` holds the explicit inverses of the factor's U_nb-wide diagonal blocks (potrf_tall)
    } fcache;
    DevBuf Ukeep;
    int64_t cache_hits = 0;
    DistState* dist = nullptr;   // multi-GPU state (dist.cuh), created by b2gp_dist_init
};

}  // namespace

static std::vector<std::pair<b2gp_ctx*, Extra*>> g_extras;
static Extra* extra_of(b2gp_ctx* ctx) {
    for (auto& p : g_extras)
        if (p.first == ctx) return p.second;
    return nullptr;
}

static inline bool dev_ptrs(unsigned flags) { return (flags & B2GP_FLAG_DEVICE_PTRS) != 0; }

static int free_buf(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    return 0;
}

struct CallTimer {
    b2gp_ctx* ctx;
    Extra* ex;
    int64_t launches0;
    CallTimer(b2gp_ctx* c) : ctx(c), ex(extra_of(c)), launches0(c->launches) {}
    std::chrono::steady_clock::time_point t_begin;
    int begin(cudaStream_t st) {
        t_begin = std::chrono::steady_clock::now();
        ex->last = b2gp_timing{};
        ex->pool.reset();
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_begin, st));
        return B2GP_OK;
    }
    // host time spent queueing work so far (call before any copy into pageable memory, which blocks the host)
    void mark_enqueued() {
        ex->last.host_enqueue_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    }
    int end(cudaStream_t st, b2gp_timing* out) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_end, st));
        if (ex->last.host_enqueue_ms == 0.0) mark_enqueued();
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_end));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_begin, ctx->ev_end));
        ex->last.total_ms = ms;
        ex->last.launches = ctx->launches - launches0;
        if (out) *out = ex->last;
        return B2GP_OK;
    }
};


// ------------------------------------------------------------------------------------------ lifecycle
extern "C" int b2gp_version(void) { return B2GP_VERSION; }

extern "C" int b2gp_ctx_create(int device, b2gp_ctx** out) {
    if (!out) return B2GP_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return B2GP_ERR_CUDA;
    if (device < 0 || device >= ndev) return B2GP_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return B2GP_ERR_CUDA;
    b2gp_ctx* ctx = new b2gp_ctx();
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return B2GP_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->mem_bytes = prop.totalGlobalMem;
    for (int i = 0; i < B2GP_MAX_STREAMS; ++i) {
        if (cudaStreamCreateWithFlags(&ctx->slots[i].stream, cudaStreamNonBlocking) != cudaSuccess) return B2GP_ERR_CUDA;
        for (int e = 0; e < 8; ++e) cudaEventCreate(&ctx->slots[i].ev[e]);
    }
    cudaEventCreate(&ctx->ev_begin);
    cudaEventCreate(&ctx->ev_end);
    cudaEventCreate(&ctx->ev_a);
    cudaEventCreate(&ctx->ev_b);
    Extra* ex = new Extra();
    for (int i = 0; i < B2GP_MAX_STREAMS; ++i) cudaEventCreateWithFlags(&ex->slot_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&ex->inputs_ready, cudaEventDisableTiming);
    g_extras.emplace_back(ctx, ex);
    *out = ctx;
    return B2GP_OK;
}

extern "C" int b2gp_ctx_destroy(b2gp_ctx* ctx) {
    if (!ctx) return B2GP_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    Extra* ex = extra_of(ctx);
    for (int i = 0; i < B2GP_MAX_STREAMS; ++i) {
        Slot& s = ctx->slots[i];
        free_buf(s.A);
        free_buf(s.Vt);
        free_buf(s.Linv);
        free_buf(s.cov);
        free_buf(s.LinvC);
        free_buf(s.misc);
        free_buf(s.panelU);
        free_buf(s.oz.planesA);
        free_buf(s.oz.planesB);
        free_buf(s.oz.scaleA);
        free_buf(s.oz.scaleB);
        free_buf(s.oz.prof);
        for (auto& l : s.oz.lists) free_buf(l.dev);
        for (int e = 0; e < 8; ++e) cudaEventDestroy(s.ev[e]);
        cudaStreamDestroy(s.stream);
    }
    for (auto& b : ctx->d_in) free_buf(b);
    for (auto& b : ctx->d_out) free_buf(b);
    free_buf(ctx->d_info);
    free_buf(ctx->last_linv);
    cudaEventDestroy(ctx->ev_begin);
    cudaEventDestroy(ctx->ev_end);
    cudaEventDestroy(ctx->ev_a);
    cudaEventDestroy(ctx->ev_b);
    if (ex) {
        ex->pool.destroy();
        for (int i = 0; i < B2GP_MAX_STREAMS; ++i) cudaEventDestroy(ex->slot_done[i]);
        cudaEventDestroy(ex->inputs_ready);
        free_buf(ex->theta1);
        free_buf(ex->Ukeep);
        for (auto& b : ex->eb) free_buf(b);
        for (auto& b : ex->f32_in) free_buf(b);
        for (auto& b : ex->f32_out) free_buf(b);
        free_buf(ex->potrf_buf);
        for (auto& b : ex->gemm_buf) free_buf(b);
        for (void* p : ex->user_allocs) cudaFree(p);
        for (size_t i = 0; i < g_extras.size(); ++i)
            if (g_extras[i].first == ctx) {
                g_extras.erase(g_extras.begin() + i);
                break;
            }
        delete ex;
    }
    delete ctx;
    return B2GP_OK;
}

static std::string g_null_err = "null context";
extern "C" const char* b2gp_last_error(const b2gp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_null_err.c_str(); }

extern "C" int b2gp_set_option(b2gp_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return B2GP_ERR_ARG;
    if (strcmp(key, "streams") == 0) {
        ARG_CHECK(ctx, value >= 1 && value <= B2GP_MAX_STREAMS);
        ctx->n_streams = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "ozaki") == 0) {
        ARG_CHECK(ctx, value == -1 || value == 0 || value == 6 || value == 7);
        ctx->ozaki = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "trsm_strip") == 0) {
        ARG_CHECK(ctx, value == 0 || value == 256 || value == 512 || value == 1024);
        ctx->trsm_strip = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "oz_cluster") == 0) {
        ARG_CHECK(ctx, value == 1 || value == 2);
        ctx->oz_cluster = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "enqueue_threads") == 0) {
        ctx->enqueue_threads = value != 0;
        return B2GP_OK;
    }
    if (strcmp(key, "big_grid") == 0) {
        ctx->big_grid = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "oz_min_tiles") == 0) {
        ctx->oz_min_tiles = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "panel") == 0) {   // diagonal-block width of the tall-panel factorisation; 0 = recursive scheme only
        ARG_CHECK(ctx, value == 0 || value == 128 || value == 256 || value == 512 || value == 1024);
        ctx->panel = (int)value;
        extra_of(ctx)->fcache.valid = false;
        return B2GP_OK;
    }
    if (strcmp(key, "tall_min") == 0) {
        ARG_CHECK(ctx, value >= 256);
        ctx->tall_min = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "oz_debug") == 0) {   // timing experiments only: 1 skips the C read-modify-write, 2 also the staging barriers
        ARG_CHECK(ctx, value >= 0 && value <= 2);
        ctx->oz_debug = (int)value;
        return B2GP_OK;
    }
    if (strcmp(key, "tma") == 0) {
        ctx->use_tma = value ? 1 : 0;
        return B2GP_OK;
    }
    if (strcmp(key, "drop_factor_cache") == 0) {
        extra_of(ctx)->fcache.valid = false;
        return B2GP_OK;
    }
    return set_err(ctx, B2GP_ERR_ARG, "b2gp_set_option", "unknown key", __FILE__, __LINE__);
}

extern "C" int b2gp_device_info(b2gp_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, size_t* mem_bytes) {
    if (!ctx) return B2GP_ERR_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (mem_bytes) *mem_bytes = ctx->mem_bytes;
    return B2GP_OK;
}

extern "C" int64_t b2gp_debug_cache_hits(b2gp_ctx* ctx) { return ctx ? extra_of(ctx)->cache_hits : -1; }

extern "C" int b2gp_last_timing(b2gp_ctx* ctx, b2gp_timing* out) {
    if (!ctx || !out) return B2GP_ERR_ARG;
    *out = extra_of(ctx)->last;
    return B2GP_OK;
}

// ------------------------------------------------------------------------------------------ memory
extern "C" int b2gp_dev_alloc(b2gp_ctx* ctx, size_t bytes, void** dptr) {
    if (!ctx || !dptr) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 16);
    if (e != cudaSuccess) return set_err(ctx, B2GP_ERR_NOMEM, "cudaMalloc", cudaGetErrorString(e), __FILE__, __LINE__);
    extra_of(ctx)->user_allocs.push_back(p);
    *dptr = p;
    return B2GP_OK;
}

extern "C" int b2gp_dev_free(b2gp_ctx* ctx, void* dptr) {
    if (!ctx) return B2GP_ERR_ARG;
    if (!dptr) return B2GP_OK;
    auto& v = extra_of(ctx)->user_allocs;
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i] == dptr) {
            v.erase(v.begin() + i);
            CUDA_TRY(ctx, cudaFree(dptr));
            return B2GP_OK;
        }
    return set_err(ctx, B2GP_ERR_ARG, "b2gp_dev_free", "pointer not owned by this ctx", __FILE__, __LINE__);
}

extern "C" int b2gp_host_alloc(b2gp_ctx* ctx, size_t bytes, void** hptr) {
    if (!ctx || !hptr) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaHostAlloc(hptr, bytes ? bytes : 16, cudaHostAllocDefault));
    return B2GP_OK;
}

extern "C" int b2gp_host_free(b2gp_ctx* ctx, void* hptr) {
    if (!ctx) return B2GP_ERR_ARG;
    if (hptr) CUDA_TRY(ctx, cudaFreeHost(hptr));
    return B2GP_OK;
}

extern "C" int b2gp_h2d(b2gp_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
    return B2GP_OK;
}
extern "C" int b2gp_d2h(b2gp_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (!ctx) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return B2GP_OK;
}
extern "C" int b2gp_sync(b2gp_ctx* ctx) {
    if (!ctx) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaDeviceSynchronize());
    return B2GP_OK;
}

// stage a host array on the device (returns the device pointer) or pass a device pointer through
static int stage_in(b2gp_ctx* ctx, cudaStream_t st, DevBuf& buf, const void* src, size_t bytes, bool is_dev,
                    const double** out) {
    if (!src) {
        *out = nullptr;
        return B2GP_OK;
    }
    if (is_dev) {
        *out = (const double*)src;
        return B2GP_OK;
    }
    RET_IF(ensure(ctx, buf, bytes));
    CUDA_TRY(ctx, cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st));
    *out = (const double*)buf.p;
    return B2GP_OK;
}

// ---- fp32 I/O (B2GP_FLAG_F32): the reference's default precision is float32 (gpax/utils/utils.py:19-21), so callers hand
// over float arrays and expect float results.  Inputs are widened on the device right after the copy, outputs narrowed
// right before it (round to nearest); the Gram builds, the factorisation and the solves stay fp64 in between.
__global__ void cvt_f32_f64_kernel(double* __restrict__ dst, const float* __restrict__ src, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}
__global__ void cvt_f64_f32_kernel(float* __restrict__ dst, int64_t ldd, const double* __restrict__ src, int64_t lds, int64_t rows,
                                   int64_t cols) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols, c = i % cols;
        dst[r * ldd + c] = (float)src[r * lds + c];
    }
}
static inline bool f32_io(unsigned flags) { return (flags & B2GP_FLAG_F32) != 0; }

// stage_in for `count` elements that are doubles, or floats when f32: the result is always a device array of doubles
static int stage_in_t(b2gp_ctx* ctx, cudaStream_t st, DevBuf& buf, DevBuf& tmp, const void* src, size_t count, bool is_dev, bool f32,
                      const double** out) {
    if (!f32) return stage_in(ctx, st, buf, src, count * 8, is_dev, out);
    if (!src) {
        *out = nullptr;
        return B2GP_OK;
    }
    const float* fsrc = (const float*)src;
    if (!is_dev) {
        RET_IF(ensure(ctx, tmp, count * 4));
        CUDA_TRY(ctx, cudaMemcpyAsync(tmp.p, src, count * 4, cudaMemcpyHostToDevice, st));
        fsrc = (const float*)tmp.p;
    }
    RET_IF(ensure(ctx, buf, count * 8));
    cvt_f32_f64_kernel<<<grid_for((int64_t)count), 256, 0, st>>>((double*)buf.p, fsrc, (int64_t)count);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    *out = (const double*)buf.p;
    return B2GP_OK;
}

// device doubles [rows, cols] (leading dimension lds) -> the caller's float array (host or device, leading dimension ldd)
static int store_out_f32(b2gp_ctx* ctx, cudaStream_t st, DevBuf& tmp, void* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows,
                         int64_t cols, bool is_dev) {
    if (rows <= 0 || cols <= 0) return B2GP_OK;
    float* fdst = (float*)dst;
    int64_t ldt = ldd;
    if (!is_dev) {
        RET_IF(ensure(ctx, tmp, (size_t)rows * cols * 4));
        fdst = (float*)tmp.p;
        ldt = cols;
    }
    cvt_f64_f32_kernel<<<grid_for(rows * cols), 256, 0, st>>>(fdst, ldt, src, lds, rows, cols);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if (!is_dev)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, (size_t)ldd * 4, fdst, (size_t)ldt * 4, (size_t)cols * 4, (size_t)rows, cudaMemcpyDeviceToHost, st));
    return B2GP_OK;
}

// ------------------------------------------------------------------------------------------ gram
extern "C" int b2gp_gram(b2gp_ctx* ctx, int kind, const double* X, int64_t n, const double* Z, int64_t m, int d,
                         const double* lengthscale, double scale, double period, double diag_add, int same_xz, double* K,
                         int64_t ldk, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= 0 && kind <= B2GP_KERNEL_NNGP_RELU);
    ARG_CHECK(ctx, X && Z && K && lengthscale);
    ARG_CHECK(ctx, n >= 0 && m >= 0 && d >= 1 && d <= GRAM_MAX_D && ldk >= m);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    // theta of a single "draw": lengthscale (always a host pointer: d values), scale, noise := diag_add, period
    double th[GRAM_MAX_D + 3];
    for (int k = 0; k < d; ++k) th[k] = lengthscale[k];
    th[d] = scale;
    th[d + 1] = diag_add;
    th[d + 2] = period;
    RET_IF(ensure(ctx, ex->theta1, sizeof th));
    CUDA_TRY(ctx, cudaMemcpyAsync(ex->theta1.p, th, (d + 3) * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));  // th is a stack buffer
    const bool dev = dev_ptrs(flags), f32 = f32_io(flags);
    const double *dX, *dZ;
    RET_IF(stage_in_t(ctx, st, ctx->d_in[0], ex->f32_in[0], X, (size_t)n * d, dev, f32, &dX));
    if (Z == X && (!dev || f32))
        dZ = dX;
    else
        RET_IF(stage_in_t(ctx, st, ctx->d_in[1], ex->f32_in[1], Z, (size_t)m * d, dev, f32, &dZ));
    double* dK = K;
    int64_t ld = ldk;
    if (!dev || f32) {
        ld = round_up(m, 2);
        RET_IF(ensure(ctx, ctx->d_out[0], (size_t)n * ld * 8));
        dK = (double*)ctx->d_out[0].p;
    }
    const int lower = (flags & B2GP_FLAG_LOWER_ONLY) && same_xz && n == m;
    if (lower && (!dev || f32)) CUDA_TRY(ctx, cudaMemsetAsync(dK, 0, (size_t)n * ld * 8, st));
    RET_IF(launch_gram(ctx, st, kind, dX, n, dZ, m, d, (const double*)ex->theta1.p, 1.0, 0.0, same_xz ? 1 : 0, lower, dK, ld));
    if (f32)
        RET_IF(store_out_f32(ctx, st, ex->f32_out[0], K, ldk, dK, ld, n, m, dev));
    else if (!dev)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(K, (size_t)ldk * 8, dK, (size_t)ld * 8, (size_t)m * 8, (size_t)n, cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    ex->last.gram_bytes = 8.0 * (double)n * (double)m + 8.0 * (double)(n + m) * d;
    return B2GP_OK;
}

// Multi-task Gram matrices (gpax/kernels/mtkernels.py:19-58 index_kernel, 61-125 MultitaskKernel):
//   K[i, j] = (k_data(x_i, z_j) + jitter [same point, same_xz]) * B[tX_i, tZ_j]  (+ noise_task[tX_i] + jitter on i == j when same_xz)
// -- the reference calls the data kernel with noise 0 but its usual diagonal rule, so k_data carries `jitter` where the two
// points coincide (mtkernels.py:103, 167); `group` consecutive rows are one data point (1, or the task count for the
// Kronecker form, whose jitter therefore lands on the whole T x T diagonal block).
// B = W W^T + diag(v) (T x T, formed by the caller: T^2 numbers).  The data kernel is the fused Gram kernel; the task factor
// is applied in place by one elementwise pass.  MultivariateKernel's Kronecker form (mtkernels.py:128-192) is the same
// thing on inputs repeated once per task with the task index cycling fastest (the shell does that).
__global__ void mt_task_kernel(double* K, int64_t ld, int64_t n, int64_t m, const int* __restrict__ tX, const int* __restrict__ tZ,
                               const double* __restrict__ B, int T, const double* __restrict__ noise_task, double jitter, int same_xz,
                               int group) {
    const int64_t total = n * m;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = idx / m, j = idx % m;
        double kd = K[i * ld + j];
        if (same_xz && i / group == j / group) kd += jitter;
        double v = kd * B[(int64_t)tX[i] * T + tZ[j]];
        if (same_xz && i == j) v += (noise_task ? noise_task[tX[i]] : 0.0) + jitter;
        K[i * ld + j] = v;
    }
}

extern "C" int b2gp_gram_multitask(b2gp_ctx* ctx, int kind, const double* X, const int* taskX, int64_t n, const double* Z,
                                   const int* taskZ, int64_t m, int d, const double* lengthscale, double scale, double period,
                                   const double* B, int T, const double* noise_task, double jitter, int same_xz, int group,
                                   double* K, int64_t ldk, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= 0 && kind <= B2GP_KERNEL_NNGP_RELU);
    ARG_CHECK(ctx, X && Z && taskX && taskZ && B && K && lengthscale);
    ARG_CHECK(ctx, n >= 1 && m >= 1 && T >= 1 && group >= 1 && d >= 1 && d <= GRAM_MAX_D && ldk >= m);
    ARG_CHECK(ctx, !(flags & (B2GP_FLAG_DEVICE_PTRS | B2GP_FLAG_F32)));      // host fp64 arrays (a callable-kernel building block)
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    double th[GRAM_MAX_D + 3];
    for (int k = 0; k < d; ++k) th[k] = lengthscale[k];
    th[d] = scale;
    th[d + 1] = 0.0;
    th[d + 2] = period;
    RET_IF(ensure(ctx, ex->theta1, sizeof th));
    CUDA_TRY(ctx, cudaMemcpyAsync(ex->theta1.p, th, (d + 3) * sizeof(double), cudaMemcpyHostToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    const double *dX, *dZ, *dB, *dn = nullptr, *dtx, *dtz;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], X, (size_t)n * d * 8, false, &dX));
    RET_IF(stage_in(ctx, st, ctx->d_in[1], Z, (size_t)m * d * 8, false, &dZ));
    RET_IF(stage_in(ctx, st, ctx->d_in[2], B, (size_t)T * T * 8, false, &dB));
    if (noise_task) RET_IF(stage_in(ctx, st, ctx->d_in[4], noise_task, (size_t)T * 8, false, &dn));
    RET_IF(stage_in(ctx, st, ctx->d_in[5], taskX, (size_t)n * 4, false, &dtx));
    RET_IF(stage_in(ctx, st, ctx->d_in[6], taskZ, (size_t)m * 4, false, &dtz));
    const int64_t ld = round_up(m, 2);
    RET_IF(ensure(ctx, ctx->d_out[0], (size_t)n * ld * 8));
    double* dK = (double*)ctx->d_out[0].p;
    RET_IF(launch_gram(ctx, st, kind, dX, n, dZ, m, d, (const double*)ex->theta1.p, 0.0, 0.0, 0, 0, dK, ld));
    mt_task_kernel<<<grid_for(n * m), 256, 0, st>>>(dK, ld, n, m, (const int*)dtx, (const int*)dtz, dB, T, dn, jitter, same_xz ? 1 : 0, group);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CUDA_TRY(ctx, cudaMemcpy2DAsync(K, (size_t)ldk * 8, dK, (size_t)ld * 8, (size_t)m * 8, (size_t)n, cudaMemcpyDeviceToHost, st));
    return tm.end(st, nullptr);
}

// ------------------------------------------------------------------------------------------ potrf / trsm / gemm
extern "C" int b2gp_potrf(b2gp_ctx* ctx, int64_t n, double* A, int64_t lda, int* info, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, A && info && n >= 0 && lda >= n);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    *info = 0;
    if (n == 0) return tm.end(st, nullptr);
    const bool dev = dev_ptrs(flags);
    double* dA = A;
    int64_t ld = lda;
    if (!dev) {
        ld = round_up(n, 2);
        RET_IF(ensure(ctx, ex->potrf_buf, (size_t)n * ld * 8));
        dA = (double*)ex->potrf_buf.p;
        CUDA_TRY(ctx, cudaMemcpy2DAsync(dA, (size_t)ld * 8, A, (size_t)lda * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, st));
    }
    RET_IF(ensure(ctx, ctx->last_linv, (size_t)linv_bytes(n)));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_info.p, 0, 8, st));
    RET_IF(potrf_auto(ctx, st, dA, ld, n, 0, (double*)ctx->last_linv.p, (int*)ctx->d_info.p));
    ctx->last_n = n;
    if (!dev) {
        // the strict upper triangle of the caller's array is documented as untouched: stage the factor on the
        // host and write back j <= i only
        std::vector<double> tmp((size_t)n * n);
        CUDA_TRY(ctx, cudaMemcpy2DAsync(tmp.data(), (size_t)n * 8, dA, (size_t)ld * 8, (size_t)n * 8, (size_t)n, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        for (int64_t i = 0; i < n; ++i) memcpy(A + i * lda, tmp.data() + i * n, (size_t)(i + 1) * 8);
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(info, ctx->d_info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    ex->last.flops = (double)n * (double)n * (double)n / 3.0;
    ex->last.potrf_ms = ex->last.total_ms;
    return B2GP_OK;
}

extern "C" int b2gp_trsm_lower(b2gp_ctx* ctx, int64_t n, int64_t nrhs, const double* L, int64_t ldl, double* B, int64_t ldb,
                               unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, L && B && n >= 0 && nrhs >= 0 && ldl >= n && ldb >= n);
    if (ctx->last_n != n || !ctx->last_linv.p)
        return set_err(ctx, B2GP_ERR_ARG, "b2gp_trsm_lower", "call b2gp_potrf on this factor first (same ctx, same n)", __FILE__, __LINE__);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    if (n == 0 || nrhs == 0) return tm.end(st, nullptr);
    const bool dev = dev_ptrs(flags);
    const double* dL = L;
    double* dB = B;
    int64_t ll = ldl, lb = ldb;
    if (!dev) {
        ll = round_up(n, 2);
        lb = ll;
        RET_IF(ensure(ctx, ex->gemm_buf[0], (size_t)n * ll * 8));
        RET_IF(ensure(ctx, ex->gemm_buf[1], (size_t)nrhs * lb * 8));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(ex->gemm_buf[0].p, (size_t)ll * 8, L, (size_t)ldl * 8, (size_t)n * 8, (size_t)n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(ctx, cudaMemcpy2DAsync(ex->gemm_buf[1].p, (size_t)lb * 8, B, (size_t)ldb * 8, (size_t)n * 8, (size_t)nrhs, cudaMemcpyHostToDevice, st));
        dL = (const double*)ex->gemm_buf[0].p;
        dB = (double*)ex->gemm_buf[1].p;
    }
    RET_IF(trsm_rec(ctx, st, dB, lb, nrhs, dL, ll, n, (const double*)ctx->last_linv.p));
    if (!dev)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(B, (size_t)ldb * 8, dB, (size_t)lb * 8, (size_t)n * 8, (size_t)nrhs, cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    ex->last.flops = (double)n * (double)n * (double)nrhs;
    ex->last.trsm_ms = ex->last.total_ms;
    return B2GP_OK;
}

extern "C" int b2gp_gemm_nt(b2gp_ctx* ctx, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                            const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower_only, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, A && B && C && m >= 0 && n >= 0 && k >= 0 && lda >= k && ldb >= k && ldc >= n);
    ARG_CHECK(ctx, !lower_only || m == n);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const bool dev = dev_ptrs(flags);
    const double *dA = A, *dB = B;
    double* dC = C;
    int64_t la = lda, lb = ldb, lc = ldc;
    if (!dev) {
        la = lb = round_up(k > 0 ? k : 1, 2);
        lc = round_up(n > 0 ? n : 1, 2);
        RET_IF(ensure(ctx, ex->gemm_buf[0], (size_t)(m + 1) * la * 8));
        RET_IF(ensure(ctx, ex->gemm_buf[1], (size_t)(n + 1) * lb * 8));
        RET_IF(ensure(ctx, ex->gemm_buf[2], (size_t)(m + 1) * lc * 8));
        if (m && k) CUDA_TRY(ctx, cudaMemcpy2DAsync(ex->gemm_buf[0].p, (size_t)la * 8, A, (size_t)lda * 8, (size_t)k * 8, (size_t)m, cudaMemcpyHostToDevice, st));
        if (n && k) CUDA_TRY(ctx, cudaMemcpy2DAsync(ex->gemm_buf[1].p, (size_t)lb * 8, B, (size_t)ldb * 8, (size_t)k * 8, (size_t)n, cudaMemcpyHostToDevice, st));
        if (m && n) CUDA_TRY(ctx, cudaMemcpy2DAsync(ex->gemm_buf[2].p, (size_t)lc * 8, C, (size_t)ldc * 8, (size_t)n * 8, (size_t)m, cudaMemcpyHostToDevice, st));
        dA = (const double*)ex->gemm_buf[0].p;
        dB = (A == B && lda == ldb && m == n) ? dA : (const double*)ex->gemm_buf[1].p;
        dC = (double*)ex->gemm_buf[2].p;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, st));
    RET_IF(gemm_nt(ctx, st, m, n, k, alpha, dA, la, dB, lb, beta, dC, lc, lower_only != 0));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, st));
    if (!dev && m && n)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(C, (size_t)ldc * 8, dC, (size_t)lc * 8, (size_t)n * 8, (size_t)m, cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    ex->last.epilogue_ms = ms;  // kernel-only time of the GEMM launch
    ex->last.flops = (lower_only ? 1.0 : 2.0) * (double)m * (double)n * (double)k;
    return B2GP_OK;
}

// ------------------------------------------------------------------------------------------ posterior
namespace {
struct StageEvents {
    cudaEvent_t e[6];
};
}

// diag(A) += v   (per-point noise variances: mngp.py:96, hskgp.py:147)
__global__ void add_diag_vec_kernel(double* A, int64_t ld, int64_t n, const double* __restrict__ v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[i * ld + i] += v[i];
}

// The posterior with everything that may vary per draw: the training inputs (xtr_stride doubles between draws; 0 =
// shared), the test inputs (xnew_stride), the targets (yres_stride) and an optional vector of per-point noise variances
// added to the diagonal of k_XX (nv_stride between draws; 0 = shared).  b2gp_posterior is the all-shared special case.
static int posterior_impl(b2gp_ctx* ctx, int kind, const double* Xtr, int64_t xtr_stride, int64_t N, const double* yres,
                          int64_t yres_stride, const double* Xnew, int64_t xnew_stride, int64_t P, int d, int64_t S,
                          const double* theta, const double* noise_vec, int64_t nv_stride, int noiseless, double jitter,
                          unsigned flags, double* mean, double* var, double* cov, const double* eps, int64_t n_samp,
                          double* y_sampled, int* info, b2gp_timing* timing) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, xtr_stride == 0 || xtr_stride >= N * d);
    ARG_CHECK(ctx, xnew_stride == 0 || xnew_stride >= P * d);
    ARG_CHECK(ctx, nv_stride == 0 || nv_stride >= N);
    ARG_CHECK(ctx, Xtr && yres && Xnew && theta && info);
    ARG_CHECK(ctx, N >= 1 && P >= 1 && S >= 1 && d >= 1 && d <= GRAM_MAX_D);
    ARG_CHECK(ctx, yres_stride == 0 || yres_stride >= N);
    const bool want_mean = flags & B2GP_OUT_MEAN, want_var = flags & B2GP_OUT_VAR;
    const bool want_cov = flags & B2GP_OUT_COV, want_samp = flags & B2GP_OUT_SAMPLE;
    ARG_CHECK(ctx, !want_mean || mean);
    ARG_CHECK(ctx, !want_var || var);
    ARG_CHECK(ctx, !want_cov || cov);
    ARG_CHECK(ctx, !want_samp || (eps && y_sampled && n_samp >= 1));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    const bool dev = dev_ptrs(flags), f32 = f32_io(flags);
    const int nslots = (int)(S < ctx->n_streams ? S : ctx->n_streams);
    cudaStream_t st0 = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st0));
    auto slot_stream = [&](int q) { return ctx->slots[q].stream; };

    // ---- inputs
    const int nth = d + 3;
    const double *dXtr, *dy, *dXnew, *dtheta, *deps = nullptr, *dnv = nullptr;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, st0));
    // with B2GP_FLAG_F32 the data arrays (X, y, X_new, noise_vec, eps) are floats; theta stays double
    RET_IF(stage_in_t(ctx, st0, ctx->d_in[0], ex->f32_in[0], Xtr, (size_t)(xtr_stride ? S * xtr_stride : N * d), dev, f32, &dXtr));
    RET_IF(stage_in_t(ctx, st0, ctx->d_in[1], ex->f32_in[1], yres, (size_t)(yres_stride ? S * yres_stride : N), dev, f32, &dy));
    RET_IF(stage_in_t(ctx, st0, ctx->d_in[2], ex->f32_in[2], Xnew, (size_t)(xnew_stride ? S * xnew_stride : P * d), dev, f32, &dXnew));
    if (noise_vec) RET_IF(stage_in_t(ctx, st0, ctx->d_in[6], ex->f32_in[6], noise_vec, (size_t)(nv_stride ? S * nv_stride : N), dev, f32, &dnv));
    RET_IF(stage_in(ctx, st0, ctx->d_in[3], theta, (size_t)S * nth * 8, dev, &dtheta));
    if (want_samp) RET_IF(stage_in_t(ctx, st0, ctx->d_in[4], ex->f32_in[4], eps, (size_t)S * n_samp * P, dev, f32, &deps));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, st0));
    CUDA_TRY(ctx, cudaEventRecord(ex->inputs_ready, st0));

    // ---- outputs
    double *dmean = mean, *dvar = var, *dcov = cov, *dsamp = y_sampled;
    if (!dev || f32) {
        if (want_mean) {
            RET_IF(ensure(ctx, ctx->d_out[0], (size_t)S * P * 8));
            dmean = (double*)ctx->d_out[0].p;
        }
        if (want_var) {
            RET_IF(ensure(ctx, ctx->d_out[1], (size_t)S * P * 8));
            dvar = (double*)ctx->d_out[1].p;
        }
        if (want_cov) {
            RET_IF(ensure(ctx, ctx->d_out[2], (size_t)S * P * P * 8));
            dcov = (double*)ctx->d_out[2].p;
        }
        if (want_samp) {
            RET_IF(ensure(ctx, ctx->d_out[3], (size_t)S * n_samp * P * 8));
            dsamp = (double*)ctx->d_out[3].p;
        }
    }
    RET_IF(ensure(ctx, ctx->d_info, (size_t)2 * S * sizeof(int)));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, (size_t)2 * S * sizeof(int), st0));

    // ---- per-slot workspaces
    // the right-hand-side rows [k_pX; y^T] live directly under k_XX in the slot's matrix (potrf_tall solves them with the
    // factorisation's own panel GEMMs)
    const int64_t ldA = round_up(N, 8), ldV = ldA, ldC = round_up(P, 8);
    const bool need_cov = want_cov || want_samp;
    for (int q = 0; q < nslots; ++q) {
        Slot& sl = ctx->slots[q];
        const size_t needA = (size_t)(N + P + 1) * ldA * 8;
        if (q == 0 && ex->fcache.valid && ex->fcache.N == N && sl.A.p && sl.A.cap < needA) {
            // slot 0's matrix holds the cached factor and this call brings more test points than the one that made it:
            // grow the buffer AROUND the factor (a plain ensure() would free it and the reuse below would read garbage)
            DevBuf grown;
            RET_IF(ensure(ctx, grown, needA));
            CUDA_TRY(ctx, cudaMemcpyAsync(grown.p, sl.A.p, (size_t)N * ldA * 8, cudaMemcpyDeviceToDevice, st0));
            CUDA_TRY(ctx, cudaStreamSynchronize(st0));
            free_buf(sl.A);
            sl.A = grown;
        }
        RET_IF(ensure(ctx, sl.A, needA));
        RET_IF(ensure(ctx, sl.Linv, (size_t)linv_bytes(N)));
        if (need_cov) RET_IF(ensure(ctx, sl.cov, (size_t)P * ldC * 8));
        if (want_samp) RET_IF(ensure(ctx, sl.LinvC, (size_t)linv_bytes(P)));
        if (!want_mean) RET_IF(ensure(ctx, sl.misc, (size_t)P * 8));
    }
    // inputs and the memset of dinfo were queued on st0: order the other streams behind them
    CUDA_TRY(ctx, cudaEventRecord(ex->inputs_ready, st0));
    for (int q = 0; q < nslots; ++q)
        if (slot_stream(q) != st0) CUDA_TRY(ctx, cudaStreamWaitEvent(slot_stream(q), ex->inputs_ready, 0));

    // host copy of theta: the accuracy-aware digit-plane count of the int8 path is chosen per draw (oz_auto_planes)
    std::vector<double> htheta;
    if (ctx->ozaki == -1) {
        htheta.resize((size_t)S * nth);
        if (dev) {
            CUDA_TRY(ctx, cudaMemcpyAsync(htheta.data(), dtheta, (size_t)S * nth * 8, cudaMemcpyDeviceToHost, st0));
            CUDA_TRY(ctx, cudaStreamSynchronize(st0));
        } else {
            memcpy(htheta.data(), theta, (size_t)S * nth * 8);
        }
    }
    const double noise_mult_new = noiseless ? 0.0 : 1.0;  // gp.py:260-261
    std::vector<StageEvents> sev;
    if (timing) sev.resize((size_t)S);

    // factor reuse (see Extra::fcache): same kind / N / d / jitter / theta / training inputs as the previous
    // single-draw host-pointer call -> skip the Gram build and the factorisation
    // The cache is invalid for the whole duration of the call: any early return (allocation failure, launch error)
    // leaves it so, and it is re-validated -- together with the factor's `info` -- only after the call's work has
    // completed on the device (end of this function).
    bool reuse = false;
    const bool cacheable = (S == 1 && !dev && !noise_vec);
    if (cacheable) {
        auto& fc = ex->fcache;
        const size_t xbytes = (size_t)N * d * (f32 ? 4 : 8);
        reuse = fc.valid && fc.kind == kind && fc.N == N && fc.d == d && fc.jitter == jitter && fc.X.size() == xbytes &&
                memcmp(fc.theta.data(), theta, (size_t)nth * 8) == 0 && memcmp(fc.X.data(), Xtr, xbytes) == 0;
        if (reuse) ex->cache_hits++;
    }
    ex->fcache.valid = false;
    // a cacheable call that factors by the tall-panel scheme also keeps the diagonal blocks' explicit inverses
    const bool keepU = cacheable && !reuse && use_tall(ctx, N);
    if (keepU) RET_IF(ensure(ctx, ex->Ukeep, (size_t)ceil_div(N, (int64_t)ctx->panel) * ctx->panel * ctx->panel * 8));

    // One draw's whole pipeline, queued on its slot's stream.  Returns a B2GP_* code.
    auto enqueue_draw = [&](int64_t s) -> int {
        Slot& sl = ctx->slots[s % nslots];
        cudaStream_t st = slot_stream((int)(s % nslots));
        double* A = (double*)sl.A.p;
        double* Vt = A + N * ldA;
        double* Linv = (double*)sl.Linv.p;
        const double* th = dtheta + s * nth;
        const double* dXtr_s = dXtr + s * xtr_stride;
        const double* dXnew_s = dXnew + s * xnew_stride;
        const bool fused_solve = !reuse && use_tall(ctx, N);   // the P-side solve rides along with the factorisation
        int* inf = dinfo + s;
        int* inf2 = dinfo + S + s;
        if (timing) {
            for (int e = 0; e < 6; ++e) sev[s].e[e] = ex->pool.get();
            CUDA_TRY(ctx, cudaEventRecord(sev[s].e[0], st));
        }
        // factorisation and P-side solve: 6 or 7 digit planes from the trace bound on cond(K); covariance / sampling: 7
        sl.oz_planes = (htheta.empty() || noise_vec) ? 7 : oz_auto_planes((double)N, htheta[s * nth + d], htheta[s * nth + d + 1], jitter);
        auto rhs_rows = [&]() -> int {
            // k_pX = kernel(X_new, X_train, params, jitter=0.0)  (gp.py:268); same-shape inputs add 0 there
            RET_IF(launch_gram(ctx, st, kind, dXnew_s, P, dXtr_s, N, d, th, 0.0, 0.0, 0, 0, Vt, ldV));
            CUDA_TRY(ctx, cudaMemcpyAsync(Vt + P * ldV, dy + (yres_stride ? s * yres_stride : 0), (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
            return B2GP_OK;
        };
        if (!reuse) {
            // k_XX = kernel(X_train, X_train, params, noise, jitter)  (gp.py:269) -- lower triangle only
            RET_IF(launch_gram(ctx, st, kind, dXtr_s, N, dXtr_s, N, d, th, 1.0, jitter, 1, 1, A, ldA));
            if (dnv) {
                add_diag_vec_kernel<<<grid_for(N), 256, 0, st>>>(A, ldA, N, dnv + s * nv_stride);
                CUDA_TRY(ctx, cudaGetLastError());
                ctx->launches++;
            }
            if (fused_solve) RET_IF(rhs_rows());
            if (timing) CUDA_TRY(ctx, cudaEventRecord(sev[s].e[1], st));
            // factor instead of jnp.linalg.inv (gp.py:271); with the tall-panel scheme also [V^T; w^T] = [k_pX; y^T] L^{-T}
            if (fused_solve)
                RET_IF(potrf_tall(ctx, st, sl, A, ldA, N, P + 1, Linv, inf, 0, keepU ? (double*)ex->Ukeep.p : nullptr));
            else
                RET_IF(potrf_rec(ctx, st, A, ldA, N, Linv, inf, 0));
        } else {
            if (timing) CUDA_TRY(ctx, cudaEventRecord(sev[s].e[1], st));
            CUDA_TRY(ctx, cudaMemcpyAsync(inf, &ex->fcache.info, sizeof(int), cudaMemcpyHostToDevice, st));
        }
        if (timing) CUDA_TRY(ctx, cudaEventRecord(sev[s].e[2], st));
        if (!fused_solve) RET_IF(rhs_rows());
        if (timing) CUDA_TRY(ctx, cudaEventRecord(sev[s].e[3], st));
        // [V^T; w^T] = [k_pX; y^T] L^{-T}
        if (!fused_solve) {
            if (reuse && ex->fcache.U_nb > 0 && ex->fcache.U_nb == ctx->panel && ctx->ozaki != 0)
                RET_IF(trsm_tall(ctx, st, Vt, ldV, P + 1, A, ldA, N, (const double*)ex->Ukeep.p, ex->fcache.U_nb));
            else
                RET_IF(trsm_rec(ctx, st, Vt, ldV, P + 1, A, ldA, N, Linv));
        }
        if (timing) CUDA_TRY(ctx, cudaEventRecord(sev[s].e[4], st));
        // mean / var
        double* mean_s = want_mean ? dmean + s * P : (double*)sl.misc.p;
        if (want_mean || want_var || want_samp) {
            rowdot_kernel<<<(unsigned)P, RD_THREADS, 0, st>>>(Vt, ldV, N, P, kind, d, th, noise_mult_new, jitter, inf, mean_s,
                                                           want_var ? dvar + s * P : nullptr);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
        }
        if (need_cov) {
            sl.oz_planes = 7;
            // cov = k_pp - V^T V  (gp.py:267, 272), lower tiles then mirrored -> exactly symmetric
            double* C = want_cov ? dcov + s * P * P : (double*)sl.cov.p;
            const int64_t ldc = want_cov ? P : ldC;
            RET_IF(launch_gram(ctx, st, kind, dXnew_s, P, dXnew_s, P, d, th, noise_mult_new, jitter, 1, 1, C, ldc));
            RET_IF(gemm_nt(ctx, st, P, P, N, -1.0, Vt, ldV, Vt, ldV, 1.0, C, ldc, true));
            dim3 g2((unsigned)ceil_div(P, 32), (unsigned)ceil_div(P, 32)), b2(32, 32);
            mirror_lower_kernel<<<g2, b2, 0, st>>>(C, ldc, P);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
            if (want_samp) {
                // y = mean + chol(cov) eps  (gp.py:292)
                double* CL = (double*)sl.cov.p;
                if (want_cov) {
                    copy2d_kernel<<<grid_for(P * P), 256, 0, st>>>(CL, ldC, C, ldc, P, P);
                    CUDA_TRY(ctx, cudaGetLastError());
                    ctx->launches++;
                }
                RET_IF(potrf_rec(ctx, st, CL, ldC, P, (double*)sl.LinvC.p, inf2, 0));
                zero_upper_kernel<<<g2, b2, 0, st>>>(CL, ldC, P);
                double* Y = dsamp + s * n_samp * P;
                bcast_rows_kernel<<<grid_for(n_samp * P), 256, 0, st>>>(Y, P, n_samp, P, mean_s);
                CUDA_TRY(ctx, cudaGetLastError());
                ctx->launches += 2;
                RET_IF(gemm_nt(ctx, st, n_samp, P, P, 1.0, deps + s * n_samp * P, P, CL, ldC, 1.0, Y, P, false));
                nan_if_bad_kernel<<<grid_for(n_samp * P), 256, 0, st>>>(Y, P, n_samp, P, inf, inf2);
                CUDA_TRY(ctx, cudaGetLastError());
                ctx->launches++;
            }
            if (want_cov) {
                nan_if_bad_kernel<<<grid_for(P * P), 256, 0, st>>>(C, ldc, P, P, inf, nullptr);
                CUDA_TRY(ctx, cudaGetLastError());
                ctx->launches++;
            }
        }
        if (timing) CUDA_TRY(ctx, cudaEventRecord(sev[s].e[5], st));
        return B2GP_OK;
    };
    // A draw is ~1.4k launches at N=16384 and the driver lets the host run only ~1k launches ahead of the device, so a
    // single queueing thread feeds the slots one after the other and their streams barely overlap.  One host thread
    // per slot keeps every stream's queue full (the slots share nothing but read-only inputs).
    if (ctx->enqueue_threads && nslots > 1 && S > nslots && !timing) {
        std::vector<int> rcs((size_t)nslots, B2GP_OK);
        std::vector<std::thread> workers;
        for (int q = 0; q < nslots; ++q)
            workers.emplace_back([&, q] {
                if (cudaSetDevice(ctx->device) != cudaSuccess) {
                    rcs[q] = B2GP_ERR_CUDA;
                    return;
                }
                for (int64_t s = q; s < S && rcs[q] == B2GP_OK; s += nslots) rcs[q] = enqueue_draw(s);
            });
        for (auto& w : workers) w.join();
        for (int q = 0; q < nslots; ++q) RET_IF(rcs[q]);
    } else {
        for (int64_t s = 0; s < S; ++s) RET_IF(enqueue_draw(s));
    }
    tm.mark_enqueued();
    // ---- join the slots on stream 0
    for (int q = 0; q < nslots; ++q) {
        if (slot_stream(q) == st0) continue;
        CUDA_TRY(ctx, cudaEventRecord(ex->slot_done[q], slot_stream(q)));
        CUDA_TRY(ctx, cudaStreamWaitEvent(st0, ex->slot_done[q], 0));
    }
    cudaEvent_t ev_c = ctx->slots[0].ev[0], ev_d = ctx->slots[0].ev[1];
    CUDA_TRY(ctx, cudaEventRecord(ev_c, st0));
    std::vector<int> hinfo((size_t)2 * S);
    CUDA_TRY(ctx, cudaMemcpyAsync(hinfo.data(), dinfo, (size_t)2 * S * sizeof(int), cudaMemcpyDeviceToHost, st0));
    if (f32) {
        if (want_mean) RET_IF(store_out_f32(ctx, st0, ex->f32_out[0], mean, P, dmean, P, S, P, dev));
        if (want_var) RET_IF(store_out_f32(ctx, st0, ex->f32_out[1], var, P, dvar, P, S, P, dev));
        if (want_cov) RET_IF(store_out_f32(ctx, st0, ex->f32_out[2], cov, P, dcov, P, S * P, P, dev));
        if (want_samp) RET_IF(store_out_f32(ctx, st0, ex->f32_out[3], y_sampled, P, dsamp, P, S * n_samp, P, dev));
    } else if (!dev) {
        if (want_mean) CUDA_TRY(ctx, cudaMemcpyAsync(mean, dmean, (size_t)S * P * 8, cudaMemcpyDeviceToHost, st0));
        if (want_var) CUDA_TRY(ctx, cudaMemcpyAsync(var, dvar, (size_t)S * P * 8, cudaMemcpyDeviceToHost, st0));
        if (want_cov) CUDA_TRY(ctx, cudaMemcpyAsync(cov, dcov, (size_t)S * P * P * 8, cudaMemcpyDeviceToHost, st0));
        if (want_samp) CUDA_TRY(ctx, cudaMemcpyAsync(y_sampled, dsamp, (size_t)S * n_samp * P * 8, cudaMemcpyDeviceToHost, st0));
    }
    CUDA_TRY(ctx, cudaEventRecord(ev_d, st0));
    RET_IF(tm.end(st0, nullptr));
    for (int q = 0; q < B2GP_MAX_STREAMS; ++q) ctx->slots[q].oz_planes = 7;   // other entry points: the conservative count
    for (int64_t s = 0; s < S; ++s) info[s] = hinfo[s] != 0 ? hinfo[s] : -hinfo[S + s];
    if (cacheable) {
        auto& fc = ex->fcache;
        if (!reuse) {
            fc.kind = kind;
            fc.N = N;
            fc.d = d;
            fc.jitter = jitter;
            fc.theta.assign(theta, theta + nth);
            fc.X.assign((const char*)Xtr, (const char*)Xtr + (size_t)N * d * (f32 ? 4 : 8));
            fc.info = hinfo[0];
            fc.U_nb = keepU ? ctx->panel : 0;
        }
        fc.valid = true;
    }

    b2gp_timing& t = ex->last;
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    t.h2d_ms = ms;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ev_c, ev_d));
    t.d2h_ms = ms;
    if (timing) {
        for (int64_t s = 0; s < S; ++s) {
            float a = 0, b = 0, c = 0, e = 0, f = 0;
            cudaEventElapsedTime(&a, sev[s].e[0], sev[s].e[1]);
            cudaEventElapsedTime(&b, sev[s].e[1], sev[s].e[2]);
            cudaEventElapsedTime(&c, sev[s].e[2], sev[s].e[3]);
            cudaEventElapsedTime(&e, sev[s].e[3], sev[s].e[4]);
            cudaEventElapsedTime(&f, sev[s].e[4], sev[s].e[5]);
            t.gram_ms += a + c;
            t.potrf_ms += b;
            t.trsm_ms += e;
            t.epilogue_ms += f;
        }
    }
    const double n = (double)N, p = (double)P;
    t.flops = (double)S * (n * n * n / 3.0 + n * n * (p + 1.0) + 4.0 * n * p + (need_cov ? n * p * p : 0.0));
    t.gram_bytes = (double)S * (8.0 * n * n / 2.0 + 8.0 * n * p + (need_cov ? 8.0 * p * p : 0.0));
    if (timing) *timing = t;
    return B2GP_OK;
}

extern "C" int b2gp_posterior(b2gp_ctx* ctx, int kind, const double* Xtr, int64_t N, const double* yres, int64_t yres_stride,
                              const double* Xnew, int64_t P, int d, int64_t S, const double* theta, int noiseless, double jitter,
                              unsigned flags, double* mean, double* var, double* cov, const double* eps, int64_t n_samp,
                              double* y_sampled, int* info, b2gp_timing* timing) {
    return posterior_impl(ctx, kind, Xtr, 0, N, yres, yres_stride, Xnew, 0, P, d, S, theta, nullptr, 0, noiseless, jitter, flags, mean,
                          var, cov, eps, n_samp, y_sampled, info, timing);
}

extern "C" int b2gp_posterior_batch(b2gp_ctx* ctx, int kind, const double* Xtr, int64_t xtr_stride, int64_t N, const double* yres,
                                    int64_t yres_stride, const double* Xnew, int64_t xnew_stride, int64_t P, int d, int64_t S,
                                    const double* theta, const double* noise_vec, int64_t noise_vec_stride, int noiseless,
                                    double jitter, unsigned flags, double* mean, double* var, double* cov, const double* eps,
                                    int64_t n_samp, double* y_sampled, int* info, b2gp_timing* timing) {
    return posterior_impl(ctx, kind, Xtr, xtr_stride, N, yres, yres_stride, Xnew, xnew_stride, P, d, S, theta, noise_vec,
                          noise_vec_stride, noiseless, jitter, flags, mean, var, cov, eps, n_samp, y_sampled, info, timing);
}

// ------------------------------------------------------------------------------------------ sparse posterior
// out[r, c] = in[c, r]
__global__ void transpose_kernel(double* __restrict__ out, int64_t ldo, const double* __restrict__ in, int64_t ldi,
                                 int64_t rows_in, int64_t cols_in) {
    __shared__ double tile[32][33];
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows_in && c < cols_in) ? in[r * ldi + c] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int64_t r = c0 + i, c = r0 + threadIdx.x;  // out coordinates
        if (r < cols_in && c < rows_in) out[r * ldo + c] = tile[threadIdx.x][i];
    }
}

// generic <row, w> and |row|^2 reductions: dot[p] = <R[p,:], w>, nrm[p] = |R[p,:]|^2
__global__ void __launch_bounds__(RD_THREADS)
rowdot2_kernel(const double* __restrict__ R, int64_t ld, int64_t len, const double* __restrict__ w, double wscale,
               double* __restrict__ dot, double* __restrict__ nrm) {
    __shared__ double red1[RD_THREADS / 32], red2[RD_THREADS / 32];
    const double* row = R + (int64_t)blockIdx.x * ld;
    double s1 = 0.0, s2 = 0.0;
    for (int64_t k = threadIdx.x; k < len; k += RD_THREADS) {
        const double v = row[k];
        if (w) s1 = fma(v, w[k], s1);
        s2 = fma(v, v, s2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red1[threadIdx.x >> 5] = s1;
        red2[threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < RD_THREADS / 32; ++i) {
            a += red1[i];
            b += red2[i];
        }
        if (dot) dot[blockIdx.x] = a * wscale;
        if (nrm) nrm[blockIdx.x] = b;
    }
}

// var[p] = kdiag - q[p] + r[p];  NaN when the factorisations failed
__global__ void sparse_var_kernel(double* var, double* mean, const double* q, const double* r, int64_t P, int kind, int d,
                                  const double* theta, double noise_mult, double jitter, const int* info, const int* info2) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const bool bad = (*info != 0) || (*info2 != 0);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (var) {
        const double kd = cov_self(kind, theta[d]) + (theta[d + 1] * noise_mult + jitter);
        var[p] = bad ? nan : (kd - q[p]) + r[p];
    }
    if (mean && bad) mean[p] = nan;
}

__global__ void scale_by_inv_noise_kernel(double* v, int64_t n, const double* theta, int d) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i] / theta[d + 1];
}

// Partial Nystrom statistics of a shard of the training set (all device pointers):
//   Luu = chol(Kuu + jitter I), W = Luu^{-1} K(Xu, Xtr_shard),  Kpart = W W^T / noise (lower),  cpart = W y / noise.
// Summed over shards these are the K (before "+ I") and W D^{-1} y of sparse_gp.py:198-204.
static int sparse_partial_dev(b2gp_ctx* ctx, Slot& sl, int kind, const double* dXu, int64_t M, const double* dXtr, int64_t N,
                              const double* dy, int d, const double* dth, double jitter, double noise_h, double* Luu, int64_t ldM,
                              double* LinvU, double* Kpart, int64_t ldk, double* cpart, int* dinfo) {
    cudaStream_t st = sl.stream;
    const int64_t ldN = round_up(N, 8);
    RET_IF(ensure(ctx, sl.Vt, (size_t)N * ldM * 8));
    RET_IF(ensure(ctx, sl.cov, (size_t)M * ldN * 8));
    double* Wt = (double*)sl.Vt.p;
    double* W = (double*)sl.cov.p;
    // Kuu = kernel(Xu, Xu, params, **kwargs): noise defaults to 0, so the diagonal gets jitter only (sparse_gp.py:193)
    RET_IF(launch_gram(ctx, st, kind, dXu, M, dXu, M, d, dth, 0.0, jitter, 1, 1, Luu, ldM));
    RET_IF(potrf_auto(ctx, st, Luu, ldM, M, 0, LinvU, dinfo));                                  // sparse_gp.py:194
    // W^T = K_fu Luu^{-T}  (W = Luu^{-1} Kuf, sparse_gp.py:195-197), one training point per row
    RET_IF(launch_gram(ctx, st, kind, dXtr, N, dXu, M, d, dth, 0.0, 0.0, 0, 0, Wt, ldM));
    RET_IF(trsm_rec(ctx, st, Wt, ldM, N, Luu, ldM, M, LinvU));   // tall right-hand sides: int8 panel GEMMs (potrf.cuh)
    {
        dim3 g((unsigned)ceil_div(M, 32), (unsigned)ceil_div(N, 32)), b(32, 8);
        transpose_kernel<<<g, b, 0, st>>>(W, ldN, Wt, ldM, N, M);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    // W D^{-1} W^T with D = noise * 1  (sparse_gp.py:198-199).  Accumulated onto a zeroed matrix (beta = 1) so that the
    // product -- M^2 N flops, the bulk of the sparse posterior -- qualifies for the int8 tcgen05 path (k = N is split into
    // launches of <= 16384 by the dispatcher)
    CUDA_TRY(ctx, cudaMemsetAsync(Kpart, 0, (size_t)M * ldk * 8, st));
    RET_IF(gemm_nt(ctx, st, M, M, N, 1.0 / noise_h, W, ldN, W, ldN, 1.0, Kpart, ldk, true));
    // W D^{-1} y  (sparse_gp.py:203-204)
    rowdot2_kernel<<<(unsigned)M, RD_THREADS, 0, st>>>(W, ldN, N, dy, 1.0 / noise_h, cpart, nullptr);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B2GP_OK;
}

// Posterior from the summed statistics: K = Ksum + I, L = chol(K), then sparse_gp.py:206-217.
static int sparse_finish_dev(b2gp_ctx* ctx, Slot& sl, int kind, const double* dXu, int64_t M, const double* Luu, int64_t ldM,
                             const double* LinvU, double* Kmat, int64_t ldk, double* LinvK, const double* cvec,
                             const double* dXnew, int64_t P, int d, const double* dth, int noiseless, double jitter,
                             bool want_var, bool want_cov, double* dmean, double* dvar, double* C, int64_t ldc, int* dinfo) {
    cudaStream_t st = sl.stream;
    RET_IF(ensure(ctx, sl.LinvC, (size_t)2 * (P + 1) * ldM * 8));
    RET_IF(ensure(ctx, sl.misc, (size_t)(2 * P + 16) * 8));
    double* Wst = (double*)sl.LinvC.p;          // P x ldM          Ws^T = K_su Luu^{-T}
    double* R = Wst + (P + 1) * ldM;            // (P+1) x ldM      rows 0..P-1: (L^{-1} Ws)^T; row P: L^{-1} c
    double* qv = (double*)sl.misc.p;            // |Ws^T[p]|^2
    double* rv = qv + P;                        // |R[p]|^2
    add_diag_kernel<<<grid_for(M), 256, 0, st>>>(Kmat, ldk, M, 1.0);                            // sparse_gp.py:200
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    RET_IF(potrf_auto(ctx, st, Kmat, ldk, M, 0, LinvK, dinfo + 1));                             // sparse_gp.py:201
    // Ws^T = K_su Luu^{-T}  (sparse_gp.py:206-207)
    RET_IF(launch_gram(ctx, st, kind, dXnew, P, dXu, M, d, dth, 0.0, 0.0, 0, 0, Wst, ldM));
    RET_IF(trsm_rec(ctx, st, Wst, ldM, P, Luu, ldM, M, LinvU));
    // pack = [c | Ws]; L^{-1} pack  (sparse_gp.py:208-212)
    copy2d_kernel<<<grid_for(P * M), 256, 0, st>>>(R, ldM, Wst, ldM, P, M);
    CUDA_TRY(ctx, cudaMemcpyAsync(R + P * ldM, cvec, (size_t)M * 8, cudaMemcpyDeviceToDevice, st));
    ctx->launches++;
    RET_IF(trsm_rec(ctx, st, R, ldM, P + 1, Kmat, ldk, M, LinvK));
    // mean = (L^{-1} c)^T (L^{-1} Ws)  (sparse_gp.py:213)
    rowdot2_kernel<<<(unsigned)P, RD_THREADS, 0, st>>>(R, ldM, M, R + P * ldM, 1.0, dmean, rv);
    rowdot2_kernel<<<(unsigned)P, RD_THREADS, 0, st>>>(Wst, ldM, M, nullptr, 1.0, nullptr, qv);
    sparse_var_kernel<<<grid_for(P), 256, 0, st>>>(want_var ? dvar : nullptr, dmean, qv, rv, P, kind, d, dth,
                                                   noiseless ? 0.0 : 1.0, jitter, dinfo, dinfo + 1);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 3;
    if (want_cov) {
        // cov = Kss - Ws^T Ws + (L^{-1}Ws)^T (L^{-1}Ws)  (sparse_gp.py:215-217)
        RET_IF(launch_gram(ctx, st, kind, dXnew, P, dXnew, P, d, dth, noiseless ? 0.0 : 1.0, jitter, 1, 1, C, ldc));
        RET_IF(gemm_nt(ctx, st, P, P, M, -1.0, Wst, ldM, Wst, ldM, 1.0, C, ldc, true));
        RET_IF(gemm_nt(ctx, st, P, P, M, 1.0, R, ldM, R, ldM, 1.0, C, ldc, true));
        dim3 g2((unsigned)ceil_div(P, 32), (unsigned)ceil_div(P, 32)), b2(32, 32);
        mirror_lower_kernel<<<g2, b2, 0, st>>>(C, ldc, P);
        nan_if_bad_kernel<<<grid_for(P * P), 256, 0, st>>>(C, ldc, P, P, dinfo, dinfo + 1);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 2;
    }
    return B2GP_OK;
}

extern "C" int b2gp_sparse_posterior(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, const double* Xtr, int64_t N,
                                     const double* yres, const double* Xnew, int64_t P, int d, const double* theta, int noiseless,
                                     double jitter, unsigned flags, double* mean, double* var, double* cov, int* info,
                                     b2gp_timing* timing) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, Xu && Xtr && yres && Xnew && theta && info);
    ARG_CHECK(ctx, M >= 1 && N >= 1 && P >= 1 && d >= 1 && d <= GRAM_MAX_D);
    const bool want_mean = flags & B2GP_OUT_MEAN, want_var = flags & B2GP_OUT_VAR, want_cov = flags & B2GP_OUT_COV;
    ARG_CHECK(ctx, !want_mean || mean);
    ARG_CHECK(ctx, !want_var || var);
    ARG_CHECK(ctx, !want_cov || cov);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    const bool dev = dev_ptrs(flags), f32 = f32_io(flags);
    Slot& sl = ctx->slots[0];
    cudaStream_t st = sl.stream;
    ex->fcache.valid = false;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const int nth = d + 3;
    const double *dXu, *dXtr, *dy, *dXnew, *dth;
    RET_IF(stage_in_t(ctx, st, ctx->d_in[0], ex->f32_in[0], Xtr, (size_t)N * d, dev, f32, &dXtr));
    RET_IF(stage_in_t(ctx, st, ctx->d_in[1], ex->f32_in[1], yres, (size_t)N, dev, f32, &dy));
    RET_IF(stage_in_t(ctx, st, ctx->d_in[2], ex->f32_in[2], Xnew, (size_t)P * d, dev, f32, &dXnew));
    RET_IF(stage_in(ctx, st, ctx->d_in[3], theta, (size_t)nth * 8, dev, &dth));
    RET_IF(stage_in_t(ctx, st, ctx->d_in[5], ex->f32_in[5], Xu, (size_t)M * d, dev, f32, &dXu));
    double noise_h = 0.0;
    if (dev) {
        CUDA_TRY(ctx, cudaMemcpyAsync(&noise_h, dth + d + 1, 8, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
    } else {
        noise_h = theta[d + 1];
    }
    const int64_t ldM = round_up(M, 8), ldC = round_up(P, 8);
    RET_IF(ensure(ctx, sl.A, (size_t)2 * M * ldM * 8));
    RET_IF(ensure(ctx, sl.Linv, (size_t)2 * linv_bytes(M)));
    RET_IF(ensure(ctx, ctx->d_out[0], (size_t)(2 * P + M + 16) * 8));
    if (want_cov && (!dev || f32)) RET_IF(ensure(ctx, ctx->d_out[2], (size_t)P * ldC * 8));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, 16, st));
    double* Luu = (double*)sl.A.p;
    double* Kmat = Luu + M * ldM;
    double* LinvU = (double*)sl.Linv.p;
    double* LinvK = LinvU + linv_bytes(M) / 8;
    double* mv = (double*)ctx->d_out[0].p;
    double* vv = mv + P;
    double* cvec = vv + P;
    RET_IF(sparse_partial_dev(ctx, sl, kind, dXu, M, dXtr, N, dy, d, dth, jitter, noise_h, Luu, ldM, LinvU, Kmat, ldM, cvec, dinfo));
    const bool direct = dev && !f32;     // results written straight into the caller's (device, fp64) arrays
    double* dmean = (want_mean && direct) ? mean : mv;
    double* dvar = (want_var && direct) ? var : vv;
    double* C = direct ? cov : (double*)ctx->d_out[2].p;
    const int64_t ldc = direct ? P : ldC;
    RET_IF(sparse_finish_dev(ctx, sl, kind, dXu, M, Luu, ldM, LinvU, Kmat, ldM, LinvK, cvec, dXnew, P, d, dth, noiseless, jitter,
                             want_var, want_cov, dmean, dvar, C, ldc, dinfo));
    if (want_cov && f32)
        RET_IF(store_out_f32(ctx, st, ex->f32_out[2], cov, P, C, ldc, P, P, dev));
    else if (want_cov && !dev)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(cov, (size_t)P * 8, C, (size_t)ldc * 8, (size_t)P * 8, (size_t)P, cudaMemcpyDeviceToHost, st));
    int hinfo[2] = {0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(hinfo, dinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (f32) {
        if (want_mean) RET_IF(store_out_f32(ctx, st, ex->f32_out[0], mean, P, dmean, P, 1, P, dev));
        if (want_var) RET_IF(store_out_f32(ctx, st, ex->f32_out[1], var, P, dvar, P, 1, P, dev));
    } else if (!dev) {
        if (want_mean) CUDA_TRY(ctx, cudaMemcpyAsync(mean, dmean, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
        if (want_var) CUDA_TRY(ctx, cudaMemcpyAsync(var, dvar, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    }
    RET_IF(tm.end(st, nullptr));
    info[0] = hinfo[0] != 0 ? hinfo[0] : -hinfo[1];
    const double m = (double)M, n = (double)N, p = (double)P;
    ex->last.flops = 2.0 * m * m * m / 3.0 + 2.0 * m * m * n + 2.0 * m * m * (p + 1.0) + (want_cov ? 2.0 * m * p * p : 0.0);
    ex->last.gram_bytes = 8.0 * (m * n + m * m / 2.0 + m * p);
    if (timing) *timing = ex->last;
    return B2GP_OK;
}

// ---- sharded sparse path ("N-sharded sparse GP"): each rank calls _partial on its shard of
// the training set, the M x M matrix and the M-vector are summed across ranks (NCCL all-reduce by the caller),
// every rank calls _finish.  All array pointers are DEVICE pointers; theta is a HOST pointer (d+3 values).
extern "C" int b2gp_sparse_partial(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, const double* Xtr, int64_t N,
                                   const double* yres, int d, const double* theta, double jitter, double* Kpart, int64_t ldk,
                                   double* cpart, int* info) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, Xu && Xtr && yres && theta && Kpart && cpart && info);
    ARG_CHECK(ctx, M >= 1 && N >= 1 && d >= 1 && d <= GRAM_MAX_D && ldk >= M);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    extra_of(ctx)->fcache.valid = false;
    Slot& sl = ctx->slots[0];
    cudaStream_t st = sl.stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const double* dth;
    RET_IF(stage_in(ctx, st, ctx->d_in[3], theta, (size_t)(d + 3) * 8, false, &dth));
    const int64_t ldM = round_up(M, 8);
    RET_IF(ensure(ctx, sl.A, (size_t)2 * M * ldM * 8));
    RET_IF(ensure(ctx, sl.Linv, (size_t)2 * linv_bytes(M)));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, 16, st));
    RET_IF(sparse_partial_dev(ctx, sl, kind, Xu, M, Xtr, N, yres, d, dth, jitter, theta[d + 1], (double*)sl.A.p, ldM,
                              (double*)sl.Linv.p, Kpart, ldk, cpart, dinfo));
    CUDA_TRY(ctx, cudaMemcpyAsync(info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    return tm.end(st, nullptr);
}

extern "C" int b2gp_sparse_finish(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, double* Ksum, int64_t ldk,
                                  const double* csum, const double* Xnew, int64_t P, int d, const double* theta, int noiseless,
                                  double jitter, unsigned flags, double* mean, double* var, double* cov, int* info) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, Xu && Ksum && csum && Xnew && theta && mean && info);
    ARG_CHECK(ctx, M >= 1 && P >= 1 && d >= 1 && d <= GRAM_MAX_D && ldk >= M);
    const bool want_var = flags & B2GP_OUT_VAR, want_cov = flags & B2GP_OUT_COV;
    ARG_CHECK(ctx, !want_var || var);
    ARG_CHECK(ctx, !want_cov || cov);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    extra_of(ctx)->fcache.valid = false;
    Slot& sl = ctx->slots[0];
    cudaStream_t st = sl.stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const double* dth;
    RET_IF(stage_in(ctx, st, ctx->d_in[3], theta, (size_t)(d + 3) * 8, false, &dth));
    const int64_t ldM = round_up(M, 8);
    RET_IF(ensure(ctx, sl.A, (size_t)2 * M * ldM * 8));
    RET_IF(ensure(ctx, sl.Linv, (size_t)2 * linv_bytes(M)));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, 16, st));
    double* Luu = (double*)sl.A.p;
    double* LinvU = (double*)sl.Linv.p;
    double* LinvK = LinvU + linv_bytes(M) / 8;
    // Luu is rebuilt here (M^3/3 flops) so that _finish does not depend on ctx state left by _partial
    RET_IF(launch_gram(ctx, st, kind, Xu, M, Xu, M, d, dth, 0.0, jitter, 1, 1, Luu, ldM));
    RET_IF(potrf_rec(ctx, st, Luu, ldM, M, LinvU, dinfo, 0));
    RET_IF(sparse_finish_dev(ctx, sl, kind, Xu, M, Luu, ldM, LinvU, Ksum, ldk, LinvK, csum, Xnew, P, d, dth, noiseless, jitter,
                             want_var, want_cov, mean, var, cov, P, dinfo));
    int hinfo[2] = {0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(hinfo, dinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    info[0] = hinfo[0] != 0 ? hinfo[0] : -hinfo[1];
    return B2GP_OK;
}

// ---- building blocks of the block-cyclic multi-GPU factorisation (device pointers only)
// Factor an n x n block and export its inverted 128x128 diagonal blocks (ceil(n/128) * 128*128 doubles) to Linv_out.
extern "C" int b2gp_potrf_inv(b2gp_ctx* ctx, int64_t n, double* A, int64_t lda, double* Linv_out, int* info) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, A && Linv_out && info && n >= 1 && lda >= n);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_info.p, 0, 8, st));
    RET_IF(potrf_rec(ctx, st, A, lda, n, Linv_out, (int*)ctx->d_info.p, 0));
    CUDA_TRY(ctx, cudaMemcpyAsync(info, ctx->d_info.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    return tm.end(st, nullptr);
}

// B (nrhs rows of length n) <- B L^{-T} with the inverted diagonal blocks supplied by the caller
extern "C" int b2gp_trsm_inv(b2gp_ctx* ctx, int64_t n, int64_t nrhs, const double* L, int64_t ldl, const double* Linv,
                             double* B, int64_t ldb) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, L && Linv && B && n >= 1 && nrhs >= 0 && ldl >= n && ldb >= n);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    RET_IF(trsm_rec(ctx, st, B, ldb, nrhs, L, ldl, n, Linv));
    return tm.end(st, nullptr);
}

// dot[r] (+)= scale * <R[r, 0:len), w>,  nrm[r] (+)= |R[r, 0:len)|^2   (either output may be NULL)
__global__ void accumulate_kernel(double* dst, const double* src, int64_t n, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = accumulate ? dst[i] + src[i] : src[i];
}
extern "C" int b2gp_rowdot(b2gp_ctx* ctx, int64_t rows, int64_t len, const double* R, int64_t ldr, const double* w,
                           double scale, double* dot, double* nrm, int accumulate) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, R && rows >= 0 && len >= 0 && ldr >= len && (dot == nullptr || w != nullptr));
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    if (rows > 0) {
        RET_IF(ensure(ctx, ctx->slots[0].misc, (size_t)(2 * rows + 16) * 8));
        double* t1 = (double*)ctx->slots[0].misc.p;
        double* t2 = t1 + rows;
        rowdot2_kernel<<<(unsigned)rows, RD_THREADS, 0, st>>>(R, ldr, len, w, scale, dot ? t1 : nullptr, nrm ? t2 : nullptr);
        if (dot) accumulate_kernel<<<grid_for(rows), 256, 0, st>>>(dot, t1, rows, accumulate);
        if (nrm) accumulate_kernel<<<grid_for(rows), 256, 0, st>>>(nrm, t2, rows, accumulate);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 3;
    }
    return tm.end(st, nullptr);
}

extern "C" int b2gp_copy2d(b2gp_ctx* ctx, double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, dst && src && rows >= 0 && cols >= 0 && ldd >= cols && lds >= cols);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    if (rows && cols)
        CUDA_TRY(ctx, cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)cols * 8, (size_t)rows,
                                        cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return B2GP_OK;
}

// ------------------------------------------------------------------------------------------ fit side
// value and gradient (w.r.t. log lengthscale[d], log k_scale, log noise, log period) of the exact-GP log marginal
// likelihood -- see mll.cuh.  X[N,d], yres[N] host or device pointers (flags); theta is a HOST pointer (d+3);
// value, grad[d+3] and the optional alpha[N] = K^{-1} yres are HOST outputs.
// g[i] = 1/2 (alpha_i^2 - Kinv_ii): d log N(y; 0, K) / d K_ii, the gradient w.r.t. a per-point noise variance
__global__ void mll_diag_grad_kernel(const double* __restrict__ alpha, const double* __restrict__ Kinv, int64_t ldk, int64_t n,
                                     double* __restrict__ g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) g[i] = 0.5 * (alpha[i] * alpha[i] - Kinv[i * ldk + i]);
}

static int mll_impl(b2gp_ctx* ctx, int kind, const double* X, int64_t N, const double* yres, int d, const double* theta,
                    const double* noise_vec, double jitter, unsigned flags, double* value, double* grad, double* alpha_out,
                    double* grad_noise_vec, int* info) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, !grad_noise_vec || grad);
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, X && yres && theta && value && info);
    ARG_CHECK(ctx, N >= 1 && d >= 1 && d <= MLL_MAX_D);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    ex->fcache.valid = false;
    const bool dev = dev_ptrs(flags);
    Slot& sl = ctx->slots[0];
    cudaStream_t st = sl.stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const int nth = d + 3;
    const double *dX, *dy, *dth, *dnv = nullptr;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], X, (size_t)N * d * 8, dev, &dX));
    RET_IF(stage_in(ctx, st, ctx->d_in[1], yres, (size_t)N * 8, dev, &dy));
    RET_IF(stage_in(ctx, st, ctx->d_in[3], theta, (size_t)nth * 8, false, &dth));
    if (noise_vec) RET_IF(stage_in(ctx, st, ctx->d_in[6], noise_vec, (size_t)N * 8, dev, &dnv));
    const int64_t ld = round_up(N, 8);
    const int64_t tiles = ceil_div(N, MLL_TILE);
    RET_IF(ensure(ctx, sl.A, (size_t)(N + 1) * ld * 8));
    RET_IF(ensure(ctx, sl.Linv, (size_t)linv_bytes(N)));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    RET_IF(ensure(ctx, sl.misc, (size_t)(3 * ld + 64 + tiles * tiles * nth) * 8));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, 8, st));
    double* A = (double*)sl.A.p;
    double* Linv = (double*)sl.Linv.p;
    double* w = A + N * ld;              // y rides under K as a right-hand-side row: L^{-1} y after the factorisation
    double* alpha = (double*)sl.misc.p + ld;   // K^{-1} y
    double* sc = alpha + ld;             // [0] sum log L_ii, [1] |w|^2, [8..8+nth) grad
    double* partial = sc + 64;
    RET_IF(launch_gram(ctx, st, kind, dX, N, dX, N, d, dth, 1.0, jitter, 1, 1, A, ld));
    if (dnv) {   // k + diag(measured_noise) / k + diag(exp(log_var)): mngp.py:96, hskgp.py:147
        add_diag_vec_kernel<<<grid_for(N), 256, 0, st>>>(A, ld, N, dnv);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(w, dy, (size_t)N * 8, cudaMemcpyDeviceToDevice, st));
    // the scheme of the posterior (tall-panel int8 at N >= 2048); w = L^{-1} y falls out of the panel solves instead of
    // a separate chain of 2 N / 128 strip launches for one row
    RET_IF(potrf_auto(ctx, st, A, ld, N, 1, Linv, dinfo));
    logdiag_kernel<<<1, 256, 0, st>>>(A, ld, N, sc);
    rowdot2_kernel<<<1, RD_THREADS, 0, st>>>(w, ld, N, nullptr, 1.0, nullptr, sc + 1);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    if (grad || alpha_out) {
        RET_IF(ensure(ctx, sl.cov, (size_t)N * ld * 8));
        double* Bt = (double*)sl.cov.p;  // (L^{-1})^T
        set_identity_kernel<<<grid_for(N * N), 256, 0, st>>>(Bt, ld, N);
        CUDA_TRY(ctx, cudaGetLastError());
        RET_IF(trsm_rec(ctx, st, Bt, ld, N, A, ld, N, Linv));
        // alpha = L^{-T} w : alpha_i = <Bt[i,:], w>
        rowdot2_kernel<<<(unsigned)N, RD_THREADS, 0, st>>>(Bt, ld, N, w, 1.0, alpha, nullptr);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 2;
        if (grad) {
            RET_IF(ensure(ctx, sl.Vt, (size_t)N * ld * 8));
            double* Kinv = (double*)sl.Vt.p;
            // K^{-1} = L^{-T} L^{-1} accumulated onto zeros: beta = 1 is what the int8 tensor-core path takes (7 planes here)
            CUDA_TRY(ctx, cudaMemsetAsync(Kinv, 0, (size_t)N * ld * 8, st));
            RET_IF(gemm_nt(ctx, st, N, N, N, 1.0, Bt, ld, Bt, ld, 1.0, Kinv, ld, true));
            dim3 g((unsigned)tiles, (unsigned)tiles);
            mll_grad_kernel<<<g, MLL_THREADS, 0, st>>>(dX, N, d, kind, dth, alpha, Kinv, ld, partial);
            mll_finish_kernel<<<1, 32, 0, st>>>(partial, tiles * tiles, nth, sc + 8);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches += 2;
            if (grad_noise_vec) {
                mll_diag_grad_kernel<<<grid_for(N), 256, 0, st>>>(alpha, Kinv, ld, N, w);   // w is free again
                CUDA_TRY(ctx, cudaGetLastError());
                ctx->launches++;
                CUDA_TRY(ctx, cudaMemcpyAsync(grad_noise_vec, w, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
            }
        }
    }
    double hsc[8 + MLL_MAX_D + 3];
    CUDA_TRY(ctx, cudaMemcpyAsync(hsc, sc, sizeof hsc, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (alpha_out) CUDA_TRY(ctx, cudaMemcpyAsync(alpha_out, alpha, (size_t)N * 8, cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    *value = -0.5 * hsc[1] - hsc[0] - 0.5 * (double)N * 1.8378770664093453;  // log(2 pi)
    if (grad)
        for (int k = 0; k < nth; ++k) grad[k] = hsc[8 + k];
    if (*info != 0) {
        *value = NAN;
        if (grad)
            for (int k = 0; k < nth; ++k) grad[k] = NAN;
        if (grad_noise_vec)
            for (int64_t i = 0; i < N; ++i) grad_noise_vec[i] = NAN;
    }
    ex->last.flops = (double)N * N * N * (grad ? 1.0 / 3 + 1.0 + 1.0 : 1.0 / 3);
    return B2GP_OK;
}

extern "C" int b2gp_mll(b2gp_ctx* ctx, int kind, const double* X, int64_t N, const double* yres, int d, const double* theta,
                        double jitter, unsigned flags, double* value, double* grad, double* alpha_out, int* info) {
    return mll_impl(ctx, kind, X, N, yres, d, theta, nullptr, jitter, flags, value, grad, alpha_out, nullptr, info);
}

extern "C" int b2gp_mll_v(b2gp_ctx* ctx, int kind, const double* X, int64_t N, const double* yres, int d, const double* theta,
                          const double* noise_vec, double jitter, unsigned flags, double* value, double* grad, double* alpha_out,
                          double* grad_noise_vec, int* info) {
    return mll_impl(ctx, kind, X, N, yres, d, theta, noise_vec, jitter, flags, value, grad, alpha_out, grad_noise_vec, info);
}

// value and gradient of the VFE bound of the sparse GP (see sparse_elbo.cuh): d/dlog(lengthscale[d], k_scale, noise, period)
// in grad_theta[d+3] and d/dXu in grad_Xu[M,d].  Xu, X, yres follow `flags`; theta is a HOST pointer; outputs are HOST.
extern "C" int b2gp_sparse_elbo(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, const double* X, int64_t N, const double* yres,
                                int d, const double* theta, double jitter, unsigned flags, double* value, double* grad_theta,
                                double* grad_Xu, int* info) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, Xu && X && yres && theta && value && grad_theta && grad_Xu && info);
    ARG_CHECK(ctx, M >= 1 && N >= 1 && d >= 1 && d <= MLL_MAX_D);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    ex->fcache.valid = false;
    const bool dev = dev_ptrs(flags);
    Slot& sl = ctx->slots[0];
    cudaStream_t st = sl.stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const int nth = d + 3;
    const double noise = theta[d + 1], scale = theta[d];
    const double *dXu, *dX, *dy, *dth;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], X, (size_t)N * d * 8, dev, &dX));
    RET_IF(stage_in(ctx, st, ctx->d_in[1], yres, (size_t)N * 8, dev, &dy));
    RET_IF(stage_in(ctx, st, ctx->d_in[3], theta, (size_t)nth * 8, false, &dth));
    RET_IF(stage_in(ctx, st, ctx->d_in[5], Xu, (size_t)M * d * 8, dev, &dXu));
    const int64_t ldM = round_up(M, 8), ldN = round_up(N, 8);
    RET_IF(ensure(ctx, sl.A, (size_t)2 * M * ldM * 8));
    RET_IF(ensure(ctx, sl.Linv, (size_t)2 * linv_bytes(M)));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    for (int i = 0; i < 6; ++i) RET_IF(ensure(ctx, ex->eb[i], (size_t)M * ldM * 8));
    RET_IF(ensure(ctx, ex->eb[6], (size_t)N * ldM * 8));
    RET_IF(ensure(ctx, ex->eb[7], (size_t)N * ldM * 8));
    RET_IF(ensure(ctx, ex->eb[8], (size_t)M * ldN * 8));
    RET_IF(ensure(ctx, ex->eb[9], (size_t)(4 * ldM + 3 * ldN + 64 + M * (nth + d)) * 8));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, 16, st));
    double* Luu = (double*)sl.A.p;
    double* Cm = Luu + M * ldM;
    double* LinvU = (double*)sl.Linv.p;
    double* LinvC = LinvU + linv_bytes(M) / 8;
    double *BtU = (double*)ex->eb[0].p, *BtC = (double*)ex->eb[1].p, *Cinv = (double*)ex->eb[2].p;
    double *T1 = (double*)ex->eb[3].p, *T2 = (double*)ex->eb[4].p, *T3 = (double*)ex->eb[5].p;
    double *E = (double*)ex->eb[6].p, *GKuft = (double*)ex->eb[7].p, *GKuf = (double*)ex->eb[8].p;
    double* vec = (double*)ex->eb[9].p;
    double *bvec = vec, *u = vec + ldM, *beta = vec + 2 * ldM, *tmpM = vec + 3 * ldM;
    double *tmpN = vec + 4 * ldM, *alpha = tmpN + ldN, *tmpN2 = alpha + ldN;
    double* scal = tmpN2 + ldN;          // [0] sum log LC_ii, [1] u'u, [2] y'y, [3] |W|_F^2, [4] tr(C^-1), [5] alpha'alpha, [8..] chain
    double* partial = scal + 64;         // M x (d+3)
    double* gXu = partial + M * nth;     // M x d
    dim3 b32(32, 32), gMM((unsigned)ceil_div(M, 32), (unsigned)ceil_div(M, 32));

    // forward pieces shared with the posterior: Luu, W (both layouts), W W^T / noise, W y / noise
    RET_IF(sparse_partial_dev(ctx, sl, kind, dXu, M, dX, N, dy, d, dth, jitter, noise, Luu, ldM, LinvU, Cm, ldM, bvec, dinfo));
    double* Wt = (double*)sl.Vt.p;
    double* W = (double*)sl.cov.p;
    add_diag_kernel<<<grid_for(M), 256, 0, st>>>(Cm, ldM, M, 1.0);
    RET_IF(potrf_rec(ctx, st, Cm, ldM, M, LinvC, dinfo + 1, 0));
    CUDA_TRY(ctx, cudaMemcpyAsync(u, bvec, (size_t)M * 8, cudaMemcpyDeviceToDevice, st));
    RET_IF(trsm_rec(ctx, st, u, ldM, 1, Cm, ldM, M, LinvC));
    logdiag_kernel<<<1, 256, 0, st>>>(Cm, ldM, M, scal + 0);
    rowdot2_kernel<<<1, RD_THREADS, 0, st>>>(u, ldM, M, nullptr, 1.0, nullptr, scal + 1);
    rowdot2_kernel<<<1, RD_THREADS, 0, st>>>(dy, ldN, N, nullptr, 1.0, nullptr, scal + 2);
    rowdot2_kernel<<<(unsigned)M, RD_THREADS, 0, st>>>(W, ldN, N, nullptr, 1.0, nullptr, tmpM);
    vecsum_kernel<<<1, 256, 0, st>>>(tmpM, M, scal + 3);
    // C^{-1} = BtC BtC^T with BtC = (LC^{-1})^T, beta = C^{-1} b
    set_identity_kernel<<<grid_for(M * M), 256, 0, st>>>(BtC, ldM, M);
    RET_IF(trsm_rec(ctx, st, BtC, ldM, M, Cm, ldM, M, LinvC));
    rowdot2_kernel<<<(unsigned)M, RD_THREADS, 0, st>>>(BtC, ldM, M, u, 1.0, beta, tmpM);
    vecsum_kernel<<<1, 256, 0, st>>>(tmpM, M, scal + 4);
    RET_IF(gemm_nt(ctx, st, M, M, M, 1.0, BtC, ldM, BtC, ldM, 0.0, Cinv, ldM, true));
    mirror_lower_kernel<<<gMM, b32, 0, st>>>(Cinv, ldM, M);
    // alpha = (y - W^T beta) / noise
    rowdot2_kernel<<<(unsigned)N, RD_THREADS, 0, st>>>(Wt, ldM, M, beta, 1.0, tmpN, nullptr);
    elbo_alpha_kernel<<<grid_for(N), 256, 0, st>>>(alpha, dy, tmpN, N, noise);
    rowdot2_kernel<<<1, RD_THREADS, 0, st>>>(alpha, ldN, N, nullptr, 1.0, nullptr, scal + 5);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 14;
    // the clip of the trace term decides a coefficient of the reverse pass: fetch the scalars now
    double hs[8];
    CUDA_TRY(ctx, cudaMemcpyAsync(hs, scal, sizeof hs, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    double kd = scale;
    if (kind == B2GP_KERNEL_MATERN52) {
        const double r = sqrt(1e-12), s5r = 2.23606797749979 * r;
        kd = scale * (1.0 + s5r) * exp(-s5r);
    }
    const double T = (double)N * kd - hs[3];
    const double coef = (T > 0.0) ? 1.0 : 0.0;
    // dELBO/dW^T = alpha beta^T + (coef W^T - W^T C^{-1}) / noise
    RET_IF(gemm_nt(ctx, st, N, M, M, 1.0, Wt, ldM, Cinv, ldM, 0.0, E, ldM, false));
    elbo_gw_kernel<<<grid_for(N * M), 256, 0, st>>>(E, ldM, Wt, ldM, alpha, beta, N, M, coef, noise);
    // dELBO/dKuf^T = (dELBO/dW^T) Luu^{-1}
    set_identity_kernel<<<grid_for(M * M), 256, 0, st>>>(BtU, ldM, M);
    RET_IF(trsm_rec(ctx, st, BtU, ldM, M, Luu, ldM, M, LinvU));
    RET_IF(gemm_nt(ctx, st, N, M, M, 1.0, E, ldM, BtU, ldM, 0.0, GKuft, ldM, false));
    {
        dim3 g((unsigned)ceil_div(M, 32), (unsigned)ceil_div(N, 32)), b(32, 8);
        transpose_kernel<<<g, b, 0, st>>>(GKuf, ldN, GKuft, ldM, N, M);
    }
    // H^T = W G_Kuf^T; G_L = -tril(H); dELBO/dKuu = Luu^{-T} Phi(Luu^T G_L) Luu^{-1}
    RET_IF(gemm_nt(ctx, st, M, M, N, 1.0, W, ldN, GKuf, ldN, 0.0, T1, ldM, false));
    tri_kernel<<<gMM, b32, 0, st>>>(T2, ldM, T1, ldM, M, 1);              // T2 = G_L^T = -triu(H^T)
    tri_kernel<<<gMM, b32, 0, st>>>(T3, ldM, Luu, ldM, M, 0);             // T3 = tril(Luu)
    {
        dim3 b(32, 8);
        transpose_kernel<<<gMM, b, 0, st>>>(T1, ldM, T3, ldM, M, M);      // T1 = Luu^T
    }
    RET_IF(gemm_nt(ctx, st, M, M, M, 1.0, T1, ldM, T2, ldM, 0.0, T3, ldM, false));   // T3 = Luu^T G_L
    tri_kernel<<<gMM, b32, 0, st>>>(T2, ldM, T3, ldM, M, 2);              // T2 = Phi(.)
    RET_IF(gemm_nt(ctx, st, M, M, M, 1.0, BtU, ldM, T2, ldM, 0.0, T1, ldM, false));  // T1 = BtU P^T
    RET_IF(gemm_nt(ctx, st, M, M, M, 1.0, BtU, ldM, T1, ldM, 0.0, T3, ldM, false));  // T3 = dELBO/dKuu
    tri_kernel<<<gMM, b32, 0, st>>>(T2, ldM, T3, ldM, M, 3);              // T2 = T3 + T3^T
    // contract with the kernel derivatives
    elbo_chain_kernel<<<(unsigned)M, 256, 0, st>>>(kind, d, dth, dXu, (int)M, dX, N, GKuf, ldN, GKuf, ldN, 0, 0, partial, gXu);
    elbo_chain_kernel<<<(unsigned)M, 256, 0, st>>>(kind, d, dth, dXu, (int)M, dXu, M, T3, ldM, T2, ldM, 1, 1, partial, gXu);
    colsum_kernel<<<1, 32, 0, st>>>(partial, M, nth, scal + 8);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 12;
    double hs2[8 + MLL_MAX_D + 3];
    int hinfo[2] = {0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(hs2, scal, sizeof hs2, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(grad_Xu, gXu, (size_t)M * d * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(hinfo, dinfo, sizeof hinfo, cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    *info = hinfo[0] != 0 ? hinfo[0] : -hinfo[1];
    const double n = (double)N, m = (double)M;
    const double loglik = -0.5 * (n * 1.8378770664093453 + n * log(noise) + 2.0 * hs2[0] + hs2[2] / noise - hs2[1]);
    *value = loglik - 0.5 * (T > 0.0 ? T / noise : 0.0);
    for (int k = 0; k < nth; ++k) grad_theta[k] = hs2[8 + k];
    grad_theta[d] += (T > 0.0) ? -0.5 * n * kd / noise : 0.0;
    const double trSinv = (n - m + hs2[4]) / noise;
    grad_theta[d + 1] = noise * (-0.5 * trSinv + 0.5 * hs2[5] + ((T > 0.0) ? 0.5 * T / (noise * noise) : 0.0));
    if (*info != 0) {
        *value = NAN;
        for (int k = 0; k < nth; ++k) grad_theta[k] = NAN;
    }
    return B2GP_OK;
}

// ------------------------------------------------------------------------------------------ MVN sampling
// y[s, i, :] = mean[s, :] + chol(cov[s]) eps[s, i, :]  -- numpyro.distributions.MultivariateNormal(mean, cov).sample
// (call sites gpax/models/gp.py:292, gpax/models/hskgp.py via ExactGP._predict, gpax/acquisition/base_acq.py:221) for
// covariances the caller modified after the posterior call (VarNoiseGP adds the predicted noise to the diagonal).
extern "C" int b2gp_mvn_sample(b2gp_ctx* ctx, const double* mean, const double* cov, int64_t S, int64_t P, const double* eps,
                               int64_t n, double* y, int* info, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, mean && cov && eps && y && info && S >= 1 && P >= 1 && n >= 1);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool dev = dev_ptrs(flags);
    Slot& sl = ctx->slots[0];
    cudaStream_t st = sl.stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const double *dmean, *dcov, *deps;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], mean, (size_t)S * P * 8, dev, &dmean));
    RET_IF(stage_in(ctx, st, ctx->d_in[1], cov, (size_t)S * P * P * 8, dev, &dcov));
    RET_IF(stage_in(ctx, st, ctx->d_in[2], eps, (size_t)S * n * P * 8, dev, &deps));
    double* dy = y;
    if (!dev) {
        RET_IF(ensure(ctx, ctx->d_out[3], (size_t)S * n * P * 8));
        dy = (double*)ctx->d_out[3].p;
    }
    const int64_t ldC = round_up(P, 8);
    RET_IF(ensure(ctx, sl.cov, (size_t)P * ldC * 8));
    RET_IF(ensure(ctx, sl.LinvC, (size_t)linv_bytes(P)));
    RET_IF(ensure(ctx, ctx->d_info, (size_t)S * sizeof(int)));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, (size_t)S * sizeof(int), st));
    double* CL = (double*)sl.cov.p;
    dim3 g2((unsigned)ceil_div(P, 32), (unsigned)ceil_div(P, 32)), b2(32, 32);
    for (int64_t s = 0; s < S; ++s) {
        copy2d_kernel<<<grid_for(P * P), 256, 0, st>>>(CL, ldC, dcov + s * P * P, P, P, P);
        RET_IF(potrf_rec(ctx, st, CL, ldC, P, (double*)sl.LinvC.p, dinfo + s, 0));
        zero_upper_kernel<<<g2, b2, 0, st>>>(CL, ldC, P);
        double* Y = dy + s * n * P;
        bcast_rows_kernel<<<grid_for(n * P), 256, 0, st>>>(Y, P, n, P, dmean + s * P);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches += 3;
        RET_IF(gemm_nt(ctx, st, n, P, P, 1.0, deps + s * n * P, P, CL, ldC, 1.0, Y, P, false));
        nan_if_bad_kernel<<<grid_for(n * P), 256, 0, st>>>(Y, P, n, P, dinfo + s, nullptr);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(info, dinfo, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (!dev) CUDA_TRY(ctx, cudaMemcpyAsync(y, dy, (size_t)S * n * P * 8, cudaMemcpyDeviceToHost, st));
    return tm.end(st, nullptr);
}

// ------------------------------------------------------------------------------------------ acquisition epilogues
// see include/b200gp.h; kernels in acq.cuh
extern "C" int b2gp_acq_moments(b2gp_ctx* ctx, int kind, const double* mean, const double* var, int64_t R, int64_t P, int have_best,
                                double best_f, double param, int maximize, double* out, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= ACQ_EI && kind <= ACQ_POI);
    ARG_CHECK(ctx, var && out && R >= 1 && P >= 1);
    ARG_CHECK(ctx, mean || kind == ACQ_UE);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool dev = dev_ptrs(flags);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const double *dmean, *dvar;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], mean, (size_t)R * P * 8, dev, &dmean));
    RET_IF(stage_in(ctx, st, ctx->d_in[1], var, (size_t)R * P * 8, dev, &dvar));
    double* dout = out;
    if (!dev) {
        RET_IF(ensure(ctx, ctx->d_out[0], (size_t)R * P * 8));
        dout = (double*)ctx->d_out[0].p;
    }
    RET_IF(ensure(ctx, ctx->slots[0].misc, (size_t)(R + 16) * 8));
    double* dbest = (double*)ctx->slots[0].misc.p;
    if (kind == ACQ_EI || kind == ACQ_POI) {
        if (have_best) {
            std::vector<double> hb((size_t)R, best_f);
            CUDA_TRY(ctx, cudaMemcpyAsync(dbest, hb.data(), (size_t)R * 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(ctx, cudaStreamSynchronize(st));
        } else {
            acq_best_kernel<<<(unsigned)R, 256, 0, st>>>(dmean, P, P, maximize, dbest);
            ctx->launches++;
        }
    }
    acq_moments_kernel<<<grid_for(R * P), 256, 0, st>>>(kind, dmean, dvar, P, R, P, dbest, param, maximize, dout, P);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if (!dev) CUDA_TRY(ctx, cudaMemcpyAsync(out, dout, (size_t)R * P * 8, cudaMemcpyDeviceToHost, st));
    return tm.end(st, nullptr);
}

extern "C" int b2gp_acq_samples(b2gp_ctx* ctx, int kind, const double* y, int64_t R, int64_t P, int have_best, double best_f,
                                double param, int maximize, double* out, double* mean_out, double* var_out, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, kind >= ACQ_EI && kind <= ACQ_POI);
    ARG_CHECK(ctx, y && out && R >= 1 && P >= 1);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool dev = dev_ptrs(flags);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const double* dy;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], y, (size_t)R * P * 8, dev, &dy));
    RET_IF(ensure(ctx, ctx->d_out[0], (size_t)3 * P * 8));
    double* dm = (double*)ctx->d_out[0].p;
    double* dv = dm + P;
    double* dout = dev ? out : dv + P;
    RET_IF(ensure(ctx, ctx->slots[0].misc, 16 * 8));
    double* dbest = (double*)ctx->slots[0].misc.p;
    sample_moments_kernel<<<(unsigned)ceil_div(P, 128), 128, 0, st>>>(dy, R, P, dm, dv);
    ctx->launches++;
    if (kind == ACQ_EI || kind == ACQ_POI) {
        if (have_best) {
            CUDA_TRY(ctx, cudaMemcpyAsync(dbest, &best_f, 8, cudaMemcpyHostToDevice, st));
            CUDA_TRY(ctx, cudaStreamSynchronize(st));
        } else {
            acq_best_kernel<<<1, 256, 0, st>>>(dm, P, P, maximize, dbest);
            ctx->launches++;
        }
    }
    acq_moments_kernel<<<grid_for(P), 256, 0, st>>>(kind, dm, dv, P, 1, P, dbest, param, maximize, dout, P);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    const cudaMemcpyKind kd = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (!dev) CUDA_TRY(ctx, cudaMemcpyAsync(out, dout, (size_t)P * 8, kd, st));
    if (mean_out) CUDA_TRY(ctx, cudaMemcpyAsync(mean_out, dm, (size_t)P * 8, kd, st));
    if (var_out) CUDA_TRY(ctx, cudaMemcpyAsync(var_out, dv, (size_t)P * 8, kd, st));
    return tm.end(st, nullptr);
}

extern "C" int b2gp_kg(b2gp_ctx* ctx, const double* mean, const double* cov, int64_t P, const double* ysim, int64_t n,
                       double diag_sub, double noise_plus_jitter, int maximize, double* out, unsigned flags) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, mean && cov && ysim && out && P >= 1 && n >= 1);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const bool dev = dev_ptrs(flags);
    cudaStream_t st = ctx->slots[0].stream;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const double *dmean, *dcov, *dys;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], mean, (size_t)P * 8, dev, &dmean));
    RET_IF(stage_in(ctx, st, ctx->d_in[1], cov, (size_t)P * P * 8, dev, &dcov));
    RET_IF(stage_in(ctx, st, ctx->d_in[2], ysim, (size_t)n * P * 8, dev, &dys));
    double* dout = out;
    if (!dev) {
        RET_IF(ensure(ctx, ctx->d_out[0], (size_t)P * 8));
        dout = (double*)ctx->d_out[0].p;
    }
    RET_IF(ensure(ctx, ctx->slots[0].misc, 16 * 8));
    double* dbest = (double*)ctx->slots[0].misc.p;
    acq_best_kernel<<<1, 256, 0, st>>>(dmean, P, P, maximize, dbest);
    kg_kernel<<<(unsigned)P, 256, 0, st>>>(dmean, dcov, P, dys, (int)n, P, diag_sub, noise_plus_jitter, maximize, dbest, dout);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    if (!dev) CUDA_TRY(ctx, cudaMemcpyAsync(out, dout, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    return tm.end(st, nullptr);
}

// ------------------------------------------------------------------------------------------ debug
// Development aid (not part of include/b200gp.h): run the leaf kernel on a device block with per-phase
// clock64 / globaltimer stamps.  prof_host receives 2*16 values (cycles, ns) per stamp.
extern "C" int b2gp_debug_leaf(b2gp_ctx* ctx, int n, double* A_dev, int64_t lda, double* linv_dev, long long* prof_host) {
    if (!ctx) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    RET_IF(ensure(ctx, ctx->d_info, 64 + 128 * 8));
    int* dinfo = (int*)ctx->d_info.p;
    long long* dprof = (long long*)((char*)ctx->d_info.p + 64);
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_info.p, 0, 64 + 128 * 8, st));
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PD_SMEM));
        attr.done(ctx->device);
    }
    potrf_diag_kernel<<<1, PD_THREADS, PD_SMEM, st>>>(A_dev, lda, n, linv_dev, dinfo, 0, dprof);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaMemcpyAsync(prof_host, dprof, 64 * 8, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    return B2GP_OK;
}

// Development aid: time one GEMM/SYRK launch with an explicit tile configuration (device pointers).
extern "C" int b2gp_debug_gemm_cfg(b2gp_ctx* ctx, int cfg, int64_t m, int64_t n, int64_t k, const double* A, int64_t lda,
                                   const double* B, int64_t ldb, double* C, int64_t ldc, int lower, double* ms_out) {
    if (!ctx) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    GemmArgs a;
    a.m = (int)m; a.n = (int)n; a.k = (int)k; a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
    a.alpha = -1.0; a.beta = 1.0; a.lower_only = lower;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, st));
    int rc = B2GP_ERR_ARG;
    switch (cfg) {
        case 0: rc = launch_gemm_cfg<128, 128, 2, 4, 4, 1>(ctx, st, a); break;
        case 1: rc = launch_gemm_cfg<128, 128, 4, 4, 4, 1>(ctx, st, a); break;
        case 2: rc = launch_gemm_cfg<128, 128, 2, 4, 3, 1>(ctx, st, a); break;
        case 3: rc = launch_gemm_cfg<128, 128, 4, 2, 4, 1>(ctx, st, a); break;
        case 4: rc = launch_gemm_cfg<128, 128, 2, 8, 4, 1>(ctx, st, a); break;
        case 5: rc = launch_gemm_cfg<128, 128, 4, 4, 3, 1>(ctx, st, a); break;
        default: break;
    }
    RET_IF(rc);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, st));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_b));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    *ms_out = ms;
    return B2GP_OK;
}

// Development aid: C[m,n] += alpha * A B^T through the int8 tcgen05 path (ozaki.cuh), device pointers, S = 7 or 8 digit planes.
extern "C" int b2gp_debug_ozaki(b2gp_ctx* ctx, int S, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                                const double* B, int64_t ldb, double* C, int64_t ldc, int lower, double* ms_out) {
    if (!ctx) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    Extra* ex = extra_of(ctx);
    RET_IF(ensure(ctx, ctx->slots[0].oz.prof, 148 * 8 * 8));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->slots[0].oz.prof.p, 0, 148 * 8 * 8, st));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, st));
    int rc;
    if (S == 6)
        rc = ozaki_gemm_nt<6>(ctx, st, ctx->slots[0].oz, m, n, k, alpha, A, lda, B, ldb, C, ldc, lower != 0);
    else
        rc = ozaki_gemm_nt<7>(ctx, st, ctx->slots[0].oz, m, n, k, alpha, A, lda, B, ldb, C, ldc, lower != 0);
    RET_IF(rc);
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, st));
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_b));
    float ms = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    if (ms_out) *ms_out = ms;
    {
        long long h[148 * 8];
        CUDA_TRY(ctx, cudaMemcpy(h, ctx->slots[0].oz.prof.p, sizeof h, cudaMemcpyDeviceToHost));
        double w = 0, e = 0, u = 0, t = 0, mf = 0, me = 0, mt = 0;
        for (int i = 0; i < 148; ++i) {
            w += h[8 * i]; e += h[8 * i + 1]; u += h[8 * i + 2]; t += h[8 * i + 3];
            mf += h[8 * i + 4]; me += h[8 * i + 5]; mt += h[8 * i + 6];
        }
        if (t > 0 && getenv("B2GP_OZ_PROF"))
            fprintf(stderr, "[oz prof] per tile: epilogue waits for MMA %.0f cyc, epilogue %.0f cyc (C update %.0f); MMA issuer: total %.0f, waits for TMA %.0f, "
                            "for TMEM drain %.0f; tiles %.0f, k-blocks/tile %lld\n",
                    w / t, e / t, u / t, mt / t, mf / t, me / t, t, (long long)((k + 31) / 32));
    }
    return B2GP_OK;
}

// Development aid / roofline denominator: the measured int8 tcgen05 ceiling in TOP/s (see oz_i8_peak_kernel), best of `reps`.
extern "C" int b2gp_debug_i8_peak(b2gp_ctx* ctx, int iters, int reps, double* tops_out, double* ms_out) {
    if (!ctx || !tops_out || iters < 4 || reps < 1) return B2GP_ERR_ARG;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    const int smem = 200 * 1024;   // one CTA per SM
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(oz_i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr.done(ctx->device);
    }
    double best = 1e30;
    for (int r = 0; r < reps + 1; ++r) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_a, st));
        oz_i8_peak_kernel<<<ctx->sm_count, 128, smem, st>>>(iters, nullptr);
        CUDA_TRY(ctx, cudaGetLastError());
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev_b, st));
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_b));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
        if (r > 0 && ms < best) best = ms;
    }
    *tops_out = 2.0 * 128 * 256 * 32 * (double)iters * ctx->sm_count / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return B2GP_OK;
}

// ------------------------------------------------------------------------------------------ multi-GPU (dist.cuh)
extern "C" int b2gp_dist_unique_id(void* id128) {
    if (!id128) return B2GP_ERR_ARG;
    NcclApi* n = nccl_api();
    if (!n->handle || !n->error.empty()) return B2GP_ERR_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (n->GetUniqueId(&id) != ncclSuccess) return B2GP_ERR_CUDA;
    memcpy(id128, &id, 128);
    return B2GP_OK;
}

static int dist_free(b2gp_ctx* ctx, DistState* ds) {
    NcclApi* n = nccl_api();
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (DevBuf* b : {&ds->Aloc, &ds->PB[0], &ds->PB[1], &ds->UB, &ds->EB, &ds->Xrows, &ds->Zcols, &ds->yloc, &ds->red, &ds->linv, &ds->updA, &ds->updB,
                      &ds->updSA, &ds->updSB, &ds->wseg})
        free_buf(*b);
    for (auto& st : ds->steps) {
        free_buf(st.g1);
        free_buf(st.g2);
        free_buf(st.bmap);
    }
    for (cudaEvent_t e : {ds->ev_u, ds->ev_ubc[0], ds->ev_ubc[1], ds->ev_chunk, ds->ev_comm[0], ds->ev_comm[1], ds->ev_done, ds->ev_e,
                          ds->ev_early[0], ds->ev_early[1], ds->ev_g2})
        if (e) cudaEventDestroy(e);
    if (ds->ms) cudaStreamDestroy(ds->ms);
    if (ds->dq) cudaStreamDestroy(ds->dq);
    if (n->handle) {
        if (ds->rowc) n->CommDestroy(ds->rowc);
        if (ds->colc) n->CommDestroy(ds->colc);
        if (ds->world) n->CommDestroy(ds->world);
    }
    delete ds;
    return B2GP_OK;
}

extern "C" int b2gp_dist_finalize(b2gp_ctx* ctx) {
    if (!ctx) return B2GP_ERR_ARG;
    Extra* ex = extra_of(ctx);
    if (ex->dist) {
        dist_free(ctx, ex->dist);
        ex->dist = nullptr;
    }
    return B2GP_OK;
}

extern "C" int b2gp_dist_init(b2gp_ctx* ctx, const void* id128, int rank, int nranks, int grid_rows, int grid_cols) {
    if (!ctx) return B2GP_ERR_ARG;
    ARG_CHECK(ctx, id128 && nranks >= 1 && rank >= 0 && rank < nranks);
    ARG_CHECK(ctx, grid_rows >= 1 && grid_cols >= 1 && grid_rows * grid_cols == nranks);
    NcclApi* n = nccl_api();
    if (!n->handle || !n->error.empty())
        return set_err(ctx, B2GP_ERR_UNSUPPORTED, "b2gp_dist_init", n->error.c_str(), __FILE__, __LINE__);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    Extra* ex = extra_of(ctx);
    if (ex->dist) RET_IF(b2gp_dist_finalize(ctx));
    DistState* ds = new DistState();
    ex->dist = ds;
    ds->rank = rank;
    ds->nranks = nranks;
    ds->g.pr = grid_rows;
    ds->g.pc = grid_cols;
    ds->g.myrow = rank / grid_cols;     // row-major process grid
    ds->g.mycol = rank % grid_cols;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    NCCL_TRY(ctx, n->CommInitRank(&ds->world, nranks, id, rank));
    // row communicator: the pc processes of my grid row, ranked by column; column communicator: the pr processes of my
    // grid column, ranked by row
    NCCL_TRY(ctx, n->CommSplit(ds->world, ds->g.myrow, ds->g.mycol, &ds->rowc, nullptr));
    NCCL_TRY(ctx, n->CommSplit(ds->world, ds->g.mycol, ds->g.myrow, &ds->colc, nullptr));
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ds->ms, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ds->dq, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&ds->ev_u, &ds->ev_ubc[0], &ds->ev_ubc[1], &ds->ev_chunk, &ds->ev_comm[0], &ds->ev_comm[1], &ds->ev_done, &ds->ev_e,
                           &ds->ev_early[0], &ds->ev_early[1], &ds->ev_g2})
        CUDA_TRY(ctx, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    ds->ready = true;
    return B2GP_OK;
}

extern "C" int b2gp_dist_info(b2gp_ctx* ctx, int* rank, int* nranks, int* grid_rows, int* grid_cols) {
    if (!ctx) return B2GP_ERR_ARG;
    DistState* ds = extra_of(ctx)->dist;
    if (!ds || !ds->ready) return set_err(ctx, B2GP_ERR_ARG, "b2gp_dist_info", "call b2gp_dist_init first", __FILE__, __LINE__);
    if (rank) *rank = ds->rank;
    if (nranks) *nranks = ds->nranks;
    if (grid_rows) *grid_rows = ds->g.pr;
    if (grid_cols) *grid_cols = ds->g.pc;
    return B2GP_OK;
}

// Pure index algebra of the block-cyclic layout (no GPU, no NCCL): for the process at (row, col) of a pr x pc grid,
// out[0] = local tile rows, out[1] = local tile columns, out[2] = rows of the panel of step k it holds, out[3] = slot
// rows of the panel buffer at step k, out[4] = first local tile row after k, out[5] = first local tile column after k.
extern "C" int b2gp_dist_layout(int64_t T, int64_t R, int64_t nb, int pr, int pc, int row, int col, int64_t k, int64_t* out) {
    if (!out || T < 1 || R < 0 || nb < 1 || pr < 1 || pc < 1 || row < 0 || row >= pr || col < 0 || col >= pc || k < 0) return B2GP_ERR_ARG;
    BcGrid g;
    g.pr = pr;
    g.pc = pc;
    g.myrow = row;
    g.mycol = col;
    g.nb = nb;
    g.T = T;
    g.R = R;
    out[0] = g.lr(row);
    out[1] = g.lc(col);
    out[2] = g.panel_rows(k, row);
    out[3] = g.slot_rows(k);
    out[4] = g.first_row_after(k, row);
    out[5] = g.first_col_after(k, col);
    return B2GP_OK;
}

// staircase tile lists and B-operand row maps of every step, for the current grid / problem shape
static int dist_build_steps(b2gp_ctx* ctx, DistState* ds, cudaStream_t st, int CL) {
    const BcGrid& g = ds->g;
    if (ds->cache_T == g.T && ds->cache_R == g.R && ds->cache_nb == g.nb && ds->cache_cl == CL) return B2GP_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(st));   // lists of a previous shape may still be in use
    for (auto& s : ds->steps) {
        free_buf(s.g1);
        free_buf(s.g2);
        free_buf(s.bmap);
    }
    ds->steps.assign((size_t)g.T, DistStep());
    const int64_t nb = g.nb, t128 = nb / 128;          // 128-row groups per tile
    const int64_t colw = CL == 2 ? 128 : 64;           // columns covered by one list entry
    const int64_t ent_per_tile = nb / colw;
    for (int64_t k = 0; k < g.T; ++k) {
        DistStep& s = ds->steps[(size_t)k];
        const int64_t li0 = g.first_row_after(k, g.myrow), lj0 = g.first_col_after(k, g.mycol);
        const int64_t ntr = g.lr(g.myrow) - li0, ntc = g.lc(g.mycol) - lj0;    // local tile rows / columns of the trailing matrix
        if (ntr <= 0 || ntc <= 0) continue;
        // B operand: the panel tiles (gj, k) of my tile columns gj > k, found in slot gj % pr of the panel buffer
        const int64_t slot = g.slot_rows(k);
        std::vector<int64_t> bmap((size_t)(ntc * t128));
        for (int64_t c = 0; c < ntc; ++c) {
            const int64_t gj = (lj0 + c) * g.pc + g.mycol;
            const int r = (int)(gj % g.pr);
            const int64_t src = (int64_t)r * slot + (gj / g.pr - g.first_row_after(k, r)) * nb;
            for (int64_t q = 0; q < t128; ++q) bmap[(size_t)(c * t128 + q)] = src + q * 128;
        }
        RET_IF(ensure(ctx, s.bmap, bmap.size() * 8));
        CUDA_TRY(ctx, cudaMemcpyAsync(s.bmap.p, bmap.data(), bmap.size() * 8, cudaMemcpyHostToDevice, st));
        // staircase: 128-row tile ti of local tile row a (global gi) x column entry of local tile column c (global gj), gi >= gj
        std::vector<int2> l1, l2;
        const int64_t next_col = (k + 1 < g.T && (k + 1) % g.pc == g.mycol) ? (k + 1) / g.pc - lj0 : -1;   // local index (in the trailing matrix) of tile column k+1
        const int G = 8;
        const int64_t rows128 = ntr * t128;
        for (int64_t b0 = 0; b0 < rows128; b0 += G) {
            const int64_t b1 = std::min(b0 + G, rows128);
            for (int64_t c = 0; c < ntc; ++c) {
                const int64_t gj = (lj0 + c) * g.pc + g.mycol;
                for (int64_t e = 0; e < ent_per_tile; ++e)
                    for (int64_t ti = b0; ti < b1; ++ti) {
                        const int64_t gi = (li0 + ti / t128) * g.pr + g.myrow;
                        if (gi < gj) continue;
                        if (gi == k + 1 && gj == k + 1) continue;    // the next diagonal tile takes panel k on the diagonal stream (diag_factor)
                        (c == next_col ? l1 : l2).push_back(make_int2((int)ti, (int)(c * ent_per_tile + e)));
                    }
            }
        }
        s.n1 = (int64_t)l1.size();
        s.n2 = (int64_t)l2.size();
        if (s.n1) {
            RET_IF(ensure(ctx, s.g1, l1.size() * sizeof(int2)));
            CUDA_TRY(ctx, cudaMemcpyAsync(s.g1.p, l1.data(), l1.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
        }
        if (s.n2) {
            RET_IF(ensure(ctx, s.g2, l2.size() * sizeof(int2)));
            CUDA_TRY(ctx, cudaMemcpyAsync(s.g2.p, l2.data(), l2.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
        }
        CUDA_TRY(ctx, cudaStreamSynchronize(st));   // the host vectors die here
    }
    ds->cache_T = g.T;
    ds->cache_R = g.R;
    ds->cache_nb = g.nb;
    ds->cache_cl = CL;
    return B2GP_OK;
}

// Exact-GP posterior (mean + diagonal variance) with k_XX distributed over the process grid.  COLLECTIVE: every rank
// calls it with the same (replicated) inputs -- X, y, X_new and theta are a few hundred KB, only K is big.  HOST pointers.
extern "C" int b2gp_dist_posterior(b2gp_ctx* ctx, int kind, const double* Xtr, int64_t N, const double* yres, const double* Xnew,
                                   int64_t P, int d, const double* theta, int noiseless, double jitter, int64_t nb, unsigned flags,
                                   double* mean, double* var, int* info, b2gp_timing* timing) {
    if (!ctx) return B2GP_ERR_ARG;
    Extra* ex = extra_of(ctx);
    DistState* ds = ex->dist;
    if (!ds || !ds->ready) return set_err(ctx, B2GP_ERR_ARG, "b2gp_dist_posterior", "call b2gp_dist_init first", __FILE__, __LINE__);
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, Xtr && yres && Xnew && theta && mean && info);
    ARG_CHECK(ctx, N >= 1 && P >= 1 && d >= 1 && d <= GRAM_MAX_D);
    ARG_CHECK(ctx, nb >= 128 && nb <= 1024 && nb % 128 == 0 && N % nb == 0);
    ARG_CHECK(ctx, !(flags & B2GP_FLAG_DEVICE_PTRS) && !(flags & (B2GP_OUT_COV | B2GP_OUT_SAMPLE)));
    if (ctx->ozaki == 0) return set_err(ctx, B2GP_ERR_UNSUPPORTED, "b2gp_dist_posterior", "needs the int8 path (ozaki != 0)", __FILE__, __LINE__);
    const bool want_var = flags & B2GP_OUT_VAR;
    ARG_CHECK(ctx, !want_var || var);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    NcclApi* nc = nccl_api();
    ex->fcache.valid = false;
    Slot& sl = ctx->slots[0];
    cudaStream_t cs = sl.stream, ms = ds->ms;
    BcGrid& g = ds->g;
    g.nb = nb;
    g.T = N / nb;
    g.R = ceil_div(P + 1, nb);
    const int pr = g.pr, pc = g.pc, myrow = g.myrow, mycol = g.mycol;
    const int64_t T = g.T, Lr = g.lr(myrow), Lc = g.lc(mycol), ld = Lc * nb;
    const int CL = (ctx->oz_cluster == 2) ? 2 : 1;
    const int nth = d + 3;
    CallTimer tm(ctx);
    RET_IF(tm.begin(cs));
    RET_IF(dist_build_steps(ctx, ds, cs, CL));

    // ---- inputs: the rows / columns of X this process's tiles need, gathered on the host (a few hundred KB)
    std::vector<double> xr((size_t)(Lr * nb * d)), zc((size_t)std::max<int64_t>(Lc * nb * d, 1)), yl((size_t)std::max<int64_t>(Lc * nb, 1));
    for (int64_t li = 0; li < Lr; ++li) {
        const int64_t gi = li * pr + myrow;
        for (int64_t r = 0; r < nb; ++r) {
            const double* src;
            if (gi < T)
                src = Xtr + (gi * nb + r) * d;
            else {
                const int64_t p = (gi - T) * nb + r;
                src = Xnew + (p < P ? p : 0) * d;       // padding rows repeat a valid point; their results are never read
            }
            memcpy(&xr[(size_t)((li * nb + r) * d)], src, (size_t)d * 8);
        }
    }
    for (int64_t lj = 0; lj < Lc; ++lj) {
        const int64_t gj = lj * pc + mycol;
        memcpy(&zc[(size_t)(lj * nb * d)], Xtr + gj * nb * d, (size_t)(nb * d) * 8);
        memcpy(&yl[(size_t)(lj * nb)], yres + gj * nb, (size_t)nb * 8);
    }
    RET_IF(ensure(ctx, ds->Xrows, xr.size() * 8));
    RET_IF(ensure(ctx, ds->Zcols, zc.size() * 8));
    RET_IF(ensure(ctx, ds->yloc, yl.size() * 8));
    RET_IF(ensure(ctx, ctx->d_in[3], (size_t)nth * 8));
    RET_IF(ensure(ctx, ds->Aloc, (size_t)std::max<int64_t>(Lr * nb * ld, 1) * 8));
    RET_IF(ensure(ctx, ds->UB, (size_t)2 * nb * nb * 8));     // U of two consecutive steps
    RET_IF(ensure(ctx, ds->EB, (size_t)2 * nb * nb * 8));     // early tiles of two consecutive steps
    RET_IF(ensure(ctx, ds->linv, (size_t)linv_bytes(nb)));
    RET_IF(ensure(ctx, ds->red, (size_t)(4 * P + 64) * 8));
    RET_IF(ensure(ctx, ds->wseg, (size_t)std::max<int64_t>(ld, 1) * 8));
    const int64_t slot0 = g.slot_rows(0);
    for (int b = 0; b < 2; ++b) RET_IF(ensure(ctx, ds->PB[b], (size_t)std::max<int64_t>(pr * slot0 * nb, 1) * 8));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemcpyAsync(ds->Xrows.p, xr.data(), xr.size() * 8, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(ds->Zcols.p, zc.data(), zc.size() * 8, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(ds->yloc.p, yl.data(), yl.size() * 8, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_in[3].p, theta, (size_t)nth * 8, cudaMemcpyHostToDevice, cs));
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, 16, cs));
    CUDA_TRY(ctx, cudaStreamSynchronize(cs));   // the host staging vectors are pageable
    const double* dth = (const double*)ctx->d_in[3].p;
    double* A = (double*)ds->Aloc.p;
    sl.oz_planes = ctx->ozaki > 0 ? ctx->ozaki : oz_auto_planes((double)N, theta[d], theta[d + 1], jitter);

    // ---- the local matrix, generated in place: k(rows, columns) for every local tile, then the diagonal term and y
    if (Lr > 0 && Lc > 0) {
        RET_IF(launch_gram(ctx, cs, kind, (const double*)ds->Xrows.p, Lr * nb, (const double*)ds->Zcols.p, Lc * nb, d, dth, 0.0, 0.0, 0, 0, A, ld));
        int64_t gi0 = -1, step = 0, count = 0;
        for (int64_t gi = 0; gi < T; ++gi)
            if (gi % pr == myrow && gi % pc == mycol) {
                if (gi0 < 0)
                    gi0 = gi;
                else if (step == 0)
                    step = gi - gi0;
                ++count;
            }
        if (count > 0) {
            dist_diag_kernel<<<grid_for(count * nb), 256, 0, cs>>>(A, ld, nb, pr, pc, gi0, step ? step : 1, count, dth, d, jitter);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
        }
        const int64_t gy = T + P / nb;                      // tile row holding the y^T right-hand side (global rhs row P)
        if (gy % pr == myrow)
            CUDA_TRY(ctx, cudaMemcpyAsync(A + ((gy / pr) * nb + P % nb) * ld, ds->yloc.p, (size_t)ld * 8, cudaMemcpyDeviceToDevice, cs));
    }
    cudaEvent_t ev_f0 = ctx->slots[0].ev[2], ev_f1 = ctx->slots[0].ev[3];
    CUDA_TRY(ctx, cudaEventRecord(ev_f0, cs));

    // optional phase profile (B2GP_DIST_PROF=1): CUDA events on the streams, summed per phase after the call
    const bool prof = getenv("B2GP_DIST_PROF") != nullptr;
    enum { PH_SLICE, PH_G1, PH_POTRF, PH_UWAIT, PH_PANEL, PH_G2, PH_COMMWAIT, PH_UBCAST, PH_ROWBCAST, PH_ALLGATHER, PH_EARLY, PH_N };
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pev[PH_N];
    auto mark = [&](cudaStream_t s) -> cudaEvent_t {
        if (!prof) return nullptr;
        cudaEvent_t e = ex->pool.get();
        cudaEventRecord(e, s);
        return e;
    };
    auto span = [&](int ph, cudaEvent_t a, cudaEvent_t b) {
        if (prof && a && b) pev[ph].emplace_back(a, b);
    };
    cudaStream_t dq = ds->dq;                         // the diagonal tiles are factored on their own stream, one step ahead
    double* UBs[2] = {(double*)ds->UB.p, (double*)ds->UB.p + nb * nb};
    double* EBs[2] = {(double*)ds->EB.p, (double*)ds->EB.p + nb * nb};
    auto diag_tile = [&](int64_t k) { return A + ((k / pr) * nb) * ld + (k / pc) * nb; };

    // ---- (a), (b) of step k on stream s: the owner folds in the tile (k, k-1) it received ahead of the panel
    // (`early`), factors the diagonal tile and forms U = L_kk^{-T}; every rank of the process column joins the broadcast
    // of U on the communication stream.  ev_ubc[k & 1] marks "U_k is here" for the panel solve.
    auto diag_factor = [&](int64_t k, cudaStream_t s, bool early) -> int {
        const int kr = (int)(k % pr), kc = (int)(k % pc);
        if (mycol != kc) return B2GP_OK;
        double* U = UBs[k & 1];
        if (myrow == kr) {
            cudaEvent_t p0 = mark(s);
            double* D = diag_tile(k);
            if (early) {   // D -= E E^T, E = L tile (k, k-1): from the early broadcast, or in place when this rank solved it itself
                const bool local = (pc == 1);
                const double* E = local ? A + (g.first_row_after(k - 1, myrow) * nb) * ld + ((k - 1) / pc) * nb : EBs[(k - 1) & 1];
                RET_IF(gemm_nt(ctx, s, nb, nb, nb, -1.0, E, local ? ld : nb, E, local ? ld : nb, 1.0, D, ld, true));
            }
            RET_IF(potrf_rec(ctx, s, D, ld, nb, (double*)ds->linv.p, dinfo, k * nb));
            set_identity_kernel<<<grid_for(nb * nb), 256, 0, s>>>(U, nb, nb);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
            RET_IF(trsm_rec(ctx, s, U, nb, nb, D, ld, nb, (const double*)ds->linv.p));      // U = L_kk^{-T}
            span(PH_POTRF, p0, mark(s));
        }
        if (pr > 1) {
            CUDA_TRY(ctx, cudaEventRecord(ds->ev_u, s));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ms, ds->ev_u, 0));
            cudaEvent_t m0 = mark(ms);
            NCCL_TRY(ctx, nc->Broadcast(U, U, (size_t)(nb * nb), ncclDouble, kr, ds->colc, ms));
            span(PH_UBCAST, m0, mark(ms));
            CUDA_TRY(ctx, cudaEventRecord(ds->ev_ubc[k & 1], ms));
        } else {
            CUDA_TRY(ctx, cudaEventRecord(ds->ev_ubc[k & 1], s));
        }
        return B2GP_OK;
    };
    // ---- (c) .. (f) of step k on the compute stream: panel solve once U is here, the tile (k+1, k) sent ahead to the owner
    // of the next diagonal tile (row communicator), pack, panel exchange
    auto panel_front_b = [&](int64_t k) -> int {
        const int kc = (int)(k % pc);
        const int64_t li0 = g.first_row_after(k, myrow), rows = g.panel_rows(k, myrow), slot = g.slot_rows(k);
        double* PBk = (double*)ds->PB[k & 1].p;
        double* rp = A + (li0 * nb) * ld + (k / pc) * nb;
        if (mycol == kc) {
            cudaEvent_t p1 = mark(cs);
            CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ds->ev_ubc[k & 1], 0));
            cudaEvent_t p0 = mark(cs);
            span(PH_UWAIT, p1, p0);
            if (rows > 0) RET_IF(ozaki_dispatch(ctx, cs, rows, nb, nb, 1.0, rp, ld, UBs[k & 1], nb, rp, ld, false, true, true, true));
        }
        if (k + 1 < T && myrow == (int)((k + 1) % pr)) {      // my process row holds tile (k+1, k) -- the first tile of its panel rows
            if (pc > 1) {
                if (mycol == kc) {
                    copy2d_kernel<<<grid_for(nb * nb), 256, 0, cs>>>(EBs[k & 1], nb, rp, ld, nb, nb);
                    CUDA_TRY(ctx, cudaGetLastError());
                    ctx->launches++;
                }
                CUDA_TRY(ctx, cudaEventRecord(ds->ev_e, cs));
                CUDA_TRY(ctx, cudaStreamWaitEvent(ms, ds->ev_e, 0));
                cudaEvent_t m0 = mark(ms);
                NCCL_TRY(ctx, nc->Broadcast(EBs[k & 1], EBs[k & 1], (size_t)(nb * nb), ncclDouble, kc, ds->rowc, ms));
                span(PH_EARLY, m0, mark(ms));
                CUDA_TRY(ctx, cudaEventRecord(ds->ev_early[k & 1], ms));
            } else {
                CUDA_TRY(ctx, cudaEventRecord(ds->ev_early[k & 1], cs));
            }
        }
        if (mycol == kc) {
            cudaEvent_t p0 = mark(cs);
            if (rows > 0) {
                copy2d_kernel<<<grid_for(rows * nb), 256, 0, cs>>>(PBk + (int64_t)myrow * slot * nb, nb, rp, ld, rows, nb);
                CUDA_TRY(ctx, cudaGetLastError());
                ctx->launches++;
            }
            span(PH_PANEL, p0, mark(cs));
        }
        if (slot > 0) {
            CUDA_TRY(ctx, cudaEventRecord(ds->ev_chunk, cs));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ms, ds->ev_chunk, 0));
            double* mine = PBk + (int64_t)myrow * slot * nb;
            cudaEvent_t m0 = mark(ms);
            if (pc > 1) NCCL_TRY(ctx, nc->Broadcast(mine, mine, (size_t)(slot * nb), ncclDouble, kc, ds->rowc, ms));
            cudaEvent_t m1 = mark(ms);
            span(PH_ROWBCAST, m0, m1);
            if (pr > 1) NCCL_TRY(ctx, nc->AllGather(mine, PBk, (size_t)(slot * nb), ncclDouble, ds->colc, ms));
            span(PH_ALLGATHER, m1, mark(ms));
        }
        CUDA_TRY(ctx, cudaEventRecord(ds->ev_comm[k & 1], ms));
        return B2GP_OK;
    };
    // ---- diagonal look-ahead: tile j's last update (from panel j-1) and its factorisation run on the diagonal stream as
    // soon as the tile (j, j-1) has arrived and the compute stream has applied panels < j-1 to it (ev_g2)
    auto diag_ahead = [&](int64_t j) -> int {
        if (j >= T) return B2GP_OK;
        const bool in_col = mycol == (int)(j % pc), owner = in_col && myrow == (int)(j % pr);
        if (!in_col) return B2GP_OK;
        if (owner) {
            CUDA_TRY(ctx, cudaEventRecord(ds->ev_g2, cs));
            CUDA_TRY(ctx, cudaStreamWaitEvent(dq, ds->ev_g2, 0));
            CUDA_TRY(ctx, cudaStreamWaitEvent(dq, ds->ev_early[(j - 1) & 1], 0));
        }
        return diag_factor(j, dq, true);
    };
    OzOperand opA, opB;
    OzMode upd_mode;
    auto update_slices = [&](int64_t k) -> int {   // digit planes of the panel of step k for this process's rows and columns
        const DistStep& s = ds->steps[(size_t)k];
        const int64_t rows = g.panel_rows(k, myrow), cols = (Lc - g.first_col_after(k, mycol)) * nb, slot = g.slot_rows(k);
        if (rows <= 0 || cols <= 0 || s.n1 + s.n2 == 0) return B2GP_OK;
        const double* PBk = (const double*)ds->PB[k & 1].p;
        if (sl.oz_planes == 6) {
            RET_IF(oz_slice_launch<6>(ctx, cs, PBk + (int64_t)myrow * slot * nb, nb, rows, nb, ds->updA, ds->updSA, false, nullptr, &opA));
            RET_IF(oz_slice_launch<6>(ctx, cs, PBk, nb, cols, nb, ds->updB, ds->updSB, false, (const int64_t*)s.bmap.p, &opB));
        } else {
            RET_IF(oz_slice_launch<7>(ctx, cs, PBk + (int64_t)myrow * slot * nb, nb, rows, nb, ds->updA, ds->updSA, false, nullptr, &opA));
            RET_IF(oz_slice_launch<7>(ctx, cs, PBk, nb, cols, nb, ds->updB, ds->updSB, false, (const int64_t*)s.bmap.p, &opB));
        }
        return B2GP_OK;
    };
    auto update_part = [&](int64_t k, int part) -> int {   // (g1) / (g2) of step k
        const DistStep& s = ds->steps[(size_t)k];
        const int64_t cnt = part == 1 ? s.n1 : s.n2;
        if (cnt == 0) return B2GP_OK;
        const int64_t li0 = g.first_row_after(k, myrow), lj0 = g.first_col_after(k, mycol);
        const int64_t rows = g.panel_rows(k, myrow), cols = (Lc - lj0) * nb;
        double* C = A + (li0 * nb) * ld + lj0 * nb;
        const int2* list = (const int2*)(part == 1 ? s.g1.p : s.g2.p);
        if (sl.oz_planes == 6) return oz_mma_launch<6>(ctx, cs, sl.oz, opA, opB, rows, cols, nb, -1.0, C, ld, false, upd_mode, list, cnt);
        return oz_mma_launch<7>(ctx, cs, sl.oz, opA, opB, rows, cols, nb, -1.0, C, ld, false, upd_mode, list, cnt);
    };

    // ---- the factorisation (with the right-hand-side rows riding below).  Compute stream, step k: wait for panel k;
    // slices; g1 (tile column k+1 without its diagonal tile, which the diagonal stream owns); panel solve / early tile /
    // exchange of step k+1; g2; then hand tile k+2 to the diagonal stream.
    // the persistent int8 kernels leave a few SMs to the NCCL kernels and the diagonal-tile kernels running beside them
    const int big_grid_saved = ctx->big_grid;
    if (ctx->big_grid == 0) ctx->big_grid = ctx->sm_count - 16;
    CUDA_TRY(ctx, cudaEventRecord(ds->ev_g2, cs));
    CUDA_TRY(ctx, cudaStreamWaitEvent(dq, ds->ev_g2, 0));          // the diagonal stream starts behind the Gram build
    RET_IF(diag_factor(0, cs, false));
    RET_IF(panel_front_b(0));
    RET_IF(diag_ahead(1));
    for (int64_t k = 0; k < T; ++k) {
        cudaEvent_t q0 = mark(cs);
        CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ds->ev_comm[k & 1], 0));
        cudaEvent_t q1 = mark(cs);
        span(PH_COMMWAIT, q0, q1);
        RET_IF(update_slices(k));
        cudaEvent_t q2 = mark(cs);
        span(PH_SLICE, q1, q2);
        RET_IF(update_part(k, 1));
        span(PH_G1, q2, mark(cs));
        if (k + 1 < T) RET_IF(panel_front_b(k + 1));
        cudaEvent_t q3 = mark(cs);
        RET_IF(update_part(k, 2));
        span(PH_G2, q3, mark(cs));
        RET_IF(diag_ahead(k + 2));
    }
    CUDA_TRY(ctx, cudaEventRecord(ds->ev_u, dq));                   // the compute stream ends behind the diagonal stream
    CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ds->ev_u, 0));
    ctx->big_grid = big_grid_saved;
    CUDA_TRY(ctx, cudaEventRecord(ev_f1, cs));

    // ---- epilogue: mean[p] = <V[p, :], w>, var[p] = k(x,x) + noise_p + jitter - |V[p, :]|^2, summed over the process grid
    double* red = (double*)ds->red.p;        // [0, P): dot, [P, 2P): |V|^2
    CUDA_TRY(ctx, cudaMemsetAsync(red, 0, (size_t)(2 * P) * 8, cs));
    {
        const int64_t gy = T + P / nb;
        double* wseg = (double*)ds->wseg.p;
        if (gy % pr == myrow && Lc > 0)
            CUDA_TRY(ctx, cudaMemcpyAsync(wseg, A + ((gy / pr) * nb + P % nb) * ld, (size_t)ld * 8, cudaMemcpyDeviceToDevice, cs));
        if (pr > 1 && Lc > 0) {
            CUDA_TRY(ctx, cudaEventRecord(ds->ev_u, cs));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ms, ds->ev_u, 0));
            NCCL_TRY(ctx, nc->Broadcast(wseg, wseg, (size_t)ld, ncclDouble, (int)(gy % pr), ds->colc, ms));
            CUDA_TRY(ctx, cudaEventRecord(ds->ev_ubc[0], ms));
            CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ds->ev_ubc[0], 0));
        }
        for (int64_t li = 0; li < Lr && Lc > 0; ++li) {
            const int64_t gi = li * pr + myrow;
            if (gi < T) continue;
            const int64_t p0 = (gi - T) * nb, np = std::min<int64_t>(nb, P - p0);
            if (np <= 0) continue;
            rowdot2_kernel<<<(unsigned)np, RD_THREADS, 0, cs>>>(A + (li * nb) * ld, ld, ld, wseg, 1.0, red + p0, red + P + p0);
            CUDA_TRY(ctx, cudaGetLastError());
            ctx->launches++;
        }
        CUDA_TRY(ctx, cudaEventRecord(ds->ev_chunk, cs));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ms, ds->ev_chunk, 0));
        NCCL_TRY(ctx, nc->AllReduce(red, red, (size_t)(2 * P), ncclDouble, ncclSum, ds->world, ms));
        NCCL_TRY(ctx, nc->AllReduce(dinfo, dinfo, 1, ncclInt, ncclMax, ds->world, ms));
        CUDA_TRY(ctx, cudaEventRecord(ds->ev_done, ms));
        CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ds->ev_done, 0));
        dist_finish_kernel<<<grid_for(P), 256, 0, cs>>>(red, want_var ? red + 2 * P : nullptr, red + P, P, kind, d, dth, noiseless ? 0.0 : 1.0,
                                                         jitter, dinfo);
        CUDA_TRY(ctx, cudaGetLastError());
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(mean, red, (size_t)P * 8, cudaMemcpyDeviceToHost, cs));
    if (want_var) CUDA_TRY(ctx, cudaMemcpyAsync(var, red + 2 * P, (size_t)P * 8, cudaMemcpyDeviceToHost, cs));
    CUDA_TRY(ctx, cudaMemcpyAsync(info, dinfo, sizeof(int), cudaMemcpyDeviceToHost, cs));
    RET_IF(tm.end(cs, nullptr));
    sl.oz_planes = 7;
    float ms_f = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&ms_f, ev_f0, ev_f1));
    ex->last.potrf_ms = ms_f;
    if (prof) {
        static const char* names[PH_N] = {"update slices", "g1 (next tile column)", "[diag stream] early update + potrf + U", "wait for U", "panel solve + pack",
                                          "g2 (rest of the update)", "wait for the panel exchange", "[comm stream] U broadcast",
                                          "[comm stream] panel row broadcast", "[comm stream] panel column all-gather",
                                          "[comm stream] early tile broadcast"};
        char line[2048];
        int off = snprintf(line, sizeof line, "[dist prof] rank %d (%d,%d) N=%lld nb=%lld factorisation %.2f ms:", ds->rank, myrow, mycol,
                           (long long)N, (long long)nb, ms_f);
        for (int ph = 0; ph < PH_N; ++ph) {
            double tot = 0.0;
            for (auto& pe : pev[ph]) {
                float f = 0.f;
                cudaEventElapsedTime(&f, pe.first, pe.second);
                tot += f;
            }
            if (off < (int)sizeof line) off += snprintf(line + off, sizeof line - off, " %s %.2f;", names[ph], tot);
        }
        fprintf(stderr, "%s\n", line);     // one write per rank: the ranks share a terminal
    }
    const double n = (double)N, p = (double)P;
    ex->last.flops = n * n * n / 3.0 + n * n * (p + 1.0) + 4.0 * n * p;
    if (timing) *timing = ex->last;
    return B2GP_OK;
}

// N-sharded sparse (Nystrom / VFE) posterior -- gpax/models/sparse_gp.py:173-223 with the training set split over the
// ranks (config 5).  COLLECTIVE, HOST pointers: every rank passes ITS shard (Xtr_shard[N_shard, d], y_shard) and the same
// Xu, X_new, theta.  Per rank: Luu, W = Luu^{-1} K(Xu, shard) and the statistics W W^T / noise, W y / noise
// (sparse_gp.py:193-199, 203-204 restricted to the shard); ONE in-library NCCL all-reduce of the M x M matrix and the
// M-vector; then the M x M Cholesky and the P-side solves replicated on every rank (sparse_gp.py:200-217).
extern "C" int b2gp_dist_sparse_posterior(b2gp_ctx* ctx, int kind, const double* Xu, int64_t M, const double* Xtr_shard,
                                          int64_t N_shard, const double* y_shard, const double* Xnew, int64_t P, int d,
                                          const double* theta, int noiseless, double jitter, unsigned flags, double* mean,
                                          double* var, int* info, b2gp_timing* timing) {
    if (!ctx) return B2GP_ERR_ARG;
    Extra* ex = extra_of(ctx);
    DistState* ds = ex->dist;
    if (!ds || !ds->ready) return set_err(ctx, B2GP_ERR_ARG, "b2gp_dist_sparse_posterior", "call b2gp_dist_init first", __FILE__, __LINE__);
    ARG_CHECK(ctx, kind >= 0 && kind <= 2);
    ARG_CHECK(ctx, Xu && Xtr_shard && y_shard && Xnew && theta && mean && info);
    ARG_CHECK(ctx, M >= 1 && N_shard >= 1 && P >= 1 && d >= 1 && d <= GRAM_MAX_D);
    ARG_CHECK(ctx, !(flags & B2GP_FLAG_DEVICE_PTRS) && !(flags & (B2GP_OUT_COV | B2GP_OUT_SAMPLE)));
    const bool want_var = flags & B2GP_OUT_VAR;
    ARG_CHECK(ctx, !want_var || var);
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    NcclApi* nc = nccl_api();
    ex->fcache.valid = false;
    Slot& sl = ctx->slots[0];
    cudaStream_t st = sl.stream, ms = ds->ms;
    CallTimer tm(ctx);
    RET_IF(tm.begin(st));
    const int nth = d + 3;
    const double *dXu, *dXtr, *dy, *dXnew, *dth;
    RET_IF(stage_in(ctx, st, ctx->d_in[0], Xtr_shard, (size_t)N_shard * d * 8, false, &dXtr));
    RET_IF(stage_in(ctx, st, ctx->d_in[1], y_shard, (size_t)N_shard * 8, false, &dy));
    RET_IF(stage_in(ctx, st, ctx->d_in[2], Xnew, (size_t)P * d * 8, false, &dXnew));
    RET_IF(stage_in(ctx, st, ctx->d_in[3], theta, (size_t)nth * 8, false, &dth));
    RET_IF(stage_in(ctx, st, ctx->d_in[5], Xu, (size_t)M * d * 8, false, &dXu));
    const int64_t ldM = round_up(M, 8);
    RET_IF(ensure(ctx, sl.A, (size_t)(2 * M * ldM + ldM) * 8));      // Luu | K (+ the M-vector right behind it: one all-reduce)
    RET_IF(ensure(ctx, sl.Linv, (size_t)2 * linv_bytes(M)));
    RET_IF(ensure(ctx, ctx->d_out[0], (size_t)(2 * P + 16) * 8));
    RET_IF(ensure(ctx, ctx->d_info, 64));
    int* dinfo = (int*)ctx->d_info.p;
    CUDA_TRY(ctx, cudaMemsetAsync(dinfo, 0, 16, st));
    double* Luu = (double*)sl.A.p;
    double* Kmat = Luu + M * ldM;
    double* cvec = Kmat + M * ldM;
    double* LinvU = (double*)sl.Linv.p;
    double* LinvK = LinvU + linv_bytes(M) / 8;
    double* mv = (double*)ctx->d_out[0].p;
    double* vv = mv + P;
    cudaEvent_t e0 = sl.ev[2], e1 = sl.ev[3], e2 = sl.ev[4];
    CUDA_TRY(ctx, cudaEventRecord(e0, st));
    RET_IF(sparse_partial_dev(ctx, sl, kind, dXu, M, dXtr, N_shard, dy, d, dth, jitter, theta[d + 1], Luu, ldM, LinvU, Kmat, ldM, cvec, dinfo));
    CUDA_TRY(ctx, cudaEventRecord(e1, st));
    if (ds->nranks > 1) {
        CUDA_TRY(ctx, cudaEventRecord(ds->ev_chunk, st));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ms, ds->ev_chunk, 0));
        NCCL_TRY(ctx, nc->AllReduce(Kmat, Kmat, (size_t)(M * ldM + M), ncclDouble, ncclSum, ds->world, ms));   // sparse_gp.py:199, 204 summed over shards
        NCCL_TRY(ctx, nc->AllReduce(dinfo, dinfo, 1, ncclInt, ncclMax, ds->world, ms));
        CUDA_TRY(ctx, cudaEventRecord(ds->ev_done, ms));
        CUDA_TRY(ctx, cudaStreamWaitEvent(st, ds->ev_done, 0));
    }
    CUDA_TRY(ctx, cudaEventRecord(e2, st));
    RET_IF(sparse_finish_dev(ctx, sl, kind, dXu, M, Luu, ldM, LinvU, Kmat, ldM, LinvK, cvec, dXnew, P, d, dth, noiseless, jitter, want_var,
                             false, mv, vv, nullptr, P, dinfo));
    int hinfo[2] = {0, 0};
    CUDA_TRY(ctx, cudaMemcpyAsync(hinfo, dinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(mean, mv, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    if (want_var) CUDA_TRY(ctx, cudaMemcpyAsync(var, vv, (size_t)P * 8, cudaMemcpyDeviceToHost, st));
    RET_IF(tm.end(st, nullptr));
    info[0] = hinfo[0] != 0 ? hinfo[0] : -hinfo[1];
    float f = 0.f;
    CUDA_TRY(ctx, cudaEventElapsedTime(&f, e0, e1));
    ex->last.potrf_ms = f;          // per-rank statistics (Gram, Luu, W, W W^T)
    CUDA_TRY(ctx, cudaEventElapsedTime(&f, e1, e2));
    ex->last.trsm_ms = f;           // the all-reduce
    const double m = (double)M, n = (double)N_shard * ds->nranks, p = (double)P;
    ex->last.flops = 2.0 * m * m * m / 3.0 + 2.0 * m * m * n + 2.0 * m * m * (p + 1.0);
    if (timing) *timing = ex->last;
    return B2GP_OK;
}
