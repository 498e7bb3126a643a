// ozaki.cuh -- fp64 GEMM/SYRK on the int8 tcgen05 tensor cores by integer splitting (Ozaki scheme).
//
// sm_100a has no fp64 tcgen05.mma; what it has is kind::i8 with exact int32 accumulation in TMEM at ~16x the
// DMMA rate.  C[m,n] += alpha * A[m,k] B[n,k]^T is evaluated as follows.
//   1. slice (oz_slice_kernel): every row of A (and B) is scaled by a power of two 2^-e_i so that |a| < 1, rounded ONCE
//      to the integer I = rint(a 2^(8S-2-e_i)) (|I| <= 2^(8S-2)) and cut into S base-256 digits in two's-complement
//      style, lowest first: d = signed low byte of I in [-128, 127], I <- (I - d) >> 8; the leading digit is what is left
//      (6 bits + a carry).  Full-range int8 digits: 8 bits per plane instead of the 7 of a round-to-nearest base-128 split,
//      so S = 7 planes carry 54 bits (what 8 planes did before) and 28 instead of 36 products; S = 6 carry 46 bits.
//         a = 2^e_i * sum_q d_q 2^-w(q),   w(q) = 6 + 8 q,  q = 0 .. S-1   (exact to half a unit of the last digit)
//      The digits are stored as S int8 planes [S][rows][k], K-major.
//   2. products (oz_mma_kernel): A_p B_q^T is an int8 GEMM; its int32 result is EXACT (|d d'| <= 2^14, at most S pairs
//      per accumulator, k <= OZ_K_MAX = 16384 per launch: 7 * 2^14 * 2^14 < 2^31; longer k is split by the dispatcher).
//      Its weight 2^-(12 + 8 (p+q)) depends on t = p+q only, so all pairs of one weight class accumulate into the
//      same TMEM accumulator.  Classes t >= S are below the target precision and are dropped: S(S+1)/2 products.
//   3. recombine (epilogue of the same kernel): acc = sum_t 2^-(12+8t) D_t in fp64, smallest class first, and
//      C[i,j] += alpha * 2^(e_i + f_j) * acc.
// Kernel organisation (one CTA per SM, persistent):
//   * output tile 128 x 64: S accumulators of 64 TMEM columns = 512 columns for S = 8 (all of TMEM);
//   * a pipeline stage is one 32-byte k-block (one MMA K step) of ALL planes: S boxes [128 rows x 32 B] of A and S
//     boxes [64 rows x 32 B] of B (TMA, 32-byte swizzle), 4 stages in flight -- each plane is fetched once per
//     k-block and used by every pair it takes part in, which keeps L2 traffic per fp64-equivalent flop at the DMMA
//     kernel's level;
//   * warp 0 lane 0: TMA producer; warp 1 lane 0: MMA issuer.  The S(S+1)/2 products of a stage are issued as
//     12 (S = 8) wide tcgen05.mma: A plane p against the stacked B planes 0..S-1-p (M = 128, N up to 256, K = 32),
//     landing in the adjacent accumulators p..S-1; then commit to the stage's `empty` mbarrier;
//   * warps 2-5: epilogue (tcgen05.ld 32x32b, one TMEM lane = one row each).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_tma.cuh"

constexpr int OZ_BM = 128, OZ_BN = 64, OZ_KB = 32;   // 32-byte k-blocks (one MMA K step)
// stages in flight: as many as fit beside the epilogue's staging tile -- 5 x 36 KB with 6 planes, 4 x 42 KB with 7
__host__ __device__ constexpr int oz_stages_for(int S) { return S <= 6 ? 5 : 4; }
constexpr int OZ_THREADS = 192;
constexpr int OZ_K_MAX = 16384;           // int32 accumulation bound of one launch (see above)
constexpr int OZ_A_TILE = OZ_BM * OZ_KB;  // 4096 B
constexpr int OZ_B_TILE = OZ_BN * OZ_KB;  // 2048 B

struct OzArgs {
    int m, n, kb_count;        // kb_count = padded k / OZ_KB
    int rowsA_pad, rowsB_pad;  // rows of one digit plane (multiple of 128)
    const double* sa;          // 2^e_i per row of A
    const double* sb;          // 2^f_j per row of B
    double* C;
    int64_t ldc;
    double alpha;
    int lower_only;
    int tiles_m, tiles_n;
    int num_tiles;
    const int2* tile_list;  // [num_tiles] (ti, tj) in execution order (host-built, L2-friendly)
    int overwrite;  // 1: C = alpha A B^T (C is not read: beta = 0; C may alias the fp64 source of A, which was copied to planes)
    int ktri;       // 1: B is lower triangular (B[c][k] = 0 for k > c): a column tile needs only the k-blocks up to its last row
    int debug;  // 0 normal; 1 skip the global read-modify-write; 2 also skip staging barriers (timing experiments only)
    long long* prof;  // optional [gridDim.x][8]: epilogue cycles (wait acc_full, whole, C update), tiles, MMA issuer (wait TMA, wait drain, total)
};

// ---------------------------------------------------------------------------------------------- slicing
// one CTA per row: row maximum -> exponent, then S digits per element; planes[p][row][kk]
// trans != 0: the operand is the TRANSPOSE of the stored matrix, element (row, kk) = A[kk * lda + row] (small operands only:
// the reads are strided)
template <int S>
__global__ void __launch_bounds__(256) oz_slice_kernel(const double* __restrict__ A, int64_t lda, int rows, int k, int8_t* __restrict__ planes,
                                                       int rows_pad, int kpad, double* __restrict__ scale, int trans,
                                                       const int64_t* __restrict__ rowmap128) {
    __shared__ double red[8];
    const int row = blockIdx.x;
    const int tid = threadIdx.x;
    const int64_t sr = trans ? 1 : lda, sk = trans ? lda : 1;   // strides of (row, kk)
    if (rowmap128) A += (rowmap128[row >> 7] - (int64_t)(row & ~127)) * lda;   // source row = rowmap128[row / 128] + row % 128
    int e = 0;
    if (row < rows) {
        double mx = 0.0;
        for (int kk = tid; kk < k; kk += 256) mx = fmax(mx, fabs(A[(int64_t)row * sr + (int64_t)kk * sk]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((tid & 31) == 0) red[tid >> 5] = mx;
        __syncthreads();
        mx = red[0];
        for (int w = 1; w < 8; ++w) mx = fmax(mx, red[w]);
        if (mx > 0.0 && mx < 1e300) frexp(mx, &e);  // mx = f * 2^e, f in [0.5, 1)
        if (e < -900) e = -900;                      // 2^(8S-2-e) below must stay finite
        if (tid == 0) scale[row] = ldexp(1.0, e);
    }
    const double up = ldexp(1.0, 8 * S - 2 - e);     // exact power of two: a * up is the exact scaling
    // 4 consecutive k per thread -> one 32-bit store per plane
    for (int k4 = tid * 4; k4 < kpad; k4 += 1024) {
        int packed[S];
#pragma unroll
        for (int p = 0; p < S; ++p) packed[p] = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = k4 + j;
            long long I = (row < rows && kk < k) ? __double2ll_rn(A[(int64_t)row * sr + (int64_t)kk * sk] * up) : 0ll;
#pragma unroll
            for (int p = S - 1; p > 0; --p) {
                const int d = (int)(signed char)(I & 0xff);          // signed low byte
                packed[p] |= (d & 0xff) << (8 * j);
                I = (I - d) >> 8;
            }
            packed[0] |= ((int)I & 0xff) << (8 * j);
        }
#pragma unroll
        for (int p = 0; p < S; ++p)
            *reinterpret_cast<int*>(planes + ((int64_t)p * rows_pad + row) * kpad + k4) = packed[p];
    }
}

// ---------------------------------------------------------------------------------------------- tcgen05 helpers
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// commit that arrives on the barrier at the same shared-memory offset in every CTA of `mask` (cluster of 2)
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
                 : "memory");
}
// TMA box load delivered to the same shared-memory offset (and signalling the same barrier offset) in every CTA of `mask`
__device__ __forceinline__ void tma_load_2d_mc(uint32_t smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32, M = 128, N = 64, K = 32
__device__ __forceinline__ void tc_mma_i8(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(taddr),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor, K-major, 32-byte swizzle (rows of 32 B): 8-row groups are 256 B apart
__device__ __forceinline__ uint64_t oz_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);   // start address
    d |= (uint64_t)0 << 16;                   // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(256 >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                   // descriptor version (sm_100)
    d |= (uint64_t)6 << 61;                   // SWIZZLE_32B
    return d;
}
// exact int32 -> double without I2F.F64 (measured: the conversion instruction runs at ~1 per 5 cycles per SM on
// this part and was the whole epilogue): 2^52 + 2^31 + v is assembled in the mantissa, the constant subtracted.
__device__ __forceinline__ double i32_to_f64(int v) {
    return __hiloint2double(0x43300000, v ^ 0x80000000) - 4503601774854144.0;
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 32 consecutive accumulator columns of this thread's TMEM lane.  The registers are valid only after tc_wait_ld32,
// which names them as operands so that no use can be scheduled above the wait.
__device__ __forceinline__ void tc_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld32(int (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]),
                   "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]),
                   "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]),
                   "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// k-blocks a tile of column pair / column tile `ty` runs over: all of them, or -- B lower triangular -- those up to
// the last row of B the tile (the CTA pair's tiles, which share the A stages) touches
__device__ __forceinline__ int oz_kb_count(const OzArgs& p, int CL, int ty) {
    if (!p.ktri) return p.kb_count;
    const int lim = ((CL * ty + CL) * OZ_BN + OZ_KB - 1) / OZ_KB;
    return lim < p.kb_count ? lim : p.kb_count;
}

__device__ __forceinline__ void oz_tile_of(int x, int lower_only, int tiles_n, int& ti, int& tj) {
    if (lower_only) {  // 128 x 64 tiles of the lower triangle: row block ti owns column tiles 0 .. 2 ti + 1
        int t = (int)((sqrt(4.0 * (double)x + 1.0) - 1.0) * 0.5);
        while ((int64_t)(t + 1) * (t + 2) <= x) ++t;
        while ((int64_t)t * (t + 1) > x) --t;
        ti = t;
        tj = x - t * (t + 1);
    } else {
        ti = x / tiles_n;
        tj = x % tiles_n;
    }
}

// CL = 2: the two CTAs of a cluster work on column tiles 2 j and 2 j + 1 of the same row block.  They need the same A
// digit planes, so each CTA fetches half of them and TMA multicasts every box into both shared memories (the L2 -> SM
// operand traffic, which is what bounds this kernel, drops from 48 to 32 KB per CTA and k-block).  A stage may be
// refilled only when BOTH CTAs have consumed it: the `empty` barriers count two commits, one multicast from each.
template <int S, int CL>
__global__ void __launch_bounds__(OZ_THREADS, 1)
oz_mma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const OzArgs p) {
    const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u;
    const int tile_first = (int)blockIdx.x / CL, tile_step = (int)gridDim.x / CL;
    constexpr int STAGE_BYTES = S * (OZ_A_TILE + OZ_B_TILE);
    constexpr int OZ_STAGES = oz_stages_for(S);
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bars = base + OZ_STAGES * STAGE_BYTES;  // full[2], empty[2], acc_full, acc_empty
    const uint32_t tmem_slot = bars + 16 * OZ_STAGES + 16;
    uint8_t* sgen = smem_raw + (base - raw);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + OZ_STAGES * STAGE_BYTES + 16 * OZ_STAGES + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t full0 = bars, empty0 = bars + 8 * OZ_STAGES, acc_full = bars + 16 * OZ_STAGES, acc_empty = bars + 16 * OZ_STAGES + 8;

    if (tid == 0) {
        for (int s = 0; s < OZ_STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, CL);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();   // the peer's barriers exist before anything is multicast at them
    tc_fence_after();
    const uint32_t tmem = *tmem_slot_gen;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = tile_first; tile < p.num_tiles; tile += tile_step) {
                const int2 tt = p.tile_list[tile];
                const int row0 = tt.x * OZ_BM, col0 = (CL * tt.y + (int)crank) * OZ_BN;
                const int kbn = oz_kb_count(p, CL, tt.y);
                for (int kb = 0; kb < kbn; ++kb, ++it) {
                    const uint32_t s = it % OZ_STAGES, ph = (it / OZ_STAGES) & 1u;
                    mbar_wait(empty0 + 8 * s, ph ^ 1u);
                    mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);   // bytes landing in THIS CTA's stage, whoever fetches them
                    const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll
                    for (int q = 0; q < S; ++q) {
                        if (CL == 1)
                            tma_load_2d(st + q * OZ_A_TILE, &mapA, kb * OZ_KB, q * p.rowsA_pad + row0, full0 + 8 * s);
                        else if ((q < (S + 1) / 2) == (crank == 0))
                            tma_load_2d_mc(st + q * OZ_A_TILE, &mapA, kb * OZ_KB, q * p.rowsA_pad + row0, full0 + 8 * s, (uint16_t)3);
                        tma_load_2d(st + S * OZ_A_TILE + q * OZ_B_TILE, &mapB, kb * OZ_KB, q * p.rowsB_pad + col0, full0 + 8 * s);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            uint32_t it = 0, tcount = 0;
            long long m_full = 0, m_empty = 0;
            const long long m_t0 = clock64();
            for (int tile = tile_first; tile < p.num_tiles; tile += tile_step, ++tcount) {
                const long long e0 = clock64();
                mbar_wait(acc_empty, (tcount & 1u) ^ 1u);   // epilogue of the previous tile has drained TMEM
                m_empty += clock64() - e0;
                tc_fence_after();
                const int kbn = oz_kb_count(p, CL, p.tile_list[tile].y);
                for (int kb = 0; kb < kbn; ++kb, ++it) {
                    const uint32_t s = it % OZ_STAGES, ph = (it / OZ_STAGES) & 1u;
                    const long long f0 = p.prof ? clock64() : 0;
                    mbar_wait(full0 + 8 * s, ph);
                    if (p.prof) m_full += clock64() - f0;
                    tc_fence_after();
                    const uint32_t st = base + s * STAGE_BYTES;
                    // Digit plane pp of A meets planes q = 0 .. S-1-pp of B, and the product (pp, q) belongs to weight class
                    // pp + q = accumulator columns 64 (pp + q).  The B planes of a stage are contiguous in shared memory
                    // (a [64 S] x 32 B K-major operand) and so are the accumulators, so ONE MMA with N = 64 (S - pp) does
                    // all of plane pp's products (two when N > 256).  Besides cutting the instruction count 3x this
                    // reads the 4 KB A plane from shared memory once per 2-4 products instead of once per product:
                    // with N = 64 per MMA the operand reads (6 KB / 32 cycles) exceed the 128 B/clk of shared memory.
#pragma unroll
                    for (int pp = 0; pp < S; ++pp) {
                        const uint64_t ad = oz_smem_desc(st + pp * OZ_A_TILE);
                        const uint32_t acc = (kb > 0 || pp > 0) ? 1u : 0u;   // plane 0 touches every class first
                        constexpr uint32_t IDESC_BASE = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(OZ_BM >> 4) << 24);
                        const int nq = S - pp;                                // number of B planes
                        const int n1 = nq > 4 ? 4 : nq;
                        tc_mma_i8(tmem + 64u * pp, ad, oz_smem_desc(st + S * OZ_A_TILE), IDESC_BASE | ((uint32_t)(n1 * OZ_BN >> 3) << 17), acc);
                        if (nq > 4)
                            tc_mma_i8(tmem + 64u * (pp + 4), ad, oz_smem_desc(st + S * OZ_A_TILE + 4 * OZ_B_TILE),
                                      IDESC_BASE | ((uint32_t)((nq - 4) * OZ_BN >> 3) << 17), acc);
                    }
                    if (CL == 2)
                        tc_commit_mc(empty0 + 8 * s, (uint16_t)3);
                    else
                        tc_commit(empty0 + 8 * s);
                }
                tc_commit(acc_full);
            }
            if (p.prof) {
                p.prof[8 * blockIdx.x + 4] = m_full;               // MMA issuer: waiting for TMA data
                p.prof[8 * blockIdx.x + 5] = m_empty;              // ... for the epilogue to drain TMEM
                p.prof[8 * blockIdx.x + 6] = clock64() - m_t0;     // ... whole loop
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue: warps 2..5, TMEM lane quarter = warp % 4
        const int quarter = warp & 3;
        uint32_t tcount = 0;
        long long c_wait = 0, c_ld = 0, c_upd = 0;
        for (int tile = tile_first; tile < p.num_tiles; tile += tile_step, ++tcount) {
            const int2 tt = p.tile_list[tile];
            const int ti = tt.x, tj = CL * tt.y + (int)crank;
            const int row = ti * OZ_BM + 32 * quarter + lane;
            const int col0 = tj * OZ_BN;
            // coalesced mapping of the C update: 16 consecutive threads cover the 128 bytes of one row of a 16-column chunk
            const int et = tid - 64;                      // 0..127 among the epilogue threads
            const int cc = et & 15;
            const long long t0 = clock64();
            mbar_wait(acc_full, tcount & 1u);
            tc_fence_after();
            const long long t1 = clock64();
            c_wait += t1 - t0;
            // pull the tile of C towards L2 now: the drain below takes longer than an HBM round trip
            if (p.debug == 0 && !p.overwrite && (cc == 0 || cc == 15)) {
#pragma unroll
                for (int c16 = 0; c16 < 4; ++c16) {
                    const int gc = col0 + 16 * c16 + cc;
#pragma unroll
                    for (int it = 0; it < 16; ++it) {
                        const int gr = ti * OZ_BM + it * 8 + (et >> 4);
                        if (gr < p.m && gc < p.n && !(p.lower_only && col0 + 16 * c16 > gr)) prefetch_l2(p.C + (int64_t)gr * p.ldc + gc);
                    }
                }
            }
            // staging area for one 128 x 16 chunk (stride 17 doubles) + the row scales of this tile
            double* stg = reinterpret_cast<double*>(sgen + OZ_STAGES * STAGE_BYTES + 128);
            double* sa_s = stg + 128 * 17;
            const int lrow = 32 * quarter + lane;         // row of the tile this thread drains from TMEM
            asm volatile("bar.sync 1, 128;" ::: "memory");   // the previous tile's last chunk is done with sa_s / stg
            sa_s[lrow] = (row < p.m) ? p.sa[row] * p.alpha : 0.0;
            // ---- drain: all S accumulators of this row, smallest weight class first, 32 columns per tcgen05.ld; the
            // load of the next block is in flight while the current one is converted and accumulated
            double acc[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = 0.0;
            {
                int v[2][32];
                const uint32_t trow = tmem + ((uint32_t)(32 * quarter) << 16);
                tc_ld32(trow + 64u * (S - 1), v[0]);
                tc_wait_ld32(v[0]);
#pragma unroll
                for (int nb = 0; nb < 2 * S; ++nb) {
                    const int t = S - 1 - (nb >> 1), h = nb & 1;
                    if (nb + 1 < 2 * S) tc_ld32(trow + 64u * (S - 1 - ((nb + 1) >> 1)) + 32u * ((nb + 1) & 1), v[(nb + 1) & 1]);
                    const double w = __longlong_as_double((long long)(1023 - (12 + 8 * t)) << 52);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[32 * h + j] = fma(i32_to_f64(v[nb & 1][j]), w, acc[32 * h + j]);
                    if (nb + 1 < 2 * S) tc_wait_ld32(v[(nb + 1) & 1]);
                }
            }
            // TMEM is drained: the next tile's MMAs run while this one is written out
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            const long long t2 = clock64();
#pragma unroll
            for (int c16 = 0; c16 < 4; ++c16) {
                if (c16 > 0 && p.debug < 2) asm volatile("bar.sync 1, 128;" ::: "memory");   // previous chunk's readers are done with stg
#pragma unroll
                for (int j = 0; j < 16; ++j) stg[lrow * 17 + j] = acc[16 * c16 + j];
                if (p.debug < 2) asm volatile("bar.sync 1, 128;" ::: "memory");
                const int gc = col0 + 16 * c16 + cc;
                const double sbv = (gc < p.n) ? p.sb[gc] : 0.0;
                if (p.debug == 0) {
                    double oldv[16];
#pragma unroll
                    for (int it = 0; it < 16; ++it) {
                        const int gr = ti * OZ_BM + it * 8 + (et >> 4);
                        oldv[it] = (!p.overwrite && gr < p.m && gc < p.n && !(p.lower_only && gc > gr)) ? p.C[(int64_t)gr * p.ldc + gc] : 0.0;
                    }
#pragma unroll
                    for (int it = 0; it < 16; ++it) {
                        const int lr = it * 8 + (et >> 4);
                        const int gr = ti * OZ_BM + lr;
                        if (gr < p.m && gc < p.n && !(p.lower_only && gc > gr))
                            p.C[(int64_t)gr * p.ldc + gc] = fma(sa_s[lr] * sbv, stg[lr * 17 + cc], oldv[it]);
                    }
                }
            }
            c_upd += clock64() - t2;
            c_ld += clock64() - t1;
        }
        if (p.prof && warp == 2 && lane == 0) {
            p.prof[8 * blockIdx.x + 0] = c_wait;
            p.prof[8 * blockIdx.x + 1] = c_ld;     // whole epilogue (drain + combine + C update)
            p.prof[8 * blockIdx.x + 2] = c_upd;    // of which: C update
            p.prof[8 * blockIdx.x + 3] = tcount;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();   // no commit of mine may still be on its way to a peer that has exited
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------- peak probe
// The int8 tensor-pipe ceiling, measured instead of assumed: every CTA (one per SM) issues `iters` back-to-back
// tcgen05.mma kind::i8 M128 N256 K32 on operands that stay in shared memory (no TMA, no epilogue), alternating between
// the two halves of TMEM, then one commit.  ops = 2 * 128 * 256 * 32 per instruction.  What oz_mma_kernel can reach at
// most; its roofline denominator in bench.py.
__global__ void __launch_bounds__(128, 1) oz_i8_peak_kernel(int iters, int* sink) {
    extern __shared__ uint8_t pk_raw[];
    const uint32_t raw = (uint32_t)__cvta_generic_to_shared(pk_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    const uint32_t bar = base + 4096 + 8192;
    const uint32_t slot = bar + 16;
    uint8_t* gen = pk_raw + (base - raw);
    for (int i = threadIdx.x; i < (4096 + 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gen)[i] = 0x01010101u * (uint32_t)(i & 3);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the operands -> async proxy (MMA)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + 4096 + 8192 + 16);
    if (threadIdx.x == 0) {
        constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24) | ((uint32_t)(256 >> 3) << 17);
        const uint64_t ad = oz_smem_desc(base), bd = oz_smem_desc(base + 4096);
        for (int i = 0; i < iters; ++i) tc_mma_i8(tmem + 256u * (uint32_t)(i & 1), ad, bd, IDESC, i >= 2 ? 1u : 0u);
        tc_commit(bar);
        mbar_wait(bar, 0);
        tc_fence_after();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        int v[16];
        tc_ld16(tmem + ((uint32_t)0 << 16), v);
        tc_wait_ld();
        if (sink && v[0] == 0x7fffffff) sink[0] = v[1];   // keeps the accumulators observable
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------- host side
static bool make_tmap_u8(CUtensorMap* map, const int8_t* base, int64_t rows_total, int64_t kpad, int box_rows) {
    PFN_tmapEncodeTiled enc = tmap_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows_total};
    cuuint64_t gstride[1] = {(cuuint64_t)kpad};
    cuuint32_t box[2] = {(cuuint32_t)OZ_KB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// C[m,n] += alpha * A[m,k] B[n,k]^T  (lower_only: j <= i, A == B allowed) through the int8 tensor cores
// Modes (OzMode): overwrite -- C = alpha A B^T, C not read (it may alias A's fp64 storage: the operands are copied to digit
// planes first); transB -- B is given transposed (element (j, kk) at B[kk * ldb + j]); ktri -- B is lower triangular.
// lower_only with m > n is the "trapezoid" of a factorisation with right-hand-side rows riding below the matrix: rows
// < n update j <= i only, the rows below update all n columns.
struct OzMode {
    bool overwrite = false, transB = false, ktri = false;
};
// One sliced operand: S digit planes [S][rows_pad][kpad] + the row scales.
struct OzOperand {
    const int8_t* planes = nullptr;
    const double* scale = nullptr;
    int64_t rows_pad = 0;
};

// slice `rows` rows of length k (source row r at src + rowmap128[r / 128] * ld + (r % 128) * ld when a row map is given -- a
// device array with one source row index per group of 128 rows, for operands gathered from block-cyclic storage --
// else at src + r * ld) into `planes` / `scale` (capacity checked by the caller)
template <int S>
static int oz_slice_launch(b2gp_ctx* ctx, cudaStream_t st, const double* src, int64_t ld, int64_t rows, int64_t k, DevBuf& planes,
                           DevBuf& scale, bool trans, const int64_t* rowmap128, OzOperand* out) {
    const int64_t kpad = round_up(k, OZ_KB), rp = round_up(rows, 128);
    RET_IF(ensure(ctx, planes, (size_t)S * rp * kpad));
    RET_IF(ensure(ctx, scale, (size_t)rp * 8));
    oz_slice_kernel<S><<<(unsigned)rp, 256, 0, st>>>(src, ld, (int)rows, (int)k, (int8_t*)planes.p, (int)rp, (int)kpad, (double*)scale.p,
                                                     trans ? 1 : 0, rowmap128);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    out->planes = (const int8_t*)planes.p;
    out->scale = (const double*)scale.p;
    out->rows_pad = rp;
    return B2GP_OK;
}

// the band-ordered tile list of an m x n (lower-only / trapezoid) product, cached per shape in `w`
static int oz_default_list(b2gp_ctx* ctx, cudaStream_t st, OzWork& w, int tiles_m, int tiles_n, bool lower_only, int CL,
                           const int2** list, int64_t* count) {
    // Tile order.  One round of the persistent loop runs sm_count consecutive list entries concurrently and, tiles
    // being equally long, in k-lockstep: if those tiles form a compact block of (row block, column block) pairs, each
    // operand block is fetched from HBM once per round and served to the other tiles from L2.  Bands of G row blocks,
    // column-major inside a band: a round covers ~G x (sm_count/G) tiles = G + sm_count/G distinct operand blocks.
    const int pairs_n = (tiles_n + CL - 1) / CL;   // list entries per row block: column tiles (CL = 1) or pairs of them
    OzTileList* tl = nullptr;
    for (auto& l : w.lists)
        if (l.tm == tiles_m && l.tn == tiles_n && l.lower == (lower_only ? 1 : 0) && l.cl == CL) tl = &l;
    if (!tl) {
        // one factorisation cycles through ~2 N / panel distinct shapes; they repeat from draw to draw
        tl = &w.lists[w.next_list];
        w.next_list = (w.next_list + 1) % OZ_LISTS;
        tl->host.clear();
        const int G = 8;
        for (int b0 = 0; b0 < tiles_m; b0 += G) {
            const int b1 = b0 + G < tiles_m ? b0 + G : tiles_m;
            // lower: row block ti owns column tiles 0 .. 2 ti + 1, i.e. pairs 0 .. ti
            const int last = lower_only ? (CL == 2 ? b1 - 1 : 2 * (b1 - 1) + 1) : pairs_n - 1;
            const int jmax = last < pairs_n - 1 ? last : pairs_n - 1;
            for (int tj = 0; tj <= jmax; ++tj)
                for (int ti = b0; ti < b1; ++ti)
                    if (!lower_only || tj <= (CL == 2 ? ti : 2 * ti + 1)) tl->host.push_back(make_int2(ti, tj));
        }
        tl->tm = tiles_m;
        tl->tn = tiles_n;
        tl->lower = lower_only ? 1 : 0;
        tl->cl = CL;
        tl->count = (int64_t)tl->host.size();
        // a list may still be in use by a kernel queued earlier on this stream: the copy is stream-ordered behind it
        RET_IF(ensure(ctx, tl->dev, tl->host.size() * sizeof(int2)));
        CUDA_TRY(ctx, cudaMemcpyAsync(tl->dev.p, tl->host.data(), tl->host.size() * sizeof(int2), cudaMemcpyHostToDevice, st));
    }
    *list = (const int2*)tl->dev.p;
    *count = tl->count;
    return B2GP_OK;
}

// C[m,n] (+)= alpha A B^T from sliced operands.  `list` (device, `count` entries of (row tile, column pair / tile)) overrides
// the default tile order: the distributed factorisation passes the staircase of a block-cyclic trailing matrix.
template <int S>
static int oz_mma_launch(b2gp_ctx* ctx, cudaStream_t st, OzWork& w, const OzOperand& oa, const OzOperand& ob, int64_t m, int64_t n,
                         int64_t k, double alpha, double* C, int64_t ldc, bool lower_only, OzMode mode, const int2* list, int64_t count) {
    const int64_t kpad = round_up(k, OZ_KB);
    CUtensorMap mapA, mapB;
    if (!make_tmap_u8(&mapA, oa.planes, (int64_t)S * oa.rows_pad, kpad, OZ_BM) || !make_tmap_u8(&mapB, ob.planes, (int64_t)S * ob.rows_pad, kpad, OZ_BN))
        return B2GP_ERR_UNSUPPORTED;
    OzArgs a;
    a.m = (int)m;
    a.n = (int)n;
    a.kb_count = (int)(kpad / OZ_KB);
    a.rowsA_pad = (int)oa.rows_pad;
    a.rowsB_pad = (int)ob.rows_pad;
    a.sa = oa.scale;
    a.sb = ob.scale;
    a.C = C;
    a.ldc = ldc;
    a.alpha = alpha;
    a.lower_only = lower_only ? 1 : 0;
    a.overwrite = mode.overwrite ? 1 : 0;
    a.ktri = mode.ktri ? 1 : 0;
    a.tiles_m = (int)ceil_div(m, OZ_BM);
    a.tiles_n = (int)ceil_div(n, OZ_BN);
    const int CL = (ctx->oz_cluster == 2) ? 2 : 1;
    if (!list) RET_IF(oz_default_list(ctx, st, w, a.tiles_m, a.tiles_n, lower_only, CL, &list, &count));
    if (count <= 0) return B2GP_OK;
    const int64_t tiles = count;
    a.num_tiles = (int)tiles;
    a.tile_list = list;
    a.prof = nullptr;
    if (w.prof.p) a.prof = (long long*)w.prof.p;
    a.debug = ctx->oz_debug;
    constexpr int smem_bytes = oz_stages_for(S) * S * (OZ_A_TILE + OZ_B_TILE) + 128 + (128 * 17 + 128) * 8 + 1024;
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        CUDA_TRY(ctx, cudaFuncSetAttribute(oz_mma_kernel<S, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        CUDA_TRY(ctx, cudaFuncSetAttribute(oz_mma_kernel<S, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr.done(ctx->device);
    }
    const int nsm = persist_sms(ctx);
    if (CL == 2) {
        const int ncl = nsm / 2;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * (tiles < ncl ? tiles : ncl)));
        cfg.blockDim = dim3(OZ_THREADS);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, oz_mma_kernel<S, 2>, mapA, mapB, a));
    } else {
        const int grid = (int)(tiles < nsm ? tiles : nsm);
        oz_mma_kernel<S, 1><<<grid, OZ_THREADS, smem_bytes, st>>>(mapA, mapB, a);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    return B2GP_OK;
}

template <int S>
static int ozaki_gemm_nt(b2gp_ctx* ctx, cudaStream_t st, OzWork& w, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                         int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc, bool lower_only, OzMode mode = OzMode()) {
    if (m <= 0 || n <= 0 || k <= 0) return B2GP_OK;
    if (k > OZ_K_MAX) return B2GP_ERR_UNSUPPORTED;  // int32 accumulation bound (the dispatcher splits longer k)
    const bool same = (A == B && lda == ldb && n <= m && !mode.transB);   // B's rows are the first n rows of A: one set of planes
    OzOperand oa, ob;
    RET_IF(oz_slice_launch<S>(ctx, st, A, lda, m, k, w.planesA, w.scaleA, false, nullptr, &oa));
    if (same)
        ob = oa;
    else
        RET_IF(oz_slice_launch<S>(ctx, st, B, ldb, n, k, w.planesB, w.scaleB, mode.transB, nullptr, &ob));
    return oz_mma_launch<S>(ctx, st, w, oa, ob, m, n, k, alpha, C, ldc, lower_only, mode, nullptr, 0);
}

static int ozaki_dispatch(b2gp_ctx* ctx, cudaStream_t st, int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                          const double* B, int64_t ldb, double* C, int64_t ldc, bool lower_only, bool overwrite, bool transB, bool ktri) {
    Slot* sl = slot_of(ctx, st);
    OzWork* w = sl ? &sl->oz : &ctx->slots[0].oz;
    const int planes = oz_planes_for(ctx, st);
    for (int64_t k0 = 0; k0 < k; k0 += OZ_K_MAX) {
        const int64_t kc = k - k0 < OZ_K_MAX ? k - k0 : OZ_K_MAX;
        OzMode mode;
        mode.overwrite = overwrite && k0 == 0;      // later k-chunks accumulate onto the first
        mode.transB = transB;
        mode.ktri = ktri && k <= OZ_K_MAX;
        const double* Bk = transB ? B + k0 * ldb : B + k0;
        int rc;
        if (planes == 6)
            rc = ozaki_gemm_nt<6>(ctx, st, *w, m, n, kc, alpha, A + k0, lda, Bk, ldb, C, ldc, lower_only, mode);
        else
            rc = ozaki_gemm_nt<7>(ctx, st, *w, m, n, kc, alpha, A + k0, lda, Bk, ldb, C, ldc, lower_only, mode);
        if (rc != B2GP_OK) return rc;
    }
    return B2GP_OK;
}
