// Nested SPD projection on the 5th-generation tensor cores: tcgen05.mma.kind::tf32 with the accumulators in TMEM.
// Included by nested_project.cu (same pack buffer, same entry point).
//
// STATUS: parity-green on the B200 (tests/test_nested_gpu.py::test_projection_tcgen05_kernel_parity) and OPT-IN
// (GABO_PROJECT_KERNEL=tc): it is slower than the mma.sync kernel on this shape.  Measured at N = 2^20 (scripts/dev_tc.py):
// 0.413 ms (2.29 TB/s) against 0.196 ms for the mma.sync kernel.
//
// What the measurements say (scripts/micro/umma_rate.cu, umma_rate2.cu; profiles/r02b_umma_rate*.log):
//  * tcgen05.mma.kind::tf32 from shared memory, issued by ONE elected lane of a warp-uniform branch with the products unrolled:
//    M 64 x N 16 x K 8 = 23 cycles, N 64 = 32, N 128 = 64, N 256 = 128 (M = 128 costs the same from N = 128 on: 4090 FLOP/clk/SM,
//    the dense TF32 peak; kind::f16 K = 16 takes the same cycles = 8180 FLOP/clk/SM), same whether or not consecutive products
//    share an accumulator or the operand addresses advance.  Issued from `if (tid == 0)` inside a rolled loop -- the first form
//    of this kernel -- ptxas wraps every product in an ELECT / BRA.U.ANY retry loop and the same product costs 110 - 210 cycles:
//    that, not the tensor core, was the 0.643 ms of the first version.
//  * With the products at 23 cycles (81 per 64-row tile = 1.9 k cycles, under the 2.4 k cycles HBM needs per tile) the kernel
//    is bound by the split / pack pass: without the MMAs it takes 0.314 ms.  The rows arrive with an 840-byte pitch (no tensor
//    map can deliver them in the canonical core-matrix layout: the pitch is not a multiple of 16 bytes), so every value is read
//    from the raw tile, split and written twice (hi, lo) by the CUDA cores, and then read three times by the tensor core
//    (hi twice, lo once): ~7.9 bytes of shared-memory traffic per input byte, 3.3 k cycles per tile at 128 B/clk -- already above
//    the HBM time before any latency.  The mma.sync kernel reads every value ONCE into a register fragment (the operator
//    lives in registers), which is why it wins here although its tensor pipe is the slow one.
//  * A form that could win keeps the operator as the A operand in TMEM ([hi; lo] stacked to M = 64) and streams the rows as
//    the N operand (hi and lo read once each): 4.7 bytes of shared-memory traffic per input byte, best case ~0.9 of HBM.
//    Not built this round.
//
// Why it was built: the ablation of the mma.sync kernel (scripts/micro/project_variants.cu, profiles/r02_*) shows that kernel is bound by
// the LEGACY tensor pipe -- with one HMMA.1688.TF32 per k-step instead of the three of 3xTF32 it streams at 0.98 of HBM, with
// three it stops at 0.74 (8.7 cycles per HMMA per SM sub-partition).  tcgen05.mma issues a whole 64 x 16 x 8 product from ONE
// thread at >= 8 cycles per dispatch for the SM, ~25x fewer tensor-pipe cycles per row, so the 3xTF32 triple is free again.
//
// Data flow per CTA (one per SM, 256 threads):
//   raw ring      2 stages x (64 rows x dvh floats), filled by the TMA bulk-copy engine exactly as in the mma.sync kernel
//   split + pack  all 8 warps: a 64-row x 72-column chunk of the raw tile is split into tf32 hi / lo (LOP3 + FADD per value) and
//                 written in the canonical K-major no-swizzle UMMA layout: 8-row x 16-byte core matrices, row blocks 128 B
//                 apart (SBO), k-chunks 1024 B apart (LBO).  Lane <-> row: the 64-bit reads of the 840-byte rows and the 128-bit
//                 writes are both bank-conflict free.  Two chunk buffers, so packing chunk c+1 overlaps the MMAs of chunk c.
//   MMA           thread 0: per k-step (8 columns) three tcgen05.mma M64 N16 K8: hi*hi into accumulator 0, lo*hi and hi*lo
//                 into accumulator 1 (small terms kept apart, as before); tcgen05.commit to an mbarrier frees the chunk buffer.
//                 The operator (16 x dvh, hi and lo, canonical layout) is built once per CTA from the pack buffer.
//   epilogue      warps 0-3: tcgen05.ld 32x32b.x16 of both accumulators (M = 64 puts rows 16w .. 16w+15 on lanes 0..15 of TMEM
//                 quadrant w), sum, store.  Accumulators are double-buffered in TMEM (64 columns), so the epilogue of tile t
//                 runs under the MMAs of tile t+1.
#pragma once
#include "common.cuh"

namespace gabo {
namespace tc {

constexpr int kRows = 64;            // rows per tile = M of the MMA
constexpr int kChunkCols = 72;       // columns per packed chunk (9 k-steps of 8)
constexpr int kChunkKc = kChunkCols / 4;   // 16-byte k-chunks per packed chunk
constexpr int kThreads = 256;
constexpr int kN = 16;               // output Mandel entries per MMA (dvl <= 16)
constexpr int kPhases = 4;            // independent accumulators per product: k-step s accumulates into phase s % kPhases
constexpr int kAccPerTile = 3 * kPhases;   // (hi*hi, lo*hi, hi*lo) x phases, 16 TMEM columns each
constexpr int kTileCols = 256;       // TMEM columns reserved per tile in flight (12 x 16 = 192 used)
constexpr int kAccCols = 512;        // TMEM columns allocated: 2 tiles in flight

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48), layout NONE
    return static_cast<uint64_t>((smem_addr >> 4) & 0x3fffu) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fffu) << 16) |
           (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((kN >> 3) << 17) | ((kRows >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool elect_one() {          // one lane of a converged warp
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(e));
    return e != 0;
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory carve-up (bytes): raw ring | chunk buffers (hi, lo) x 2 | operator (hi, lo) | barriers + TMEM slot
struct Smem {
    int raw_stage, chunk_half, op_half, nchunks, kp;
    size_t raw_off, chunk_off, op_off, bar_off, total;
    __host__ __device__ Smem(int dvh) {
        nchunks = (dvh + kChunkCols - 1) / kChunkCols;
        kp = nchunks * kChunkCols;                       // padded K
        raw_stage = kRows * dvh * 4;                     // multiple of 256
        chunk_half = kRows * kChunkCols * 4;             // 18432
        op_half = kN * kp * 4;
        raw_off = 0;
        chunk_off = raw_off + 2 * static_cast<size_t>(raw_stage);
        op_off = chunk_off + 4 * static_cast<size_t>(chunk_half);
        bar_off = op_off + 2 * static_cast<size_t>(op_half);
        total = bar_off + 128;
    }
};

// op_can: canonical operator image in the pack buffer: [hi | lo], each kN x kp floats in core-matrix order
//   float index of (n, k) = ((k / 4) * 2 + n / 8) * 32 + (n % 8) * 4 + k % 4
__global__ void __launch_bounds__(kThreads, 1)
    nested_project_tc_kernel(const float* __restrict__ x, int64_t n, int dvh, int dvl, const float* __restrict__ op_can,
                             float* __restrict__ y) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const Smem L(dvh);
    float* raw = reinterpret_cast<float*>(smem + L.raw_off);
    unsigned char* chunk = smem + L.chunk_off;          // [buf][hi/lo]
    float* op = reinterpret_cast<float*>(smem + L.op_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bar_off);   // [0,1] raw full, [2,3] chunk free, [4,5] accumulator full
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t tiles = (n + kRows - 1) / kRows;

    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(&bars[i], 1);
        fence_mbar_init();
    }
    if (warp == 0) {   // TMEM allocation: one warp, address through shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kAccCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int e = tid; e < 2 * L.op_half / 16; e += kThreads)
        reinterpret_cast<float4*>(op)[e] = __ldg(reinterpret_cast<const float4*>(op_can) + e);
    tc_fence_before();
    fence_proxy_async();                                 // operator image: generic-proxy writes -> tensor-core reads
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t full_bytes = static_cast<uint32_t>(L.raw_stage);
    auto issue_tma = [&](int64_t tile, int stage) {      // thread 0 only; full tiles only
        if (tile < tiles && (tile + 1) * kRows <= n) {
            mbar_expect_tx(&bars[stage], full_bytes);
            tma_load_1d(reinterpret_cast<unsigned char*>(raw) + static_cast<size_t>(stage) * L.raw_stage,
                        x + tile * kRows * dvh, full_bytes, &bars[stage]);
        }
    };
    if (tid == 0) {
        issue_tma(blockIdx.x, 0);
        issue_tma(static_cast<int64_t>(blockIdx.x) + gridDim.x, 1);
    }
    const uint32_t op_hi = smem_u32(op), op_lo = op_hi + L.op_half;
    // packing role of this thread: row r, k-chunks kc0, kc0 + 4, ...
    const int r = tid & 63, kc0 = tid >> 6;
    const bool even = (dvh & 1) == 0;
    const uint32_t row_dst = static_cast<uint32_t>(((r >> 3) * 8 + (r & 7)) * 16);   // inside a k-chunk's 1024-byte block

    auto epilogue = [&](int64_t tile, int par) {         // warps 0-3: rows 16 warp .. 16 warp + 15 on TMEM lanes 32 warp + 0..15
        if (warp < 4) {
            float a[16], b[16];
            const uint32_t t0 = tmem_base + (static_cast<uint32_t>(32 * warp) << 16) + static_cast<uint32_t>(par * kTileCols);
#pragma unroll
            for (int c = 0; c < 16; ++c) a[c] = 0.0f;
            // small products first (accumulators kPhases .. 3 kPhases - 1), then the hi * hi partial tiles
#pragma unroll 1
            for (int q = kAccPerTile - 1; q >= 0; --q) {
                tmem_ld16(t0 + static_cast<uint32_t>(q * 16), b);
#pragma unroll
                for (int c = 0; c < 16; ++c) a[c] += b[c];
            }
            const int64_t row = tile * kRows + 16 * warp + lane;
            if (lane < 16 && row < n) {
                float* dst = y + row * dvl;
                for (int c = 0; c < dvl; ++c) __stcs(dst + c, a[c]);
            }
        }
        tc_fence_before();
    };

    int it = 0;
    int64_t prev_tile = -1;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int st = it & 1, par = it & 1;
        const int rows = static_cast<int>(imin(kRows, n - tile * kRows));
        float* A = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(raw) + static_cast<size_t>(st) * L.raw_stage);
        if (rows == kRows) {
            mbar_wait(&bars[st], (it >> 1) & 1);
        } else {   // ragged last tile: plain cooperative copy, zero fill
            const float* src = x + tile * kRows * dvh;
            for (int e = tid; e < rows * dvh; e += kThreads) A[e] = src[e];
            for (int e = rows * dvh + tid; e < kRows * dvh; e += kThreads) A[e] = 0.0f;
            __syncthreads();
        }
        for (int c = 0; c < L.nchunks; ++c) {
            const int g = it * L.nchunks + c, buf = g & 1;
            if (g >= 2) mbar_wait(&bars[2 + buf], ((g - 2) >> 1) & 1);     // the MMAs that read this buffer are done
            unsigned char* hi = chunk + static_cast<size_t>(buf) * 2 * L.chunk_half;
            unsigned char* lo = hi + L.chunk_half;
            const float* rowp = A + r * dvh + c * kChunkCols;
#pragma unroll
            for (int j = 0; j < (kChunkKc + 3) / 4; ++j) {
                const int kc = kc0 + 4 * j;
#if defined(GABO_TC_ABLATE) && GABO_TC_ABLATE == 2      // ablation 2: no split / pack pass (MMAs on stale buffers)
                if (kc < 0) {
#else
                if (kc < kChunkKc) {
#endif
                    const int col = c * kChunkCols + 4 * kc;
                    float v[4];
                    if (even && col + 3 < dvh) {     // 8-byte aligned pairs only when the row pitch is even
                        const float2 u0 = *reinterpret_cast<const float2*>(rowp + 4 * kc);
                        const float2 u1 = *reinterpret_cast<const float2*>(rowp + 4 * kc + 2);
                        v[0] = u0.x; v[1] = u0.y; v[2] = u1.x; v[3] = u1.y;
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[q] = (col + q < dvh) ? rowp[4 * kc + q] : 0.0f;
                    }
                    float h[4], l[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        h[q] = __uint_as_float(__float_as_uint(v[q]) & 0xffffe000u);
                        l[q] = v[q] - h[q];
                    }
                    const uint32_t off = static_cast<uint32_t>(kc) * 1024u + row_dst;
                    *reinterpret_cast<float4*>(hi + off) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(lo + off) = make_float4(l[0], l[1], l[2], l[3]);
                }
            }
            fence_proxy_async();                         // packed chunk: generic-proxy writes -> tensor-core reads
            tc_fence_before();
            __syncthreads();
            // issue pattern matters (scripts/micro/umma_rate*.cu): from ONE elected lane of a warp-uniform branch, unrolled, a
            // 64 x 16 x 8 product costs 23 cycles; from `if (tid == 0)` with a rolled loop ptxas wraps every tcgen05.mma in an
            // ELECT / BRA.U.ANY retry loop and the same product costs 110 - 210 cycles
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                const uint32_t a_hi = smem_u32(hi), a_lo = smem_u32(lo);
                // kPhases accumulators per product type, used round-robin over the k-steps (kept from the first version; the
                // micro-benchmark shows dependent products pipeline just as well), added up by the epilogue
                const uint32_t d_tile = tmem_base + static_cast<uint32_t>(par * kTileCols);
#pragma unroll
                for (int s = 0; s < kChunkCols / 8; ++s) {
                    const int ks = c * (kChunkCols / 8) + s;                       // global k-step
                    const uint64_t dah = umma_desc(a_hi + 2048u * s, 1024u, 128u);
                    const uint64_t dal = umma_desc(a_lo + 2048u * s, 1024u, 128u);
                    const uint64_t dbh = umma_desc(op_hi + 512u * ks, 256u, 128u);
                    const uint64_t dbl = umma_desc(op_lo + 512u * ks, 256u, 128u);
                    const uint32_t acc = (ks >= kPhases) ? 1u : 0u;
                    const uint32_t col = d_tile + static_cast<uint32_t>((ks % kPhases) * 16);
#if !defined(GABO_TC_ABLATE) || GABO_TC_ABLATE != 1   // ablation 1 (scripts/micro/project_variants.cu): no MMAs
                    umma_tf32(col, dah, dbh, acc);                                 // hi * hi
                    umma_tf32(col + kPhases * 16, dal, dbh, acc);                  // lo(x) * hi(P)
                    umma_tf32(col + 2 * kPhases * 16, dah, dbl, acc);              // hi(x) * lo(P)
#endif
                }
                umma_commit(&bars[2 + buf]);
                if (c == L.nchunks - 1) umma_commit(&bars[4 + par]);
            }
            if (c == 0 && prev_tile >= 0) {              // epilogue of the previous tile under this tile's MMAs
                mbar_wait(&bars[4 + (par ^ 1)], ((it - 1) >> 1) & 1);
                tc_fence_after();
                epilogue(prev_tile, par ^ 1);
            }
        }
        // every thread is past its reads of the raw stage (the barrier of the last chunk): refill it
        if (warp == 0) __syncwarp();
        if (tid == 0) {
            fence_proxy_async();
            issue_tma(tile + 2 * static_cast<int64_t>(gridDim.x), st);
        }
        prev_tile = tile;
    }
    if (prev_tile >= 0) {
        const int par = (it - 1) & 1;
        mbar_wait(&bars[4 + par], ((it - 1) >> 1) & 1);
        tc_fence_after();
        epilogue(prev_tile, par);
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kAccCols));
    }
}

}  // namespace tc
}  // namespace gabo
