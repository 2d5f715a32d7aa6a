// Nested SPD projection Y = W^T X W of HD-GaBO (P1 of SURVEY.md section 8) as ONE dense contraction in Mandel
// coordinates:  y[n x dvl] = x[n x dvh] * P^T  with  P[(a,b),(p,q)] = m_ab (W_pa W_qb + [p != q] W_qa W_pb) / m_pq.
//
// Replaces projection_from_spd_to_nested_spd (BoManifolds/nested_mappings/nested_spd_utils.py:13-48: W repeated n
// times, two bmm) together with the Mandel unpack / pack either side of it (kernels_nested_spd.py:122-127).
//
// Roofline: HBM read, 4 (dvh + dvl) bytes per matrix (900 B for SPD(20) -> SPD(5)) against 2 dvh dvl flop; the FP32
// pipe alone cannot keep up with HBM at that intensity, so the products run on the tensor cores as 3xTF32
// (a_hi b_hi + a_lo b_hi + a_hi b_lo, error ~ 2^-21) with fp32 accumulation -- the only GEMM-shaped op on the path.
// Layout: persistent CTAs (one per SM), 64-row tiles of x streamed through a 3-stage shared-memory ring by the TMA
// bulk-copy engine (rows are contiguous, a tile is one 1-D copy), four warps x m16 rows, P pre-split into hi/lo and
// pre-arranged per lane so that a k-step needs one 16-byte shared load per 8 output columns.
#include "spd_common.cuh"

namespace gabo {
namespace {

constexpr int kTileRows = 64;
constexpr int kStages = 3;
constexpr int kThreadsP = 128;

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], const float (&a)[4], float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
          "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

__device__ __forceinline__ void mandel_rc_dev(int d, int pos, int& r, int& c) {
    int k = 0, len = d;
    while (pos >= len) {
        pos -= len;
        --len;
        ++k;
    }
    r = pos;
    c = pos + k;
}

// pack[(s * 32 + lane) * NT + j][4] = { hi(b0), hi(b1), lo(b0), lo(b1) },  b0 = P[8j + g][8s + 2t], b1 = P[8j + g][8s + 2t + 1]
__global__ void projection_pack_kernel(const double* __restrict__ w, int D, int d, int ksteps, int nt,
                                       float* __restrict__ pack) {
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    const int total = ksteps * 32 * nt * 2;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int which = e & 1;
        const int j = (e >> 1) % nt;
        const int lane = ((e >> 1) / nt) & 31;
        const int s = ((e >> 1) / nt) >> 5;
        const int g = lane >> 2, t = lane & 3;
        const int o = 8 * j + g;           // low Mandel index (output column)
        const int i = 8 * s + 2 * t + which;  // high Mandel index (k)
        double v = 0.0;
        if (o < dvl && i < dvh) {
            int a, b, p, q;
            mandel_rc_dev(d, o, a, b);
            mandel_rc_dev(D, i, p, q);
            const double mab = (a == b) ? 1.0 : 1.4142135623730951;
            if (p == q) v = mab * w[p * d + a] * w[p * d + b];
            else v = mab * (w[p * d + a] * w[q * d + b] + w[q * d + a] * w[p * d + b]) / 1.4142135623730951;
        }
        const float vf = static_cast<float>(v);
        const float hi = to_tf32(vf);
        const float lo = to_tf32(static_cast<float>(v - static_cast<double>(hi)));
        float* dst = pack + (static_cast<int64_t>((s * 32 + lane) * nt + j)) * 4;
        dst[which] = hi;
        dst[2 + which] = lo;
    }
}

template <int NT, bool EVEN>
__global__ void __launch_bounds__(kThreadsP, 1)
    nested_project_kernel(const float* __restrict__ x, int64_t n, int dvh, int dvl, int ksteps,
                          const float* __restrict__ pack, float* __restrict__ y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int stage_floats = kTileRows * dvh;  // 64 * dvh * 4 bytes: a multiple of 256
    float* As = reinterpret_cast<float*>(smem_raw);
    float* Bp = As + kStages * stage_floats;
    float* Os = Bp + ksteps * 32 * NT * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(Os + ((kTileRows * dvl + 3) & ~3));

    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    for (int e = threadIdx.x; e < ksteps * 32 * NT; e += kThreadsP)
        reinterpret_cast<float4*>(Bp)[e] = __ldg(reinterpret_cast<const float4*>(pack) + e);
    __syncthreads();

    const uint32_t full_bytes = static_cast<uint32_t>(stage_floats) * 4u;
    auto issue = [&](int64_t tile, int stage) {  // thread 0 only; full tiles only
        if (tile < tiles && (tile + 1) * kTileRows <= n) {
            mbar_expect_tx(&bars[stage], full_bytes);
            tma_load_1d(As + stage * stage_floats, x + tile * kTileRows * dvh, full_bytes, &bars[stage]);
        }
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) issue(static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(s) * gridDim.x, s);
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    uint32_t phase_bits = 0u;
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++it) {
        const int stage = it % kStages;
        const int rows = static_cast<int>(imin(kTileRows, n - tile * kTileRows));
        float* A = As + stage * stage_floats;
        if (rows == kTileRows) {
            mbar_wait(&bars[stage], (phase_bits >> stage) & 1u);
            phase_bits ^= (1u << stage);
        } else {  // ragged last tile: plain cooperative copy
            const float* src = x + tile * kTileRows * dvh;
            for (int e = threadIdx.x; e < rows * dvh; e += kThreadsP) A[e] = src[e];
            for (int e = rows * dvh + threadIdx.x; e < stage_floats; e += kThreadsP) A[e] = 0.0f;
            __syncthreads();
        }

        // Three accumulators per output tile (hi*hi, lo*hi, hi*lo): with a single one every k-step appended three
        // DEPENDENT mma.sync to the same registers (81 in a row for dvh = 210), and with one warp per scheduler that
        // chain, not HBM, set the tile time.  The small terms are summed first at the end.
        float acc[NT][4], acc_lh[NT][4], acc_hl[NT][4];
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[j][q] = 0.0f;
                acc_lh[j][q] = 0.0f;
                acc_hl[j][q] = 0.0f;
            }
        const float* r0 = A + (warp * 16 + g) * dvh;
        const float* r1 = r0 + 8 * dvh;
        const float4* bp = reinterpret_cast<const float4*>(Bp) + lane * NT;
#pragma unroll 3
        for (int s = 0; s < ksteps; ++s) {
            const int col = 8 * s + 2 * t;
            float a[4];  // a0:(g, k=col) a1:(g+8, col) a2:(g, col+1) a3:(g+8, col+1)
            if (EVEN) {
                const float2 u = *reinterpret_cast<const float2*>(r0 + col);
                const float2 v = *reinterpret_cast<const float2*>(r1 + col);
                a[0] = u.x; a[2] = u.y; a[1] = v.x; a[3] = v.y;
            } else {
                a[0] = r0[col]; a[2] = r0[col + 1]; a[1] = r1[col]; a[3] = r1[col + 1];
            }
            if (col >= dvh) { a[0] = 0.0f; a[1] = 0.0f; }
            if (col + 1 >= dvh) { a[2] = 0.0f; a[3] = 0.0f; }
            float ah[4], al[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ah[q] = to_tf32(a[q]);
                al[q] = to_tf32(a[q] - ah[q]);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                const float4 b = bp[s * 32 * NT + j];  // {hi b0, hi b1, lo b0, lo b1}
                mma_tf32(acc_lh[j], al, b.x, b.y);
                mma_tf32(acc_hl[j], ah, b.z, b.w);
                mma_tf32(acc[j], ah, b.x, b.y);
            }
        }
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[j][q] += acc_lh[j][q] + acc_hl[j][q];
        __syncthreads();  // every warp is done with this stage (and with the previous tile's output staging)
        if (threadIdx.x == 0) {
            fence_proxy_async();
            issue(tile + static_cast<int64_t>(kStages) * gridDim.x, stage);
        }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int c0 = 8 * j + 2 * t;
            float* o0 = Os + (warp * 16 + g) * dvl;
            float* o1 = o0 + 8 * dvl;
            if (c0 < dvl) { o0[c0] = acc[j][0]; o1[c0] = acc[j][2]; }
            if (c0 + 1 < dvl) { o0[c0 + 1] = acc[j][1]; o1[c0 + 1] = acc[j][3]; }
        }
        __syncthreads();
        float* dst = y + tile * kTileRows * dvl;
        for (int e = threadIdx.x; e < rows * dvl; e += kThreadsP) __stcs(dst + e, Os[e]);
    }
}

int ksteps_for(int dvh) { return (dvh + 7) / 8; }
int ntiles_for(int dvl) { return (dvl + 7) / 8; }

template <int NT>
int launch_nt(const float* x, int64_t n, int dvh, int dvl, const float* pack, float* y, cudaStream_t s) {
    const int ksteps = ksteps_for(dvh);
    const size_t smem = sizeof(float) * (static_cast<size_t>(kStages) * kTileRows * dvh + static_cast<size_t>(ksteps) * 32 * NT * 4 +
                                         ((kTileRows * dvl + 3) & ~3)) + 8 * kStages + 16;
    GABO_REQUIRE(smem <= 227 * 1024, GABO_E_UNSUPPORTED,
                 "gabo_nested_spd_project: Mandel length %d needs %zu bytes of shared memory (> 227 KB)", dvh, smem);
    const int64_t tiles = (n + kTileRows - 1) / kTileRows;
    const unsigned grid = static_cast<unsigned>(imin(tiles, sm_count()));
    const bool even = (dvh % 2) == 0;
    if (even) {
        auto kern = nested_project_kernel<NT, true>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kThreadsP, smem, s>>>(x, n, dvh, dvl, ksteps, pack, y);
    } else {
        auto kern = nested_project_kernel<NT, false>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        kern<<<grid, kThreadsP, smem, s>>>(x, n, dvh, dvl, ksteps, pack, y);
    }
    return check_launch("nested_project_kernel");
}

}  // namespace
}  // namespace gabo

extern "C" int64_t gabo_nested_projection_pack_size(int D, int d) {
    if (D < 1 || d < 1 || d > D || d > GABO_MAX_SPD_DIM) return -1;
    return static_cast<int64_t>(gabo::ksteps_for(D * (D + 1) / 2)) * 32 * gabo::ntiles_for(d * (d + 1) / 2) * 4;
}

extern "C" int gabo_nested_projection_matrix(const double* w, int D, int d, float* p_pack, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(w && p_pack, GABO_E_ARG, "gabo_nested_projection_matrix: null pointer");
    GABO_REQUIRE(D >= 1 && d >= 1 && d <= D && d <= GABO_MAX_SPD_DIM, GABO_E_ARG,
                 "gabo_nested_projection_matrix: need 1 <= d <= min(D, %d), got D=%d d=%d", GABO_MAX_SPD_DIM, D, d);
    GABO_REQUIRE(aligned16(p_pack), GABO_E_ALIGN, "gabo_nested_projection_matrix: pack must be 16-byte aligned");
    const int ksteps = ksteps_for(D * (D + 1) / 2), nt = ntiles_for(d * (d + 1) / 2);
    const int total = ksteps * 32 * nt * 2;
    projection_pack_kernel<<<(total + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, D, d, ksteps, nt,
                                                                                              p_pack);
    return check_launch("projection_pack_kernel");
}

extern "C" int gabo_nested_spd_project(const float* x_mandel, int64_t n, int D, int d, const float* p_pack,
                                       float* y_mandel, void* stream) {
    using namespace gabo;
    GABO_REQUIRE(n >= 0, GABO_E_ARG, "gabo_nested_spd_project: negative size");
    if (n == 0) return GABO_OK;
    GABO_REQUIRE(x_mandel && p_pack && y_mandel, GABO_E_ARG, "gabo_nested_spd_project: null pointer");
    GABO_REQUIRE(D >= 1 && d >= 1 && d <= D && d <= GABO_MAX_SPD_DIM, GABO_E_ARG,
                 "gabo_nested_spd_project: need 1 <= d <= min(D, %d), got D=%d d=%d", GABO_MAX_SPD_DIM, D, d);
    GABO_REQUIRE(aligned16(x_mandel) && aligned16(p_pack), GABO_E_ALIGN,
                 "gabo_nested_spd_project: x and pack must be 16-byte aligned");
    const int dvh = D * (D + 1) / 2, dvl = d * (d + 1) / 2;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    switch (ntiles_for(dvl)) {
        case 1: return launch_nt<1>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 2: return launch_nt<2>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 3: return launch_nt<3>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 4: return launch_nt<4>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
        case 5: return launch_nt<5>(x_mandel, n, dvh, dvl, p_pack, y_mandel, s);
    }
    return GABO_E_ARG;
}
