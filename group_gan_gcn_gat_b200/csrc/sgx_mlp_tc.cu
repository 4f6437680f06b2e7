// The two-layer context MLP (sgx_mlp.cu: out = ReLU(W2 ReLU(W1 [xa ; xb] + b1) + b2), make_mlp of sgan/models.py:7-20 at
// :898, :165-166, :990) with both linear maps on tcgen05 -- the tile scheme of the graph kernels (sgx_graph_tc.cuh): a
// tile = 128 pedestrians per four-warp group, thread = pedestrian = TMEM lane, fp16 hi + lo operand splits with per-row
// power-of-two scaling (fp32-grade), 3 K/16 MMAs per layer, the 64-wide hidden row never leaves registers / TMEM.
// Inference only (the nn.Sequential runs under autograd).  In the SGAN-P forward this operator sits between pooling and
// decoder: mma.sync version 0.050 ms on 206 k pedestrians (16 % of the HBM roofline).
#include "sgx_graph_tc.cuh"

namespace sgx {
namespace mtc {

using namespace gtile;

#ifndef MTC_GROUPS
#define MTC_GROUPS 4
#endif
constexpr int GROUPS = MTC_GROUPS;
constexpr int TCOLS = 64;            // TMEM columns per group (N <= 64)
constexpr int NTHREADS = GROUPS * 128;
constexpr int HID = 64;

template <int IN, int OUTP>
struct Cfg {
    static constexpr int K1 = (IN + 15) / 16 * 16, N1 = HID;      // W1: IN -> 64
    static constexpr int K2 = HID, N2 = OUTP <= 16 ? 16 : 32;      // W2: 64 -> OUT (padded to the MMA N: multiples of 16)
    static constexpr int KC = (K1 > K2 ? K1 : K2) / 8;            // K cores of the widest operand (hi), as many lo
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = OFF_W1 + 4 * K1 * N1;
    static constexpr int OFF_B1 = OFF_W2 + 4 * K2 * N2;           // float[64]
    static constexpr int OFF_B2 = OFF_B1 + HID * 4;               // float[32]
    static constexpr int OFF_WS = OFF_B2 + 128;                   // float[2] inverse weight scales, uint[2] max |w| bits
    static constexpr int OFF_BAR = OFF_WS + 64;
    static constexpr int OFF_GRP = (OFF_BAR + 64 + 127) / 128 * 128;
    static constexpr int GRP_BYTES = 2 * KC * CORE;
    static constexpr int SMEM_TOTAL = OFF_GRP + GROUPS * GRP_BYTES + 128;
    static constexpr int S1 = 0, S2 = N1 * K1, STAGE_FLOATS = S2 + N2 * K2;
    static_assert(STAGE_FLOATS * 4 <= GROUPS * GRP_BYTES, "weight staging lives in the group buffers");
    static_assert(SMEM_TOTAL <= 227 * 1024 && IN % 4 == 0 && IN <= 48 && OUTP <= 32, "dims");
};

template <int IN, int OUTP>
__global__ void __launch_bounds__(NTHREADS, 1)
mlp2_tc_kernel(const float* __restrict__ xa, int da, const float* __restrict__ xb, int db, int64_t batch,
               const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
               const float* __restrict__ b2, int OUT, float* __restrict__ out) {
    using C = Cfg<IN, OUTP>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    float* s_b1 = reinterpret_cast<float*>(smem + C::OFF_B1);
    float* s_b2 = reinterpret_cast<float*>(smem + C::OFF_B2);
    float* s_winv = reinterpret_cast<float*>(smem + C::OFF_WS);
    uint32_t* s_wmax = reinterpret_cast<uint32_t*>(smem + C::OFF_WS + 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + GROUPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp >> 2, wq = warp & 3;

    pdl_trigger();
    if (threadIdx.x == 0) {
        for (int g = 0; g < GROUPS; ++g) mbar_init(&bars[g], 1);
        fence_barrier_init();
    }
    if (threadIdx.x < 2) s_wmax[threadIdx.x] = 0u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_wait();          // barriers and TMEM are set up while the previous kernel of the chain drains; no global access yet

    // ---- weight images (W1, W2 are [out][in] row-major = the K-major B operand already), biases ----
    {
        float* stage = reinterpret_cast<float*>(smem + C::OFF_GRP);
        for (int e = threadIdx.x; e < C::STAGE_FLOATS; e += NTHREADS) stage[e] = 0.f;
        __syncthreads();
        uint32_t m1 = 0u, m2 = 0u;
#pragma unroll 4
        for (int e = threadIdx.x; e < HID * IN; e += NTHREADS) {
            const float v = W1[e];
            stage[C::S1 + (e / IN) * C::K1 + e % IN] = v;
            m1 = max(m1, __float_as_uint(v) & 0x7fffffffu);
        }
#pragma unroll 4
        for (int e = threadIdx.x; e < OUT * HID; e += NTHREADS) {
            const float v = W2[e];
            stage[C::S2 + e] = v;
            m2 = max(m2, __float_as_uint(v) & 0x7fffffffu);
        }
        m1 = __reduce_max_sync(0xffffffffu, m1);
        m2 = __reduce_max_sync(0xffffffffu, m2);
        if (lane == 0) { if (m1) atomicMax(&s_wmax[0], m1); if (m2) atomicMax(&s_wmax[1], m2); }
        if (threadIdx.x < HID) s_b1[threadIdx.x] = b1[threadIdx.x];
        if (threadIdx.x < 32) s_b2[threadIdx.x] = threadIdx.x < OUT ? b2[threadIdx.x] : 0.f;
        __syncthreads();
        build_image(smem + C::OFF_W1, stage + C::S1, C::N1, C::K1, s_wmax[0], &s_winv[0], threadIdx.x, NTHREADS);
        build_image(smem + C::OFF_W2, stage + C::S2, C::N2, C::K2, s_wmax[1], &s_winv[1], threadIdx.x, NTHREADS);
        fence_proxy_async();
        __syncthreads();
    }

    uint8_t* abuf = smem + C::OFF_GRP + grp * C::GRP_BYTES;
    const int row = wq * 32 + lane;
    uint8_t* arow = abuf + row * 16;
    uint64_t* bar = &bars[grp];
    const uint32_t a_s = sbase + C::OFF_GRP + grp * C::GRP_BYTES;
    const uint32_t d_tmem = tmem + (uint32_t)grp * TCOLS;
    const uint32_t d_mine = d_tmem + ((uint32_t)(wq * 32) << 16);
    uint32_t parity = 0;
    const float winv1 = s_winv[0], winv2 = s_winv[1];

    auto run_layer = [&](auto issue) {
        fence_proxy_async();
        tc_fence_before();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
        if (wq == 0) {
            tc_fence_after();
            if (elect_one()) issue();
            __syncwarp();
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        tc_fence_after();
    };

    // [xa ; xb] row of one pedestrian (zeros past the batch)
    float4 xq[IN / 4];
    const int qa = da >> 2;
    auto load_x = [&](int64_t p) {
#pragma unroll
        for (int c = 0; c < IN / 4; ++c) xq[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p < batch) {
            const float4* ra = reinterpret_cast<const float4*>(xa + p * da);
            const float4* rb = db > 0 ? reinterpret_cast<const float4*>(xb + p * db) - qa : ra;   // indexed by the cat column quad
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) xq[c] = (c < qa) ? ra[c] : rb[c];
        }
    };
    const int64_t n_tiles = (batch + 127) / 128;
    int64_t tile = (int64_t)blockIdx.x * GROUPS + grp;
    const int64_t tile_step = (int64_t)gridDim.x * GROUPS;
    if (tile < n_tiles) load_x(tile * 128 + row);

    for (; tile < n_tiles; tile += tile_step) {
        const int64_t p = tile * 128 + row;
        float sc;
        {
            float xv[IN];
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) { xv[4 * c] = xq[c].x; xv[4 * c + 1] = xq[c].y; xv[4 * c + 2] = xq[c].z; xv[4 * c + 3] = xq[c].w; }
            sc = row_to_operand<IN, C::K1>(arow, xv) * winv1;
        }
        run_layer([&]() { issue_layer<C::K1, C::N1>(d_tmem, a_s, sbase + C::OFF_W1, bar); });
        {
            float h[HID];
            {
                uint32_t v0[32];
                tmem_ld32(d_mine, v0);
                tmem_wait_ld();
#pragma unroll
                for (int f = 0; f < 32; ++f) h[f] = fmaxf(fmaf(__uint_as_float(v0[f]), sc, s_b1[f]), 0.f);
            }
            {
                uint32_t v1[32];
                tmem_ld32(d_mine + 32, v1);
                tmem_wait_ld();
#pragma unroll
                for (int f = 0; f < 32; ++f) h[32 + f] = fmaxf(fmaf(__uint_as_float(v1[f]), sc, s_b1[32 + f]), 0.f);
            }
            sc = row_to_operand<HID, C::K2>(arow, h) * winv2;
        }
        if (tile + tile_step < n_tiles) load_x((tile + tile_step) * 128 + row);       // in flight during the second round trip
        run_layer([&]() { issue_layer<C::K2, C::N2>(d_tmem, a_s, sbase + C::OFF_W2, bar); });
        if (OUTP >= 16) {
            uint32_t v0[32];
            tmem_ld32(d_mine, v0);            // (N2 <= 32 columns are written; the rest of the 32 is stale and unused)
            tmem_wait_ld();
            if (p < batch) {
                float4* orow = reinterpret_cast<float4*>(out + p * OUT);
#pragma unroll
                for (int f = 0; f < OUTP / 4; ++f)
                    orow[f] = make_float4(fmaxf(fmaf(__uint_as_float(v0[4 * f]), sc, s_b2[4 * f]), 0.f),
                                          fmaxf(fmaf(__uint_as_float(v0[4 * f + 1]), sc, s_b2[4 * f + 1]), 0.f),
                                          fmaxf(fmaf(__uint_as_float(v0[4 * f + 2]), sc, s_b2[4 * f + 2]), 0.f),
                                          fmaxf(fmaf(__uint_as_float(v0[4 * f + 3]), sc, s_b2[4 * f + 3]), 0.f));
            }
        } else {                              // the discriminator's single score
            uint32_t v0[2];
            tmem_ld2(d_mine, v0);
            tmem_wait_ld();
            if (p < batch) out[p * OUT] = fmaxf(fmaf(__uint_as_float(v0[0]), sc, s_b2[0]), 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

}  // namespace mtc

template <int IN, int OUTP>
int mlp2_tc_forward(const float* xa, int da, const float* xb, int db, int64_t batch, const float* W1, const float* b1,
                    const float* W2, const float* b2, int OUT, float* out, cudaStream_t st) {
    using C = mtc::Cfg<IN, OUTP>;
    auto kern = mtc::mlp2_tc_kernel<IN, OUTP>;
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL));
    int dev = 0, sms = 148;
    SGX_CUDA(cudaGetDevice(&dev));
    SGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t n_tiles = (batch + 127) / 128;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n_tiles + mtc::GROUPS - 1) / mtc::GROUPS, sms));
    SGX_CUDA(launch_pdl(kern, dim3(grid), dim3(mtc::NTHREADS), C::SMEM_TOTAL, st, true, xa, da, xb, db, batch, W1, b1, W2, b2,
                        OUT, out));
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

#define MLP_TC_INST(I, O)                                                                                             \
    template int mlp2_tc_forward<I, O>(const float*, int, const float*, int, int64_t, const float*, const float*,     \
                                       const float*, const float*, int, float*, cudaStream_t);
MLP_TC_INST(32, 24) MLP_TC_INST(40, 24) MLP_TC_INST(48, 24)
MLP_TC_INST(32, 32) MLP_TC_INST(40, 32) MLP_TC_INST(48, 32)
MLP_TC_INST(32, 8) MLP_TC_INST(40, 8) MLP_TC_INST(48, 8)

}  // namespace sgx
