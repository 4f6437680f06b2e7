// GCNModule forward (sgan/models.py:583-712; GCN 552-580) with the five linear maps on the 5th-gen tensor cores
// (tcgen05.mma kind::f16, accumulators in TMEM) -- the single-launch path for batches whose scenes fit a warp chunk
// (<= 32 pedestrians), dims (32 | 40) / 72 / 16 / (24 | 32).
//
// Same tile scheme as the GATEncoder kernel (sgx_gat_tc.cu, sgx_graph_tc.cuh): a tile = 128 pedestrians = four warp
// chunks, thread = pedestrian = TMEM lane, fp16 hi + lo operand splits with per-row power-of-two scaling (fp32-grade
// accuracy), one batch of 3 K/16 MMAs per linear map.  The normalised adjacencies of the reference are segmented means
// (A_intra = 1/|group| on the group, A_inter = 1/G on the scene's groups), identical for every row of a group / scene,
// so the chain per pedestrian is
//     M1 = mean of x over the group -> [W0] -> M2 = |g|-fold sum of relu(.)/|g| -> [W1] -> X1 = relu(.),
//     Xg = |g|-fold sum of X1/|g| -> N1 = mean of Xg over the scene's groups -> [V0] -> N2 = G-fold sum of relu(.)/G
//     -> [V1] -> Y = relu(.) -> cat(X1, Y/|g|) -> [Wo] + bo
// (the k-fold sums add the k identical terms one by one, which is what the reference's A @ H does in fp32).  Every
// member of a group / scene computes its group's / scene's rows itself -- identical operand rows give identical
// accumulator rows -- so the only values that cross lanes are the x rows (group mean) and the Xg rows (scene mean),
// both through the warp's own 512-byte pieces of the operand buffer.
#include "sgx_graph_tc.cuh"

namespace sgx {
namespace gctc {

using namespace gtile;

#ifndef GCTC_GROUPS
#define GCTC_GROUPS 3
#endif
constexpr int GROUPS = GCTC_GROUPS;
constexpr int NTHREADS = GROUPS * 128;
constexpr int HID = 72, OUT = 16;
constexpr int ABUF = 20 * CORE;                       // A operand (K <= 80: 10 hi + 10 lo cores) / fp32 rows
constexpr int GRP_BYTES = ABUF;

template <int IN, int FIN>
struct Cfg {
    static constexpr int K1 = (IN + 15) / 16 * 16, N1 = 80;     // W0: IN -> 72
    static constexpr int K2 = 80, N2 = 16;                      // W1: 72 -> 16
    static constexpr int K3 = 16, N3 = 80;                      // V0: 16 -> 72
    static constexpr int K4 = 80, N4 = 16;                      // V1: 72 -> 16
    static constexpr int K5 = 32, N5 = 32;                      // Wo^T: 32 -> FIN
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_W2 = OFF_W1 + 4 * K1 * N1;
    static constexpr int OFF_W3 = OFF_W2 + 4 * K2 * N2;
    static constexpr int OFF_W4 = OFF_W3 + 4 * K3 * N3;
    static constexpr int OFF_W5 = OFF_W4 + 4 * K4 * N4;
    static constexpr int OFF_BO = OFF_W5 + 4 * K5 * N5;         // float[32]: bo
    static constexpr int OFF_WS = OFF_BO + 128;                 // float[8]: inverse weight scales; uint[8]: max |w| bits
    static constexpr int OFF_BAR = OFF_WS + 64;                 // GROUPS mbarriers + the TMEM base slot
    static constexpr int OFF_GRP = (OFF_BAR + 64 + 127) / 128 * 128;
    static constexpr int SMEM_TOTAL = OFF_GRP + GROUPS * GRP_BYTES + 128;
    static constexpr int S1 = 0, S2 = S1 + N1 * K1, S3 = S2 + N2 * K2, S4 = S3 + N3 * K3, S5 = S4 + N4 * K4,
                         STAGE_FLOATS = S5 + N5 * K5;
    static_assert(STAGE_FLOATS * 4 <= GROUPS * GRP_BYTES, "the fp32 staging of the weight prep lives in the group buffers");
    static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory");
    static_assert(OFF_BAR % 16 == 0 && 8 * (GROUPS + 2) <= 64, "the prepared blob is one bulk copy; barriers + TMEM slot fit 64 bytes");
    static_assert(FIN <= 32 && FIN % 4 == 0 && IN % 4 == 0 && IN <= 48, "dims");
};

// ReLU that keeps a NaN a NaN like torch.relu (fmaxf would return the other operand)
__device__ __forceinline__ float relu_nan(float v) { return v < 0.f ? 0.f : v; }

// k-fold sum of one term, added one by one like the reference's A @ H over k identical rows (F terms at once; the
// hidden row goes through in blocks of <= 32 so that term + accumulator stay in registers)
template <int F>
__device__ __forceinline__ void repeat_block(const uint32_t (&raw)[F], float sc, float scale2, int times, float* __restrict__ h) {
    float term[F], acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { term[f] = scale2 * relu_nan(__uint_as_float(raw[f]) * sc); acc[f] = 0.f; }
    for (int t = 0; t < times; ++t) {
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] += term[f];
    }
#pragma unroll
    for (int f = 0; f < F; ++f) h[f] = acc[f];
}
template <int F>
__device__ __forceinline__ void repeat_rows(float (&v)[F], int times) {
    float acc[F];
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
    for (int t = 0; t < times; ++t) {
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] += v[f];
    }
#pragma unroll
    for (int f = 0; f < F; ++f) v[f] = acc[f];
}

// the thread's 72-wide accumulator row -> registers: h <- `times`-fold sum of scale2 * relu(row)
__device__ __forceinline__ void hidden_from_tmem(uint32_t d_mine, float sc, float scale2, int times, float (&h)[HID]) {
    {
        uint32_t v0[32];
        tmem_ld32(d_mine, v0);
        tmem_wait_ld();
        repeat_block<32>(v0, sc, scale2, times, h);
    }
    {
        uint32_t v1[32];
        tmem_ld32(d_mine + 32, v1);
        tmem_wait_ld();
        repeat_block<32>(v1, sc, scale2, times, h + 32);
    }
    {
        uint32_t v2[8];
        tmem_ld8(d_mine + 64, v2);
        tmem_wait_ld();
        repeat_block<8>(v2, sc, scale2, times, h + 64);
    }
}

template <int IN, int FIN>
__global__ void __launch_bounds__(NTHREADS, 1)
gcn_fused_tc_kernel(const float* __restrict__ x, const int32_t* __restrict__ leader, const int32_t* __restrict__ gsize,
                    const float* __restrict__ labels, const int32_t* __restrict__ ped_start, const int32_t* __restrict__ ped_end,
                    const int32_t* __restrict__ scene_start, const int32_t* __restrict__ chunk_scene, int n_chunks,
                    const float* __restrict__ W0, const float* __restrict__ W1, const float* __restrict__ V0,
                    const float* __restrict__ V1, const float* __restrict__ Wo, const float* __restrict__ bo,
                    float* __restrict__ out, const uint8_t* __restrict__ prep, uint8_t* __restrict__ prep_out) {
    using C = Cfg<IN, FIN>;
    pdl_trigger();       // (sgx_common.cuh: the next kernel of the forward may set itself up while this one drains)
    pdl_wait();          // metadata, x and possibly the weight blob come from launches just before this one
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    float* s_bo = reinterpret_cast<float*>(smem + C::OFF_BO);
    float* s_winv = reinterpret_cast<float*>(smem + C::OFF_WS);
    uint32_t* s_wmax = reinterpret_cast<uint32_t*>(smem + C::OFF_WS + 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + GROUPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp >> 2, wq = warp & 3;
    const int n_tiles = (n_chunks + 3) >> 2;
    const int tile_step = gridDim.x * GROUPS;

    // the first tile's chunk bounds are fetched before the weight prep so that their latency hides behind it
    int tile = blockIdx.x * GROUPS + grp;
    int p0 = 0, p1 = 0;
    if (tile < n_tiles && tile * 4 + wq < n_chunks) {
        p0 = scene_start[chunk_scene[tile * 4 + wq]];
        p1 = scene_start[chunk_scene[tile * 4 + wq + 1]];
    }
    int cs0 = 0, cs1 = 0;                                    // scene bounds of the NEXT tile's chunk (loaded one tile ahead)
    if (tile + tile_step < n_tiles && (tile + tile_step) * 4 + wq < n_chunks) {
        cs0 = chunk_scene[(tile + tile_step) * 4 + wq];
        cs1 = chunk_scene[(tile + tile_step) * 4 + wq + 1];
    }

    uint64_t* wbar = bars + GROUPS + 1;                      // completion of the prepared blob's bulk copy
    if (threadIdx.x == 0) {
        for (int g = 0; g < GROUPS; ++g) mbar_init(&bars[g], 1);
        mbar_init(wbar, 1);
        fence_barrier_init();
        if (prep != nullptr) {                               // in flight behind the TMEM allocation and the first tile's loads
            mbar_expect_tx(wbar, C::OFF_BAR);
            bulk_g2s(smem, prep, C::OFF_BAR, wbar);
        }
    }
    if (prep == nullptr && threadIdx.x < 8) s_wmax[threadIdx.x] = 0u;      // (the blob copy owns the region otherwise)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---------------- weight images: copied from a prepared blob (sgx_gcn_module_tc_prep, cached per weight version by
    // the host side), or built here: fp32 staging [n][k] (in the group buffers), per-matrix scale, hi/lo split ----------------
    if (prep == nullptr) {
        float* stage = reinterpret_cast<float*>(smem + C::OFF_GRP);
        for (int e = threadIdx.x; e < C::STAGE_FLOATS; e += NTHREADS) stage[e] = 0.f;
        __syncthreads();
        uint32_t mx[5] = {0u, 0u, 0u, 0u, 0u};
        auto put = [&](int q, int idx, float v) {
            stage[idx] = v;
            const uint32_t b = __float_as_uint(v) & 0x7fffffffu;
#pragma unroll
            for (int qq = 0; qq < 5; ++qq) if (qq == q) mx[qq] = max(mx[qq], b);
        };
#pragma unroll 8
        for (int e = threadIdx.x; e < IN * HID; e += NTHREADS) put(0, C::S1 + (e % HID) * C::K1 + e / HID, W0[e]);
#pragma unroll 4
        for (int e = threadIdx.x; e < HID * OUT; e += NTHREADS) put(1, C::S2 + (e % OUT) * C::K2 + e / OUT, W1[e]);
#pragma unroll 4
        for (int e = threadIdx.x; e < OUT * HID; e += NTHREADS) put(2, C::S3 + (e % HID) * C::K3 + e / HID, V0[e]);
#pragma unroll 4
        for (int e = threadIdx.x; e < HID * OUT; e += NTHREADS) put(3, C::S4 + (e % OUT) * C::K4 + e / OUT, V1[e]);
#pragma unroll 2
        for (int e = threadIdx.x; e < FIN * 2 * OUT; e += NTHREADS) put(4, C::S5 + e, Wo[e]);     // [n][k] already
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const uint32_t r = __reduce_max_sync(0xffffffffu, mx[q]);
            if (lane == 0 && r) atomicMax(&s_wmax[q], r);
        }
        if (threadIdx.x < FIN) s_bo[threadIdx.x] = bo[threadIdx.x];
        __syncthreads();
        build_image(smem + C::OFF_W1, stage + C::S1, C::N1, C::K1, s_wmax[0], &s_winv[0], threadIdx.x, NTHREADS);
        build_image(smem + C::OFF_W2, stage + C::S2, C::N2, C::K2, s_wmax[1], &s_winv[1], threadIdx.x, NTHREADS);
        build_image(smem + C::OFF_W3, stage + C::S3, C::N3, C::K3, s_wmax[2], &s_winv[2], threadIdx.x, NTHREADS);
        build_image(smem + C::OFF_W4, stage + C::S4, C::N4, C::K4, s_wmax[3], &s_winv[3], threadIdx.x, NTHREADS);
        build_image(smem + C::OFF_W5, stage + C::S5, C::N5, C::K5, s_wmax[4], &s_winv[4], threadIdx.x, NTHREADS);
        fence_proxy_async();
        __syncthreads();
    }

    if (prep_out != nullptr)                                 // prep launch: hand the images out (no tiles: n_chunks = 0)
        for (int e = threadIdx.x; e < C::OFF_BAR / 16; e += NTHREADS)
            reinterpret_cast<uint4*>(prep_out)[e] = reinterpret_cast<const uint4*>(smem)[e];

    // ---------------- tile groups ----------------
    uint8_t* abuf = smem + C::OFF_GRP + grp * GRP_BYTES;     // A operand / fp32 rows of the group's tile
    const int row = wq * 32 + lane;                          // row of the tile = TMEM lane
    uint8_t* arow = abuf + row * 16;                         // this pedestrian's 16 bytes of every core
    const uint8_t* wrows = abuf + wq * 512;                  // first row of this warp, per core
    uint64_t* bar = &bars[grp];
    const uint32_t a_s = sbase + C::OFF_GRP + grp * GRP_BYTES;
    const uint32_t d_tmem = tmem + (uint32_t)grp * 128u;
    const uint32_t d_mine = d_tmem + ((uint32_t)(wq * 32) << 16);
    uint32_t parity = 0;

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

    // per-lane metadata (raw words: nothing is computed from a prefetched word before the tile that uses it) and x row
    struct Meta { int b, e, lead, gs; };
    auto load_meta = [&](int p0_, int np_) {
        Meta m{p0_, p0_, p0_ + lane, 1};
        if (lane < np_) {
            const int p = p0_ + lane;
            m.b = ped_start[p]; m.e = ped_end[p];
            if (labels != nullptr) m.lead = __float_as_int(labels[p]);      // group structure derived in the kernel
            else { m.lead = leader[p]; m.gs = gsize[p]; }
        }
        return m;
    };
    float4 xq[IN / 4];
    auto load_x = [&](int p0_, int np_) {
#pragma unroll
        for (int c = 0; c < IN / 4; ++c) xq[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < np_) {
            const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)(p0_ + lane) * IN);
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) xq[c] = xr[c];
        }
    };
    Meta mt = load_meta(p0, p1 - p0);
    load_x(p0, p1 - p0);
    if (prep != nullptr) mbar_wait(wbar, 0);
    const float winv1 = s_winv[0], winv2 = s_winv[1], winv3 = s_winv[2], winv4 = s_winv[3], winv5 = s_winv[4];

    for (; tile < n_tiles; tile += tile_step) {
        const int np = p1 - p0;
        const bool live = lane < np;
        const int p = p0 + lane;
        const int sb = mt.b - p0, se = mt.e - p0;
        int my_lead, k;
        uint32_t group_mask;
        group_structure(labels != nullptr, live, lane, mt.lead, mt.gs, mt.b, p0, my_lead, k, group_mask);
        const float a = __frcp_rn((float)k);
        const bool is_lead = live && (my_lead == lane);
        const uint32_t scene_mask = (se >= 32 ? 0xffffffffu : ((1u << se) - 1u)) & ~((1u << sb) - 1u);
        const uint32_t lead_ballot = __ballot_sync(0xffffffffu, is_lead);
        const uint32_t leader_mask = live ? (lead_ballot & scene_mask) : (1u << lane);
        const int G = __popc(leader_mask);
        const float cg = __frcp_rn((float)G);
        // ---- x rows -> shared (fp32 quads of the own row), group mean M1 ----
#pragma unroll
        for (int c = 0; c < IN / 4; ++c) *reinterpret_cast<float4*>(arow + c * CORE) = xq[c];
        __syncwarp();
        float sc;
        {
            float m1[IN];
#pragma unroll
            for (int c = 0; c < IN; ++c) m1[c] = 0.f;
            for (uint32_t mm = group_mask; mm; mm &= mm - 1) {
                const uint8_t* r = wrows + (__ffs(mm) - 1) * 16;
#pragma unroll
                for (int c = 0; c < IN / 4; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(r + c * CORE);
                    m1[4 * c] = fmaf(a, v.x, m1[4 * c]); m1[4 * c + 1] = fmaf(a, v.y, m1[4 * c + 1]);
                    m1[4 * c + 2] = fmaf(a, v.z, m1[4 * c + 2]); m1[4 * c + 3] = fmaf(a, v.w, m1[4 * c + 3]);
                }
            }
            __syncwarp();                                    // every lane is done reading the x rows
            sc = row_to_operand<IN, C::K1>(arow, m1) * winv1;
        }
        // the next tile's chunk bounds (its scene indices were loaded during the previous tile) and the scene indices of
        // the tile after it: no load address depends on a word still in flight
        const int ntile = tile + tile_step, nntile = ntile + tile_step;
        int p0n = 0, p1n = 0;
        if (ntile < n_tiles && ntile * 4 + wq < n_chunks) {
            p0n = scene_start[cs0];
            p1n = scene_start[cs1];
        }
        if (nntile < n_tiles && nntile * 4 + wq < n_chunks) {
            cs0 = chunk_scene[nntile * 4 + wq];
            cs1 = chunk_scene[nntile * 4 + wq + 1];
        }
        // ---- intra GCN: H1 = M1 W0, M2 = k-fold sum of relu(H1)/k, X1 = relu(M2 W1) ----
        run_layer([&]() { issue_layer<C::K1, C::N1>(d_tmem, a_s, sbase + C::OFF_W1, bar); });
        {
            float h[HID];
            hidden_from_tmem(d_mine, sc, a, k, h);
            sc = row_to_operand<HID, C::K2>(arow, h) * winv2;
        }
        run_layer([&]() { issue_layer<C::K2, C::N2>(d_tmem, a_s, sbase + C::OFF_W2, bar); });
        float x1[OUT];
        {
            uint32_t v0[16];
            tmem_ld16(d_mine, v0);
            tmem_wait_ld();
            float xg[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) { x1[o] = relu_nan(__uint_as_float(v0[o]) * sc); xg[o] = a * x1[o]; }
            repeat_rows<OUT>(xg, k);
            store_core_row<OUT>(arow, xg);                   // Xg rows (the operand of the last layer is consumed)
        }
        __syncwarp();
        // ---- inter GCN: N1 = mean of the scene's group states, N2 = G-fold sum of relu(N1 V0)/G, Y = relu(N2 V1) ----
        {
            float n1[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) n1[o] = 0.f;
            for (uint32_t mm = leader_mask; mm; mm &= mm - 1) {
                const uint8_t* r = wrows + (__ffs(mm) - 1) * 16;
#pragma unroll
                for (int c = 0; c < OUT / 4; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(r + c * CORE);
                    n1[4 * c] = fmaf(cg, v.x, n1[4 * c]); n1[4 * c + 1] = fmaf(cg, v.y, n1[4 * c + 1]);
                    n1[4 * c + 2] = fmaf(cg, v.z, n1[4 * c + 2]); n1[4 * c + 3] = fmaf(cg, v.w, n1[4 * c + 3]);
                }
            }
            __syncwarp();                                    // every lane is done reading the Xg rows
            sc = row_to_operand<OUT, C::K3>(arow, n1) * winv3;
        }
        run_layer([&]() { issue_layer<C::K3, C::N3>(d_tmem, a_s, sbase + C::OFF_W3, bar); });
        {
            float h[HID];
            hidden_from_tmem(d_mine, sc, cg, G, h);
            sc = row_to_operand<HID, C::K4>(arow, h) * winv4;
        }
        run_layer([&]() { issue_layer<C::K4, C::N4>(d_tmem, a_s, sbase + C::OFF_W4, bar); });
        {
            uint32_t v0[16];
            tmem_ld16(d_mine, v0);
            tmem_wait_ld();
            float cat[2 * OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) { cat[o] = x1[o]; cat[OUT + o] = a * relu_nan(__uint_as_float(v0[o]) * sc); }
            sc = row_to_operand<2 * OUT, C::K5>(arow, cat) * winv5;
        }
        // the next tile's metadata and x row: in flight during the last round trip and the output stores
        const Meta mtn = load_meta(p0n, p1n - p0n);
        load_x(p0n, p1n - p0n);
        run_layer([&]() { issue_layer<C::K5, C::N5>(d_tmem, a_s, sbase + C::OFF_W5, bar); });
        {
            uint32_t v0[32];
            tmem_ld32(d_mine, v0);
            tmem_wait_ld();
            if (live) {
                float4* orow = reinterpret_cast<float4*>(out + (int64_t)p * FIN);
#pragma unroll
                for (int f = 0; f < FIN / 4; ++f)
                    orow[f] = make_float4(fmaf(__uint_as_float(v0[4 * f]), sc, s_bo[4 * f]),
                                          fmaf(__uint_as_float(v0[4 * f + 1]), sc, s_bo[4 * f + 1]),
                                          fmaf(__uint_as_float(v0[4 * f + 2]), sc, s_bo[4 * f + 2]),
                                          fmaf(__uint_as_float(v0[4 * f + 3]), sc, s_bo[4 * f + 3]));
            }
        }
        p0 = p0n; p1 = p1n; mt = mtn;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

}  // namespace gctc

template <int IN, int FIN>
int64_t gcn_tc_prep_bytes() { return gctc::Cfg<IN, FIN>::OFF_BAR; }

template <int IN, int FIN>
int gcn_fused_tc_forward(const float* x, const int32_t* leader, const int32_t* gsize, const float* labels, const int32_t* ps, const int32_t* pe,
                         const int32_t* scene_start, const int32_t* chunk_scene, int n_chunks, const float* W0,
                         const float* W1, const float* V0, const float* V1, const float* Wo, const float* bo, float* out,
                         cudaStream_t st, const void* prep, void* prep_out) {
    using C = gctc::Cfg<IN, FIN>;
    auto kern = gctc::gcn_fused_tc_kernel<IN, FIN>;
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL));
    int dev = 0, sms = 148;
    SGX_CUDA(cudaGetDevice(&dev));
    SGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_tiles = (n_chunks + 3) / 4;
    const int grid = std::max(1, std::min((n_tiles + gctc::GROUPS - 1) / gctc::GROUPS, sms));
    SGX_CUDA(launch_pdl(kern, dim3(grid), dim3(gctc::NTHREADS), C::SMEM_TOTAL, st, true, x, leader, gsize, labels, ps, pe,
                        scene_start, chunk_scene, n_chunks, W0, W1, V0, V1, Wo, bo, out, (const uint8_t*)prep,
                        (uint8_t*)prep_out));
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

#define GCN_TC_INST(I, F)                                                                                                    \
    template int64_t gcn_tc_prep_bytes<I, F>();                                                                              \
    template int gcn_fused_tc_forward<I, F>(const float*, const int32_t*, const int32_t*, const float*, const int32_t*,     \
                                            const int32_t*, const int32_t*, const int32_t*, int, const float*, const float*, \
                                            const float*, const float*, const float*, const float*, float*, cudaStream_t,   \
                                            const void*, void*);
GCN_TC_INST(32, 24)
GCN_TC_INST(32, 32)
GCN_TC_INST(40, 24)
GCN_TC_INST(40, 32)

}  // namespace sgx
