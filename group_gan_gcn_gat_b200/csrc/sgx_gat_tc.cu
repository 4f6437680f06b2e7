// GATEncoder forward (sgan/models.py:254-294; GraphAttentionLayer 198-220, GAT 231-237) with the five linear maps on
// the 5th-gen tensor cores (tcgen05.mma kind::f16, accumulators in TMEM) -- the single-launch path for batches whose
// scenes fit a warp chunk (<= 32 pedestrians), n_heads = 1, dims 40 / 72 / 16 / 24.
//
// The mma.sync kernel (sgx_gat.cu, gat_fused_mma_kernel) spends ~40 % of its instruction stream on the warp-level GEMMs
// (fragment loads, 3xTF32 splits, 816 HMMA per chunk) and is issue bound.  Here a TILE is 128 pedestrians = four warp
// chunks; a tile GROUP of four warps (thread = pedestrian = TMEM lane) walks the layer chain
//     x -> [Wi] -> attention (group) -> ELU -> [Wio] -> attention -> ELU, log_softmax = x1 -> mean over the group = Xg
//       -> [We] -> attention (leaders of the scene) -> ELU -> [Weo] -> attention -> ELU, log_softmax = Yg
//       -> cat(x1, Yg[leader] / |group|) -> [Wo] + bo
// and every [W] is ONE batch of tcgen05.mma over the tile's 128 rows:
//   * fp32-grade accuracy from fp16 operand splits, v = hi + lo (11 + 11 significant bits), three products
//     lo.hi + hi.lo + hi.hi accumulated in fp32 (~7e-7 relative, the level of the 3xTF32 kernel it replaces).  fp16 has a
//     narrow exponent range, so every activation ROW is multiplied by its own power of two (row maximum -> [2^13, 2^14))
//     before the split and every weight matrix by one power of two; both are undone exactly on the way out of TMEM.
//   * operands in the no-swizzle K-major canonical layout: K core kc (8 halves = 16 bytes) of row r at kc * 2048 + r * 16.
//     The SAME 40 KB buffer is, alternately, the A operand of a layer (fp16 cores) and the fp32 row buffer the attention
//     reads its neighbours' Wh rows from (feature quad f of row r at f * 2048 + r * 16): a layer's A operand is dead once
//     its MMAs have completed, and the Wh rows are dead once the tile's warps have built the next A operand.  Both uses
//     touch only the 512-byte pieces of the pedestrian's own warp, so the hand-over needs a __syncwarp, not a barrier.
//   * the attention scores ride along as two extra weight columns (W a1, W a2), as in the mma.sync kernel.
//   * per layer: build the operand rows, fence.proxy.async, a 128-thread named barrier, one elected thread issues
//     3 K/16 MMAs + tcgen05.commit, the group waits on its mbarrier, tcgen05.ld brings the thread's own row back.
// One CTA per SM holds FOUR tile groups (16 warps) that share one 44 KB set of weight images and interleave their MMA
// round trips; TMEM: 128 columns per group.  Every member of a group computes its group's rows of the inter level
// (identical operand rows give identical accumulator rows), so no lane is ever predicated off and nothing is zeroed.  The weight images are built in-kernel from the fp32 parameters (a few
// microseconds per CTA, overlapped across SMs) so the entry point keeps its stateless signature.
#include <type_traits>
#include "sgx_gat_fused.cuh"
#include "sgx_graph_tc.cuh"

namespace sgx {
namespace gtc {

#ifndef GTC_GROUPS
#define GTC_GROUPS 3           // measured: 4 groups (GTC_HALVES, 128 registers, no spills) 98.3 us, 3 groups 96.3 us per call
#endif
#ifndef GTC_HALVES
#define GTC_HALVES (GTC_GROUPS >= 4)
#endif
#ifndef GTC_FASTPATH
#define GTC_FASTPATH 1
#endif
constexpr int GROUPS = GTC_GROUPS;
constexpr int NTHREADS = GROUPS * 128;
constexpr int IN = 40, FIN = 24;
using namespace gtile;

// weight images [hi cores | lo cores], core kc of output column n at kc * N * 16 + n * 16
constexpr int N1 = 80, K1 = 48;          // Wi  + 2 score columns: 40 (padded to 48) -> 72 + 2
constexpr int N2 = 32, K2 = 80;          // Wio + 2 score columns: 72 (padded to 80) -> 16 + 2
constexpr int N3 = 80, K3 = 16;          // We
constexpr int N4 = 32, K4 = 80;          // Weo
constexpr int N5 = 32, K5 = 32;          // Wo^T: cat(32) -> 24
constexpr int OFF_W1 = 0;
constexpr int OFF_W2 = OFF_W1 + 4 * K1 * N1;
constexpr int OFF_W3 = OFF_W2 + 4 * K2 * N2;
constexpr int OFF_W4 = OFF_W3 + 4 * K3 * N3;
constexpr int OFF_W5 = OFF_W4 + 4 * K4 * N4;
constexpr int OFF_BO = OFF_W5 + 4 * K5 * N5;          // float[32]: bo
constexpr int OFF_WA = OFF_BO + 128;                  // float[2][16]: We ae1, We ae2 (the inter level's score vectors, fp32)
constexpr int OFF_WS = OFF_WA + 128;                  // float[8]: inverse weight scales; uint[8]: max |w| bits
constexpr int OFF_BAR = OFF_WS + 64;                  // GROUPS mbarriers + the TMEM base slot
constexpr int OFF_GRP = OFF_BAR + 64;
constexpr int ABUF = 20 * CORE;                       // A operand (K <= 80: 10 hi + 10 lo cores) / 72-wide fp32 rows
constexpr int NCORE = 16;                             // the 16-wide fp32 rows (Wh2 / x1 / Wh4) live in cores 16..19
constexpr int GRP_BYTES = ABUF + 128 * 8;             // + (s, t) per row
constexpr int SMEM_TOTAL = OFF_GRP + GROUPS * GRP_BYTES + 128;
constexpr int STAGE_FLOATS = N1 * K1 + N2 * K2 + N3 * K3 + N4 * K4 + N5 * K5;
static_assert(OFF_GRP % 128 == 0 && GRP_BYTES % 128 == 0, "operand buffers must stay 128-byte aligned");
static_assert(STAGE_FLOATS * 4 <= GROUPS * GRP_BYTES, "the fp32 staging of the weight prep lives in the group buffers");
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory");
static_assert(OFF_BAR % 16 == 0 && 8 * (GROUPS + 2) <= 64, "the prepared blob is one bulk copy; barriers + TMEM slot fit 64 bytes");

// attention of one node over the lanes set in `mask`; neighbour rows in the core layout (quad f of lane q at
// f * CORE + q * 16 from `wrows`, the first row of this warp), scores (s, t) per lane in `st`.  Same term order as
// attend_mask (sgx_gat_fused.cuh).
template <int F>
__device__ __forceinline__ void attend_core(const uint8_t* __restrict__ wrows, const float2* __restrict__ st, uint32_t mask,
                                            float s_i, float alpha, float (&hp)[F]) {
    // max_j lrelu(s_i + t_j) = lrelu(s_i + max_j t_j) for alpha >= 0 (both roundings are monotone): the first pass is one
    // shared-memory read and one max per neighbour
    float m = -INFINITY;
    if (alpha >= 0.f) {
        for (uint32_t mm = mask; mm; mm &= mm - 1) m = fmaxf(m, st[__ffs(mm) - 1].y);
        m = lrelu(s_i + m, alpha);
    } else {
        for (uint32_t mm = mask; mm; mm &= mm - 1) m = fmaxf(m, lrelu(s_i + st[__ffs(mm) - 1].y, alpha));
    }
    float den = 0.f;
#pragma unroll
    for (int f = 0; f < F; ++f) hp[f] = 0.f;
    // two neighbours per trip: their score -> exp chains are independent, the sums keep the ascending-lane order
    for (uint32_t mm = mask; mm;) {
        const int q0 = __ffs(mm) - 1;
        mm &= mm - 1;
        const bool two = mm != 0u;
        const int q1 = two ? __ffs(mm) - 1 : q0;
        mm &= mm - 1;
        const float w0 = fexp(lrelu(s_i + st[q0].y, alpha) - m);
        const float w1 = fexp(lrelu(s_i + st[q1].y, alpha) - m);
        const uint8_t* row0 = wrows + q0 * 16;
        const uint8_t* row1 = wrows + q1 * 16;
        den += w0;
#pragma unroll
        for (int f = 0; f < F / 4; ++f) {
            const float4 v = *reinterpret_cast<const float4*>(row0 + f * CORE);
            hp[4 * f] = fmaf(w0, v.x, hp[4 * f]); hp[4 * f + 1] = fmaf(w0, v.y, hp[4 * f + 1]);
            hp[4 * f + 2] = fmaf(w0, v.z, hp[4 * f + 2]); hp[4 * f + 3] = fmaf(w0, v.w, hp[4 * f + 3]);
        }
        if (two) {
            den += w1;
#pragma unroll
            for (int f = 0; f < F / 4; ++f) {
                const float4 v = *reinterpret_cast<const float4*>(row1 + f * CORE);
                hp[4 * f] = fmaf(w1, v.x, hp[4 * f]); hp[4 * f + 1] = fmaf(w1, v.y, hp[4 * f + 1]);
                hp[4 * f + 2] = fmaf(w1, v.z, hp[4 * f + 2]); hp[4 * f + 3] = fmaf(w1, v.w, hp[4 * f + 3]);
            }
        }
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int f = 0; f < F; ++f) hp[f] *= inv;
}

// ---- the 72-wide attention layers in two feature blocks (GTC_HALVES): 72 accumulators per thread cap the kernel at 168
// registers = three tile groups per SM; in blocks of 32 + 40 features (the exponentials are recomputed per block, the first
// block's ELU outputs wait in 32 spare TMEM columns of the thread's own lane) the kernel fits 128 registers = FOUR groups.
// Built to test whether a fourth group hides more latency: it does not (98.3 us against 96.3 us, bit-identical results),
// so the kernel is not bound by the number of resident warps; kept as a compile-time variant (-DGTC_GROUPS=4). ----
template <int Q0, int NQ>          // feature quads [Q0, Q0 + NQ) of the neighbour rows
__device__ __forceinline__ void attend_block(const uint8_t* __restrict__ wrows, const float2* __restrict__ st, uint32_t mask,
                                             float s_i, float m, float alpha, float (&hp)[4 * NQ]) {
    float den = 0.f;
#pragma unroll
    for (int f = 0; f < 4 * NQ; ++f) hp[f] = 0.f;
    for (uint32_t mm = mask; mm; mm &= mm - 1) {
        const int q = __ffs(mm) - 1;
        const float w = fexp(lrelu(s_i + st[q].y, alpha) - m);
        den += w;
        const uint8_t* row = wrows + q * 16 + Q0 * CORE;
#pragma unroll
        for (int f = 0; f < NQ; ++f) {
            const float4 v = *reinterpret_cast<const float4*>(row + f * CORE);
            hp[4 * f] = fmaf(w, v.x, hp[4 * f]); hp[4 * f + 1] = fmaf(w, v.y, hp[4 * f + 1]);
            hp[4 * f + 2] = fmaf(w, v.z, hp[4 * f + 2]); hp[4 * f + 3] = fmaf(w, v.w, hp[4 * f + 3]);
        }
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int f = 0; f < 4 * NQ; ++f) hp[f] = felu(hp[f] * inv);
}

// hi / lo K cores [C0, C0 + N/8) of an operand with KC hi cores from N values
template <int N, bool SCALED>
__device__ __forceinline__ void write_core_block(uint8_t* __restrict__ arow, int c0, int kc_total, const float (&v)[N], float s) {
#pragma unroll
    for (int kc = 0; kc < N / 8; ++kc) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a0 = v[kc * 8 + 2 * j], a1 = v[kc * 8 + 2 * j + 1];
            if (SCALED) { a0 *= s; a1 *= s; }
            const uint32_t h = pack_f16_rn(a0, a1);
            float l0, l1;
            sub_f16x2(h, a0, a1, l0, l1);
            hi[j] = h;
            lo[j] = pack_f16_rn(l0, l1);
        }
        *reinterpret_cast<uint4*>(arow + (c0 + kc) * CORE) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(arow + (kc_total + c0 + kc) * CORE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// attention + ELU of a 72-wide layer -> the K = 80 operand of the next linear map; returns the inverse row scale
__device__ __forceinline__ float wide_attend_to_operand(const uint8_t* __restrict__ wrows, const float2* __restrict__ st,
                                                        uint32_t mask, float s_i, float alpha, uint8_t* __restrict__ arow,
                                                        uint32_t spare_tmem) {
    float m = -INFINITY;
    for (uint32_t mm = mask; mm; mm &= mm - 1) m = fmaxf(m, lrelu(s_i + st[__ffs(mm) - 1].y, alpha));
    float mxb = 0.f;
    {
        float hb[32];                                        // features 40..71
        attend_block<10, 8>(wrows, st, mask, s_i, m, alpha, hb);
        uint32_t raw[32];
#pragma unroll
        for (int f = 0; f < 32; ++f) { mxb = fmaxf(mxb, fabsf(hb[f])); raw[f] = __float_as_uint(hb[f]); }
        tmem_st32(spare_tmem, raw);
    }
    float ha[40];                                            // features 0..39
    attend_block<0, 10>(wrows, st, mask, s_i, m, alpha, ha);
    float mx = mxb;
#pragma unroll
    for (int f = 0; f < 40; ++f) mx = fmaxf(mx, fabsf(ha[f]));
    tmem_wait_st();
    __syncwarp();                                            // every lane is done reading the Wh rows
    const bool unscaled = GTC_FASTPATH && __all_sync(0xffffffffu, scale_free(mx));
    float s = 1.f, inv = 1.f;
    if (!unscaled) pow2_scale(mx, s, inv);
    if (unscaled) write_core_block<40, false>(arow, 0, 10, ha, 1.f); else write_core_block<40, true>(arow, 0, 10, ha, s);
    {
        uint32_t raw[32];
        tmem_ld32(spare_tmem, raw);
        tmem_wait_ld();
        float hb[32];
#pragma unroll
        for (int f = 0; f < 32; ++f) hb[f] = __uint_as_float(raw[f]);
        if (unscaled) write_core_block<32, false>(arow, 5, 10, hb, 1.f); else write_core_block<32, true>(arow, 5, 10, hb, s);
    }
    *reinterpret_cast<uint4*>(arow + 9 * CORE) = make_uint4(0u, 0u, 0u, 0u);      // K padding 72..79
    *reinterpret_cast<uint4*>(arow + 19 * CORE) = make_uint4(0u, 0u, 0u, 0u);
    return inv;
}

// the thread's accumulator row (72 + 2 columns) out of TMEM -> fp32 quads of its own row + (s, t)
__device__ __forceinline__ void wide_from_tmem(uint32_t d_mine, float sc, uint8_t* __restrict__ arow, float2& sv) {
    uint32_t v0[32], v1[32], v2[8], v3[2];
    tmem_ld32(d_mine, v0);
    tmem_ld32(d_mine + 32, v1);
    tmem_ld8(d_mine + 64, v2);
    tmem_ld2(d_mine + 72, v3);
    tmem_wait_ld();
    if (GTC_FASTPATH && __all_sync(0xffffffffu, sc == 1.f)) {
#pragma unroll
        for (int f = 0; f < 8; ++f) *reinterpret_cast<uint4*>(arow + f * CORE) = make_uint4(v0[4 * f], v0[4 * f + 1], v0[4 * f + 2], v0[4 * f + 3]);
#pragma unroll
        for (int f = 0; f < 8; ++f) *reinterpret_cast<uint4*>(arow + (8 + f) * CORE) = make_uint4(v1[4 * f], v1[4 * f + 1], v1[4 * f + 2], v1[4 * f + 3]);
#pragma unroll
        for (int f = 0; f < 2; ++f) *reinterpret_cast<uint4*>(arow + (16 + f) * CORE) = make_uint4(v2[4 * f], v2[4 * f + 1], v2[4 * f + 2], v2[4 * f + 3]);
        sv = make_float2(__uint_as_float(v3[0]), __uint_as_float(v3[1]));
    } else {
#pragma unroll
        for (int f = 0; f < 8; ++f)
            *reinterpret_cast<float4*>(arow + f * CORE) =
                make_float4(__uint_as_float(v0[4 * f]) * sc, __uint_as_float(v0[4 * f + 1]) * sc,
                            __uint_as_float(v0[4 * f + 2]) * sc, __uint_as_float(v0[4 * f + 3]) * sc);
#pragma unroll
        for (int f = 0; f < 8; ++f)
            *reinterpret_cast<float4*>(arow + (8 + f) * CORE) =
                make_float4(__uint_as_float(v1[4 * f]) * sc, __uint_as_float(v1[4 * f + 1]) * sc,
                            __uint_as_float(v1[4 * f + 2]) * sc, __uint_as_float(v1[4 * f + 3]) * sc);
#pragma unroll
        for (int f = 0; f < 2; ++f)
            *reinterpret_cast<float4*>(arow + (16 + f) * CORE) =
                make_float4(__uint_as_float(v2[4 * f]) * sc, __uint_as_float(v2[4 * f + 1]) * sc,
                            __uint_as_float(v2[4 * f + 2]) * sc, __uint_as_float(v2[4 * f + 3]) * sc);
        sv = make_float2(__uint_as_float(v3[0]) * sc, __uint_as_float(v3[1]) * sc);
    }
}
// the thread's 72 accumulator columns -> ELU -> the K = 80 operand of the next linear map; returns the inverse row scale.
// Two passes over TMEM in blocks (never 72 live values): the row scale comes from a bound on max |ELU(v)| that needs no
// exponential, v for v > 0 and min(|v|, 1) otherwise (within a factor 1.6 of the true maximum: at most one bit of the
// fp16 headroom), then each block is read again, activated and split.
__device__ __forceinline__ float wide_elu_to_operand(uint32_t d_mine, float sc, uint8_t* __restrict__ arow) {
    float mx = 0.f;
    {
        uint32_t v0[32], v1[32], v2[8];
        tmem_ld32(d_mine, v0);
        tmem_ld32(d_mine + 32, v1);
        tmem_ld8(d_mine + 64, v2);
        tmem_wait_ld();
        auto bound = [&](uint32_t raw) {
            const float v = __uint_as_float(raw) * sc;
            mx = fmaxf(mx, v > 0.f ? v : fminf(-v, 1.f));     // NaN rows: fmaxf drops the NaN, the values carry it
        };
#pragma unroll
        for (int f = 0; f < 32; ++f) { bound(v0[f]); bound(v1[f]); }
#pragma unroll
        for (int f = 0; f < 8; ++f) bound(v2[f]);
    }
    const bool unscaled = GTC_FASTPATH && __all_sync(0xffffffffu, scale_free(mx));
    float s = 1.f, inv = 1.f;
    if (!unscaled) pow2_scale(mx, s, inv);
    auto block = [&](auto tag, int col, int core) {
        constexpr int NB = decltype(tag)::value;
        uint32_t raw[NB];
        if constexpr (NB == 32) tmem_ld32(d_mine + col, raw); else tmem_ld8(d_mine + col, raw);
        tmem_wait_ld();
        float h[NB];
#pragma unroll
        for (int f = 0; f < NB; ++f) h[f] = felu(__uint_as_float(raw[f]) * sc);
        if (unscaled) write_core_block<NB, false>(arow, core, 10, h, 1.f); else write_core_block<NB, true>(arow, core, 10, h, s);
    };
    block(std::integral_constant<int, 32>{}, 0, 0);
    block(std::integral_constant<int, 32>{}, 32, 4);
    block(std::integral_constant<int, 8>{}, 64, 8);
    *reinterpret_cast<uint4*>(arow + 9 * CORE) = make_uint4(0u, 0u, 0u, 0u);      // K padding 72..79
    *reinterpret_cast<uint4*>(arow + 19 * CORE) = make_uint4(0u, 0u, 0u, 0u);
    return inv;
}
// ... (16 + 2 columns) -> fp32 quads NCORE.. of its own row + (s, t)
__device__ __forceinline__ void narrow_from_tmem(uint32_t d_mine, float sc, uint8_t* __restrict__ nrow, float2& sv) {
    uint32_t v0[16], v3[2];
    tmem_ld16(d_mine, v0);
    tmem_ld2(d_mine + 16, v3);
    tmem_wait_ld();
#pragma unroll
    for (int f = 0; f < 4; ++f)
        *reinterpret_cast<float4*>(nrow + f * CORE) =
            make_float4(__uint_as_float(v0[4 * f]) * sc, __uint_as_float(v0[4 * f + 1]) * sc,
                        __uint_as_float(v0[4 * f + 2]) * sc, __uint_as_float(v0[4 * f + 3]) * sc);
    sv = make_float2(__uint_as_float(v3[0]) * sc, __uint_as_float(v3[1]) * sc);
}

__global__ void __launch_bounds__(NTHREADS, 1)
gat_fused_tc_kernel(const float* __restrict__ x, const int32_t* __restrict__ leader, const int32_t* __restrict__ gsize,
                    const float* __restrict__ labels, const int32_t* __restrict__ ped_start, const int32_t* __restrict__ ped_end,
                    const int32_t* __restrict__ scene_start, const int32_t* __restrict__ chunk_scene, int n_chunks,
                    const float* __restrict__ Wi, const float* __restrict__ ai, const float* __restrict__ Wio,
                    const float* __restrict__ aio, const float* __restrict__ We, const float* __restrict__ ae,
                    const float* __restrict__ Weo, const float* __restrict__ aeo, const float* __restrict__ Wo,
                    const float* __restrict__ bo, float alpha, float* __restrict__ out,
                    const uint8_t* __restrict__ prep, uint8_t* __restrict__ prep_out) {
    pdl_trigger();       // (sgx_common.cuh: the next kernel of the forward may set itself up while this one drains)
    pdl_wait();          // metadata, x and possibly the weight blob come from launches just before this one
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    float* s_bo = reinterpret_cast<float*>(smem + OFF_BO);
    float* s_wa = reinterpret_cast<float*>(smem + OFF_WA);
    float* s_winv = reinterpret_cast<float*>(smem + OFF_WS);
    uint32_t* s_wmax = reinterpret_cast<uint32_t*>(smem + OFF_WS + 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + GROUPS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp >> 2, wq = warp & 3;
    const int n_tiles = (n_chunks + 3) >> 2;
    const int tile_step = gridDim.x * GROUPS;

    // the first tile's metadata is fetched before the weight prep so that its latency hides behind it
    int tile = blockIdx.x * GROUPS + grp;
    int p0 = 0, p1 = 0;                                      // first pedestrian of the chunk, one past its last
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
            mbar_expect_tx(wbar, OFF_BAR);
            bulk_g2s(smem, prep, OFF_BAR, wbar);
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

    // ---------------- weight images: copied from a prepared blob (sgx_gat_encoder_tc_prep, cached per weight version by
    // the host side), or built here: fp32 staging [n][k] (in the group buffers), per-matrix scale, hi/lo split ----------------
    if (prep == nullptr) {
        float* stage = reinterpret_cast<float*>(smem + OFF_GRP);
        constexpr int S1 = 0, S2 = S1 + N1 * K1, S3 = S2 + N2 * K2, S4 = S3 + N3 * K3, S5 = S4 + N4 * K4;
        for (int e = threadIdx.x; e < STAGE_FLOATS; e += NTHREADS) stage[e] = 0.f;
        __syncthreads();
        uint32_t mx[5] = {0u, 0u, 0u, 0u, 0u};
        auto put = [&](int q, int idx, float v) {
            stage[idx] = v;
            const uint32_t b = __float_as_uint(v) & 0x7fffffffu;
#pragma unroll
            for (int qq = 0; qq < 5; ++qq) if (qq == q) mx[qq] = max(mx[qq], b);
        };
        // plain entries, global reads coalesced along the parameter's rows
#pragma unroll 8
        for (int e = threadIdx.x; e < IN * HID; e += NTHREADS) put(0, S1 + (e % HID) * K1 + e / HID, Wi[e]);
#pragma unroll 4
        for (int e = threadIdx.x; e < HID * OUT; e += NTHREADS) put(1, S2 + (e % OUT) * K2 + e / OUT, Wio[e]);
#pragma unroll 4
        for (int e = threadIdx.x; e < OUT * HID; e += NTHREADS) put(2, S3 + (e % HID) * K3 + e / HID, We[e]);
#pragma unroll 4
        for (int e = threadIdx.x; e < HID * OUT; e += NTHREADS) put(3, S4 + (e % OUT) * K4 + e / OUT, Weo[e]);
#pragma unroll 2
        for (int e = threadIdx.x; e < FIN * 2 * OUT; e += NTHREADS) put(4, S5 + e, Wo[e]);     // [n][k] already
        // score columns W a1, W a2: the 72-long dots by warps, the 16-long ones by threads
        for (int d = warp; d < 2 * (IN + OUT); d += NTHREADS / 32) {
            const bool first = d < 2 * IN;
            const int k = (first ? d : d - 2 * IN) >> 1, which = d & 1;
            const float* wrow = (first ? Wi : We) + k * HID;
            const float* av = (first ? ai : ae) + which * HID;
            float part = 0.f;
#pragma unroll
            for (int c = lane; c < HID; c += 32) part = fmaf(wrow[c], av[c], part);
            part = warp_sum(part);
            if (lane == 0) {
                put(first ? 0 : 2, (first ? S1 + (HID + which) * K1 : S3 + (HID + which) * K3) + k, part);
                if (!first) s_wa[which * OUT + k] = part;
            }
        }
        for (int d = threadIdx.x; d < 4 * HID; d += NTHREADS) {
            const bool first = d < 2 * HID;
            const int k = (first ? d : d - 2 * HID) >> 1, which = d & 1;
            const float* wrow = (first ? Wio : Weo) + k * OUT;
            const float* av = (first ? aio : aeo) + which * OUT;
            float v = 0.f;
#pragma unroll
            for (int c = 0; c < OUT; ++c) v = fmaf(wrow[c], av[c], v);
            put(first ? 1 : 3, (first ? S2 + (OUT + which) * K2 : S4 + (OUT + which) * K4) + k, v);
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const uint32_t r = __reduce_max_sync(0xffffffffu, mx[q]);
            if (lane == 0 && r) atomicMax(&s_wmax[q], r);
        }
        if (threadIdx.x < FIN) s_bo[threadIdx.x] = bo[threadIdx.x];
        __syncthreads();
        // units of one (matrix, column n, K core kc): 8 values -> one hi and one lo 16-byte core
        constexpr int U1 = 0, U2 = U1 + N1 * K1 / 8, U3 = U2 + N2 * K2 / 8, U4 = U3 + N3 * K3 / 8, U5 = U4 + N4 * K4 / 8,
                      UE = U5 + N5 * K5 / 8;
        for (int u = threadIdx.x; u < UE; u += NTHREADS) {
            int q, N, K, off, sb, ub;
            if (u < U2) { q = 0; N = N1; K = K1; off = OFF_W1; sb = S1; ub = U1; }
            else if (u < U3) { q = 1; N = N2; K = K2; off = OFF_W2; sb = S2; ub = U2; }
            else if (u < U4) { q = 2; N = N3; K = K3; off = OFF_W3; sb = S3; ub = U3; }
            else if (u < U5) { q = 3; N = N4; K = K4; off = OFF_W4; sb = S4; ub = U4; }
            else { q = 4; N = N5; K = K5; off = OFF_W5; sb = S5; ub = U5; }
            const int KC = K / 8;
            const int kc = (u - ub) / N, n = (u - ub) % N;
            const float wm = __uint_as_float(s_wmax[q]);
            float s = 1.f, inv = 1.f;
            if (!scale_free(wm)) pow2_scale(wm, s, inv);
            const float* src = stage + sb + n * K + kc * 8;
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a0 = src[2 * j] * s, a1 = src[2 * j + 1] * s;
                const uint32_t h = pack_f16_rn(a0, a1);
                float l0, l1;
                sub_f16x2(h, a0, a1, l0, l1);
                hi[j] = h;
                lo[j] = pack_f16_rn(l0, l1);
            }
            *reinterpret_cast<uint4*>(smem + off + kc * N * 16 + n * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(smem + off + (KC + kc) * N * 16 + n * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            if (n == 0 && kc == 0) s_winv[q] = inv;
        }
        fence_proxy_async();
        __syncthreads();
    }

    if (prep_out != nullptr)                                 // prep launch: hand the images out (no tiles: n_chunks = 0)
        for (int e = threadIdx.x; e < OFF_BAR / 16; e += NTHREADS)
            reinterpret_cast<uint4*>(prep_out)[e] = reinterpret_cast<const uint4*>(smem)[e];

    // ---------------- tile groups ----------------
    uint8_t* abuf = smem + OFF_GRP + grp * GRP_BYTES;        // A operand / fp32 rows of the group's tile
    float2* st = reinterpret_cast<float2*>(abuf + ABUF) + wq * 32;
    const int row = wq * 32 + lane;                          // row of the tile = TMEM lane
    uint8_t* arow = abuf + row * 16;                         // this pedestrian's 16 bytes of every core
    uint8_t* nrow = arow + NCORE * CORE;
    const uint8_t* wrows_a = abuf + wq * 512;                // first row of this warp, per core
    const uint8_t* wrows_n = wrows_a + NCORE * CORE;
    uint64_t* bar = &bars[grp];
    const uint32_t a_s = sbase + OFF_GRP + grp * GRP_BYTES;
    const uint32_t d_tmem = tmem + (uint32_t)grp * 128u;
    const uint32_t d_mine = d_tmem + ((uint32_t)(wq * 32) << 16);
    uint32_t parity = 0;

    // hand the operand rows to the tensor core, wait for the accumulator
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

    // per-lane metadata and the x row of a chunk; lanes past the chunk get an all-zero row and a one-lane neighbourhood
    // (raw values: nothing is computed from a prefetched word before the tile that uses it, a dependent instruction
    // would stall the in-order warp on the load)
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
        int my_lead, gs;
        uint32_t group_mask;
        group_structure(labels != nullptr, live, lane, mt.lead, mt.gs, mt.b, p0, my_lead, gs, group_mask);
        const float inv_g = __frcp_rn((float)gs);
        const bool is_lead = live && (my_lead == lane);
        const uint32_t scene_mask = (se >= 32 ? 0xffffffffu : ((1u << se) - 1u)) & ~((1u << sb) - 1u);
        const uint32_t lead_ballot = __ballot_sync(0xffffffffu, is_lead);
        const uint32_t leader_mask = live ? (lead_ballot & scene_mask) : (1u << lane);
        float sc;
        {
            float xv[IN];
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) { xv[4 * c] = xq[c].x; xv[4 * c + 1] = xq[c].y; xv[4 * c + 2] = xq[c].z; xv[4 * c + 3] = xq[c].w; }
            sc = row_to_operand<IN, K1>(arow, xv) * winv1;
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

        // ---- intra GAT, layer 1: Wh1 = x Wi (+ scores) ----
        run_layer([&]() { issue_layer<K1, N1>(d_tmem, a_s, sbase + OFF_W1, bar); });
        float2 sv;
        wide_from_tmem(d_mine, sc, arow, sv);
        st[lane] = sv;
        __syncwarp();
#if GTC_HALVES
        sc = wide_attend_to_operand(wrows_a, st, group_mask, sv.x, alpha, arow, d_mine + 80) * winv2;
#else
        {
            float hp[HID];
            attend_core<HID>(wrows_a, st, group_mask, sv.x, alpha, hp);
#pragma unroll
            for (int f = 0; f < HID; ++f) hp[f] = felu(hp[f]);
            __syncwarp();                                    // every lane is done reading the Wh1 rows
            sc = row_to_operand<HID, K2>(arow, hp) * winv2;
        }
#endif
        // ---- intra GAT, out_att: Wh2 = x1a Wio (+ scores) ----
        run_layer([&]() { issue_layer<K2, N2>(d_tmem, a_s, sbase + OFF_W2, bar); });
        narrow_from_tmem(d_mine, sc, nrow, sv);
        st[lane] = sv;
        __syncwarp();
        float x1[OUT];
        attend_core<OUT>(wrows_n, st, group_mask, sv.x, alpha, x1);
        elu_logsoftmax<OUT>(x1);
        __syncwarp();                                        // every lane is done reading the Wh2 rows
        store_core_row<OUT>(nrow, x1);
        __syncwarp();
        // ---- GPool: Xg = mean of x1 over the group (every member computes its group's row) ----
        {
            float xg[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) xg[o] = 0.f;
            for (uint32_t mm = group_mask; mm; mm &= mm - 1) {
                const int q = __ffs(mm) - 1;
#pragma unroll
                for (int f = 0; f < OUT / 4; ++f) {
                    const float4 v = *reinterpret_cast<const float4*>(wrows_n + f * CORE + q * 16);
                    xg[4 * f] = fmaf(inv_g, v.x, xg[4 * f]); xg[4 * f + 1] = fmaf(inv_g, v.y, xg[4 * f + 1]);
                    xg[4 * f + 2] = fmaf(inv_g, v.z, xg[4 * f + 2]); xg[4 * f + 3] = fmaf(inv_g, v.w, xg[4 * f + 3]);
                }
            }
            // ---- inter GAT, layer 1, aggregated BEFORE the linear map: sum_j a_ij (Xg_j We) = (sum_j a_ij Xg_j) We, so the
            //      attention over the scene's leaders runs on the 16-wide Xg rows instead of the 72-wide Wh3 rows (4.5x fewer
            //      FMAs and shared-memory reads in the kernel's heaviest loop).  The scores Xg . (We a) come straight from
            //      the fp32 vectors the prep keeps; every member of a group holds its leader's row and computes the same
            //      result. ----
            float s3 = 0.f, t3 = 0.f;
#pragma unroll
            for (int o = 0; o < OUT; ++o) { s3 = fmaf(xg[o], s_wa[o], s3); t3 = fmaf(xg[o], s_wa[OUT + o], t3); }
            __syncwarp();                                    // every lane is done reading the x1 rows
            store_core_row<OUT>(nrow, xg);
            st[lane] = make_float2(s3, t3);
            __syncwarp();
            float xb[OUT];
            attend_core<OUT>(wrows_n, st, leader_mask, s3, alpha, xb);
            sc = row_to_operand<OUT, K3>(arow, xb) * winv3;
        }
        run_layer([&]() { issue_layer<K3, N3>(d_tmem, a_s, sbase + OFF_W3, bar); });
        sc = wide_elu_to_operand(d_mine, sc, arow) * winv4;
        // ---- inter GAT, out_att: Wh4 = hp Weo (+ scores) ----
        run_layer([&]() { issue_layer<K4, N4>(d_tmem, a_s, sbase + OFF_W4, bar); });
        narrow_from_tmem(d_mine, sc, nrow, sv);
        st[lane] = sv;
        __syncwarp();
        // ---- unpool: cat = [x1 | Yg / |group|] (Yg of the own group, computed by every member), out = cat Wo^T + bo ----
        {
            float cat[2 * OUT], yg[OUT];
            attend_core<OUT>(wrows_n, st, leader_mask, sv.x, alpha, yg);
            elu_logsoftmax<OUT>(yg);
#pragma unroll
            for (int o = 0; o < OUT; ++o) { cat[o] = x1[o]; cat[OUT + o] = inv_g * yg[o]; }
            __syncwarp();                                    // the Wh4 rows (cores 16..19) are dead: cat may not overlap,
            sc = row_to_operand<2 * OUT, K5>(arow, cat) * winv5;   // but the next tile's operand will
        }
        // the next tile's metadata and x row: in flight during the last round trip and the output stores
        const Meta mtn = load_meta(p0n, p1n - p0n);
        load_x(p0n, p1n - p0n);
        run_layer([&]() { issue_layer<K5, N5>(d_tmem, a_s, sbase + OFF_W5, bar); });
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
        // the next tile's first operand overwrites the buffer: every warp of the group has read its accumulator (the
        // tcgen05.wait::ld above) before it reaches the next barrier, and the fp32 rows / st are warp-private
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

}  // namespace gtc

int64_t gat_tc_prep_bytes() { return gtc::OFF_BAR; }

// labels == nullptr: group structure from leader / gsize; prep (nullable): blob of gat_tc_prep; prep_out: build-only launch
int gat_fused_tc_forward(const float* x, const int32_t* leader, const int32_t* gsize, const float* labels, const int32_t* ps, const int32_t* pe,
                         const int32_t* scene_start, const int32_t* chunk_scene, int n_chunks, const float* Wi,
                         const float* ai, const float* Wio, const float* aio, const float* We, const float* ae,
                         const float* Weo, const float* aeo, const float* Wo, const float* bo, float alpha, float* out,
                         cudaStream_t st, const void* prep, void* prep_out) {
    auto kern = gtc::gat_fused_tc_kernel;
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gtc::SMEM_TOTAL));
    int dev = 0, sms = 148;
    SGX_CUDA(cudaGetDevice(&dev));
    SGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int n_tiles = (n_chunks + 3) / 4;
    const int grid = std::max(1, std::min((n_tiles + gtc::GROUPS - 1) / gtc::GROUPS, sms));
    SGX_CUDA(launch_pdl(kern, dim3(grid), dim3(gtc::NTHREADS), gtc::SMEM_TOTAL, st, true, x, leader, gsize, labels, ps, pe,
                        scene_start, chunk_scene, n_chunks, Wi, ai, Wio, aio, We, ae, Weo, aeo, Wo, bo, alpha, out,
                        (const uint8_t*)prep, (uint8_t*)prep_out));
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

}  // namespace sgx
