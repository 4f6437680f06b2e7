// Pieces shared by the tcgen05 GATEncoder / GCNModule forwards (sgx_gat_tc.cu, sgx_gcn_tc.cu): a 128-pedestrian tile per
// four-warp group (thread = pedestrian = TMEM lane), operands as fp16 hi + lo splits in the no-swizzle K-major canonical
// layout (K core kc of row r at kc * CORE + r * 16; the same bytes double as fp32 row storage, feature quad f of row r
// at f * CORE + r * 16), per-row power-of-two scaling, 3 K/16 MMAs per linear map.
#pragma once
#include "sgx_tc.cuh"

namespace sgx {
namespace gtile {

#ifndef GTC_FASTPATH
#define GTC_FASTPATH 1
#endif
constexpr int CORE = 2048;               // bytes of one K core (16 B) over the 128 rows of a tile

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t (&v)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// power-of-two scale that brings a maximum of magnitude `m` into [2^13, 2^14) -- or exactly 1 when m is already in
// [2^-2, 2^15) (or zero): fp16 hi + lo then carries >= 22 significant bits of the row maximum without any scaling
__device__ __forceinline__ bool scale_free(float m) { return (m >= 0.25f && m < 32768.f) || m == 0.f; }
__device__ __forceinline__ void pow2_scale(float m, float& s, float& inv) {
    const int e = (int)((__float_as_uint(m) >> 23) & 0xffu);
    const int se = min(267 - e, 253);
    s = __uint_as_float((uint32_t)se << 23);
    inv = __uint_as_float((uint32_t)(254 - se) << 23);
}

template <int K, int KP, bool SCALED>
__device__ __forceinline__ void write_cores(uint8_t* __restrict__ arow, const float (&v)[K], float s) {
    constexpr int KC = KP / 8;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k0 = kc * 8 + 2 * j;
            float a0 = (k0 < K) ? v[k0 < K ? k0 : 0] : 0.f;
            float a1 = (k0 + 1 < K) ? v[k0 + 1 < K ? k0 + 1 : 0] : 0.f;
            if (SCALED) { a0 *= s; a1 *= s; }
            const uint32_t h = pack_f16_rn(a0, a1);
            float l0, l1;
            sub_f16x2(h, a0, a1, l0, l1);
            hi[j] = h;
            lo[j] = pack_f16_rn(l0, l1);
        }
        *reinterpret_cast<uint4*>(arow + kc * CORE) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(arow + (KC + kc) * CORE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// One activation row -> fp16 hi / lo K cores of the A operand; returns the factor that undoes the row's scale.  The
// rows of a warp take the unscaled path together when every row maximum is inside the scale-free range (the usual case).
template <int K, int KP>
__device__ __forceinline__ float row_to_operand(uint8_t* __restrict__ arow, const float (&v)[K]) {
    float m = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) m = fmaxf(m, fabsf(v[k]));
#if GTC_FASTPATH
    if (__all_sync(0xffffffffu, scale_free(m))) {
        write_cores<K, KP, false>(arow, v, 1.f);
        return 1.f;
    }
#endif
    float s, inv;
    pow2_scale(m, s, inv);
    write_cores<K, KP, true>(arow, v, s);
    return inv;
}

// D[128 x N] = A . B^T for one tile: 3 K/16 MMAs (lo.hi, hi.lo, hi.hi) + commit; called by ONE thread
template <int KP, int N>
__device__ __forceinline__ void issue_layer(uint32_t d_tmem, uint32_t a_s, uint32_t b_s, uint64_t* bar) {
    constexpr int KC = KP / 8, KS = KP / 16;
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t a_hi = make_desc_ns(a_s, CORE, 128), a_lo = make_desc_ns(a_s + KC * CORE, CORE, 128);
    const uint64_t b_hi = make_desc_ns(b_s, N * 16, 128), b_lo = make_desc_ns(b_s + KC * N * 16, N * 16, 128);
#pragma unroll
    for (int s = 0; s < KS; ++s)
        mma_ss(d_tmem, a_lo + (uint64_t)(s * (2 * CORE / 16)), b_hi + (uint64_t)(s * (2 * N * 16 / 16)), idesc, s > 0);
#pragma unroll
    for (int s = 0; s < KS; ++s)
        mma_ss(d_tmem, a_hi + (uint64_t)(s * (2 * CORE / 16)), b_lo + (uint64_t)(s * (2 * N * 16 / 16)), idesc, 1);
#pragma unroll
    for (int s = 0; s < KS; ++s)
        mma_ss(d_tmem, a_hi + (uint64_t)(s * (2 * CORE / 16)), b_hi + (uint64_t)(s * (2 * N * 16 / 16)), idesc, 1);
    tc_commit(bar);
}

template <int F>
__device__ __forceinline__ void store_core_row(uint8_t* __restrict__ row, const float (&v)[F]) {
#pragma unroll
    for (int f = 0; f < F / 4; ++f)
        *reinterpret_cast<float4*>(row + f * CORE) = make_float4(v[4 * f], v[4 * f + 1], v[4 * f + 2], v[4 * f + 3]);
}


// Group structure of one warp chunk (lanes <-> pedestrians of whole scenes), sgan/models.py:263-267: the members of a
// pedestrian's group are the pedestrians of ITS scene with the same non-zero label (float ==, so -0.0 and NaN labels
// stand alone), the leader is the smallest of them.  Either from the precomputed arrays (sgx_group_ids: leader index,
// size) or -- from_labels -- straight from the label word with one 64-bit match over (scene start, label bits),
// which is the same relation bit for bit.  Lanes past the chunk are their own one-lane group.
__device__ __forceinline__ void group_structure(bool from_labels, bool live, int lane, int lead_or_label, int size_word,
                                                int scene_first, int p0, int& my_lead, int& gs, uint32_t& group_mask) {
    if (from_labels) {
        const float lab = __int_as_float(lead_or_label);
        const bool grouped = live && lab != 0.f && lab == lab;
        const unsigned long long key = grouped ? (((unsigned long long)(uint32_t)scene_first << 32) | (uint32_t)lead_or_label)
                                               : (0x8000000000000000ull | (unsigned)lane);
        group_mask = __match_any_sync(0xffffffffu, key);
        my_lead = __ffs(group_mask) - 1;
        gs = __popc(group_mask);
    } else {
        my_lead = lead_or_label - p0;
        gs = size_word;
        group_mask = __match_any_sync(0xffffffffu, live ? my_lead : 32 + lane);
    }
}

// per-matrix part of the in-kernel weight prep: fp32 staging [n][K] -> fp16 hi | lo image, core kc of output column n at
// kc * N * 16 + n * 16 (hi) and (K / 8 + kc) * N * 16 + n * 16 (lo); `wmax_bits` = bits of max |w| over the matrix.
// Returns (through *winv, written by one thread) the factor that undoes the matrix scale.  Called by all threads.
__device__ __forceinline__ void build_image(uint8_t* __restrict__ img, const float* __restrict__ stage, int N, int K,
                                            uint32_t wmax_bits, float* __restrict__ winv, int tid, int nthreads) {
    const float wm = __uint_as_float(wmax_bits);
    float s = 1.f, inv = 1.f;
    if (!scale_free(wm)) pow2_scale(wm, s, inv);
    const int KC = K / 8;
    for (int u = tid; u < N * KC; u += nthreads) {
        const int kc = u / N, n = u % N;
        const float* src = stage + n * K + kc * 8;
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
        *reinterpret_cast<uint4*>(img + kc * N * 16 + n * 16) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(img + (KC + kc) * N * 16 + n * 16) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    if (tid == 0) *winv = inv;
}

}  // namespace gtile
}  // namespace sgx
