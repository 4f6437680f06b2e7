// PoolHiddenNet forward at fp32-grade accuracy ON the 5th-gen tensor cores ("tc32" precision).
//
// Reference math (sgan/models.py:530-541), per ordered pair (i,j) of a scene:
//     y = ReLU(W2 ReLU(W1 [We (P_j-P_i)+be ; h_j] + b1) + b2)   ;   out_i = max_j y
// No fp32 tensor-core MMA exists, and bf16 operands (sgx_pool_tc.cu) give 2e-2 pooled features / ~1e-3 ADE.
// Here every operand is split into TWO fp16 pieces, v = hi + lo (11 + 11 significant bits), and the products
// hi.hi + hi.lo + lo.hi (+ lo.lo where it is free) are accumulated in fp32 by tcgen05.mma kind::f16:
// measured against an fp64 evaluation the pooled features are as close as the reference's own fp32 result
// (5e-7 of the largest output; bf16 hi/lo splits would give 8e-6, not enough for the 1e-5 contract).
//
// Per tile of 128 consecutive ordered pairs of the flat (i,j)-sorted pair list:
//   GEMM1  D1[128 x 512] = X . W1p^T, K-steps of 16:   [c0] + [h_hi.W_hi] + [h_hi.W_lo] + [h_lo.W_hi]
//          c0 = the exactly folded first layer (Aeff = W1e We, c = W1e be + b1) applied to the relative position:
//          X  slots [dxh dxl dxh dyh dyl dyh s s | dxl2 dyl2 s dxl dyl 0 0 0]       (d split 3-way, s = scale)
//          W  slots [A0h A0h A0l A1h A1h A1l ch cl | A0h A1h cl2 A0l A1l 0 0 0]
//          4 hidden chunks of N = 128, each 1 + 3 H/16 MMAs (7 for H = 32), into a ring of THREE 128-column
//          TMEM buffers.
//   EPI1   in place: TMEM -> registers, hi = cvt.rz.relu.f16x2, lo = cvt.rn.relu.f16x2(v - hi), registers -> the
//          SAME TMEM columns ([16 hi | 16 lo] per 32 fp32 columns) = the A operand of GEMM2.  rz for hi keeps
//          v - hi >= 0 for v >= 0, so both ReLUs ride on the converts.
//   GEMM2  D2[128 x 16] += H_hi . W2s^T + H_lo . W2s^T with W2s = [W2_hi (8 rows) ; W2_lo (8 rows)] stacked
//          along N: columns 0-7 collect (H_hi + H_lo) W2_hi, columns 8-15 (H_hi + H_lo) W2_lo -- all four
//          product terms for two MMA streams; the row warps add the two column groups.
//   EPI2   + b2, ReLU, segmented max over rows sharing i, packed 64-bit atomicMax (value bits << 32 | j).
// fp16 range: X is multiplied by a power of two `s` chosen per call from a bound on |z| (max |h|, scene extent,
// weight norms; exact, undone in EPI2), so nothing saturates for any input whose bound is < 2^38; s = 1 for
// every realistic input and then changes no bit.
// All operand images use the NO-SWIZZLE K-major canonical layout (8 x 16 B core matrices, LBO between the two
// K cores, SBO between 8-row groups), which allows the 80-element K extent without padding to 128-byte rows.
// One CTA per SM, persistent over tiles, 18 warps: 2 x 4 row warps (X build + EPI2), 2 x 4 EPI1 warps,
// GEMM1 issuer, GEMM2 issuer.
#include <cuda_fp16.h>

#include "sgx_tc.cuh"

namespace sgx {

namespace t32 {

constexpr int HID = SGX_POOL_HIDDEN;
constexpr int TILE = 128;
constexpr int NCHUNK = 4;
constexpr int NST = 4;
constexpr int NMETA = 4;
constexpr int NTHREADS = 18 * 32;
constexpr int SLICE = 136;
constexpr int NBUF = 3;          // ring of combined D1 / H TMEM buffers (128 columns each)
constexpr int TM_D2 = 384;       // two D2 accumulators, 64 columns apart
constexpr int N2 = 16;           // GEMM2 N: 8 channels x {W2_hi, W2_lo}

template <int H>
struct Cfg {
    static constexpr int KC = 2 + 2 * (H / 8);          // 16-byte K cores per row: c0 (2) + h_hi + h_lo
    static constexpr int X_STAGE = KC * TILE * 16;
    static constexpr int W1P_BYTES = KC * HID * 16;
    static constexpr int W2P_BYTES = (HID / 8) * N2 * 16;
    static constexpr int W1P = 0;
    static constexpr int X = W1P + W1P_BYTES;
    static constexpr int W2P = X + NST * X_STAGE;
    static constexpr int META = W2P + W2P_BYTES;          // int2 [NMETA][128]
    static constexpr int SOFF = META + NMETA * TILE * 8;  // int64 [2][SLICE]
    static constexpr int SPS = SOFF + 2 * SLICE * 8;      // int32 [2][SLICE]
    static constexpr int BARS = SPS + 2 * SLICE * 4;
    static constexpr int TOTAL = BARS + 256 + 1024;
    static constexpr int KSTEPS_H = H / 16;
};

__device__ __forceinline__ float f16_round(float v) { return __half2float(__float2half_rn(v)); }

// ---------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------
// The power-of-two scale applied to X (and undone in EPI2): |z| * sc <= 2^14 given max |h|, the scene extent and the
// weight norms; 1 for every realistic input.  Evaluated identically by the h-image kernel and by every CTA.
__device__ __forceinline__ int scale_shift(const unsigned* __restrict__ cst, const unsigned* __restrict__ stat) {
    const float hmax = __uint_as_float(stat[0]), dmax = __uint_as_float(stat[1]);
    const float bound = fmaxf(fmaxf(__uint_as_float(cst[0]) * dmax + __uint_as_float(cst[1]) * hmax + __uint_as_float(cst[2]),
                                    hmax), dmax);
    const int e = (int)((__float_as_uint(bound) >> 23) & 0xffu) - 127;     // floor(log2(bound)); 128 for inf / NaN
    const int s = e + 1 - 14;
    return s < 0 ? 0 : (s > 24 ? 24 : s);
}

// per call, pass 1: max |h| and the largest coordinate distance of a pedestrian from the first pedestrian of its scene
// (stat[0], stat[1]: float bits, zeroed by the caller; NaN / inf order above every finite value)
template <int H>
__global__ void stat_kernel(const float* __restrict__ h, const float* __restrict__ pos,
                            const int32_t* __restrict__ ped_start, int64_t batch, unsigned* __restrict__ stat) {
    pdl_trigger();
    pdl_wait();
    const int64_t n4 = batch * (H / 4);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned bh = 0, bd = 0;
    // four independent 16-byte loads in flight per thread and trip (one load per trip left the kernel latency-bound:
    // 10 us for 28 MB that mostly sit in L2)
    for (int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t0 < n4; t0 += 4 * stride) {
        float4 a[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t t = t0 + u * stride;
            a[u] = t < n4 ? reinterpret_cast<const float4*>(h)[t] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t t = t0 + u * stride;
            bh = max(bh, max(max(__float_as_uint(a[u].x) & 0x7fffffffu, __float_as_uint(a[u].y) & 0x7fffffffu),
                             max(__float_as_uint(a[u].z) & 0x7fffffffu, __float_as_uint(a[u].w) & 0x7fffffffu)));
            if (t < batch) {
                const int q = ped_start[t];
                bd = max(bd, max(__float_as_uint(pos[2 * t] - pos[2 * q]) & 0x7fffffffu,
                                 __float_as_uint(pos[2 * t + 1] - pos[2 * q + 1]) & 0x7fffffffu));
            }
        }
    }
    bh = __reduce_max_sync(0xffffffffu, bh);
    bd = __reduce_max_sync(0xffffffffu, bd);
    // one pair of global atomics per CTA, not per warp: 4 700 same-address L2 atomics were most of this kernel's 13 us
    __shared__ unsigned s_h, s_d;
    if (threadIdx.x == 0) { s_h = 0u; s_d = 0u; }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        if (bh) atomicMax(&s_h, bh);
        if (bd) atomicMax(&s_d, bd);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_h) atomicMax(&stat[0], s_h);
        if (s_d) atomicMax(&stat[1], s_d);
    }
}

// per call, pass 2: fp16 hi | lo image of the SCALED h (row = [hi (H) | lo (H)] halves)
template <int H>
__global__ void prep_h_kernel(const float* __restrict__ h, int64_t batch, const unsigned* __restrict__ cst,
                              const unsigned* __restrict__ stat, __half* __restrict__ hb) {
    constexpr int G = H / 8;
    pdl_trigger();
    pdl_wait();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= batch * G) return;
    const float sc = __uint_as_float((uint32_t)(127 - scale_shift(cst, stat)) << 23);
    const int64_t p = t / G;
    const int g = (int)(t % G);
    const float4 a = reinterpret_cast<const float4*>(h + p * H)[2 * g], b = reinterpret_cast<const float4*>(h + p * H)[2 * g + 1];
    const float v[8] = {a.x * sc, a.y * sc, a.z * sc, a.w * sc, b.x * sc, b.y * sc, b.z * sc, b.w * sc};
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float h0 = f16_round(v[2 * e]), h1 = f16_round(v[2 * e + 1]);
        hi[e] = pack_f16_rn(h0, h1);
        lo[e] = pack_f16_rn(v[2 * e] - h0, v[2 * e + 1] - h1);
    }
    uint4* row = reinterpret_cast<uint4*>(hb + p * 2 * H);
    row[g] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    row[G + g] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// per weight version: the fp16 hi/lo operand images and the norms the per-call scale is derived from.
//   W1p: K core kc (16 bytes = 8 halves) of row n at  kc * (512 * 16) + n * 16
//   W2p: K core kc of row r (r < 8: W2_hi[r], r >= 8: W2_lo[r - 8]) at  kc * (16 * 16) + r * 16
template <int H>
__global__ void prep_w_kernel(const float2* __restrict__ Aeff, const float* __restrict__ c0,
                              const float* __restrict__ W1, const float* __restrict__ W2, int E, int B,
                              __half* __restrict__ W1p, __half* __restrict__ W2p, unsigned* __restrict__ cst) {
    constexpr int KC = Cfg<H>::KC;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < HID * KC) {
        const int n = t / KC, kc = t % KC;
        float v[8];
        if (kc < 2) {
            const float2 a = Aeff[n];
            const float c = c0[n];
            const float a0h = f16_round(a.x), a1h = f16_round(a.y), ch = f16_round(c);
            const float a0l = f16_round(a.x - a0h), a1l = f16_round(a.y - a1h), cl = f16_round(c - ch);
            const float cl2 = f16_round(c - ch - cl);
            if (kc == 0) {
                v[0] = a0h; v[1] = a0h; v[2] = a0l; v[3] = a1h; v[4] = a1h; v[5] = a1l; v[6] = ch; v[7] = cl;
            } else {
                v[0] = a0h; v[1] = a1h; v[2] = cl2; v[3] = a0l; v[4] = a1l; v[5] = 0.f; v[6] = 0.f; v[7] = 0.f;
            }
        } else {
            const bool lo = kc >= 2 + H / 8;
            const int m0 = (kc - 2 - (lo ? H / 8 : 0)) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float w = W1[(int64_t)n * (E + H) + E + m0 + e];
                const float wh = f16_round(w);
                v[e] = lo ? (w - wh) : wh;
            }
        }
        uint4 o = make_uint4(pack_f16_rn(v[0], v[1]), pack_f16_rn(v[2], v[3]), pack_f16_rn(v[4], v[5]), pack_f16_rn(v[6], v[7]));
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(W1p) + (size_t)kc * (HID * 16) + (size_t)n * 16) = o;
    }
    if (t < (HID / 8) * N2) {
        const int kc = t / N2, r = t % N2;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float w = ((r & 7) < B) ? W2[(int64_t)(r & 7) * HID + kc * 8 + e] : 0.f;
            const float wh = f16_round(w);
            v[e] = (r >= 8) ? (w - wh) : wh;
        }
        uint4 o = make_uint4(pack_f16_rn(v[0], v[1]), pack_f16_rn(v[2], v[3]), pack_f16_rn(v[4], v[5]), pack_f16_rn(v[6], v[7]));
        *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(W2p) + (size_t)kc * (N2 * 16) + (size_t)r * 16) = o;
    }
    if (t < HID) {   // norms for the scale bound (cst zeroed by the caller)
        const float2 a = Aeff[t];
        float l1 = 0.f;
        for (int m = 0; m < H; ++m) l1 += fabsf(W1[(int64_t)t * (E + H) + E + m]);
        atomicMax(&cst[0], __float_as_uint(fabsf(a.x) + fabsf(a.y)) & 0x7fffffffu);
        atomicMax(&cst[1], __float_as_uint(l1) & 0x7fffffffu);
        atomicMax(&cst[2], __float_as_uint(fabsf(c0[t])) & 0x7fffffffu);
    }
}

// ---------------------------------------------------------------------------------------------
// the pooling kernel
// ---------------------------------------------------------------------------------------------
template <int H, int B>
__global__ void __launch_bounds__(NTHREADS, 1)
pool_tc32_kernel(const __half* __restrict__ hb, const float* __restrict__ pos, const int32_t* __restrict__ ped_start,
                 const int64_t* __restrict__ pair_off, const int32_t* __restrict__ tile_first, int64_t n_tiles,
                 int batch, int64_t n_pairs, const __half* __restrict__ W1p, const __half* __restrict__ W2p,
                 const unsigned* __restrict__ cst, const unsigned* __restrict__ stat, const float* __restrict__ b2,
                 unsigned long long* __restrict__ packed, long long* __restrict__ stats_out) {
    using C = Cfg<H>;
    long long stats_[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // per-role wait counters (SGX_TC_STATS builds, tools/tc_stats.py)
    const long long t_begin_ = clock64();
    static_assert(B == 8, "GEMM2 stacks W2_hi / W2_lo along N = 16");
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BARS);
    uint64_t* w_full = bars + 0;
    uint64_t* x_full = bars + 1;              // [NST]
    uint64_t* x_free = x_full + NST;          // [NST]
    uint64_t* d1_full = x_free + NST;         // [NBUF]  GEMM1 of a chunk committed
    uint64_t* h_ready = d1_full + NBUF;       // [NBUF]  EPI1 wrote the hi/lo operand in place
    uint64_t* buf_free = h_ready + NBUF;      // [NBUF]  GEMM2 consumed it
    uint64_t* d2_full = buf_free + NBUF;      // [2]
    uint64_t* d2_free = d2_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_free + 2);
    int2* meta = reinterpret_cast<int2*>(smem + C::META);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = (n_tiles > blockIdx.x) ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    pdl_trigger();
    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < NST; ++s) { mbar_init(&x_full[s], 128); mbar_init(&x_free[s], 1); }
        for (int s = 0; s < NBUF; ++s) { mbar_init(&d1_full[s], 1); mbar_init(&h_ready[s], 128); mbar_init(&buf_free[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&d2_full[s], 1); mbar_init(&d2_free[s], 128); }
        fence_barrier_init();
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0u) __trap();
    constexpr uint32_t tmem = 0u;

    // The weight images were written by sgx_pool_tc32_prep, which the memset + statistics + h-image launches of this call
    // separate from this kernel: their copy starts before the wait on the previous kernel of the chain (pdl_wait).
    if (warp == 16 && lane == 0) {
        mbar_expect_tx(w_full, C::W1P_BYTES + C::W2P_BYTES);
        // bulk copies are limited by the mbarrier tx-count field: issue the W1 image in 16 KB pieces
        for (int o = 0; o < C::W1P_BYTES; o += 16384)
            bulk_g2s(smem + C::W1P + o, reinterpret_cast<const uint8_t*>(W1p) + o,
                     (uint32_t)((C::W1P_BYTES - o) < 16384 ? (C::W1P_BYTES - o) : 16384), w_full);
        bulk_g2s(smem + C::W2P, W2p, C::W2P_BYTES, w_full);
    }
    __syncwarp();
    pdl_wait();

    // per-call power-of-two scale (see scale_shift); sc = 1 unless the bound says otherwise
    const int sshift = scale_shift(cst, stat);
    const float sc = __uint_as_float((uint32_t)(127 - sshift) << 23), inv_sc = __uint_as_float((uint32_t)(127 + sshift) << 23);

    if (warp == 16) {
        // ======================= GEMM1 issuer =======================
        mbar_wait(w_full, 0);
        constexpr uint32_t idesc1 = make_idesc_f16(128, 128);
        const uint64_t w1_d = make_desc_ns(sbase + C::W1P, HID * 16, 128);
        const uint64_t x_d = make_desc_ns(sbase + C::X, TILE * 16, 128);
        int b = 0;
        uint32_t fp = 0;                                  // bit b: parity of the number of uses of buffer b so far
        for (int t = 0; t < my_tiles; ++t) {
            const int st = t & (NST - 1);
            TWAIT(&x_full[st], (uint32_t)((t / NST) & 1), 0);
            const uint64_t xa = x_d + (uint64_t)(st * (C::X_STAGE / 16));
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                TWAIT(&buf_free[b], ((fp >> b) & 1u) ^ 1u, 1);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t d1 = tmem + (uint32_t)b * 128u;
                    const uint64_t wb = w1_d + (uint64_t)(c * (128 * 16 / 16));
                    // K core kc of X at kc * TILE * 16 bytes, of W1p at kc * HID * 16 bytes (descriptor units of 16 B)
                    auto step = [&](int kca, int kcb, uint32_t acc) {
                        mma_ss(d1, xa + (uint64_t)(kca * (TILE * 16 / 16)), wb + (uint64_t)(kcb * (HID * 16 / 16)), idesc1, acc);
                    };
                    step(0, 0, 0);
#pragma unroll
                    for (int s = 0; s < C::KSTEPS_H; ++s) step(2 + 2 * s, 2 + 2 * s, 1);                       // h_hi . W_hi
#pragma unroll
                    for (int s = 0; s < C::KSTEPS_H; ++s) step(2 + 2 * s, 2 + H / 8 + 2 * s, 1);               // h_hi . W_lo
#pragma unroll
                    for (int s = 0; s < C::KSTEPS_H; ++s) step(2 + H / 8 + 2 * s, 2 + 2 * s, 1);               // h_lo . W_hi
                    tc_commit(&d1_full[b]);
                    if (c == NCHUNK - 1) tc_commit(&x_free[st]);
                }
                __syncwarp();
                fp ^= 1u << b;
                b = (b == NBUF - 1) ? 0 : b + 1;
            }
        }
    } else if (warp == 17) {
        // ======================= GEMM2 issuer =======================
        mbar_wait(w_full, 0);
        constexpr uint32_t idesc2 = make_idesc_f16(128, N2);
        const uint64_t w2_d = make_desc_ns(sbase + C::W2P, N2 * 16, 128);
        int b = 0;
        uint32_t fp = 0;
        for (int t = 0; t < my_tiles; ++t) {
            const int db = t & 1;
            TWAIT(&d2_free[db], (uint32_t)(((t >> 1) & 1) ^ 1), 2);
            const uint32_t d2 = tmem + TM_D2 + db * 64;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                TWAIT(&h_ready[b], (fp >> b) & 1u, 3);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t hbuf = tmem + (uint32_t)b * 128u;
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        // K-step kk of this chunk: hidden units 128 c + 16 kk ..; two K cores of W2p, 2 * N2 * 16 bytes
                        const uint64_t bd = w2_d + (uint64_t)((c * 8 + kk) * (2 * N2 * 16 / 16));
                        const uint32_t a_hi = hbuf + 32 * (kk >> 1) + 8 * (kk & 1);
                        mma_ts(d2, a_hi, bd, idesc2, (c > 0 || kk > 0));
                        mma_ts(d2, a_hi + 16, bd, idesc2, 1);
                    }
                    tc_commit(&buf_free[b]);
                    if (c == NCHUNK - 1) tc_commit(&d2_full[db]);
                }
                __syncwarp();
                fp ^= 1u << b;
                b = (b == NBUF - 1) ? 0 : b + 1;
            }
        }
    } else if (warp < 8) {
        // ======================= row warps: build X tiles, final epilogue =======================
        const int set = warp >> 2;
        const int rt = (warp & 3) * 32 + lane;
        const int row = rt;
        int64_t* soff = reinterpret_cast<int64_t*>(smem + C::SOFF) + set * SLICE;
        int32_t* sps = reinterpret_cast<int32_t*>(smem + C::SPS) + set * SLICE;
        float bias2[B];
#pragma unroll
        for (int bb = 0; bb < B; ++bb) bias2[bb] = b2[bb];
        const uint32_t sc_pair = pack_f16_rn(sc, sc);
        int lo = 0, hi = -1, sp = 0, sp_x = 0;
        int64_t so = 0, so_x = 0;
        int lo2 = 0, hi2 = -1;
        auto bounds = [&](int itn, int& l, int& h) {
            l = 0; h = -1;
            if (itn < my_tiles) {
                const int64_t tn = blockIdx.x + (int64_t)itn * gridDim.x;
                l = tile_first[tn];
                h = (tn + 1 < n_tiles) ? tile_first[tn + 1] : batch - 1;
            }
        };
        bounds(set, lo, hi);
        bounds(set + 2, lo2, hi2);
        if (lo + rt <= hi) { so = pair_off[lo + rt]; sp = ped_start[lo + rt]; }
        if (rt == 0 && lo + 128 <= hi) { so_x = pair_off[lo + 128]; sp_x = ped_start[lo + 128]; }
        int2 ij = make_int2(-1, 0);
        float2 pi = make_float2(0.f, 0.f), pj = make_float2(0.f, 0.f);
        constexpr int NHV = 2 * (H / 8);
        uint4 hv[NHV];
        bool valid = false;
        auto prefetch = [&](int itn) {
            const int64_t tile = blockIdx.x + (int64_t)itn * gridDim.x;
            if (lo + rt <= hi) { soff[rt] = so; sps[rt] = sp; }
            if (rt == 0 && lo + 128 <= hi) { soff[128] = so_x; sps[128] = sp_x; }
            const int lo_c = lo, hi_c = hi;
            int lo3, hi3;
            bounds(itn + 4, lo3, hi3);
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            const int64_t q = tile * TILE + row;
            ij = make_int2(-1, 0);
            pi = make_float2(0.f, 0.f); pj = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < NHV; ++c) hv[c] = make_uint4(0, 0, 0, 0);
            valid = q < n_pairs;
            if (valid) {
                int a = 0, z = hi_c - lo_c;
                while (a < z) {
                    int mid = (a + z + 1) >> 1;
                    if (soff[mid] <= q) a = mid; else z = mid - 1;
                }
                const int i = lo_c + a;
                const int j = sps[a] + (int)(q - soff[a]);
                ij = make_int2(i, j);
                pi = *reinterpret_cast<const float2*>(pos + 2 * (int64_t)i);
                pj = *reinterpret_cast<const float2*>(pos + 2 * (int64_t)j);
                const uint4* hrow = reinterpret_cast<const uint4*>(hb + (int64_t)j * 2 * H);
#pragma unroll
                for (int c = 0; c < NHV; ++c) hv[c] = hrow[c];
            }
            lo = lo2; hi = hi2; lo2 = lo3; hi2 = hi3;
            if (lo + rt <= hi) { so = pair_off[lo + rt]; sp = ped_start[lo + rt]; }
            if (rt == 0 && lo + 128 <= hi) { so_x = pair_off[lo + 128]; sp_x = ped_start[lo + 128]; }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
        };
        if (set < my_tiles) prefetch(set);
        for (int it = set; it < my_tiles + 2; it += 2) {
            if (it < my_tiles) {
                const int st = it & (NST - 1);
                uint8_t* xrow = smem + C::X + st * C::X_STAGE + row * 16;      // K core kc at + kc * TILE * 16
                TWAIT(&x_free[st], (uint32_t)(((it / NST) & 1) ^ 1), 0);
                uint4 c0 = make_uint4(0, 0, 0, 0), c1 = make_uint4(0, 0, 0, 0);
                if (valid) {
                    // P_j - P_i first (never A.P_j - A.P_i), then a 3-way fp16 split of the scaled difference
                    const float dx = (pj.x - pi.x) * sc, dy = (pj.y - pi.y) * sc;
                    const float dxh = f16_round(dx), dyh = f16_round(dy);
                    const float dxl = f16_round(dx - dxh), dyl = f16_round(dy - dyh);
                    const float dxl2 = dx - dxh - dxl, dyl2 = dy - dyh - dyl;
                    c0.x = pack_f16_rn(dxh, dxl);
                    c0.y = pack_f16_rn(dxh, dyh);
                    c0.z = pack_f16_rn(dyl, dyh);
                    c0.w = sc_pair;
                    c1.x = pack_f16_rn(dxl2, dyl2);
                    c1.y = pack_f16_rn(sc, dxl);
                    c1.z = pack_f16_rn(dyl, 0.f);
                }
                *reinterpret_cast<uint4*>(xrow) = c0;
                *reinterpret_cast<uint4*>(xrow + TILE * 16) = c1;
#pragma unroll
                for (int c = 0; c < NHV; ++c) *reinterpret_cast<uint4*>(xrow + (2 + c) * (TILE * 16)) = hv[c];
                meta[(it & (NMETA - 1)) * TILE + row] = ij;
                fence_proxy_async();
                mbar_arrive(&x_full[st]);
            }
            if (it + 2 < my_tiles) {
#ifdef SGX_TC_STATS
                const long long tp__ = clock64();
#endif
                prefetch(it + 2);
#ifdef SGX_TC_STATS
                stats_[2] += clock64() - tp__;
#endif
            }
            if (it >= 2) {
                const int t = it - 2;
                const int db = set;
                TWAIT(&d2_full[db], (uint32_t)((t >> 1) & 1), 1);
#ifdef SGX_TC_STATS
                const long long tf__ = clock64();
#endif
                tc_fence_after();
                const uint32_t d2a = tmem + ((uint32_t)((warp & 3) << 5) << 16) + TM_D2 + db * 64;
                const int2 mij = meta[(t & (NMETA - 1)) * TILE + row];
                const int key = mij.x;
                const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
                const bool head = (lane == 0) || (key_prev != key);
                const bool uniform = __all_sync(0xffffffffu, key == __shfl_sync(0xffffffffu, key, 0));
                bool same[5];
#pragma unroll
                for (int sft = 0; sft < 5; ++sft) {
                    const int okey = __shfl_down_sync(0xffffffffu, key, 1 << sft);
                    same[sft] = (lane + (1 << sft) < 32) && (okey == key);
                }
                uint32_t v[16];
                tmem_ld16(d2a, v);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&d2_free[db]);
                uint32_t bits[B];
#pragma unroll
                for (int c = 0; c < B; ++c) {
                    // (H_hi + H_lo) W2_hi  +  (H_hi + H_lo) W2_lo, scale undone, + b2, ReLU; a NaN stays a NaN (its
                    // canonical bit pattern orders above every finite value in the packed max, like torch.max)
                    const float y = fmaf(__uint_as_float(v[c]) + __uint_as_float(v[8 + c]), inv_sc, bias2[c]);
                    bits[c] = (y != y) ? 0x7fc00000u : (__float_as_uint(fmaxf(y, 0.f)) & 0x7fffffffu);
                }
                if (uniform) {
#pragma unroll
                    for (int c = 0; c < B; ++c) {
                        const uint32_t mx = __reduce_max_sync(0xffffffffu, bits[c]);
                        const uint32_t who = __ballot_sync(0xffffffffu, bits[c] == mx);
                        const int src = 31 - __clz(who);
                        const int jj = __shfl_sync(0xffffffffu, mij.y, src);
                        if (lane == 0 && key >= 0)
                            atomicMax(&packed[(int64_t)key * B + c], ((unsigned long long)mx << 32) | (unsigned)jj);
                    }
                } else {
                    unsigned long long pk[B];
#pragma unroll
                    for (int c = 0; c < B; ++c) pk[c] = ((unsigned long long)bits[c] << 32) | (unsigned)mij.y;
#pragma unroll
                    for (int sft = 0; sft < 5; ++sft) {
#pragma unroll
                        for (int c = 0; c < B; ++c) {
                            const unsigned long long other = __shfl_down_sync(0xffffffffu, pk[c], 1 << sft);
                            if (same[sft] && other > pk[c]) pk[c] = other;
                        }
                    }
                    if (head && key >= 0) {
#pragma unroll
                        for (int c = 0; c < B; ++c) atomicMax(&packed[(int64_t)key * B + c], pk[c]);
                    }
                }
#ifdef SGX_TC_STATS
                stats_[3] += clock64() - tf__;
#endif
            }
        }
    } else {
        // ======================= EPI1 warps: D1 -> ReLU -> fp16 hi | lo, in place =======================
        const int grp = (warp - 8) >> 2;               // chunks n = grp, grp + 2, ...
        const int quad = warp & 3;
        const uint32_t lane_base = (uint32_t)(quad << 5) << 16;
        const int n_chunks = my_tiles * NCHUNK;
        int b = grp;                                   // n = 3 u + b
        uint32_t u = 0;
#pragma unroll 1
        for (int n = grp; n < n_chunks; n += 2) {
            TWAIT(&d1_full[b], u & 1u, 0);
#ifdef SGX_TC_STATS
            const long long te__ = clock64();
#endif
            tc_fence_after();
            const uint32_t buf = tmem + lane_base + (uint32_t)b * 128u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t va[32], p[32];
                tmem_ld32(buf + 32 * q, va);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float x0 = __uint_as_float(va[2 * e]), x1 = __uint_as_float(va[2 * e + 1]);
                    const uint32_t hi2 = pack_f16_rz_relu(x0, x1);
                    float l0, l1;
                    sub_f16x2(hi2, x0, x1, l0, l1);           // v - hi, one mixed-precision FMA (FHFMA) per element
                    p[e] = hi2;
                    p[16 + e] = pack_f16_rn_relu(l0, l1);
                }
                tmem_st32(buf + 32 * q, p);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&h_ready[b]);
#ifdef SGX_TC_STATS
            stats_[1] += clock64() - te__;
#endif
            b += 2;
            if (b >= NBUF) { b -= NBUF; ++u; }
        }
    }
#ifdef SGX_TC_STATS
    if (stats_out && blockIdx.x == 0 && lane == 0 && (warp == 16 || warp == 17 || warp == 0 || warp == 8 || warp == 12)) {
        const int role = warp == 16 ? 0 : warp == 17 ? 1 : warp == 0 ? 2 : warp == 8 ? 3 : 4;
        for (int k = 0; k < 4; ++k) stats_out[role * 8 + k] = stats_[k];
        stats_out[role * 8 + 4] = clock64() - t_begin_;
        stats_out[role * 8 + 5] = my_tiles;
    }
#endif
    (void)stats_out; (void)t_begin_; (void)stats_;
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

}  // namespace t32

extern long long* g_tc_stats;   // sgx_pool_tc.cu

__global__ void pool_prep_kernel(const float* __restrict__ We, const float* __restrict__ be,
                                 const float* __restrict__ W1, const float* __restrict__ b1, int E, int H,
                                 float2* __restrict__ Aeff, float* __restrict__ c0);

}  // namespace sgx

using namespace sgx;

// ---- prepared weights (per weight version): [consts 256 B | W1p | W2p | Aeff | c0] ----
bool sgx_pool_tc32_supported(int E, int H, int B) { return E >= 1 && H == 32 && B == 8; }

int64_t sgx_pool_tc32_prep_bytes(int E, int H, int B) {
    (void)E; (void)H; (void)B;
    return 256 + align_up(t32::Cfg<32>::W1P_BYTES, 256) + align_up(t32::Cfg<32>::W2P_BYTES, 256) +
           align_up(t32::HID * 8, 256) + align_up(t32::HID * 4, 256);
}

int sgx_pool_tc32_prep(const float* We, const float* be, const float* W1, const float* b1, const float* W2, int E, int H,
                       int B, void* prep, cudaStream_t st) {
    SGX_UNSUPPORTED(!sgx_pool_tc32_supported(E, H, B),
                    "tc32 pooling (tcgen05, fp16 hi/lo splits) is built for h_dim 32, bottleneck 8; got (%d,%d)", H, B);
    Carver c(prep);
    unsigned* cst = c.take<unsigned>(64);
    __half* W1p = c.take<__half>(t32::Cfg<32>::W1P_BYTES / 2);
    __half* W2p = c.take<__half>(t32::Cfg<32>::W2P_BYTES / 2);
    float2* Aeff = c.take<float2>(t32::HID);
    float* c0 = c.take<float>(t32::HID);
    SGX_CUDA(cudaMemsetAsync(cst, 0, 256, st));
    pool_prep_kernel<<<2, 256, 0, st>>>(We, be, W1, b1, E, H, Aeff, c0);
    SGX_LAUNCH_CHECK();
    t32::prep_w_kernel<32><<<blocks_for(t32::HID * t32::Cfg<32>::KC, 256), 256, 0, st>>>(Aeff, c0, W1, W2, E, B, W1p, W2p, cst);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

// per-call workspace behind `packed` (the caller zeroes packed AND the first 256 bytes of ws in one memset)
int64_t sgx_pool_tc32_ws_bytes(int64_t batch, int H) { return 256 + align_up(batch * 2 * H * 2, 256); }

int sgx_pool_fwd_tc32(const float* h, const float* pos, const int32_t* ped_start, const int64_t* pair_off,
                      const int32_t* tile_first, int64_t batch, int64_t n_pairs, const void* prep, const float* b2, int E,
                      int H, int B, unsigned long long* packed, void* ws, cudaStream_t st, bool prep_in_this_call) {
    SGX_UNSUPPORTED(!sgx_pool_tc32_supported(E, H, B),
                    "tc32 pooling (tcgen05, fp16 hi/lo splits) is built for h_dim 32, bottleneck 8; got (%d,%d)", H, B);
    SGX_REQUIRE(n_pairs < ((int64_t)1 << 40), "sgx_pool_fwd_tc32: too many pairs");
    using C = t32::Cfg<32>;
    Carver pc(const_cast<void*>(prep));
    const unsigned* cst = pc.take<unsigned>(64);
    const __half* W1p = pc.take<__half>(C::W1P_BYTES / 2);
    const __half* W2p = pc.take<__half>(C::W2P_BYTES / 2);
    Carver wc(ws);
    unsigned* stat = wc.take<unsigned>(64);          // zeroed by the caller
    __half* hb = wc.take<__half>(batch * 2 * H);
    int dev = 0, sms = 148;
    SGX_CUDA(cudaGetDevice(&dev));
    SGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // (weight images built by THIS call: the statistics kernel is ordered as usual behind them, which keeps the pooling
    // kernel's early weight copy behind a completed, flushed prep_w_kernel)
    SGX_CUDA(launch_pdl(t32::stat_kernel<32>, dim3((unsigned)std::min<int64_t>(blocks_for(batch * (H / 4), 256), 4 * sms)),
                        dim3(256), 0, st, !prep_in_this_call, h, pos, ped_start, batch, stat));
    SGX_LAUNCH_CHECK();
    SGX_CUDA(launch_pdl(t32::prep_h_kernel<32>, dim3(blocks_for(batch * (H / 8), 256)), dim3(256), 0, st, true, h, batch, cst,
                        stat, hb));
    SGX_LAUNCH_CHECK();
    const int64_t n_tiles = (n_pairs + t32::TILE - 1) / t32::TILE;
    auto kern = t32::pool_tc32_kernel<32, 8>;
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, sms);
    cudaEvent_t ev0, ev1;
    profile_events(&ev0, &ev1);
    if (ev0 && ev1) SGX_CUDA(cudaEventRecord(ev0, st));
    // (with the bench's events around it the kernel is ordered as usual, so its measured time includes its whole prologue)
    SGX_CUDA(launch_pdl(kern, dim3(grid), dim3(t32::NTHREADS), C::TOTAL, st, !(ev0 && ev1), hb, pos, ped_start, pair_off,
                        tile_first, n_tiles, (int)batch, n_pairs, W1p, W2p, cst, (const unsigned*)stat, b2, packed, g_tc_stats));
    SGX_LAUNCH_CHECK();
    if (ev0 && ev1) SGX_CUDA(cudaEventRecord(ev1, st));
    return SGX_OK;
}
