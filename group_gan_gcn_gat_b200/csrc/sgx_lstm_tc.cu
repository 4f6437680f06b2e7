// Per-pedestrian LSTM recurrences on the tensor cores with fp32-level accuracy.
//
// Same math as sgx_lstm.cu (Encoder.forward sgan/models.py:62-92, Decoder.forward 142-178 without per-step
// pooling), H = 32.  The CUDA-core kernel is bound by broadcasting W_hh from shared memory (ncu: 79 % LSU
// wavefronts); here the 128 gate pre-activations of a 128-pedestrian tile are ONE small GEMM per step,
//     G[128 x 128] = [in | h] . [Wx | W_hh]^T ,
// issued as tcgen05.mma (bf16 operands, fp32 accumulate in TMEM).  bf16 alone would break the 1e-4 ADE/FDE
// parity, so every fp32 operand is split three ways (v = hi + mid + lo, each bf16) and the product keeps all
// terms down to 2^-24:  hi*hi + hi*mid + mid*hi + hi*lo + mid*mid + lo*hi.  In K-major operand rows:
//     A block0 = [ in(16) | h_hi(32) | 0(16) ]     A block1 = [ h_mid(32) | h_lo(32) ]
//     B tiles  : Ba = [Wx(16) | W_hi | 0]  Bb = [0 | W_mid | 0]  Bc = [0 | W_lo | 0]  (x block0)
//                Bd = [W_hi | W_hi]        Be = [W_mid | 0]                           (x block1)
//     in(16) = [dxh dxm dxh dxl dxm dxh | dyh dym dyh dyl dym dyh | 1 1 1 | 0] against
//              [axh axh axm axh axm axl | ayh ayh aym ayh aym ayl | bh bm bl | 0]   (embedding + biases folded)
// = 13 MMAs (M 128, N 128, K 16) per tile-step.  The epilogue (one thread per pedestrian, c and h in registers)
// reads the gates from TMEM, applies sigmoid/tanh, updates c and h, computes hidden2pos (decoder), splits the new h
// and writes the next A rows.  Three 128-ped tiles are in flight per CTA so one tile's MMAs overlap the others'
// epilogues; the kernel is bound by the 256 MUFU operations per pedestrian-step of the gate non-linearities.
#include "sgx_tc.cuh"

namespace sgx {

constexpr int LH = 32;            // hidden size this kernel is built for
constexpr int LT = 128;           // pedestrians per tile (UMMA M)
constexpr int LSLOTS = 3;         // tiles in flight per CTA (13 warps => 128 registers per thread)
constexpr int LTHREADS = (1 + 4 * LSLOTS) * 32;

struct LstmTcSmem {               // byte offsets from the 1024-aligned base
    static constexpr int B = 0;                               // 5 weight tiles x 16 KB
    static constexpr int A = B + 5 * LT * 128;                // LSLOTS x 2 blocks x 16 KB
    static constexpr int WHP = A + LSLOTS * 2 * LT * 128;     // hidden2pos: 2*32 + 2 floats
    static constexpr int BARS = WHP + 512;
    static constexpr int TOTAL = BARS + 256 + 1024;
};

constexpr float LOG2E = 1.4426950408889634f;
__device__ __forceinline__ void split3(float v, float& hi, float& mid, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(v));
    const float r = v - hi;
    mid = __bfloat162float(__float2bfloat16_rn(r));
    lo = r - mid;                                             // rounded to bf16 when packed
}

// Weight images: 5 tiles [128 gate rows x 64 k] bf16, SWIZZLE_128B K-major.
__global__ void lstm_tc_prep_kernel(const float* __restrict__ We, const float* __restrict__ be,
                                    const float* __restrict__ W_ih, const float* __restrict__ W_hh,
                                    const float* __restrict__ b_ih, const float* __restrict__ b_hh, int E,
                                    __nv_bfloat16* __restrict__ img) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= 5 * LT * 64) return;
    const int tile = t / (LT * 64), r = (t / 64) % LT, k = t % 64;
    float v = 0.f;
    // Gate rows are pre-scaled so that the GEMM delivers the ex2 argument itself: -log2(e) * a for the sigmoid gates
    // (i, f, o: e^-a = 2^(-a log2 e)), 2 log2(e) * a for the cell candidate (tanh through e^2a).  One rounding per
    // weight, the same one the multiply in the epilogue would make.
    const float gs = (r >= 2 * LH && r < 3 * LH) ? 2.f * LOG2E : -LOG2E;
    auto w3 = [&](int col, int part) {                        // part 0/1/2 = hi/mid/lo of gs * W_hh[r][col]
        float hi, mid, lo;
        split3(gs * W_hh[r * LH + col], hi, mid, lo);
        return part == 0 ? hi : part == 1 ? mid : lo;
    };
    if (tile == 0) {
        if (k < 16) {
            float ax = 0.f, ay = 0.f, b = b_ih[r] + b_hh[r];
            for (int e = 0; e < E; ++e) {
                const float w = W_ih[r * E + e];
                ax = fmaf(w, We[2 * e], ax);
                ay = fmaf(w, We[2 * e + 1], ay);
                b = fmaf(w, be[e], b);
            }
            float h3[3];
            const int part_of[6] = {0, 0, 1, 0, 1, 2};        // against d: [h m h l m h]
            if (k < 6) { split3(gs * ax, h3[0], h3[1], h3[2]); v = h3[part_of[k]]; }
            else if (k < 12) { split3(gs * ay, h3[0], h3[1], h3[2]); v = h3[part_of[k - 6]]; }
            else if (k < 15) { split3(gs * b, h3[0], h3[1], h3[2]); v = h3[k - 12]; }
        } else if (k < 48) v = w3(k - 16, 0);
    } else if (tile == 1) { if (k >= 16 && k < 48) v = w3(k - 16, 1); }
    else if (tile == 2) { if (k >= 16 && k < 48) v = w3(k - 16, 2); }
    else if (tile == 3) { v = w3(k & 31, 0); }
    else { if (k < 32) v = w3(k, 1); }
    const uint32_t off = (uint32_t)tile * LT * 128 + swz((uint32_t)r, (uint32_t)(k >> 3)) + (k & 7) * 2;
    img[off >> 1] = __float2bfloat16_rn(v);
}

// The kernel is MUFU-bound (XU pipe > 80 % busy, profiles/r01_lstm_v3_ncu_full_raw.csv).  ex2 + rcp per sigmoid / tanh
// would be 10 MUFU per hidden unit and step.  With e_x = e^-x and E_x = e^2x:
//     c' = sigmoid(f) c + sigmoid(i) tanh(g) = [c (1+e_i)(1+E_g) + (E_g-1)(1+e_f)] / [(1+e_f)(1+e_i)(1+E_g)]
//     h' = sigmoid(o) tanh(c')               = (E_c'-1) / [(1+e_o)(1+E_c')]
// is 5 ex2 + 2 rcp = 7.  The gate pre-activations arrive pre-scaled (see lstm_tc_prep_kernel), so every exponential is a
// clamp and one ex2.approx.ftz: __expf / __fdividef would add a range check and two multiplies around every MUFU.
// Clamps: e <= 2^40 (sigmoid floor 9e-13), E <= 2^30 (tanh = 1 - 2e-9, below fp32 resolution), so the triple product
// stays finite for |c| < 2^17.  (Tried, measured, dropped: one of the five exponentials as an FMA-pipe polynomial -- no
// gain, the issue slots it costs are as scarce as the XU cycles it frees; the bf16 roundings of the operand splits as
// integer arithmetic instead of cvt.rn.bf16x2 -- 6 % slower; truncating splits (mask + byte-permute, no cvt at all) -- no
// change; software-pipelined TMEM loads -- no change; four tiles in flight (17 warps => 96 registers, 250 B of spills)
// -- 20 % slower; the two reciprocals as integer seed + two Halley steps on the FMA pipe -- 1 % slower.
// Per-phase counters (tools/lstm_stats.py): of 6 590 cycles per tile-step the gate math holds the warp
// for 4 090 with three warps sharing each SMSP's XU unit: 3 x 224 MUFU x 8 cycles = 5 376 = 82 % of the step.)
constexpr float E_CLAMP = 40.f, T_CLAMP = 30.f;
__device__ __forceinline__ float ex2_clamped(float v, float hi) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fminf(v, hi)));
    return r;
}
__device__ __forceinline__ float rcp_fast(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float cell_update(float c, float e_i, float E_g, float e_f) {
    const float A = (1.f + e_i) * (1.f + E_g), F = 1.f + e_f;
    return fmaf(c, A, (E_g - 1.f) * F) * rcp_fast(F * A);
}
__device__ __forceinline__ float hidden_out(float e_o, float c) {
    const float E = ex2_clamped(2.f * LOG2E * c, T_CLAMP);
    return (E - 1.f) * rcp_fast((1.f + e_o) * (1.f + E));
}

// hi/mid/lo bf16 split of TWO values at once, all on the ALU pipe: cvt.rn.bf16x2 (F2FP) rounds both, the rounded values
// come back as floats by a shift / mask.  The scalar form (F2F.BF16.F32) runs on the XU pipe, which the gate
// non-linearities already saturate: 136 of the kernel's 392 XU operations per pedestrian-step were these conversions.
struct Split3x2 { float a[3], b[3]; uint32_t packed[3]; };
__device__ __forceinline__ Split3x2 split3_pair(float x0, float x1) {
    Split3x2 r;
    r.packed[0] = pack_bf16(x0, x1);
    r.a[0] = __uint_as_float(r.packed[0] << 16); r.b[0] = __uint_as_float(r.packed[0] & 0xFFFF0000u);
    const float ra = x0 - r.a[0], rb = x1 - r.b[0];
    r.packed[1] = pack_bf16(ra, rb);
    r.a[1] = __uint_as_float(r.packed[1] << 16); r.b[1] = __uint_as_float(r.packed[1] & 0xFFFF0000u);
    r.a[2] = ra - r.a[1]; r.b[2] = rb - r.b[1];
    r.packed[2] = pack_bf16(r.a[2], r.b[2]);
    return r;
}

// write this pedestrian's operand rows for the next step: input chunk + 3-way split of h
__device__ __forceinline__ void write_a_rows(uint8_t* blk0, uint8_t* blk1, int row, float dx, float dy,
                                             const float (&h)[LH]) {
    const Split3x2 d = split3_pair(dx, dy);
    const float xh = d.a[0], xm = d.a[1], xl = d.a[2], yh = d.b[0], ym = d.b[1], yl = d.b[2];
    uint4 c0, c1;
    c0.x = pack_bf16(xh, xm); c0.y = pack_bf16(xh, xl); c0.z = pack_bf16(xm, xh); c0.w = pack_bf16(yh, ym);
    c1.x = pack_bf16(yh, yl); c1.y = pack_bf16(ym, yh); c1.z = pack_bf16(1.f, 1.f); c1.w = pack_bf16(1.f, 0.f);
    *reinterpret_cast<uint4*>(blk0 + swz(row, 0)) = c0;
    *reinterpret_cast<uint4*>(blk0 + swz(row, 1)) = c1;
#pragma unroll
    for (int c = 0; c < 4; ++c) {                 // 8 hidden units per 16-byte chunk; nothing wider than that stays live
        uint32_t hi[4], mi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const Split3x2 q = split3_pair(h[8 * c + 2 * j], h[8 * c + 2 * j + 1]);
            hi[j] = q.packed[0]; mi[j] = q.packed[1]; lo[j] = q.packed[2];
        }
        *reinterpret_cast<uint4*>(blk0 + swz(row, 2 + c)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(blk1 + swz(row, c)) = make_uint4(mi[0], mi[1], mi[2], mi[3]);
        *reinterpret_cast<uint4*>(blk1 + swz(row, 4 + c)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// DECODER = false: inputs obs_rel [T,batch,2], zero initial state, output h_out [batch,32].
// DECODER = true : h0 (+ folded noise), c0 or 0, first input last_pos_rel, outputs pred_rel [T,batch,2] (+ h_final).
template <bool DECODER>
__global__ void __launch_bounds__(LTHREADS, 1)
lstm_tc_kernel(const float* __restrict__ seq_in, const float* __restrict__ h0, const float* __restrict__ c0,
               const float* __restrict__ z, const int32_t* __restrict__ ped_scene, int nz, int T, int batch, int n_tiles,
               const __nv_bfloat16* __restrict__ wimg, const float* __restrict__ W_hp, const float* __restrict__ b_hp,
               float* __restrict__ seq_out, float* __restrict__ h_out, long long* __restrict__ stats_out) {
    long long stats_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin_ = clock64();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LstmTcSmem::BARS);
    uint64_t* w_full = bars;
    uint64_t* a_ready = bars + 1;                 // [LSLOTS] count 128
    uint64_t* g_full = a_ready + LSLOTS;          // [LSLOTS] count 1 (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(g_full + LSLOTS);
    float* whp = reinterpret_cast<float*>(smem + LstmTcSmem::WHP);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rounds = (n_tiles + gridDim.x * LSLOTS - 1) / (gridDim.x * LSLOTS);

    pdl_trigger();
    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < LSLOTS; ++s) { mbar_init(&a_ready[s], 128); mbar_init(&g_full[s], 1); }
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    pdl_wait();          // everything above overlaps the tail of the previous kernel of the chain; no global access yet
    if (DECODER)
        for (int e = threadIdx.x; e < 2 * LH + 2; e += LTHREADS) whp[e] = e < 2 * LH ? W_hp[e] : b_hp[e - 2 * LH];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0u) __trap();
    constexpr uint32_t tmem = 0u;

    if (warp == 0) {
        // ======================= MMA issuer =======================
        if (lane == 0) {
            mbar_expect_tx(w_full, 5 * LT * 128);
            bulk_g2s(smem + LstmTcSmem::B, wimg, 5 * LT * 128, w_full);
        }
        mbar_wait(w_full, 0);
        constexpr uint32_t idesc = make_idesc(128, 128);
        const uint64_t bd = make_desc(sbase + LstmTcSmem::B), ad = make_desc(sbase + LstmTcSmem::A);
        int use = 0;
        for (int r = 0; r < rounds; ++r) {
            for (int t = 0; t < T; ++t, ++use) {
#pragma unroll
                for (int s = 0; s < LSLOTS; ++s) {
                    // a slot without a tile in this round (only the LAST round has such slots, so the slot's barriers
                    // are never used again): nothing is issued and its epilogue warps have left their loop -- the
                    // gate math of a dead tile would take its full share of the XU pipe from the live ones
                    if ((int)blockIdx.x + (r * LSLOTS + s) * (int)gridDim.x >= n_tiles) continue;
                    TWAIT(&a_ready[s], (uint32_t)(use & 1), s);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem + s * 128;
                        const uint64_t a0 = ad + (uint64_t)(s * 2 * LT * 128 / 16), a1 = a0 + LT * 128 / 16;
                        const uint64_t Ba = bd, Bb = bd + 1 * (LT * 128 / 16), Bc = bd + 2 * (LT * 128 / 16),
                                       Bd = bd + 3 * (LT * 128 / 16), Be = bd + 4 * (LT * 128 / 16);
                        mma_ss(d, a0 + 0, Ba + 0, idesc, 0);
                        mma_ss(d, a0 + 2, Ba + 2, idesc, 1);
                        mma_ss(d, a0 + 4, Ba + 4, idesc, 1);
                        mma_ss(d, a0 + 2, Bb + 2, idesc, 1);
                        mma_ss(d, a0 + 4, Bb + 4, idesc, 1);
                        mma_ss(d, a0 + 2, Bc + 2, idesc, 1);
                        mma_ss(d, a0 + 4, Bc + 4, idesc, 1);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) mma_ss(d, a1 + 2 * ks, Bd + 2 * ks, idesc, 1);
                        mma_ss(d, a1 + 0, Be + 0, idesc, 1);
                        mma_ss(d, a1 + 2, Be + 2, idesc, 1);
                        tc_commit(&g_full[s]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ======================= epilogue: one thread per pedestrian of the slot's tile =======================
        const int s = (warp - 1) >> 2;                       // slot
        const int row = ((warp & 3) << 5) | lane;            // tile row == TMEM lane (quadrant = warp % 4)
        const uint32_t gaddr = tmem + ((uint32_t)((warp & 3) << 5) << 16) + s * 128;
        uint8_t* blk0 = smem + LstmTcSmem::A + s * 2 * LT * 128;
        uint8_t* blk1 = blk0 + LT * 128;
        int use = 0;
        for (int r = 0; r < rounds; ++r) {
            const int tile = blockIdx.x + (r * LSLOTS + s) * gridDim.x;
            if (tile >= n_tiles) break;                      // (see the issuer: dead slots of the last round are skipped)
            const int p = tile * LT + row;
            const bool live = p < batch;
            float h[LH], c[LH];
            float2 d = make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < LH; ++u) { h[u] = 0.f; c[u] = 0.f; }
            if (live) {
                if (DECODER) {
                    const int hc = LH - nz;
                    const int sc = nz > 0 ? ped_scene[p] : 0;
#pragma unroll
                    for (int u = 0; u < LH; ++u) {
                        h[u] = (u < hc) ? h0[(int64_t)p * hc + u] : z[(int64_t)sc * nz + (u - hc)];
                        if (c0) c[u] = c0[(int64_t)p * LH + u];
                    }
                }
                d = *reinterpret_cast<const float2*>(seq_in + (int64_t)p * 2);     // obs_rel[0][p] / last_pos_rel[p]
            }
            write_a_rows(blk0, blk1, row, d.x, d.y, h);
            fence_proxy_async();
            mbar_arrive(&a_ready[s]);
            for (int t = 0; t < T; ++t, ++use) {
                float2 dn = make_float2(0.f, 0.f);
                if (!DECODER && live && t + 1 < T)
                    dn = *reinterpret_cast<const float2*>(seq_in + ((int64_t)(t + 1) * batch + p) * 2);
                TWAIT(&g_full[s], (uint32_t)(use & 1), 0);
                tc_fence_after();
#ifdef SGX_TC_STATS
                const long long te0_ = clock64();
#endif
#pragma unroll
                for (int half = 0; half < 2; ++half) {         // 16 hidden units at a time keeps the thread under 128 regs
                    uint32_t v[16];
                    float A[16], B[16];
                    tmem_ld16(gaddr + 0 + half * 16, v);       // input gate -> e_i
                    tmem_wait_ld();
#pragma unroll
                    for (int u = 0; u < 16; ++u) A[u] = ex2_clamped(__uint_as_float(v[u]), E_CLAMP);
                    tmem_ld16(gaddr + 64 + half * 16, v);      // cell candidate -> E_g
                    tmem_wait_ld();
#pragma unroll
                    for (int u = 0; u < 16; ++u) B[u] = ex2_clamped(__uint_as_float(v[u]), T_CLAMP);
                    tmem_ld16(gaddr + 32 + half * 16, v);      // forget gate
                    tmem_wait_ld();
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        c[half * 16 + u] = cell_update(c[half * 16 + u], A[u], B[u],
                                                       ex2_clamped(__uint_as_float(v[u]), E_CLAMP));
                    tmem_ld16(gaddr + 96 + half * 16, v);      // output gate
                    tmem_wait_ld();
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        h[half * 16 + u] = hidden_out(ex2_clamped(__uint_as_float(v[u]), E_CLAMP), c[half * 16 + u]);
                }
                tc_fence_before();
                if (DECODER) {
                    float rx = whp[2 * LH], ry = whp[2 * LH + 1];
#pragma unroll
                    for (int u = 0; u < LH; ++u) { rx = fmaf(whp[u], h[u], rx); ry = fmaf(whp[LH + u], h[u], ry); }
                    dn = make_float2(rx, ry);
                    if (live) *reinterpret_cast<float2*>(seq_out + ((int64_t)t * batch + p) * 2) = dn;
                }
#ifdef SGX_TC_STATS
                const long long te1_ = clock64();
                stats_[1] += te1_ - te0_;
#endif
                if (t + 1 < T) {
                    write_a_rows(blk0, blk1, row, dn.x, dn.y, h);
#ifdef SGX_TC_STATS
                    const long long te2_ = clock64();
                    stats_[2] += te2_ - te1_;
#endif
                    fence_proxy_async();
                    mbar_arrive(&a_ready[s]);
#ifdef SGX_TC_STATS
                    stats_[3] += clock64() - te2_;
#endif
                }
            }
            if (live && h_out) {
#pragma unroll
                for (int u = 0; u < LH; u += 4)
                    *reinterpret_cast<float4*>(h_out + (int64_t)p * LH + u) = make_float4(h[u], h[u + 1], h[u + 2], h[u + 3]);
            }
        }
    }
#ifdef SGX_TC_STATS
    if (stats_out && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 5)) {
        const int role = warp == 0 ? 0 : warp == 1 ? 1 : 2;
        for (int k = 0; k < 4; ++k) stats_out[role * 8 + k] = stats_[k];
        stats_out[role * 8 + 4] = clock64() - t_begin_;
        stats_out[role * 8 + 5] = (long long)rounds * T;
    }
#endif
    (void)stats_out; (void)stats_; (void)t_begin_;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

long long* g_lstm_stats = nullptr;

}  // namespace sgx

extern "C" void sgx_debug_lstm_stats(void* p) { sgx::g_lstm_stats = (long long*)p; }

using namespace sgx;

// workspace: the 5 weight tiles (80 KB)
int64_t sgx_lstm_tc_ws_bytes() { return 5 * LT * 128 + 256; }

int sgx_lstm_tc_run(bool decoder, const float* seq_in, const float* h0, const float* c0, const float* z,
                    const int32_t* ped_scene, int nz, int T, int64_t batch, const float* We, const float* be,
                    const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const float* W_hp,
                    const float* b_hp, int E, float* seq_out, float* h_out, void* ws, cudaStream_t st, bool ws_prepared) {
    __nv_bfloat16* img = reinterpret_cast<__nv_bfloat16*>(ws);
    if (!ws_prepared) {                    // (the host side keeps a prepared workspace per module and weight version)
        lstm_tc_prep_kernel<<<blocks_for(5 * LT * 64, 256), 256, 0, st>>>(We, be, W_ih, W_hh, b_ih, b_hh, E, img);
        SGX_LAUNCH_CHECK();
    }
    if (seq_in == nullptr) return SGX_OK;  // prep-only call (sgx_lstm_prep)
    const int n_tiles = (int)((batch + LT - 1) / LT);
    int dev = 0, sms = 148;
    SGX_CUDA(cudaGetDevice(&dev));
    SGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = std::min((n_tiles + LSLOTS - 1) / LSLOTS, sms);
    if (decoder) {
        auto kern = lstm_tc_kernel<true>;
        SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LstmTcSmem::TOTAL));
        SGX_CUDA(launch_pdl(kern, dim3(grid), dim3(LTHREADS), LstmTcSmem::TOTAL, st, true, seq_in, h0, c0, z, ped_scene, nz, T,
                            (int)batch, n_tiles, img, W_hp, b_hp, seq_out, h_out, g_lstm_stats));
    } else {
        auto kern = lstm_tc_kernel<false>;
        SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, LstmTcSmem::TOTAL));
        SGX_CUDA(launch_pdl(kern, dim3(grid), dim3(LTHREADS), LstmTcSmem::TOTAL, st, true, seq_in, nullptr, nullptr, nullptr,
                            nullptr, 0, T, (int)batch, n_tiles, img, nullptr, nullptr, nullptr, h_out, g_lstm_stats));
    }
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}
