// PoolHiddenNet forward on the 5th-gen tensor cores: tcgen05.mma + TMEM, bf16 operands, fp32 accumulate.
//
// Reference math (sgan/models.py:530-541), per ordered pair (i,j) of a scene:
//     y = ReLU(W2 ReLU(W1 [We (P_j-P_i)+be ; h_j] + b1) + b2)   ;   out_i = max_j y
// Tensor-core formulation.  A tile is 128 consecutive ordered pairs of the flat (i,j)-sorted pair list
// (any mix of scenes).  Per tile:
//   GEMM1  D1[128 x 512] = X[128 x K] . W1p[512 x K]^T        K = 16 + H  (48 for G, 64 for D)
//          X row = [dx_hi dx_lo dx_hi dy_hi dy_lo dy_hi 1 1 0.. | bf16(h_j)]
//          W1p row k = [A0_hi A0_hi A0_lo A1_hi A1_hi A1_lo c_hi c_lo 0.. | bf16(W1[k, E:])]
//          with Aeff = W1[:, :E] We and c = W1[:, :E] be + b1 folded on the host side of the kernel (exact
//          algebra; hi/lo bf16 splits keep the position term and the bias at ~16 mantissa bits).
//          Issued as 4 hidden chunks of N = 128 so the fp32 accumulator (128 TMEM columns) double-buffers.
//   EPI1   TMEM -> registers (tcgen05.ld 32x32b.x32), ReLU + bf16x2 pack (cvt.rn.relu.bf16x2.f32),
//          registers -> TMEM (tcgen05.st) as the A operand of GEMM2 -- the 128 x 512 hidden tile never
//          touches shared or global memory ("ts" mode; "ss" mode stages it in shared memory instead).
//   GEMM2  D2[128 x N2] += Hc[128 x 128] . W2p[N2 x 128]^T  per hidden chunk  (N2 = 16 for B = 8, 48 for B = 48)
//   EPI2   TMEM -> registers, + b2, ReLU, segmented warp max over rows that share i, one 64-bit
//          atomicMax (value bits << 32 | j) per (segment, channel) into the packed output.
// Weights are pre-swizzled bf16 images (SWIZZLE_128B, K-major) loaded once per persistent CTA with the
// bulk-copy engine (cp.async.bulk + mbarrier complete_tx).  One CTA per SM, 17 warps:
//   warp 0       MMA issuer (one elected lane)
//   warps 1-4    row warps, set 0: build X for even tiles of this CTA, EPI2 for the same tiles
//   warps 5-8    row warps, set 1: odd tiles (two tile builds in flight hide the dependent-load chain)
//   warps 9-12   EPI1 for even hidden chunks (TMEM D1 buffer 0)   warps 13-16  EPI1 for odd chunks (buffer 1)
#include <stdlib.h>

#include "sgx_tc.cuh"

namespace sgx {

constexpr int HID = SGX_POOL_HIDDEN;
constexpr int TILE = 128;        // pairs per tile = UMMA M
constexpr int NCHUNK = 4;        // hidden chunks of 128
constexpr int NST = 4;           // X-tile smem stages (row-warp set s owns stages s and s+2)
constexpr int NMETA = 4;         // ring of per-tile (i,j) metadata
constexpr int NTHREADS = 18 * 32;
constexpr int SLICE = 136;       // per-tile slice of pair_off / ped_start staged in smem for the pair decode

// ---------------------------------------------------------------------------------------------
// operand preparation: bf16 copy of h, pre-swizzled weight images
// ---------------------------------------------------------------------------------------------
__global__ void tc_prep_h_kernel(const float* __restrict__ h, int64_t n8, __nv_bfloat16* __restrict__ hb) {
    // eight values per thread: two 16-byte loads, one 16-byte store (n = 8 n8; H is a multiple of 8)
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = reinterpret_cast<const float4*>(h)[2 * i], b = reinterpret_cast<const float4*>(h)[2 * i + 1];
    reinterpret_cast<uint4*>(hb)[i] = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
}

// W1p image: 512 rows x 128 B.   W2p image: 8 K-blocks x N2 rows x 128 B.
__global__ void tc_prep_w_kernel(const float2* __restrict__ Aeff, const float* __restrict__ c0,
                                 const float* __restrict__ W1, const float* __restrict__ W2, int E, int H, int B, int N2,
                                 __nv_bfloat16* __restrict__ W1p, __nv_bfloat16* __restrict__ W2p) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < HID * 64) {
        int n = t / 64, k = t % 64;
        float v = 0.f;
        if (k < 16) {
            float2 a = Aeff[n];
            float c = c0[n];
            float src = (k < 3) ? a.x : (k < 6) ? a.y : c;
            float hi = __bfloat162float(__float2bfloat16_rn(src));
            float lo = src - hi;
            bool want_lo = (k == 2 || k == 5 || k == 7);
            v = (k < 8) ? (want_lo ? lo : hi) : 0.f;
        } else if (k - 16 < H) {
            v = W1[(int64_t)n * (E + H) + E + (k - 16)];
        }
        uint32_t off = swz((uint32_t)n, (uint32_t)(k >> 3)) + (k & 7) * 2;
        W1p[off >> 1] = __float2bfloat16_rn(v);
    }
    if (t < 8 * N2 * 64) {
        int kb = t / (N2 * 64), r = (t / 64) % N2, kk = t % 64;
        float v = (r < B) ? W2[(int64_t)r * HID + kb * 64 + kk] : 0.f;
        uint32_t off = (uint32_t)kb * N2 * 128 + swz((uint32_t)r, (uint32_t)(kk >> 3)) + (kk & 7) * 2;
        W2p[off >> 1] = __float2bfloat16_rn(v);
    }
}

__device__ __forceinline__ int find_ped_tc(const int64_t* __restrict__ pair_off, int lo, int hi, int64_t q) {
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pair_off[mid] <= q) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct TcSmem {   // dynamic shared memory layout (offsets from a 1024-aligned base)
    static constexpr int W1P = 0;                                  // 64 KB
    static constexpr int X = W1P + HID * 128;                      // NST x 16 KB
    static constexpr int W2P = X + NST * TILE * 128;               // 8 * N2 * 128
};

template <int H, int N2, bool TS>
struct TcCfg {
    static constexpr int KCH = (16 + H) / 16;                      // k-chunks of GEMM1
    static constexpr int W2P_BYTES = 8 * N2 * 128;
    static constexpr int HS = TcSmem::W2P + W2P_BYTES;             // "ss" mode: 2 x 32 KB hidden staging
    static constexpr int HS_BYTES = TS ? 0 : 2 * TILE * 256;
    static constexpr int META = HS + HS_BYTES;                     // int2 [NMETA][128]
    static constexpr int SOFF = META + NMETA * TILE * 8;           // int64 [2][SLICE]
    static constexpr int SPS = SOFF + 2 * SLICE * 8;               // int32 [2][SLICE]
    static constexpr int BARS = SPS + 2 * SLICE * 4;
    static constexpr int TOTAL = BARS + 256 + 1024;                // + alignment slack
    static constexpr int D2_STRIDE = 64;                           // TMEM columns per D2 buffer
    static constexpr int TM_D1 = 0, TM_H = 256, TM_D2 = 384;
};

// warp roles (18 warps): 0-3 row set 0, 4-7 row set 1, 8-11 EPI1 even chunks, 12-15 EPI1 odd chunks,
// 16 GEMM1 issuer, 17 GEMM2 issuer.  Two issuer threads because a single thread issuing 44 small MMAs per tile
// (each ~15 SASS instructions of descriptor set-up on the uniform datapath) was the measured bottleneck.
template <int H, int B, int N2, bool TS>
__global__ void __launch_bounds__(NTHREADS, 1)
pool_tc_kernel(const __nv_bfloat16* __restrict__ hb, const float* __restrict__ pos,
               const int32_t* __restrict__ ped_start, const int64_t* __restrict__ pair_off,
               const int32_t* __restrict__ tile_first, int64_t n_tiles, int batch, int64_t n_pairs,
               const __nv_bfloat16* __restrict__ W1p, const __nv_bfloat16* __restrict__ W2p,
               const float* __restrict__ b2, unsigned long long* __restrict__ packed,
               long long* __restrict__ stats_out) {
    using C = TcCfg<H, N2, TS>;
    static_assert(B % 8 == 0, "the final epilogue reduces eight columns at a time");
    long long stats_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_begin_ = clock64();
    extern __shared__ uint8_t smem_raw[];
    // align inside the 32-bit shared window so every derived address stays warp-uniform for the compiler
    const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (sbase - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::BARS);
    uint64_t* w_full = bars + 0;
    uint64_t* x_full = bars + 1;             // [NST]
    uint64_t* x_free = x_full + NST;         // [NST]
    uint64_t* d1_full = x_free + NST;        // [2]
    uint64_t* d1_free = d1_full + 2;         // [2]
    uint64_t* h_ready = d1_free + 2;         // [2]
    uint64_t* h_free = h_ready + 2;          // [2]
    uint64_t* d2_full = h_free + 2;          // [2]
    uint64_t* d2_free = d2_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_free + 2);
    int2* meta = reinterpret_cast<int2*>(smem + C::META);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int my_tiles = (n_tiles > blockIdx.x) ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (threadIdx.x == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < NST; ++s) { mbar_init(&x_full[s], 128); mbar_init(&x_free[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&d1_full[s], 1); mbar_init(&d1_free[s], 128);
            mbar_init(&h_ready[s], 128); mbar_init(&h_free[s], 1);
            mbar_init(&d2_full[s], 1); mbar_init(&d2_free[s], 128);
        }
        fence_barrier_init();
    }
    if (warp == 16) {   // TMEM: all 512 columns (one CTA per SM) => the allocation starts at lane 0 / column 0
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (*tmem_slot != 0u) __trap();      // TMEM addresses below are compile-time constants relative to base 0
    constexpr uint32_t tmem = 0u;

    if (warp == 16) {
        // ======================= GEMM1 issuer =======================
        if (lane == 0) {
            mbar_expect_tx(w_full, HID * 128 + C::W2P_BYTES);
            bulk_g2s(smem + TcSmem::W1P, W1p, HID * 128, w_full);
            bulk_g2s(smem + TcSmem::W2P, W2p, C::W2P_BYTES, w_full);
        }
        mbar_wait(w_full, 0);
        constexpr uint32_t idesc1 = make_idesc(128, 128);
        const uint64_t w1_d = make_desc(sbase + TcSmem::W1P), x_d = make_desc(sbase + TcSmem::X);
        for (int t = 0; t < my_tiles; ++t) {
            const int st = t & (NST - 1);
            TWAIT(&x_full[st], (uint32_t)((t / NST) & 1), 0);
            const uint64_t xa = x_d + (uint64_t)(st * (TILE * 128 / 16));
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                TWAIT(&d1_free[c & 1], (uint32_t)(((c >> 1) & 1) ^ 1), 1);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kc = 0; kc < C::KCH; ++kc)
                        mma_ss(tmem + C::TM_D1 + (c & 1) * 128, xa + kc * 2, w1_d + (c * 16384 + kc * 32) / 16, idesc1, kc > 0);
                    tc_commit(&d1_full[c & 1]);
                    if (c == NCHUNK - 1) tc_commit(&x_free[st]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 17) {
        // ======================= GEMM2 issuer =======================
        mbar_wait(w_full, 0);
        constexpr uint32_t idesc2 = make_idesc(128, N2);
        const uint64_t w2_d = make_desc(sbase + TcSmem::W2P), hs_d = make_desc(sbase + C::HS);
        for (int t = 0; t < my_tiles; ++t) {
            const int db = t & 1;
            TWAIT(&d2_free[db], (uint32_t)(((t >> 1) & 1) ^ 1), 2);
            const uint32_t d2 = tmem + C::TM_D2 + db * C::D2_STRIDE;
#pragma unroll
            for (int c = 0; c < NCHUNK; ++c) {
                TWAIT(&h_ready[c & 1], (uint32_t)((c >> 1) & 1), 3);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint64_t bd = w2_d + ((c * 2 + (kk >> 2)) * (N2 * 128) + (kk & 3) * 32) / 16;
                        const uint32_t accf = (c > 0 || kk > 0);
                        if (TS)
                            mma_ts(d2, tmem + C::TM_H + (c & 1) * 64 + kk * 8, bd, idesc2, accf);
                        else
                            mma_ss(d2, hs_d + ((c & 1) * (TILE * 256) + (kk >> 2) * 16384 + (kk & 3) * 32) / 16, bd, idesc2, accf);
                    }
                    tc_commit(&h_free[c & 1]);
                    if (c == NCHUNK - 1) tc_commit(&d2_full[db]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 8) {
        // ======================= row warps: build X tiles, final epilogue =======================
        const int set = warp >> 2;                   // set s handles this CTA's tiles it = 2k + s
        const int rt = (warp & 3) * 32 + lane;       // thread index inside the set (0..127)
        const int row = rt;                          // tile row == TMEM lane (quadrant = warp % 4)
        int64_t* soff = reinterpret_cast<int64_t*>(smem + C::SOFF) + set * SLICE;
        int32_t* sps = reinterpret_cast<int32_t*>(smem + C::SPS) + set * SLICE;
        float bias2[B];
#pragma unroll
        for (int b = 0; b < B; ++b) bias2[b] = b2[b];
        // The per-tile decode needs tile_first -> (pair_off, ped_start) slice -> (pos, hb) gathers: three dependent
        // global-load levels.  All three are software-pipelined: the slice one more tile ahead in registers, and the
        // gathers of the set's NEXT tile are issued before the finalize of its previous tile and consumed at the top of
        // the next iteration, so no load latency sits on the row warps' critical path (they bound small-scene batches).
        int lo = 0, hi = -1, sp = 0, sp_x = 0;       // slice of the tile whose rows are decoded next
        int64_t so = 0, so_x = 0;
        int lo2 = 0, hi2 = -1;                       // bounds of the tile after that (one more level of look-ahead)
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
        // prefetched row of the set's next tile
        int2 ij = make_int2(-1, 0);
        float2 pi = make_float2(0.f, 0.f), pj = make_float2(0.f, 0.f);
        uint4 hv[H / 8];
        bool valid = false;
        auto prefetch = [&](int itn) {
            const int64_t tile = blockIdx.x + (int64_t)itn * gridDim.x;
            // ---- publish this tile's slice of the ped tables, decode (i,j) from shared memory ----
            if (lo + rt <= hi) { soff[rt] = so; sps[rt] = sp; }
            if (rt == 0 && lo + 128 <= hi) { soff[128] = so_x; sps[128] = sp_x; }
            const int lo_c = lo, hi_c = hi;
            int lo3, hi3;
            bounds(itn + 4, lo3, hi3);                           // consumed two prefetches from now
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");
            const int64_t q = tile * TILE + row;
            ij = make_int2(-1, 0);
            pi = make_float2(0.f, 0.f); pj = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < H / 8; ++c) hv[c] = make_uint4(0, 0, 0, 0);
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
                const uint4* hrow = reinterpret_cast<const uint4*>(hb + (int64_t)j * H);
#pragma unroll
                for (int c = 0; c < H / 8; ++c) hv[c] = hrow[c];
            }
            lo = lo2; hi = hi2; lo2 = lo3; hi2 = hi3;           // the following tile's slice -> registers
            if (lo + rt <= hi) { so = pair_off[lo + rt]; sp = ped_start[lo + rt]; }
            if (rt == 0 && lo + 128 <= hi) { so_x = pair_off[lo + 128]; sp_x = ped_start[lo + 128]; }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + set) : "memory");   // slice may be overwritten next round
        };
        if (set < my_tiles) prefetch(set);
        for (int it = set; it < my_tiles + 2; it += 2) {
            if (it < my_tiles) {                                 // ---- X stage of tile `it` from the prefetched row ----
                const int st = it & (NST - 1);
                uint8_t* xrow = smem + TcSmem::X + st * TILE * 128;
                TWAIT(&x_free[st], (uint32_t)(((it / NST) & 1) ^ 1), 0);
                uint4 c0 = make_uint4(0, 0, 0, 0);
                if (valid) {
                    const float dx = pj.x - pi.x, dy = pj.y - pi.y;
                    const float dxh = __bfloat162float(__float2bfloat16_rn(dx)), dyh = __bfloat162float(__float2bfloat16_rn(dy));
                    c0.x = pack_bf16(dxh, dx - dxh);       // slots 0,1
                    c0.y = pack_bf16(dxh, dyh);            // slots 2,3
                    c0.z = pack_bf16(dy - dyh, dyh);       // slots 4,5
                    c0.w = pack_bf16(1.f, 1.f);            // slots 6,7
                }
                *reinterpret_cast<uint4*>(xrow + swz(row, 0)) = c0;
                *reinterpret_cast<uint4*>(xrow + swz(row, 1)) = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int c = 0; c < H / 8; ++c) *reinterpret_cast<uint4*>(xrow + swz(row, 2 + c)) = hv[c];
                meta[(it & (NMETA - 1)) * TILE + row] = ij;
                fence_proxy_async();
                mbar_arrive(&x_full[st]);
            }
            if (it + 2 < my_tiles) {                             // gathers in flight across the finalize below
#ifdef SGX_TC_STATS
                const long long tp__ = clock64();
#endif
                prefetch(it + 2);
#ifdef SGX_TC_STATS
                stats_[2] += clock64() - tp__;
#endif
            }
            if (it >= 2) {
                const int t = it - 2;                // the previous tile of this set; D2 buffer = t & 1 = set
                const int db = set;
                TWAIT(&d2_full[db], (uint32_t)((t >> 1) & 1), 1);
#ifdef SGX_TC_STATS
                const long long tf__ = clock64();
#endif
                tc_fence_after();
                const uint32_t d2a = tmem + ((uint32_t)((warp & 3) << 5) << 16) + C::TM_D2 + db * C::D2_STRIDE;
                const int2 mij = meta[(t & (NMETA - 1)) * TILE + row];
                const int key = mij.x;
                const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
                const bool head = (lane == 0) || (key_prev != key);
                const bool uniform = __all_sync(0xffffffffu, key == __shfl_sync(0xffffffffu, key, 0));
                bool same[5];                           // lane + 2^s belongs to the same pedestrian i (hoisted out of b)
#pragma unroll
                for (int sft = 0; sft < 5; ++sft) {
                    const int okey = __shfl_down_sync(0xffffffffu, key, 1 << sft);
                    same[sft] = (lane + (1 << sft) < 32) && (okey == key);
                }
                constexpr int NG = (B + 15) / 16;       // 16 accumulator columns at a time (register budget of the D dims)
#pragma unroll
                for (int g = 0; g < NG; ++g) {
                    uint32_t v[16];
                    tmem_ld16(d2a + 16 * g, v);
                    tmem_wait_ld();
                    if (g == NG - 1) {                  // all of D2 has been read: GEMM2 may overwrite this buffer
                        tc_fence_before();
                        mbar_arrive(&d2_free[db]);
                    }
                    auto column_bits = [&](int c) {     // c: column inside the group
                        // a NaN keeps its (canonical) bit pattern: it orders above every finite value, like torch.max
                        const float y = __uint_as_float(v[c]) + bias2[16 * g + c];
                        return (y != y) ? 0x7fc00000u : (__float_as_uint(fmaxf(y, 0.f)) & 0x7fffffffu);
                    };
                    // The uniform / segmented choice sits outside the column loops: with the branch inside, every
                    // column was its own basic block and the dependent shuffle chains ran one after the other.
                    if (uniform) {
                        // the whole warp belongs to one pedestrian i (dense crowd): one REDUX + one atomic per column
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            if (16 * g + c < B) {
                                const uint32_t bits = column_bits(c);
                                const uint32_t mx = __reduce_max_sync(0xffffffffu, bits);
                                const uint32_t who = __ballot_sync(0xffffffffu, bits == mx);
                                const int src = 31 - __clz(who);                       // ties -> larger j
                                const int jj = __shfl_sync(0xffffffffu, mij.y, src);
                                if (lane == 0 && key >= 0)
                                    atomicMax(&packed[(int64_t)key * B + 16 * g + c],
                                              ((unsigned long long)mx << 32) | (unsigned)jj);
                            }
                        }
                    } else {
#pragma unroll
                        for (int c0 = 0; c0 < 16; c0 += 8) {       // eight independent scan chains at a time
                            if (16 * g + c0 < B) {
                                unsigned long long pk[8];
#pragma unroll
                                for (int c = 0; c < 8; ++c)
                                    pk[c] = ((unsigned long long)column_bits(c0 + c) << 32) | (unsigned)mij.y;
#pragma unroll
                                for (int sft = 0; sft < 5; ++sft) {
#pragma unroll
                                    for (int c = 0; c < 8; ++c) {
                                        const unsigned long long other = __shfl_down_sync(0xffffffffu, pk[c], 1 << sft);
                                        if (same[sft] && other > pk[c]) pk[c] = other;
                                    }
                                }
                                if (head && key >= 0) {
#pragma unroll
                                    for (int c = 0; c < 8; ++c)
                                        atomicMax(&packed[(int64_t)key * B + 16 * g + c0 + c], pk[c]);
                                }
                            }
                        }
                    }
                }
#ifdef SGX_TC_STATS
                stats_[3] += clock64() - tf__;
#endif
            }
        }
    } else {
        // ======================= EPI1 warps: D1 -> ReLU -> bf16 -> H =======================
        const int grp = (warp - 8) >> 2;               // 0: even chunks (buffer 0), 1: odd chunks (buffer 1)
        const int quad = warp & 3;
        const int row = (quad << 5) | lane;
        const uint32_t lane_base = (uint32_t)(quad << 5) << 16;
        const int n_chunks = my_tiles * NCHUNK;
        int use = 0;
        for (int g = grp; g < n_chunks; g += 2, ++use) {
            TWAIT(&d1_full[grp], (uint32_t)(use & 1), 0);
            tc_fence_after();
            const uint32_t d1a = tmem + lane_base + C::TM_D1 + grp * 128;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t va[32], vb[32];
                tmem_ld32(d1a + half * 64, va);
                tmem_ld32(d1a + half * 64 + 32, vb);
                tmem_wait_ld();
                if (half == 1) {          // all of D1 has been read: the GEMM1 issuer may overwrite this accumulator
                    tc_fence_before();
                    mbar_arrive(&d1_free[grp]);
                }
                uint32_t p[32];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    p[e] = relu_pack(va[2 * e + 1], va[2 * e]);
                    p[16 + e] = relu_pack(vb[2 * e + 1], vb[2 * e]);
                }
                if (half == 0) {          // GEMM2 of chunk g-2 must have consumed this H buffer
                    TWAIT(&h_free[grp], (uint32_t)((use & 1) ^ 1), 1);
                    tc_fence_after();
                }
                if (TS) {
                    tmem_st32(tmem + lane_base + C::TM_H + grp * 64 + half * 32, p);
                } else {
                    uint8_t* hs = smem + C::HS + grp * (TILE * 256) + half * 16384;
#pragma unroll
                    for (int cc = 0; cc < 8; ++cc)
                        *reinterpret_cast<uint4*>(hs + swz(row, cc)) =
                            make_uint4(p[4 * cc], p[4 * cc + 1], p[4 * cc + 2], p[4 * cc + 3]);
                }
            }
            if (TS) tmem_wait_st(); else fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&h_ready[grp]);
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

long long* g_tc_stats = nullptr;   // debug: device buffer of 40 long longs (SGX_TC_STATS builds)

template <int H, int B, int N2, bool TS>
static int launch_tc(const __nv_bfloat16* hb, const float* pos, const int32_t* ped_start, const int64_t* pair_off,
                     const int32_t* tile_first, int64_t n_tiles, int batch, int64_t n_pairs, const __nv_bfloat16* W1p,
                     const __nv_bfloat16* W2p, const float* b2, unsigned long long* packed, cudaStream_t st) {
    using C = TcCfg<H, N2, TS>;
    auto kern = pool_tc_kernel<H, B, N2, TS>;
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::TOTAL));
    int dev = 0, sms = 148;
    SGX_CUDA(cudaGetDevice(&dev));
    SGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    unsigned grid = (unsigned)std::min<int64_t>(n_tiles, sms);
    cudaEvent_t ev0, ev1;
    profile_events(&ev0, &ev1);
    if (ev0 && ev1) SGX_CUDA(cudaEventRecord(ev0, st));
    kern<<<grid, NTHREADS, C::TOTAL, st>>>(hb, pos, ped_start, pair_off, tile_first, n_tiles, batch, n_pairs, W1p, W2p,
                                           b2, packed, g_tc_stats);
    SGX_LAUNCH_CHECK();
    if (ev0 && ev1) SGX_CUDA(cudaEventRecord(ev1, st));
    return SGX_OK;
}

__global__ void pool_prep_kernel(const float* __restrict__ We, const float* __restrict__ be,
                                 const float* __restrict__ W1, const float* __restrict__ b1, int E, int H,
                                 float2* __restrict__ Aeff, float* __restrict__ c0);

}  // namespace sgx

using namespace sgx;

static int n2_for(int B) { return B <= 16 ? 16 : 48; }

// prepared weights (per weight version): [W1p | W2p | Aeff | c0]
int64_t sgx_pool_bf16_prep_bytes(int E, int H, int B) {
    (void)E; (void)H;
    return align_up(HID * 128, 256) + align_up(8 * n2_for(B) * 128, 256) + align_up(HID * 8, 256) + align_up(HID * 4, 256);
}
int64_t sgx_pool_bf16_ws_bytes(int64_t batch, int E, int H, int B) {
    (void)E; (void)B;
    return align_up(batch * H * 2, 256);
}

// 1 when the current device can run the tcgen05 kernels of this build (compute capability 10.x)
extern "C" int sgx_has_tcgen05(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}
extern "C" void sgx_debug_tc_stats(void* dev_buf) { sgx::g_tc_stats = (long long*)dev_buf; }

static bool bf16_supported(int H, int B) { return (H == 32 && B == 8) || (H == 48 && B == 48); }

int sgx_pool_bf16_prep(const float* We, const float* be, const float* W1, const float* b1, const float* W2, int E, int H,
                       int B, void* prep, cudaStream_t st) {
    SGX_UNSUPPORTED(!bf16_supported(H, B),
                    "bf16 tensor-core pooling is built for (h_dim, bottleneck) in {(32,8), (48,48)}; got (%d,%d) -- "
                    "use precision fp32", H, B);
    const int N2 = n2_for(B);
    Carver c(prep);
    __nv_bfloat16* W1p = c.take<__nv_bfloat16>(HID * 64);
    __nv_bfloat16* W2p = c.take<__nv_bfloat16>(8 * N2 * 64);
    float2* Aeff = c.take<float2>(HID);
    float* c0 = c.take<float>(HID);
    pool_prep_kernel<<<2, 256, 0, st>>>(We, be, W1, b1, E, H, Aeff, c0);
    SGX_LAUNCH_CHECK();
    tc_prep_w_kernel<<<blocks_for(HID * 64, 256), 256, 0, st>>>(Aeff, c0, W1, W2, E, H, B, N2, W1p, W2p);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

int sgx_pool_fwd_bf16(const float* h, const float* pos, const int32_t* ped_start, const int64_t* pair_off,
                      const int32_t* tile_first, int64_t batch, int64_t n_pairs, const void* prep, const float* b2, int E,
                      int H, int B, unsigned long long* packed, void* ws, cudaStream_t st) {
    (void)E;
    SGX_UNSUPPORTED(!bf16_supported(H, B),
                    "bf16 tensor-core pooling is built for (h_dim, bottleneck) in {(32,8), (48,48)}; got (%d,%d) -- "
                    "use precision fp32", H, B);
    SGX_REQUIRE(n_pairs < ((int64_t)1 << 40), "sgx_pool_fwd_bf16: too many pairs");
    const int N2 = n2_for(B);
    Carver pc(const_cast<void*>(prep));
    const __nv_bfloat16* W1p = pc.take<__nv_bfloat16>(HID * 64);
    const __nv_bfloat16* W2p = pc.take<__nv_bfloat16>(8 * N2 * 64);
    Carver c(ws);
    __nv_bfloat16* hb = c.take<__nv_bfloat16>(batch * H);
    tc_prep_h_kernel<<<blocks_for(batch * H / 8, 256), 256, 0, st>>>(h, batch * H / 8, hb);
    SGX_LAUNCH_CHECK();
    const int64_t n_tiles = (n_pairs + TILE - 1) / TILE;
    if (H == 32) {
#ifdef SGX_AB_VARIANTS   // A/B build only: GEMM2 A operand staged in shared memory instead of TMEM
        const char* mode = getenv("SGX_POOL_TC_MODE");
        if (mode && mode[0] == 's')
            return launch_tc<32, 8, 16, false>(hb, pos, ped_start, pair_off, tile_first, n_tiles, (int)batch, n_pairs,
                                               W1p, W2p, b2, packed, st);
#endif
        return launch_tc<32, 8, 16, true>(hb, pos, ped_start, pair_off, tile_first, n_tiles, (int)batch, n_pairs, W1p,
                                          W2p, b2, packed, st);
    }
    return launch_tc<48, 48, 48, true>(hb, pos, ped_start, pair_off, tile_first, n_tiles, (int)batch, n_pairs, W1p, W2p,
                                       b2, packed, st);
}
