// GCNModule backward in ONE launch (+ a tiny reduction) for batches whose scenes fit a warp chunk (<= 32 peds) --
// what autograd does through sgan/models.py:628-712 (GCN.forward 573-580) with dense [N,N] adjacency products.
//
// A_intra has fl(1/|g|) on the row's group and A_inter = 1/G, so A.H is a segmented mean and the rows of a group (of a
// scene) are identical.  Collapsed per group g (k members, a = 1/k) and per scene s (G groups, c = 1/G):
//     forward   m1_g = mean_g x ;  h1_g = relu(m1_g W0) ;  x1_g = relu(h1_g W1) ;  n1_s = mean_s x1 ;
//               k1_s = relu(n1_s V0) ;  y_s = relu(k1_s V1) ;  out_i = [x1_g(i) | a_i y_s(i)] Wo^T + bo
//     backward  dcat = dout Wo ;  D2_s = (sum_{i in s} a_i dcat_i[16:]) * [y_s > 0] ;  E2_s = (D2_s V1^T) * [k1_s > 0] ;
//               dXg_g = c E2_s V0^T ;  D1_g = (sum_{i in g} dcat_i[:16] + dXg_g) * [x1_g > 0] ;  E_g = (D1_g W1^T) * [h1_g > 0] ;
//               dx_i = a_i E_g(i) W0^T ;   dWo += dout^T cat, dV1 += k1^T D2, dV0 += n1^T E2, dW1 += h1^T D1, dW0 += m1^T E.
// (k a and G c are 1 up to one rounding; the reference's own gradient carries the same rounding.)
// A warp PAIR per chunk of whole scenes, lanes <-> pedestrians (the main warp); group rows live at the leader's slot, scene
// rows at the scene's first slot; every linear map is a warp-level 3xTF32 tensor-core GEMM (sgx_warp_mma.cuh), one m-tile
// per warp of the pair; parameter gradients are added with red.global into the CTA's own gradient block in HBM and the
// blocks are reduced in block order (reproducible to rounding: the adds within a block are floating-point atomics).
#include "sgx_common.cuh"
#include "sgx_warp_mma.cuh"

namespace sgx {

namespace gcnb {

constexpr int HID = 72, OUT = 16;
#ifndef GCB_GLOBAL_GRADS
#define GCB_GLOBAL_GRADS 1           // 1: the CTA's gradient block lives in its HBM partial (red.global), 6 chunks in flight per SM;
#endif                               // 0: in shared memory (28 KB), 5 chunks
constexpr int PAIRS = GCB_GLOBAL_GRADS ? 6 : 5;          // warp PAIRS: two warps share a chunk's scratch (see the kernel)
constexpr int WARPS = 2 * PAIRS;
constexpr int RS = 76;                 // 72-wide rows
constexpr int RA = 20;                 // 16-wide rows
constexpr int SWW = 72;                // W0 / V0 blocks [K][72]
constexpr int SWN = 24;                // W1 / V1 blocks [72][16 -> 24]

template <int IN, int FIN>
struct Cfg {
    static constexpr int SWO = (FIN % 32 == 0) ? FIN + 8 : FIN;       // Wo^T block [32][SWO]
    static constexpr int RG = RS;                                     // grad_out rows: columns 40.. of the P rows (m1 and dM1
                                                                      // use columns < 40), so they cost no scratch of their own
    static_assert(IN <= 40 && 40 + FIN <= RS, "grad_out rows live behind the m1 columns of P");
    static constexpr int WFLOATS = IN * SWW + 2 * HID * SWN + OUT * SWW + 2 * OUT * SWO;
    static constexpr int GRAD_FLOATS = IN * HID + HID * OUT + OUT * HID + HID * OUT + FIN * 2 * OUT + FIN;   // W0 W1 V0 V1 Wo bo
    static constexpr int SCRATCH = 2 * 32 * RS + 5 * 32 * RA;
    static constexpr int SMEM = (WFLOATS + (GCB_GLOBAL_GRADS ? 0 : GRAD_FLOATS) + PAIRS * SCRATCH) * (int)sizeof(float);
};

template <int F>
__device__ __forceinline__ void store_row(float* __restrict__ row, const float (&v)[F]) {
#pragma unroll
    for (int f = 0; f < F / 4; ++f)
        reinterpret_cast<float4*>(row)[f] = make_float4(v[4 * f], v[4 * f + 1], v[4 * f + 2], v[4 * f + 3]);
}

template <int IN, int FIN>
__global__ void __launch_bounds__(WARPS * 32)
gcn_fused_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gout, const int32_t* __restrict__ leader,
                     const int32_t* __restrict__ gsize, const int32_t* __restrict__ ped_start,
                     const int32_t* __restrict__ ped_end, const int32_t* __restrict__ scene_start,
                     const int32_t* __restrict__ chunk_scene, int n_chunks, const float* __restrict__ W0,
                     const float* __restrict__ W1, const float* __restrict__ V0, const float* __restrict__ V1,
                     const float* __restrict__ Wo, float* __restrict__ grad_x, float* __restrict__ partials) {
    using C = Cfg<IN, FIN>;
    constexpr int SWO = C::SWO, RG = C::RG;
    extern __shared__ __align__(16) uint8_t raw[];
    float* sW0 = reinterpret_cast<float*>(raw);               // [IN][72]
    float* sW1 = sW0 + IN * SWW;                              // [72][24]
    float* sV0 = sW1 + HID * SWN;                             // [16][72]
    float* sV1 = sV0 + OUT * SWW;                             // [72][24]
    float* sWoT = sV1 + HID * SWN;                            // [32][SWO] = Wo^T
#if GCB_GLOBAL_GRADS
    float* gW0 = partials + (int64_t)blockIdx.x * C::GRAD_FLOATS;      // gradient block = this CTA's partial in HBM (red.global)
#else
    float* gW0 = sWoT + 2 * OUT * SWO;                        // gradient block, in the order of the partials
#endif
    float* gW1 = gW0 + IN * HID;
    float* gV0 = gW1 + HID * OUT;
    float* gV1 = gV0 + OUT * HID;
    float* gWo = gV1 + HID * OUT;
    float* gbo = gWo + FIN * 2 * OUT;
#if GCB_GLOBAL_GRADS
    float* bufs = sWoT + 2 * OUT * SWO;
#else
    float* bufs = gbo + FIN;
#endif
    for (int e = threadIdx.x; e < IN * HID; e += blockDim.x) sW0[e] = W0[e];
    for (int e = threadIdx.x; e < OUT * HID; e += blockDim.x) sV0[e] = V0[e];
    for (int e = threadIdx.x; e < HID * SWN; e += blockDim.x) {
        const int k = e / SWN, n = e % SWN;
        sW1[e] = n < OUT ? W1[k * OUT + n] : 0.f;
        sV1[e] = n < OUT ? V1[k * OUT + n] : 0.f;
    }
    for (int e = threadIdx.x; e < 2 * OUT * SWO; e += blockDim.x) {
        const int k = e / SWO, n = e % SWO;
        sWoT[e] = n < FIN ? Wo[n * 2 * OUT + k] : 0.f;
    }
    for (int e = threadIdx.x; e < C::GRAD_FLOATS; e += blockDim.x) gW0[e] = 0.f;
#if GCB_GLOBAL_GRADS
    __threadfence();
#endif
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp >> 1, role = warp & 1;
    const int g = lane >> 2, t = lane & 3;
    float* P = bufs + pair * C::SCRATCH;                      // [32][RS] x rows -> group mean m1 (leaders) -> dM1
    float* Q = P + 32 * RS;                                   // [32][RS] h1 / k1 (post-ReLU) -> E2 / E
    float* X1s = Q + 32 * RS;                                 // [32][RA] x1 at the leader slots
    float* N1s = X1s + 32 * RA;                               // [32][RA] n1 at the scene's first slot
    float* A = N1s + 32 * RA;                                 // [32][RA] x1[lead] rows / dcat[:16] / dXg (scene rows)
    float* B = A + 32 * RA;                                   // [32][RA] x2 rows / dcat[16:] / D2 / D1
    float* Ys = B + 32 * RA;                                  // [32][RA] y at the scene's first slot
    float* G = P + 40;                                        // [32][RS] grad_out rows, behind the m1 columns of P

    auto relu_to = [&](float* dst, int stride) {              // GEMM result -> ReLU -> rows of dst
        return [=](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            *reinterpret_cast<float2*>(dst + r * stride + col) = make_float2(fmaxf(c[0], 0.f), fmaxf(c[1], 0.f));
            *reinterpret_cast<float2*>(dst + (r + 8) * stride + col) = make_float2(fmaxf(c[2], 0.f), fmaxf(c[3], 0.f));
        };
    };
    auto masked_into = [&](float* dst) {                      // result * [dst > 0] -> dst (72-wide, in place)
        return [=](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            float2* q0 = reinterpret_cast<float2*>(dst + r * RS + col);
            float2* q1 = reinterpret_cast<float2*>(dst + (r + 8) * RS + col);
            const float2 y0 = *q0, y1 = *q1;
            *q0 = make_float2(y0.x > 0.f ? c[0] : 0.f, y0.y > 0.f ? c[1] : 0.f);
            *q1 = make_float2(y1.x > 0.f ? c[2] : 0.f, y1.y > 0.f ? c[3] : 0.f);
        };
    };
    auto plain_to = [&](float* dst, int stride, int ncols) {
        return [=](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            if (col < ncols) {
                *reinterpret_cast<float2*>(dst + r * stride + col) = make_float2(c[0], c[1]);
                *reinterpret_cast<float2*>(dst + (r + 8) * stride + col) = make_float2(c[2], c[3]);
            }
        };
    };
    auto grad_to = [&](float* dst, int ld, int M, int N) {
        return [=](int m0, int nt, const float (&c)[4]) {
            const int m = m0 + g, n = nt * 8 + 2 * t;
            if (n < N) {
                if (m < M) { atomicAdd(dst + m * ld + n, c[0]); if (n + 1 < N) atomicAdd(dst + m * ld + n + 1, c[1]); }
                if (m + 8 < M) { atomicAdd(dst + (m + 8) * ld + n, c[2]); if (n + 1 < N) atomicAdd(dst + (m + 8) * ld + n + 1, c[3]); }
            }
        };
    };

    // Two warps share a chunk: the MAIN warp (role 0) does everything that is lane <-> pedestrian, both take one m-tile of
    // every warp GEMM (the parameter-gradient GEMMs split their m-tiles the same way); a 64-thread named barrier of the
    // pair stands wherever rows written by one warp are read by the other.  The kernel is bound by dependent-issue latency
    // at 1.5 warps per scheduler and ~3/4 of a chunk's time is GEMMs, so the second warp costs no scratch and halves them.
    auto psync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory"); };
    const bool main_w = role == 0;
    const int n_pairs_total = gridDim.x * PAIRS;
    for (int chunk = blockIdx.x * PAIRS + pair; chunk < n_chunks; chunk += n_pairs_total) {
        const int p0 = scene_start[chunk_scene[chunk]];
        const int np = scene_start[chunk_scene[chunk + 1]] - p0;
        const bool live = lane < np;
        const int p = p0 + lane;
        int b = 0, e = 0, my_lead = lane, k = 1;
        bool is_lead = false, is_head = false;
        float a = 1.f, cg = 1.f;
        uint32_t group_mask = 0u, scene_mask = 0u, leader_mask = 0u;
        if (main_w) {
            {
                float4 xv[IN / 4], gv[FIN / 4];
#pragma unroll
                for (int c = 0; c < IN / 4; ++c) xv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int c = 0; c < FIN / 4; ++c) gv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live) {
                    b = ped_start[p] - p0; e = ped_end[p] - p0; my_lead = leader[p] - p0; k = gsize[p];
                    const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)p * IN);
                    const float4* gp = reinterpret_cast<const float4*>(gout + (int64_t)p * FIN);
#pragma unroll
                    for (int c = 0; c < IN / 4; ++c) xv[c] = xr[c];
#pragma unroll
                    for (int c = 0; c < FIN / 4; ++c) gv[c] = gp[c];
                }
#pragma unroll
                for (int c = 0; c < IN / 4; ++c) reinterpret_cast<float4*>(Q + lane * RS)[c] = xv[c];      // x rows, transient
#pragma unroll
                for (int c = 0; c < FIN / 4; ++c) reinterpret_cast<float4*>(G + lane * RG)[c] = gv[c];
            }
            is_lead = live && my_lead == lane;
            is_head = live && lane == b;
            a = __frcp_rn((float)k);
            group_mask = __match_any_sync(0xffffffffu, live ? my_lead : 32 + lane);
            scene_mask = (e >= 32 ? 0xffffffffu : ((1u << e) - 1u)) & ~((1u << b) - 1u);
            leader_mask = __ballot_sync(0xffffffffu, is_lead) & scene_mask;
            const int Gs = __popc(leader_mask);
            cg = __frcp_rn((float)(Gs > 0 ? Gs : 1));
            __syncwarp();
            if (lane < FIN) {                                     // d(bo) = column sums of the grad_out rows
                float sgo = 0.f;
                for (int r = 0; r < 32; ++r) sgo += G[r * RG + lane];
                atomicAdd(&gbo[lane], sgo);
            }
            // ---- m1 = group mean of x at the leader slots -> P ----
            float m1[IN];
#pragma unroll
            for (int c = 0; c < IN; ++c) m1[c] = 0.f;
            if (is_lead) {
                for (uint32_t mm = group_mask; mm; mm &= mm - 1) {
                    const float4* row = reinterpret_cast<const float4*>(Q + (__ffs(mm) - 1) * RS);
#pragma unroll
                    for (int c = 0; c < IN / 4; ++c) {
                        const float4 v = row[c];
                        m1[4 * c] = fmaf(a, v.x, m1[4 * c]); m1[4 * c + 1] = fmaf(a, v.y, m1[4 * c + 1]);
                        m1[4 * c + 2] = fmaf(a, v.z, m1[4 * c + 2]); m1[4 * c + 3] = fmaf(a, v.w, m1[4 * c + 3]);
                    }
                }
            }
            store_row<IN>(P + lane * RS, m1);
        }
        psync();
        // ---- forward: h1 = relu(m1 W0) -> Q ; x1 = relu(h1 W1) -> X1s ----
        warp_gemm_3xtf32<IN, HID / 8, RS, SWW>(P, sW0, lane, relu_to(Q, RS), role, 2);
        psync();
        warp_gemm_3xtf32<HID, OUT / 8, RS, SWN>(Q, sW1, lane, relu_to(X1s, RA), role, 2);
        psync();
        // ---- n1 = mean of the scene's group states at the scene's first slot -> N1s ----
        if (main_w) {
            float n1[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) n1[o] = 0.f;
            if (is_head) {
                for (uint32_t mm = leader_mask; mm; mm &= mm - 1) {
                    const int q = __ffs(mm) - 1;
#pragma unroll
                    for (int o = 0; o < OUT; ++o) n1[o] = fmaf(cg, X1s[q * RA + o], n1[o]);
                }
            }
            store_row<OUT>(N1s + lane * RA, n1);
        }
        psync();
        // ---- k1 = relu(n1 V0) -> Q ; y = relu(k1 V1) -> Ys ----
        warp_gemm_3xtf32<OUT, HID / 8, RA, SWW>(N1s, sV0, lane, relu_to(Q, RS), role, 2);
        psync();
        warp_gemm_3xtf32<HID, OUT / 8, RS, SWN>(Q, sV1, lane, relu_to(Ys, RA), role, 2);
        psync();
        // ---- cat = [x1[lead] | a y[head]] ; d(Wo) += grad_out^T cat ----
        if (main_w) {
            float c1[OUT], c2[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                c1[o] = live ? X1s[my_lead * RA + o] : 0.f;
                c2[o] = live ? a * Ys[b * RA + o] : 0.f;
            }
            store_row<OUT>(A + lane * RA, c1);
            store_row<OUT>(B + lane * RA, c2);
        }
        psync();
        warp_gemm_3xtf32_at_pair<FIN, OUT / 8, RG, RA>(G, A, lane, grad_to(gWo, 2 * OUT, FIN, OUT), role);
        warp_gemm_3xtf32_at_pair<FIN, OUT / 8, RG, RA>(G, B, lane, grad_to(gWo + OUT, 2 * OUT, FIN, OUT), role);
        psync();
        // ---- d(cat) = grad_out Wo: columns 0..15 -> A, 16..31 -> B ----
        warp_gemm_3xtf32_bt<FIN, 2 * OUT / 8, RG, SWO>(G, sWoT, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g;
            float* dst = (nt < OUT / 8 ? A : B) + (nt % (OUT / 8)) * 8 + 2 * t;
            *reinterpret_cast<float2*>(dst + r * RA) = make_float2(c[0], c[1]);
            *reinterpret_cast<float2*>(dst + (r + 8) * RA) = make_float2(c[2], c[3]);
        }, role, 2);
        psync();
        // ---- inter level: D2 = (sum_{i in scene} a_i dx2_i) * [y > 0] at the scene's first slot ----
        float dx1s[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) dx1s[o] = 0.f;
        if (main_w) {
            float d2[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) d2[o] = 0.f;
            {                                                     // x2_i = a_i y: scale every row of d(x2) by its own a_i
                float4* row = reinterpret_cast<float4*>(B + lane * RA);
                const float sc = live ? a : 0.f;
#pragma unroll
                for (int c = 0; c < OUT / 4; ++c) { float4 v = row[c]; row[c] = make_float4(sc * v.x, sc * v.y, sc * v.z, sc * v.w); }
            }
            __syncwarp();
            if (is_head) {
                for (uint32_t mm = scene_mask; mm; mm &= mm - 1) {
                    const int q = __ffs(mm) - 1;
#pragma unroll
                    for (int o = 0; o < OUT; ++o) d2[o] += B[q * RA + o];
                }
#pragma unroll
                for (int o = 0; o < OUT; ++o) d2[o] = Ys[lane * RA + o] > 0.f ? d2[o] : 0.f;
            }
            if (is_lead) {                                        // direct part of d(x1_g): the members' dcat[:16]
                for (uint32_t mm = group_mask; mm; mm &= mm - 1) {
                    const int q = __ffs(mm) - 1;
#pragma unroll
                    for (int o = 0; o < OUT; ++o) dx1s[o] += A[q * RA + o];
                }
            }
            __syncwarp();
            store_row<OUT>(B + lane * RA, d2);                    // D2 rows (zero off the scene heads)
        }
        psync();
        warp_gemm_3xtf32_at_pair<HID, OUT / 8, RS, RA>(Q, B, lane, grad_to(gV1, OUT, HID, OUT), role);    // d(V1) += k1^T D2
        psync();
        warp_gemm_3xtf32_bt<OUT, HID / 8, RA, SWN>(B, sV1, lane, masked_into(Q), role, 2);               // E2 -> Q
        psync();
        warp_gemm_3xtf32_at_pair<OUT, HID / 8, RA, RS>(N1s, Q, lane, grad_to(gV0, HID, OUT, HID), role);  // d(V0) += n1^T E2 (one m-tile)
        warp_gemm_3xtf32_bt<HID, OUT / 8, RS, SWW>(Q, sV0, lane, plain_to(A, RA, OUT), role, 2);         // E2 V0^T -> A (scene rows)
        psync();
        // ---- intra level: D1 = (direct + c E2 V0^T of my scene) * [x1 > 0] at the leader slots -> B ----
        if (main_w) {
            float d1[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                const float v = dx1s[o] + cg * A[b * RA + o];
                d1[o] = (is_lead && X1s[lane * RA + o] > 0.f) ? v : 0.f;
            }
            store_row<OUT>(B + lane * RA, d1);                    // (the pair's last reads of B were before the barrier)
        }
        // h1 again (Q was reused by the inter level); independent of D1, so the second warp starts while the first writes it
        warp_gemm_3xtf32<IN, HID / 8, RS, SWW>(P, sW0, lane, relu_to(Q, RS), role, 2);
        psync();
        warp_gemm_3xtf32_at_pair<HID, OUT / 8, RS, RA>(Q, B, lane, grad_to(gW1, OUT, HID, OUT), role);    // d(W1) += h1^T D1
        psync();
        warp_gemm_3xtf32_bt<OUT, HID / 8, RA, SWN>(B, sW1, lane, masked_into(Q), role, 2);               // E -> Q
        psync();
        warp_gemm_3xtf32_at_pair<IN, HID / 8, RS, RS>(P, Q, lane, grad_to(gW0, HID, IN, HID), role);      // d(W0) += m1^T E
        psync();
        warp_gemm_3xtf32_bt<HID, IN / 8, RS, SWW>(Q, sW0, lane, plain_to(P, RS, IN), role, 2);           // E W0^T -> P (leader rows)
        psync();
        // ---- d(x_i) = a_i (E W0^T)[leader(i)] ----
        if (main_w && live) {
            const float4* src = reinterpret_cast<const float4*>(P + my_lead * RS);
            float4* dst = reinterpret_cast<float4*>(grad_x + (int64_t)p * IN);
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) {
                const float4 v = src[c];
                dst[c] = make_float4(a * v.x, a * v.y, a * v.z, a * v.w);
            }
        }
        psync();                                              // P / Q / G are rewritten by the next chunk's loads
    }
#if !GCB_GLOBAL_GRADS
    __syncthreads();
    float* mine = partials + (int64_t)blockIdx.x * C::GRAD_FLOATS;
    for (int e = threadIdx.x; e < C::GRAD_FLOATS; e += blockDim.x) mine[e] = gW0[e];
#endif
}

template <int IN, int FIN>
__global__ void gcn_bwd_reduce_kernel(const float* __restrict__ partials, int n_blocks, float* gW0, float* gW1, float* gV0,
                                      float* gV1, float* gWo, float* gbo) {
    using C = Cfg<IN, FIN>;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= C::GRAD_FLOATS) return;
    float s = 0.f;
    for (int k = 0; k < n_blocks; ++k) s += partials[(int64_t)k * C::GRAD_FLOATS + e];
    constexpr int o1 = IN * HID, o2 = o1 + HID * OUT, o3 = o2 + OUT * HID, o4 = o3 + HID * OUT, o5 = o4 + FIN * 2 * OUT;
    if (e < o1) gW0[e] = s;
    else if (e < o2) gW1[e - o1] = s;
    else if (e < o3) gV0[e - o2] = s;
    else if (e < o4) gV1[e - o3] = s;
    else if (e < o5) gWo[e - o4] = s;
    else gbo[e - o5] = s;
}

template <int IN, int FIN>
static int launch(const float* x, const float* gout, const int32_t* leader, const int32_t* gsize, const int32_t* ps,
                  const int32_t* pe, const int32_t* scene_start, const int32_t* chunk_scene, int n_chunks, const float* W0,
                  const float* W1, const float* V0, const float* V1, const float* Wo, float* gx, float* gW0, float* gW1,
                  float* gV0, float* gV1, float* gWo, float* gbo, float* partials, cudaStream_t st) {
    using C = Cfg<IN, FIN>;
    auto kern = gcn_fused_bwd_kernel<IN, FIN>;
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    const int grid = std::min((n_chunks + PAIRS - 1) / PAIRS, 148);
    kern<<<grid, WARPS * 32, C::SMEM, st>>>(x, gout, leader, gsize, ps, pe, scene_start, chunk_scene, n_chunks, W0, W1, V0, V1,
                                            Wo, gx, partials);
    SGX_LAUNCH_CHECK();
    gcn_bwd_reduce_kernel<IN, FIN><<<blocks_for(C::GRAD_FLOATS, 256), 256, 0, st>>>(partials, grid, gW0, gW1, gV0, gV1, gWo, gbo);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

}  // namespace gcnb
}  // namespace sgx

using namespace sgx;

extern "C" int64_t sgx_gcn_module_fused_bwd_ws_bytes(void) { return (int64_t)148 * gcnb::Cfg<40, 32>::GRAD_FLOATS * 4 + 256; }

extern "C" int sgx_gcn_module_fused_bwd(const float* x, const float* grad_out, const int32_t* leader,
                                        const int32_t* group_size, const int32_t* ped_start, const int32_t* ped_end,
                                        const int32_t* scene_start, const int32_t* chunk_scene, int64_t n_chunks,
                                        const float* W0, const float* W1, const float* V0, const float* V1, const float* Wo,
                                        const float* bo, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, float* grad_x,
                                        float* grad_W0, float* grad_W1, float* grad_V0, float* grad_V1, float* grad_Wo,
                                        float* grad_bo, void* workspace, int64_t ws_bytes, void* stream) {
    (void)bo;
    SGX_REQUIRE(x && grad_out && leader && group_size && ped_start && ped_end && scene_start && chunk_scene && W0 && W1 && V0 &&
                    V1 && Wo && grad_x && grad_W0 && grad_W1 && grad_V0 && grad_V1 && grad_Wo && grad_bo && workspace,
                "sgx_gcn_module_fused_bwd: null pointer");
    SGX_REQUIRE(n_chunks > 0 && n_chunks < ((int64_t)1 << 31), "sgx_gcn_module_fused_bwd: bad chunk count");
    SGX_UNSUPPORTED(HID != 72 || OUT != 16 || !(IN == 32 || IN == 40) || !(FIN == 24 || FIN == 32),
                    "the single-launch GCNModule backward is built for input 32|40, hidden 72, out 16, final 24|32 "
                    "(got %d/%d/%d/%d)", IN, HID, OUT, FIN);
    SGX_REQUIRE(ws_bytes >= sgx_gcn_module_fused_bwd_ws_bytes(), "sgx_gcn_module_fused_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* partials = (float*)workspace;
#define SGX_GCNB(I, F)                                                                                                   \
    if (IN == I && FIN == F)                                                                                             \
        return gcnb::launch<I, F>(x, grad_out, leader, group_size, ped_start, ped_end, scene_start, chunk_scene,        \
                                  (int)n_chunks, W0, W1, V0, V1, Wo, grad_x, grad_W0, grad_W1, grad_V0, grad_V1, grad_Wo, \
                                  grad_bo, partials, st);
    SGX_GCNB(40, 24) SGX_GCNB(40, 32) SGX_GCNB(32, 24) SGX_GCNB(32, 32)
#undef SGX_GCNB
    return SGX_ERR_UNSUPPORTED;
}
