// GATEncoder backward in ONE launch (+ a tiny reduction), for batches whose scenes fit a warp chunk (<= 32 peds),
// n_heads = 1, dims 40 / 72 / 16 / 24 -- what autograd does through sgan/models.py:254-294 (GraphAttentionLayer
// 198-220, GAT 231-237), scene by scene with dense [N,N,2F] tensors.
//
// A warp PAIR per chunk of whole scenes, lanes <-> pedestrians (the main warp), exactly like the mma.sync forward
// (sgx_gat.cu): nothing but x, grad_out, the group structure and grad_x touches HBM.  Per chunk
//   1. the forward is recomputed up to Yg (the saved state between forward and backward is nothing),
//   2. the four attention layers are walked in reverse.  Each layer's backward is a ROW role (softmax statistics,
//      c_i = dhp_i.hp_i, ds_i) and a COLUMN role (dt_j and dWh_j gathered over the symmetric neighbourhood) on lane masks of
//      the actual neighbours; the layer's forward values are recomputed right before they are needed so that only two
//      72-wide row buffers are alive.  The inter level's first layer runs in its aggregated form (16-wide rows, see the
//      kernel body),
//   3. every linear map of the chain rule is a warp-level 3xTF32 tensor-core GEMM, one m-tile per warp of the pair:
//      dX = dY W^T straight from the forward's weight block (warp_gemm_3xtf32_bt) and the per-chunk parameter gradient
//      dW = input^T dY (warp_gemm_3xtf32_at_pair), added with red.global into the CTA's own gradient block in HBM.
// gat_bwd_reduce_kernel sums the (<= 148) blocks in block order; within a block the chunks' contributions arrive as
// floating-point atomics, so parameter gradients are reproducible to rounding, not bit for bit.
// The general multi-pass path (sgx_gat.cu: ~35 launches, intermediates in HBM) remains for larger scenes / more heads.
#include "sgx_common.cuh"
#include "sgx_gat_fused.cuh"
#include "sgx_warp_mma.cuh"

namespace sgx {

#ifndef GB_GLOBAL_GRADS
#define GB_GLOBAL_GRADS 1            // 1: the CTA's gradient block lives in its HBM partial (red.global), 6 chunks in flight
#endif                               // per SM; 0: in shared memory (30 KB), 5 chunks
constexpr int GB_PAIRS = GB_GLOBAL_GRADS ? 6 : 5;        // warp PAIRS: two warps share a chunk's scratch (see the kernel)
constexpr int GB_WARPS = 2 * GB_PAIRS;
constexpr int GB_IN = 40, GB_FIN = 24;
constexpr int RG = 28;                       // row stride of the grad_out rows (24 wide): conflict-free A fragments

struct GatGrad {                             // per-CTA gradient block == layout of the per-CTA partials in HBM
    float Wi[GB_IN * HID], ai[2 * HID], Wio[HID * OUT], aio[2 * OUT];
    float We[OUT * HID], ae[2 * HID], Weo[HID * OUT], aeo[2 * OUT];
    float Wo[GB_FIN * 2 * OUT], bo[GB_FIN];
    float du[2 * OUT];                       // d(We ae1), d(We ae2): the inter level's score vectors (see the reduce kernel)
};
constexpr int GB_GRAD_FLOATS = sizeof(GatGrad) / sizeof(float);
struct GatAvec { float ai[2 * HID], aio[2 * OUT], ae[2 * HID], aeo[2 * OUT], ue[2 * OUT]; };   // ue = (We ae1, We ae2)

// per-warp scratch (floats): two 72-wide row buffers, four 16-wide, scores / statistics.  The grad_out rows live in the
// tail of the first wide buffer while it holds nothing wide (between the first layer's attention and the reload of x
// for the intra level's backward): 31 KB per warp instead of 34.6 KB = one more warp per SM, and the kernel is bound by
// exactly that (1.28 warps per scheduler, profiles/r02_gat_tc_source_hotspots.md).
constexpr int GB_SCRATCH = 2 * 32 * RS + 4 * 32 * RA + 2 * 32 * 2 + 32 * 4 + 32 * 2 + 8;   // + pad: the d(a) GEMM reads its
                                             // 2-wide right operand as 8 columns (the extra ones are dropped)

__device__ __forceinline__ float lrelu_grad(float pre, float alpha) { return pre > 0.f ? 1.f : alpha; }

// ROW role of one attention layer for node i: softmax statistics over its neighbourhood and
//   c_i = sum_j alpha_ij (dhp_i . Wh_j)  (= dhp_i . hp_i),   ds_i = sum_j alpha_ij (dhp_i . Wh_j - c_i) lrelu'(s_i + t_j)
template <int F, int STRIDE>
__device__ __forceinline__ void att_bwd_row(const float* __restrict__ Wh, const float2* __restrict__ st, uint32_t mask,
                                            float s_i, float alpha, const float (&dh)[F], float4& stat, float& ds) {
    float m = -INFINITY;
    for (uint32_t mm = mask; mm; mm &= mm - 1) m = fmaxf(m, lrelu(s_i + st[__ffs(mm) - 1].y, alpha));
    float den = 0.f, c = 0.f, s1 = 0.f, s2 = 0.f;
    for (uint32_t mm = mask; mm; mm &= mm - 1) {
        const int q = __ffs(mm) - 1;
        const float pre = s_i + st[q].y;
        const float w = fexp(lrelu(pre, alpha) - m);
        const float4* row = reinterpret_cast<const float4*>(Wh + q * STRIDE);
        float dot = 0.f;
#pragma unroll
        for (int f = 0; f < F / 4; ++f) {
            const float4 v = row[f];
            dot = fmaf(dh[4 * f], v.x, dot); dot = fmaf(dh[4 * f + 1], v.y, dot);
            dot = fmaf(dh[4 * f + 2], v.z, dot); dot = fmaf(dh[4 * f + 3], v.w, dot);
        }
        const float lg = lrelu_grad(pre, alpha);
        den += w;
        c = fmaf(w, dot, c);
        s1 = fmaf(w * lg, dot, s1);
        s2 = fmaf(w, lg, s2);
    }
    const float inv = 1.f / den;
    c *= inv;
    ds = (s1 - c * s2) * inv;
    stat = make_float4(m, inv, c, 0.f);
}

// COLUMN role for node j: dt_j = sum_i d(pre_ij), dWh_j = sum_i alpha_ij dhp_i + ds_j a1 + dt_j a2 over the rows i that
// attend to j (the neighbourhood is symmetric).  dhp rows and the row statistics of every i are in shared memory.
template <int F, int SW, int SD>
__device__ __forceinline__ void att_bwd_col(const float* __restrict__ wh_own, const float* __restrict__ dhp,
                                            const float2* __restrict__ st, const float4* __restrict__ stat, uint32_t mask,
                                            float t_j, float ds_j, float alpha, const float* __restrict__ avec,
                                            float (&dwh)[F], float& dt_out) {
    float whj[F];
#pragma unroll
    for (int f = 0; f < F / 4; ++f) {
        const float4 v = reinterpret_cast<const float4*>(wh_own)[f];
        whj[4 * f] = v.x; whj[4 * f + 1] = v.y; whj[4 * f + 2] = v.z; whj[4 * f + 3] = v.w;
    }
#pragma unroll
    for (int f = 0; f < F; ++f) dwh[f] = 0.f;
    float dt = 0.f;
    for (uint32_t mm = mask; mm; mm &= mm - 1) {
        const int i = __ffs(mm) - 1;
        const float pre = st[i].x + t_j;
        const float4 si = stat[i];
        const float a_ij = fexp(lrelu(pre, alpha) - si.x) * si.y;
        const float4* row = reinterpret_cast<const float4*>(dhp + i * SD);
        float dot = 0.f;
#pragma unroll
        for (int f = 0; f < F / 4; ++f) {
            const float4 v = row[f];
            dot = fmaf(v.x, whj[4 * f], dot); dot = fmaf(v.y, whj[4 * f + 1], dot);
            dot = fmaf(v.z, whj[4 * f + 2], dot); dot = fmaf(v.w, whj[4 * f + 3], dot);
            dwh[4 * f] = fmaf(a_ij, v.x, dwh[4 * f]); dwh[4 * f + 1] = fmaf(a_ij, v.y, dwh[4 * f + 1]);
            dwh[4 * f + 2] = fmaf(a_ij, v.z, dwh[4 * f + 2]); dwh[4 * f + 3] = fmaf(a_ij, v.w, dwh[4 * f + 3]);
        }
        dt = fmaf(a_ij * lrelu_grad(pre, alpha), dot - si.z, dt);
    }
#pragma unroll
    for (int f = 0; f < F; ++f) dwh[f] = fmaf(ds_j, avec[f], fmaf(dt, avec[F + f], dwh[f]));
    dt_out = dt;
    (void)SW;
}

template <int F>
__device__ __forceinline__ void load_row(const float* __restrict__ row, float (&v)[F]) {
#pragma unroll
    for (int f = 0; f < F / 4; ++f) {
        const float4 u = reinterpret_cast<const float4*>(row)[f];
        v[4 * f] = u.x; v[4 * f + 1] = u.y; v[4 * f + 2] = u.z; v[4 * f + 3] = u.w;
    }
}

// d(hp) from d(output) for the ELU + log_softmax epilogue (GAT.forward, models.py:236-237): y = u - lse(u), u = elu(hp)
template <int F>
__device__ __forceinline__ void elu_logsoftmax_bwd(const float (&hp)[F], const float (&y)[F], const float (&dy)[F],
                                                   float (&dhp)[F]) {
    float gsum = 0.f;
#pragma unroll
    for (int f = 0; f < F; ++f) gsum += dy[f];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const float du = dy[f] - fexp(y[f]) * gsum;                 // softmax(u) = exp(y)
        dhp[f] = du * (hp[f] > 0.f ? 1.f : fexp(hp[f]));
    }
}

__global__ void __launch_bounds__(GB_WARPS * 32)
gat_fused_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gout, const int32_t* __restrict__ leader,
                     const int32_t* __restrict__ gsize, const int32_t* __restrict__ ped_start,
                     const int32_t* __restrict__ ped_end, const int32_t* __restrict__ scene_start,
                     const int32_t* __restrict__ chunk_scene, int n_chunks, const float* __restrict__ Wi,
                     const float* __restrict__ ai, const float* __restrict__ Wio, const float* __restrict__ aio,
                     const float* __restrict__ We, const float* __restrict__ ae, const float* __restrict__ Weo,
                     const float* __restrict__ aeo, const float* __restrict__ Wo, const float* __restrict__ bo, float alpha,
                     float* __restrict__ grad_x, float* __restrict__ partials) {
    constexpr int IN = GB_IN, FIN = GB_FIN;
    extern __shared__ __align__(16) uint8_t raw[];
    FusedWm& w = *reinterpret_cast<FusedWm*>(raw);
    GatAvec& av = *reinterpret_cast<GatAvec*>(raw + sizeof(FusedWm));
#if GB_GLOBAL_GRADS
    // the per-chunk parameter gradients are added straight into this CTA's partial block in HBM (fire-and-forget
    // red.global.add, ~7.5 k per chunk): the 30 KB it occupied in shared memory is a fifth warp's scratch
    GatGrad& gr = *reinterpret_cast<GatGrad*>(partials + (int64_t)blockIdx.x * GB_GRAD_FLOATS);
    float* bufs = reinterpret_cast<float*>(raw + sizeof(FusedWm) + sizeof(GatAvec));
#else
    GatGrad& gr = *reinterpret_cast<GatGrad*>(raw + sizeof(FusedWm) + sizeof(GatAvec));
    float* bufs = reinterpret_cast<float*>(raw + sizeof(FusedWm) + sizeof(GatAvec) + sizeof(GatGrad));
#endif
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    fused_load_weights<IN, FIN>(w, Wi, ai, Wio, aio, We, ae, Weo, aeo, Wo, bo);
    for (int e = threadIdx.x; e < 2 * HID; e += blockDim.x) { av.ai[e] = ai[e]; av.ae[e] = ae[e]; }
    for (int e = threadIdx.x; e < 2 * OUT; e += blockDim.x) { av.aio[e] = aio[e]; av.aeo[e] = aeo[e]; }
    for (int e = threadIdx.x; e < GB_GRAD_FLOATS; e += blockDim.x) reinterpret_cast<float*>(&gr)[e] = 0.f;
#if GB_GLOBAL_GRADS
    __threadfence();
#endif
    __syncthreads();
    if (threadIdx.x < 2 * OUT) av.ue[threadIdx.x] = w.We[(threadIdx.x % OUT) * SW1 + HID + threadIdx.x / OUT];
    __syncthreads();

    const int pair = warp >> 1, role = warp & 1;
    float* P = bufs + pair * GB_SCRATCH;                 // [32][RS]  x / Wh1; [32][RA] xbar of the inter level
    float* Q = P + 32 * RS;                              // [32][RS]  x1a / y3 -> d(hp) -> dWh of the 72-wide layers
    float* Xg = Q + 32 * RS;                             // [32][RA]  pooled group state (leader slots)
    float* A = Xg + 32 * RA;                             // [32][RA]  Wh2 / Wh4
    float* B = A + 32 * RA;                              // [32][RA]  x1 / dx2 / d(hp) -> dWh of the 16-wide layers
    float* D = B + 32 * RA;                              // [32][RA]  Yg / x2 / dXg
    float* G = P + 32 * RA;                              // [32][RG]  grad_out rows, in the tail of P (behind xbar)
    float2* stA = reinterpret_cast<float2*>(D + 32 * RA);   // (s, t) of the 72-wide layer being processed
    float2* stB = stA + 32;                              // (s, t) of the 16-wide layer
    float4* stat = reinterpret_cast<float4*>(stB + 32);  // (m, 1/den, c) per row of the layer in its backward
    float2* dstb = reinterpret_cast<float2*>(stat + 32); // (ds, dt) per row: right operand of the d(a) GEMM
    const int g = lane >> 2, t = lane & 3;

    // GEMM result routing: a 72(+2)-wide result to a wide buffer + scores, a 16(+2)-wide one to a narrow buffer + scores
    auto wide_to = [&](float* dst, float2* st) {
        return [=](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g;
            if (nt < HID / 8) {
                *reinterpret_cast<float2*>(dst + r * RS + nt * 8 + 2 * t) = make_float2(c[0], c[1]);
                *reinterpret_cast<float2*>(dst + (r + 8) * RS + nt * 8 + 2 * t) = make_float2(c[2], c[3]);
            } else if (t == 0) {
                st[r] = make_float2(c[0], c[1]);
                st[r + 8] = make_float2(c[2], c[3]);
            }
        };
    };
    auto narrow_to = [&](float* dst, float2* st) {
        return [=](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g;
            if (nt < OUT / 8) {
                *reinterpret_cast<float2*>(dst + r * RA + nt * 8 + 2 * t) = make_float2(c[0], c[1]);
                *reinterpret_cast<float2*>(dst + (r + 8) * RA + nt * 8 + 2 * t) = make_float2(c[2], c[3]);
            } else if (st != nullptr && t == 0) {
                st[r] = make_float2(c[0], c[1]);
                st[r + 8] = make_float2(c[2], c[3]);
            }
        };
    };
    // per-chunk parameter gradient fragments -> the CTA's gradient block: dst[(m) * ld + n] for m < M, n < N
    auto grad_to = [&](float* dst, int ld, int M, int N) {
        return [=](int m0, int nt, const float (&c)[4]) {
            const int m = m0 + g, n = nt * 8 + 2 * t;
#ifdef GB_EXPERIMENT_NO_ATOMICS
            if (c[0] == 123.456f)
#endif
            if (n < N) {
                if (m < M) { atomicAdd(dst + m * ld + n, c[0]); if (n + 1 < N) atomicAdd(dst + m * ld + n + 1, c[1]); }
                if (m + 8 < M) { atomicAdd(dst + (m + 8) * ld + n, c[2]); if (n + 1 < N) atomicAdd(dst + (m + 8) * ld + n + 1, c[3]); }
            }
        };
    };
    // d(a) [2][F] += (ds, dt)^T Wh: result fragment rows = feature f, columns 0 / 1 = a1 / a2
    auto avec_to = [&](float* dst, int F) {
        return [=](int m0, int nt, const float (&c)[4]) {
            (void)nt;
            if (t == 0) {
                const int f = m0 + g;
                if (f < F) { atomicAdd(dst + f, c[0]); atomicAdd(dst + F + f, c[1]); }
                if (f + 8 < F) { atomicAdd(dst + f + 8, c[2]); atomicAdd(dst + F + f + 8, c[3]); }
            }
        };
    };

    // Two warps share a chunk (as in the GCN backward): the MAIN warp (role 0) owns everything that is lane <-> pedestrian,
    // both take one m-tile of every warp GEMM, the parameter-gradient GEMMs are split along their output columns.  The
    // second warp runs the same code as a warp of dead lanes -- every per-lane loop is empty for it and every per-lane
    // store is the main warp's alone -- so the two execute the same sequence of 64-thread pair barriers, which stand
    // wherever a __syncwarp stood and after every GEMM that writes rows.
    auto psync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory"); };
    const bool main_w = role == 0;
    const int n_pairs_total = gridDim.x * GB_PAIRS;
    for (int chunk = blockIdx.x * GB_PAIRS + pair; chunk < n_chunks; chunk += n_pairs_total) {
        const int p0 = scene_start[chunk_scene[chunk]];
        const int np = scene_start[chunk_scene[chunk + 1]] - p0;
        const bool live = main_w && lane < np;
        const int p = p0 + lane;
        int b = 0, e = 0, my_lead = lane;
        float inv_g = 1.f;
        if (live) {
            b = ped_start[p] - p0; e = ped_end[p] - p0; my_lead = leader[p] - p0;
            inv_g = __frcp_rn((float)gsize[p]);
        }
        auto load_x = [&]() {                               // x rows of the chunk -> P (zero rows for dead lanes)
            float4 xv[IN / 4];
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) xv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)p * IN);
#pragma unroll
                for (int c = 0; c < IN / 4; ++c) xv[c] = xr[c];
            }
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) if (main_w) reinterpret_cast<float4*>(P + lane * RS)[c] = xv[c];
        };
        load_x();
        const bool is_lead = live && (my_lead == lane);
        const uint32_t group_mask = __match_any_sync(0xffffffffu, live ? my_lead : 32 + lane);
        const uint32_t scene_mask = (e >= 32 ? 0xffffffffu : ((1u << e) - 1u)) & ~((1u << b) - 1u);
        const uint32_t leader_mask = __ballot_sync(0xffffffffu, is_lead) & scene_mask;
        psync();

        // =============== forward recompute, part 1: x -> x1 -> Xg ===============
        warp_gemm_3xtf32<IN, HID / 8 + 1, RS, SW1>(P, w.Wi, lane, wide_to(P, stA), role, 2);
        psync();
        {
            float hp[HID];
#pragma unroll
            for (int f = 0; f < HID; ++f) hp[f] = 0.f;
            if (live) {
                attend_mask<HID, RS>(P, stA, group_mask, stA[lane].x, alpha, hp);
#pragma unroll
                for (int f = 0; f < HID; ++f) hp[f] = felu(hp[f]);
            }
            if (main_w) store_row<HID>(Q + lane * RS, hp);              // x1a
        }
        psync();
        // the Wh1 rows in P are dead until the intra level's backward reloads x: the grad_out rows move into P's tail
        {
            float4 gv[FIN / 4];
#pragma unroll
            for (int c = 0; c < FIN / 4; ++c) gv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const float4* gp = reinterpret_cast<const float4*>(gout + (int64_t)p * FIN);
#pragma unroll
                for (int c = 0; c < FIN / 4; ++c) gv[c] = gp[c];
            }
#pragma unroll
            for (int c = 0; c < FIN / 4; ++c) if (main_w) reinterpret_cast<float4*>(G + lane * RG)[c] = gv[c];
        }
        psync();
        if (main_w && lane < FIN) {                         // d(bo) = column sums of the grad_out rows
            float sgo = 0.f;
            for (int r = 0; r < 32; ++r) sgo += G[r * RG + lane];
            atomicAdd(&gr.bo[lane], sgo);
        }
        warp_gemm_3xtf32<HID, OUT / 8 + 1, RS, SW2>(Q, w.Wio, lane, narrow_to(A, stB), role, 2);
        psync();
        float x1[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) x1[o] = 0.f;
        if (live) {
            attend_mask<OUT, RA>(A, stB, group_mask, stB[lane].x, alpha, x1);
            elu_logsoftmax<OUT>(x1);
        }
        if (main_w) store_row<OUT>(B + lane * RA, x1);
        psync();
        // d(Wo)[:, :16] += grad_out^T x1
        warp_gemm_3xtf32_at_pair<FIN, OUT / 8, RG, RA>(G, B, lane, grad_to(gr.Wo, 2 * OUT, FIN, OUT), role);
        {
            float xg[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) xg[o] = 0.f;
            if (is_lead) {
                for (uint32_t mm = group_mask; mm; mm &= mm - 1) {
                    const int q = __ffs(mm) - 1;
#pragma unroll
                    for (int o = 0; o < OUT; ++o) xg[o] = fmaf(inv_g, B[q * RA + o], xg[o]);
                }
            }
            if (main_w) store_row<OUT>(Xg + lane * RA, xg);
        }
        psync();
        // =============== forward recompute, part 2 (inter level, leaders): Xg -> y3 -> Yg ===============
        // layer 3 aggregated BEFORE its linear map: sum_j a_ij (Xg_j We) = (sum_j a_ij Xg_j) We = xbar We, so the attention
        // over the scene's leaders and its whole backward run on 16-wide rows instead of 72-wide ones; the scores are
        // Xg . (We ae1), Xg . (We ae2) with the two vectors the weight block already holds as its score columns
        {
            float xg[OUT];
            load_row<OUT>(Xg + lane * RA, xg);
            float s3 = 0.f, t3 = 0.f;
#pragma unroll
            for (int o = 0; o < OUT; ++o) { s3 = fmaf(xg[o], av.ue[o], s3); t3 = fmaf(xg[o], av.ue[OUT + o], t3); }
            if (main_w) stA[lane] = make_float2(s3, t3);
        }
        psync();
        {
            float xb[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) xb[o] = 0.f;
            if (is_lead) attend_mask<OUT, RA>(Xg, stA, leader_mask, stA[lane].x, alpha, xb);
            if (main_w) store_row<OUT>(P + lane * RA, xb);              // xbar (zero rows off the leaders)
        }
        psync();
        // y3 = elu(xbar We) -> Q (zero rows off the leaders)
        warp_gemm_3xtf32<OUT, HID / 8, RA, SW1>(P, w.We, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            *reinterpret_cast<float2*>(Q + r * RS + col) = make_float2(felu(c[0]), felu(c[1]));
            *reinterpret_cast<float2*>(Q + (r + 8) * RS + col) = make_float2(felu(c[2]), felu(c[3]));
        }, role, 2);
        psync();
        warp_gemm_3xtf32<HID, OUT / 8 + 1, RS, SW2>(Q, w.Weo, lane, narrow_to(A, stB), role, 2);    // Wh4 -> A, (s4, t4) -> stB
        psync();
        float hp4[OUT], yg[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) { hp4[o] = 0.f; yg[o] = 0.f; }
        if (is_lead) {
            attend_mask<OUT, RA>(A, stB, leader_mask, stB[lane].x, alpha, hp4);
#pragma unroll
            for (int o = 0; o < OUT; ++o) yg[o] = hp4[o];
            elu_logsoftmax<OUT>(yg);
        }
        if (main_w) store_row<OUT>(D + lane * RA, yg);                  // Yg at the leader slots
        psync();
        // =============== top: out = [x1 | x2] Wo^T + bo ===============
        {
            float x2[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) x2[o] = live ? inv_g * D[my_lead * RA + o] : 0.f;
            if (main_w) store_row<OUT>(B + lane * RA, x2);              // x1 rows are no longer needed in shared memory
        }
        psync();
        // d(Wo)[:, 16:] += grad_out^T x2
        warp_gemm_3xtf32_at_pair<FIN, OUT / 8, RG, RA>(G, B, lane, grad_to(gr.Wo + OUT, 2 * OUT, FIN, OUT), role);
        psync();
        // d(cat) = grad_out Wo: columns 0..15 = d(x1) (direct part) -> D, columns 16..31 = d(x2) -> B
        warp_gemm_3xtf32_bt<FIN, 2 * OUT / 8, RG, SW2>(G, w.WoT, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g;
            float* dst = (nt < OUT / 8 ? D : B) + (nt % (OUT / 8)) * 8 + 2 * t;
            *reinterpret_cast<float2*>(dst + r * RA) = make_float2(c[0], c[1]);
            *reinterpret_cast<float2*>(dst + (r + 8) * RA) = make_float2(c[2], c[3]);
        }, role, 2);
        psync();
        float dx1[OUT];                                     // d(x1): direct part now, + pooled part after the inter level
        load_row<OUT>(D + lane * RA, dx1);
        // =============== inter out_att (layer 4) backward ===============
        float dh4[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) dh4[o] = 0.f;
        float ds = 0.f, dt = 0.f;
        if (is_lead) {
            float dyg[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) dyg[o] = 0.f;
            for (uint32_t mm = group_mask; mm; mm &= mm - 1) {          // unpool backward: x2_i = Yg[lead] / |g|
                const int q = __ffs(mm) - 1;
#pragma unroll
                for (int o = 0; o < OUT; ++o) dyg[o] = fmaf(inv_g, B[q * RA + o], dyg[o]);
            }
            elu_logsoftmax_bwd<OUT>(hp4, yg, dyg, dh4);
            float4 s4;
            att_bwd_row<OUT, RA>(A, stB, leader_mask, stB[lane].x, alpha, dh4, s4, ds);
            if (main_w) stat[lane] = s4;
        }
        psync();                                       // every leader has read the d(x2) rows of its group
        if (main_w) store_row<OUT>(B + lane * RA, dh4);
        psync();
        {
            float dwh[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) dwh[o] = 0.f;
            if (is_lead)
                att_bwd_col<OUT, RA, RA>(A + lane * RA, B, stB, stat, leader_mask, stB[lane].y, ds, alpha, av.aeo, dwh, dt);
            if (main_w) dstb[lane] = make_float2(is_lead ? ds : 0.f, is_lead ? dt : 0.f);
            psync();
            if (main_w) store_row<OUT>(B + lane * RA, dwh);             // dWh4
        }
        psync();
        warp_gemm_3xtf32_at_pair<HID, OUT / 8, RS, RA>(Q, B, lane, grad_to(gr.Weo, OUT, HID, OUT), role);      // d(Weo) += y3^T dWh4
        warp_gemm_3xtf32_at_pair<OUT, 1, RA, 2>(A, reinterpret_cast<const float*>(dstb), lane, avec_to(gr.aeo, OUT), role);
        psync();
        // d(y3) = dWh4 Weo^T, times elu'(hp3) read back from y3: d(hp3) -> Q in place
        warp_gemm_3xtf32_bt<OUT, HID / 8, RA, SW2>(B, w.Weo, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            float2* q0 = reinterpret_cast<float2*>(Q + r * RS + col);
            float2* q1 = reinterpret_cast<float2*>(Q + (r + 8) * RS + col);
            const float2 y0 = *q0, y1 = *q1;                // elu'(hp) = 1 for hp > 0 (y = hp > 0), else exp(hp) = y + 1
            *q0 = make_float2(c[0] * (y0.x > 0.f ? 1.f : y0.x + 1.f), c[1] * (y0.y > 0.f ? 1.f : y0.y + 1.f));
            *q1 = make_float2(c[2] * (y1.x > 0.f ? 1.f : y1.x + 1.f), c[3] * (y1.y > 0.f ? 1.f : y1.y + 1.f));
        }, role, 2);
        psync();
        // =============== inter layer 1 (layer 3) backward, on the aggregated form ===============
        // Q = d(hp3):  d(We) += xbar^T d(hp3),  d(xbar) = d(hp3) We^T -> D
        warp_gemm_3xtf32_at_pair<OUT, HID / 8, RA, RS>(P, Q, lane, grad_to(gr.We, HID, OUT, HID), role);
        warp_gemm_3xtf32_bt<HID, OUT / 8, RS, SW1>(Q, w.We, lane, narrow_to(D, nullptr), role, 2);
        psync();
        ds = 0.f; dt = 0.f;
        {
            float dxb[OUT];
            load_row<OUT>(D + lane * RA, dxb);
            if (is_lead) {
                float4 s3;
                att_bwd_row<OUT, RA>(Xg, stA, leader_mask, stA[lane].x, alpha, dxb, s3, ds);
                if (main_w) stat[lane] = s3;
            }
        }
        psync();
        {
            float dxg[OUT];                                 // d(Xg_j) = sum_i a_ij d(xbar_i) + ds_j We ae1 + dt_j We ae2
#pragma unroll
            for (int o = 0; o < OUT; ++o) dxg[o] = 0.f;
            if (is_lead)
                att_bwd_col<OUT, RA, RA>(Xg + lane * RA, D, stA, stat, leader_mask, stA[lane].y, ds, alpha, av.ue, dxg, dt);
            if (main_w) dstb[lane] = make_float2(is_lead ? ds : 0.f, is_lead ? dt : 0.f);
            psync();                                   // every leader has read the d(xbar) rows
            if (main_w) store_row<OUT>(D + lane * RA, dxg);             // d(Xg)
        }
        psync();
        // d(We ae1), d(We ae2) [2][16] += (ds, dt)^T Xg: folded into d(We) and d(ae) by the reduce kernel (both are linear)
        warp_gemm_3xtf32_at_pair<OUT, 1, RA, 2>(Xg, reinterpret_cast<const float*>(dstb), lane, avec_to(gr.du, OUT), role);
        // pool backward: Xg[l] = sum_{i in g} x1_i / |g|
        if (live) {
#pragma unroll
            for (int o = 0; o < OUT; ++o) dx1[o] = fmaf(inv_g, D[my_lead * RA + o], dx1[o]);
        }
        psync();
        // =============== intra level: recompute Wh1, x1a, Wh2, hp2 ===============
        load_x();
        psync();
        warp_gemm_3xtf32<IN, HID / 8 + 1, RS, SW1>(P, w.Wi, lane, wide_to(P, stA), role, 2);        // Wh1 -> P, (s1, t1) -> stA
        psync();
        {
            float hp[HID];
#pragma unroll
            for (int f = 0; f < HID; ++f) hp[f] = 0.f;
            if (live) {
                attend_mask<HID, RS>(P, stA, group_mask, stA[lane].x, alpha, hp);
#pragma unroll
                for (int f = 0; f < HID; ++f) hp[f] = felu(hp[f]);
            }
            if (main_w) store_row<HID>(Q + lane * RS, hp);              // x1a
        }
        psync();
        warp_gemm_3xtf32<HID, OUT / 8 + 1, RS, SW2>(Q, w.Wio, lane, narrow_to(A, stB), role, 2);    // Wh2 -> A, (s2, t2) -> stB
        psync();
        // =============== intra out_att (layer 2) backward ===============
        float dh2[OUT];
#pragma unroll
        for (int o = 0; o < OUT; ++o) dh2[o] = 0.f;
        ds = 0.f; dt = 0.f;
        if (live) {
            float hp2[OUT];
            attend_mask<OUT, RA>(A, stB, group_mask, stB[lane].x, alpha, hp2);
            elu_logsoftmax_bwd<OUT>(hp2, x1, dx1, dh2);
            float4 s2;
            att_bwd_row<OUT, RA>(A, stB, group_mask, stB[lane].x, alpha, dh2, s2, ds);
            if (main_w) stat[lane] = s2;
        }
        if (main_w) store_row<OUT>(B + lane * RA, dh2);
        psync();
        {
            float dwh[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) dwh[o] = 0.f;
            if (live)
                att_bwd_col<OUT, RA, RA>(A + lane * RA, B, stB, stat, group_mask, stB[lane].y, ds, alpha, av.aio, dwh, dt);
            if (main_w) dstb[lane] = make_float2(live ? ds : 0.f, live ? dt : 0.f);
            psync();
            if (main_w) store_row<OUT>(B + lane * RA, dwh);             // dWh2
        }
        psync();
        warp_gemm_3xtf32_at_pair<HID, OUT / 8, RS, RA>(Q, B, lane, grad_to(gr.Wio, OUT, HID, OUT), role);      // d(Wio) += x1a^T dWh2
        warp_gemm_3xtf32_at_pair<OUT, 1, RA, 2>(A, reinterpret_cast<const float*>(dstb), lane, avec_to(gr.aio, OUT), role);
        psync();
        // d(x1a) = dWh2 Wio^T, times elu'(hp1) read back from x1a: d(hp1) -> Q in place
        warp_gemm_3xtf32_bt<OUT, HID / 8, RA, SW2>(B, w.Wio, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            float2* q0 = reinterpret_cast<float2*>(Q + r * RS + col);
            float2* q1 = reinterpret_cast<float2*>(Q + (r + 8) * RS + col);
            const float2 y0 = *q0, y1 = *q1;
            *q0 = make_float2(c[0] * (y0.x > 0.f ? 1.f : y0.x + 1.f), c[1] * (y0.y > 0.f ? 1.f : y0.y + 1.f));
            *q1 = make_float2(c[2] * (y1.x > 0.f ? 1.f : y1.x + 1.f), c[3] * (y1.y > 0.f ? 1.f : y1.y + 1.f));
        }, role, 2);
        psync();
        // =============== intra layer 1 backward ===============
        ds = 0.f; dt = 0.f;
        {
            float dh[HID];
            load_row<HID>(Q + lane * RS, dh);
            if (live) {
                float4 s1;
                att_bwd_row<HID, RS>(P, stA, group_mask, stA[lane].x, alpha, dh, s1, ds);
                if (main_w) stat[lane] = s1;
            }
        }
        psync();
        {
            float dwh[HID];
#pragma unroll
            for (int f = 0; f < HID; ++f) dwh[f] = 0.f;
            if (live)
                att_bwd_col<HID, RS, RS>(P + lane * RS, Q, stA, stat, group_mask, stA[lane].y, ds, alpha, av.ai, dwh, dt);
            if (main_w) dstb[lane] = make_float2(live ? ds : 0.f, live ? dt : 0.f);
            psync();
            if (main_w) store_row<HID>(Q + lane * RS, dwh);             // dWh1
        }
        psync();
        warp_gemm_3xtf32_at_pair<HID, 1, RS, 2>(P, reinterpret_cast<const float*>(dstb), lane, avec_to(gr.ai, HID), role);   // needs Wh1
        psync();
        load_x();
        psync();
        warp_gemm_3xtf32_at_pair<IN, HID / 8, RS, RS>(P, Q, lane, grad_to(gr.Wi, HID, IN, HID), role);         // d(Wi) += x^T dWh1
        // d(x) = dWh1 Wi^T -> HBM
        warp_gemm_3xtf32_bt<HID, IN / 8, RS, SW1>(Q, w.Wi, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            if (r < np) *reinterpret_cast<float2*>(grad_x + (int64_t)(p0 + r) * IN + col) = make_float2(c[0], c[1]);
            if (r + 8 < np) *reinterpret_cast<float2*>(grad_x + (int64_t)(p0 + r + 8) * IN + col) = make_float2(c[2], c[3]);
        }, role, 2);
        psync();
    }
#if !GB_GLOBAL_GRADS
    __syncthreads();
    float* mine = partials + (int64_t)blockIdx.x * GB_GRAD_FLOATS;
    for (int e = threadIdx.x; e < GB_GRAD_FLOATS; e += blockDim.x) mine[e] = reinterpret_cast<const float*>(&gr)[e];
#endif
}

// sums the per-CTA gradient blocks in block order (deterministic) and scatters the ten tensors.  The inter level's first
// layer is differentiated in its aggregated form, where the scores depend on the parameters through u = We [ae1 | ae2]:
// with du = the summed d(u), d(We) += du1 ae1^T + du2 ae2^T and d(ae) = We^T du.
__global__ void gat_bwd_reduce_kernel(const float* __restrict__ partials, int n_blocks, const float* __restrict__ We,
                                      const float* __restrict__ ae, float* gWi, float* gai, float* gWio,
                                      float* gaio, float* gWe, float* gae, float* gWeo, float* gaeo, float* gWo, float* gbo) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int o1 = GB_IN * HID, o2 = o1 + 2 * HID, o3 = o2 + HID * OUT, o4 = o3 + 2 * OUT, o5 = o4 + OUT * HID,
                  o6 = o5 + 2 * HID, o7 = o6 + HID * OUT, o8 = o7 + 2 * OUT, o9 = o8 + GB_FIN * 2 * OUT, o10 = o9 + GB_FIN;
    auto total = [&](int idx) {
        float s = 0.f;
        for (int k = 0; k < n_blocks; ++k) s += partials[(int64_t)k * GB_GRAD_FLOATS + idx];
        return s;
    };
    __shared__ float du[2 * OUT];                           // every block sums the 32 d(u) entries for itself
    if (threadIdx.x < 2 * OUT) du[threadIdx.x] = total(o10 + threadIdx.x);
    __syncthreads();
    if (e >= o10) return;
    if (e >= o5 && e < o6) {                                // d(ae)[which][f] = sum_k We[k][f] du[which][k]
        const int which = (e - o5) / HID, f = (e - o5) % HID;
        float s = 0.f;
        for (int k = 0; k < OUT; ++k) s = fmaf(We[k * HID + f], du[which * OUT + k], s);
        gae[e - o5] = s;
        return;
    }
    float s = total(e);
    if (e < o1) gWi[e] = s;
    else if (e < o2) gai[e - o1] = s;
    else if (e < o3) gWio[e - o2] = s;
    else if (e < o4) gaio[e - o3] = s;
    else if (e < o5) {
        const int k = (e - o4) / HID, f = (e - o4) % HID;
        gWe[e - o4] = s + du[k] * ae[f] + du[OUT + k] * ae[HID + f];
    }
    else if (e < o7) gWeo[e - o6] = s;
    else if (e < o8) gaeo[e - o7] = s;
    else if (e < o9) gWo[e - o8] = s;
    else gbo[e - o9] = s;
}

}  // namespace sgx

using namespace sgx;

extern "C" int64_t sgx_gat_encoder_fused_bwd_ws_bytes(void) { return (int64_t)148 * GB_GRAD_FLOATS * 4 + 256; }

extern "C" int sgx_gat_encoder_fused_bwd(const float* x, const float* grad_out, const int32_t* leader,
                                         const int32_t* group_size, const int32_t* ped_start, const int32_t* ped_end,
                                         const int32_t* scene_start, const int32_t* chunk_scene, int64_t n_chunks,
                                         const float* Wi, const float* ai, const float* Wio, const float* aio,
                                         const float* We, const float* ae, const float* Weo, const float* aeo,
                                         const float* Wo, const float* bo, float alpha, int32_t n_heads, int32_t IN,
                                         int32_t HID_, int32_t OUT_, int32_t FIN, float* grad_x, float* grad_Wi,
                                         float* grad_ai, float* grad_Wio, float* grad_aio, float* grad_We, float* grad_ae,
                                         float* grad_Weo, float* grad_aeo, float* grad_Wo, float* grad_bo, void* workspace,
                                         int64_t ws_bytes, void* stream) {
    SGX_REQUIRE(x && grad_out && leader && group_size && ped_start && ped_end && scene_start && chunk_scene && Wi && ai &&
                    Wio && aio && We && ae && Weo && aeo && Wo && bo && grad_x && grad_Wi && grad_ai && grad_Wio &&
                    grad_aio && grad_We && grad_ae && grad_Weo && grad_aeo && grad_Wo && grad_bo && workspace,
                "sgx_gat_encoder_fused_bwd: null pointer");
    SGX_REQUIRE(n_chunks > 0 && n_chunks < ((int64_t)1 << 31), "sgx_gat_encoder_fused_bwd: bad chunk count");
    SGX_UNSUPPORTED(n_heads != 1 || IN != GB_IN || HID_ != HID || OUT_ != OUT || FIN != GB_FIN,
                    "the single-launch GATEncoder backward is built for n_heads 1, dims 40/72/16/24 (got heads %d, "
                    "%d/%d/%d/%d)", n_heads, IN, HID_, OUT_, FIN);
    SGX_REQUIRE(ws_bytes >= sgx_gat_encoder_fused_bwd_ws_bytes(), "sgx_gat_encoder_fused_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = (int)(sizeof(FusedWm) + sizeof(GatAvec) + (GB_GLOBAL_GRADS ? 0 : sizeof(GatGrad)) +
                           GB_PAIRS * GB_SCRATCH * sizeof(float));
    SGX_CUDA(cudaFuncSetAttribute(gat_fused_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = (int)std::min<int64_t>((n_chunks + GB_PAIRS - 1) / GB_PAIRS, 148);
    float* partials = (float*)workspace;
    gat_fused_bwd_kernel<<<grid, GB_WARPS * 32, smem, st>>>(x, grad_out, leader, group_size, ped_start, ped_end, scene_start,
                                                            chunk_scene, (int)n_chunks, Wi, ai, Wio, aio, We, ae, Weo, aeo,
                                                            Wo, bo, alpha, grad_x, partials);
    SGX_LAUNCH_CHECK();
    gat_bwd_reduce_kernel<<<blocks_for(GB_GRAD_FLOATS, 256), 256, 0, st>>>(partials, grid, We, ae, grad_Wi, grad_ai, grad_Wio,
                                                                           grad_aio, grad_We, grad_ae, grad_Weo, grad_aeo,
                                                                           grad_Wo, grad_bo);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}
