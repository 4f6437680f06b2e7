// Shared pieces of the single-launch GATEncoder kernels (forward: sgx_gat.cu, backward: sgx_gat_bwd.cu): dims, the
// shared-memory weight block with the score projections as two extra columns, mask-driven attention, small epilogues.
#pragma once
#include "sgx_common.cuh"
#include "sgx_warp_mma.cuh"

namespace sgx {

constexpr int HID = 72, OUT = 16;
__device__ __forceinline__ float lrelu(float v, float alpha) { return v > 0.f ? v : alpha * v; }
constexpr int RS = 76;                 // row stride (floats) of the 72-wide row buffer: 16 B aligned, conflict-free STS.128
constexpr int RA = 20;                 // row stride of the 16-wide leader buffer (Xg, Yg); x itself is staged in the
                                       // lane's own 72-wide row, which only that lane reads before overwriting it with Wh
// ex2.approx based exp / ELU for the fused kernel (2 ulp; the general path keeps expf / expm1f): expm1f alone was
// ~25 instructions per element, more dynamic instructions than the 40x72 GEMV it follows.
__device__ __forceinline__ float fexp(float v) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v * 1.4426950408889634f));
    return r;
}
__device__ __forceinline__ float felu(float v) { return v > 0.f ? v : fexp(v) - 1.f; }

// attention of one node over the lanes set in `mask` (its group's members / its scene's leaders), ascending lane order
// = the order of attend_smem, so the two give bit-identical sums; iterates the neighbours only instead of the scene.
template <int F, int STRIDE>
__device__ __forceinline__ void attend_mask(const float* __restrict__ rows, const float2* __restrict__ st, uint32_t mask,
                                            float s_i, float alpha, float (&hp)[F]) {
    float m = -INFINITY;
    for (uint32_t mm = mask; mm; mm &= mm - 1) m = fmaxf(m, lrelu(s_i + st[__ffs(mm) - 1].y, alpha));
    float den = 0.f;
#pragma unroll
    for (int f = 0; f < F; ++f) hp[f] = 0.f;
    for (uint32_t mm = mask; mm; mm &= mm - 1) {
        const int q = __ffs(mm) - 1;
        const float w = fexp(lrelu(s_i + st[q].y, alpha) - m);
        den += w;
        const float4* row = reinterpret_cast<const float4*>(rows + q * STRIDE);
#pragma unroll
        for (int f = 0; f < F / 4; ++f) {
            const float4 v = row[f];
            hp[4 * f] = fmaf(w, v.x, hp[4 * f]); hp[4 * f + 1] = fmaf(w, v.y, hp[4 * f + 1]);
            hp[4 * f + 2] = fmaf(w, v.z, hp[4 * f + 2]); hp[4 * f + 3] = fmaf(w, v.w, hp[4 * f + 3]);
        }
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int f = 0; f < F; ++f) hp[f] *= inv;
}


template <int F>
__device__ __forceinline__ void store_row(float* __restrict__ row, const float (&v)[F]) {
#pragma unroll
    for (int f = 0; f < F / 4; ++f)
        reinterpret_cast<float4*>(row)[f] = make_float4(v[4 * f], v[4 * f + 1], v[4 * f + 2], v[4 * f + 3]);
}

template <int F>
__device__ __forceinline__ void elu_logsoftmax(float (&v)[F]) {
    float mx = -INFINITY;
#pragma unroll
    for (int f = 0; f < F; ++f) { v[f] = felu(v[f]); mx = fmaxf(mx, v[f]); }
    float sum = 0.f;
#pragma unroll
    for (int f = 0; f < F; ++f) sum += fexp(v[f] - mx);
    const float lse = mx + logf(sum);
#pragma unroll
    for (int f = 0; f < F; ++f) v[f] -= lse;
}

constexpr int SW1 = 88;                // row stride of the [K][72+2 (+pad)] weight blocks: 88 % 32 = 24 -> conflict-free B fragments
constexpr int SW2 = 24;                // row stride of the [K][16+2 (+pad)] and [32][24] blocks
struct FusedWm {
    float Wi[40 * SW1], Wio[HID * SW2], We[OUT * SW1], Weo[HID * SW2], WoT[2 * OUT * SW2], bo[24];
};

// weight blocks [K][N + 2 score columns + zero padding]: column N = W a[:N], N+1 = W a[N:]; Wo transposed
template <int IN, int FIN>
__device__ __forceinline__ void fused_load_weights(FusedWm& w, const float* __restrict__ Wi, const float* __restrict__ ai,
                                                   const float* __restrict__ Wio, const float* __restrict__ aio,
                                                   const float* __restrict__ We, const float* __restrict__ ae,
                                                   const float* __restrict__ Weo, const float* __restrict__ aeo,
                                                   const float* __restrict__ Wo, const float* __restrict__ bo) {
    {
        auto fill = [&](float* dst, int stride, const float* W, const float* a, int K, int N) {
            for (int e = threadIdx.x; e < K * stride; e += blockDim.x) {
                const int k = e / stride, n = e % stride;
                float v = 0.f;
                if (n < N) v = W[k * N + n];
                else if (n < N + 2) {
                    const float* av = a + (n - N) * N;
                    for (int c = 0; c < N; ++c) v = fmaf(W[k * N + c], av[c], v);
                }
                dst[e] = v;
            }
        };
        fill(w.Wi, SW1, Wi, ai, IN, HID);
        fill(w.Wio, SW2, Wio, aio, HID, OUT);
        fill(w.We, SW1, We, ae, OUT, HID);
        fill(w.Weo, SW2, Weo, aeo, HID, OUT);
        for (int e = threadIdx.x; e < 2 * OUT * SW2; e += blockDim.x) {
            const int k = e / SW2, n = e % SW2;
            w.WoT[e] = (n < FIN) ? Wo[n * 2 * OUT + k] : 0.f;
        }
        for (int e = threadIdx.x; e < FIN; e += blockDim.x) w.bo[e] = bo[e];
    }
}


}  // namespace sgx
