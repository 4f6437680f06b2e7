// Dense-adjacency attention rows for the standalone GraphAttentionLayer.forward(h, adj)
// (sgan/models.py:198-210).  The encoder never takes this path (it works on the group structure,
// sgx_gat.cu); this exists so the layer keeps working for an arbitrary dense `adj`.
//   e_ij  = LeakyReLU(s_i + t_j)            (s = Wh a[:F], t = Wh a[F:], computed by sgx_gemm)
//   att_i = softmax_j(adj_ij > 0 ? e_ij : -9e15)       -- literally, so a fully masked row is uniform
// One CTA per row; the row lives in global memory (n is arbitrary).
#include "sgx_common.cuh"

namespace sgx {

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float other = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, other) : v + other;
    }
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    float r = sh[0];
    for (int w = 1; w < nw; ++w) r = is_max ? fmaxf(r, sh[w]) : r + sh[w];
    return r;
}

__global__ void __launch_bounds__(128)
dense_att_fwd_kernel(const float* __restrict__ st, const float* __restrict__ adj, int n, float alpha,
                     float* __restrict__ att) {
    __shared__ float sh[4];
    const int i = blockIdx.x;
    const float s_i = st[2 * i];
    const float* arow = adj + (int64_t)i * n;
    float* orow = att + (int64_t)i * n;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float pre = s_i + st[2 * j + 1];
        float e = arow[j] > 0.f ? (pre > 0.f ? pre : alpha * pre) : -9e15f;
        orow[j] = e;
        m = fmaxf(m, e);
    }
    m = block_reduce(m, true, sh);
    float sum = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float w = expf(orow[j] - m);
        orow[j] = w;
        sum += w;
    }
    sum = block_reduce(sum, false, sh);
    const float inv = 1.f / sum;
    for (int j = threadIdx.x; j < n; j += blockDim.x) orow[j] *= inv;
}

// datt (in) -> d(pre-activation) (out, in place); ds_i = sum_j dpre_ij
__global__ void __launch_bounds__(128)
dense_att_bwd_kernel(const float* __restrict__ st, const float* __restrict__ adj, const float* __restrict__ att,
                     int n, float alpha, float* __restrict__ datt, float* __restrict__ ds) {
    __shared__ float sh[4];
    const int i = blockIdx.x;
    const float s_i = st[2 * i];
    const float* arow = adj + (int64_t)i * n;
    const float* prow = att + (int64_t)i * n;
    float* drow = datt + (int64_t)i * n;
    float c = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) c = fmaf(prow[j], drow[j], c);
    c = block_reduce(c, false, sh);
    float acc = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        float pre = s_i + st[2 * j + 1];
        float d = arow[j] > 0.f ? prow[j] * (drow[j] - c) * (pre > 0.f ? 1.f : alpha) : 0.f;
        drow[j] = d;
        acc += d;
    }
    acc = block_reduce(acc, false, sh);
    if (threadIdx.x == 0) ds[i] = acc;
}

}  // namespace sgx

extern "C" int sgx_dense_att_fwd(const float* st, const float* adj, int64_t n, float alpha, float* att, void* stream) {
    SGX_REQUIRE(st && adj && att && n > 0 && n < (1 << 30), "sgx_dense_att_fwd: bad arguments");
    sgx::dense_att_fwd_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(st, adj, (int)n, alpha, att);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

extern "C" int sgx_dense_att_bwd(const float* st, const float* adj, const float* att, int64_t n, float alpha,
                                 float* datt_inout, float* ds, void* stream) {
    SGX_REQUIRE(st && adj && att && datt_inout && ds && n > 0 && n < (1 << 30), "sgx_dense_att_bwd: bad arguments");
    sgx::dense_att_bwd_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(st, adj, att, (int)n, alpha, datt_inout, ds);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}
