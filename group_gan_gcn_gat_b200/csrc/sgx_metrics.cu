// Displacement metrics and best-of-K segment reductions.
//
// Reference: relative_to_abs (sgan/utils.py:83-96) + displacement_error / final_displacement_error mode='raw'
// (sgan/losses.py:74-119) per sample, then evaluate_helper (scripts/evaluate_model.py:58-69): per scene, sum over
// pedestrians and min over the K samples -- a python loop over scenes with .item() calls in the reference.
// Here: one kernel per sample (thread per pedestrian walks the T steps: cumsum, distance, accumulate) writing column k
// of [batch, K] buffers, and one kernel for the per-scene sum + min over K + global sum.
#include "sgx_common.cuh"

namespace sgx {

__global__ void displacement_kernel(const float* __restrict__ pred_rel, const float* __restrict__ start_pos,
                                    const float* __restrict__ gt, int T, int batch, float* __restrict__ ade,
                                    float* __restrict__ fde, int K, int k) {
    pdl_trigger();
    pdl_wait();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    float2 pos = *reinterpret_cast<const float2*>(start_pos + 2 * (int64_t)p);
    float acc = 0.f, last = 0.f;
#pragma unroll 4
    for (int t = 0; t < T; ++t) {                    // (the loads of four steps are in flight together; the sums stay in step order)
        const float2 d = *reinterpret_cast<const float2*>(pred_rel + ((int64_t)t * batch + p) * 2);
        const float2 g = *reinterpret_cast<const float2*>(gt + ((int64_t)t * batch + p) * 2);
        pos.x += d.x;
        pos.y += d.y;
        const float ex = g.x - pos.x, ey = g.y - pos.y;
        last = sqrtf(ex * ex + ey * ey);
        acc += last;
    }
    ade[(int64_t)p * K + k] = acc;
    fde[(int64_t)p * K + k] = last;
}

// a warp per scene at a time, lanes <-> samples (K <= 32): out[0] += min_k sum_p ade[p][k], out[1] likewise for fde.
// Warps stride over the scenes and keep their partial sums; one pair of atomics per BLOCK (65 k scenes used to mean 131 k
// atomics on two addresses: most of the kernel's 62 us).
__global__ void __launch_bounds__(256)
best_of_k_kernel(const float* __restrict__ ade, const float* __restrict__ fde, const int32_t* __restrict__ scene_start,
                 int n_scenes, int K, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int n_warps = (gridDim.x * blockDim.x) >> 5;
    float ta = 0.f, tf = 0.f;                          // lane 0: this warp's sums over its scenes
    for (int scene = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; scene < n_scenes; scene += n_warps) {
        const int b = scene_start[scene], e = scene_start[scene + 1];
        float sa = 0.f, sf = 0.f;
        if (lane < K)
            for (int p = b; p < e; ++p) { sa += ade[(int64_t)p * K + lane]; sf += fde[(int64_t)p * K + lane]; }
        if (lane >= K) { sa = INFINITY; sf = INFINITY; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sa = fminf(sa, __shfl_xor_sync(0xffffffffu, sa, o));
            sf = fminf(sf, __shfl_xor_sync(0xffffffffu, sf, o));
        }
        ta += sa;
        tf += sf;
    }
    __shared__ float s_a[8], s_f[8];
    if (lane == 0) { s_a[wib] = ta; s_f[wib] = tf; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, f = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += s_a[w]; f += s_f[w]; }
        atomicAdd(&out[0], a);
        atomicAdd(&out[1], f);
    }
}

}  // namespace sgx

extern "C" int sgx_displacement_errors(const float* pred_rel, const float* start_pos, const float* gt, int32_t T,
                                       int64_t batch, float* ade, float* fde, int32_t K, int32_t k, void* stream) {
    SGX_REQUIRE(pred_rel && start_pos && gt && ade && fde, "sgx_displacement_errors: null pointer");
    SGX_REQUIRE(T >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && K >= 1 && k >= 0 && k < K,
                "sgx_displacement_errors: bad shape");
    SGX_CUDA(sgx::launch_pdl(sgx::displacement_kernel, dim3(sgx::blocks_for(batch, 256)), dim3(256), 0, (cudaStream_t)stream,
                             true, pred_rel, start_pos, gt, T, (int)batch, ade, fde, K, k));
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

extern "C" int sgx_best_of_k(const float* ade, const float* fde, const int32_t* scene_start, int64_t n_scenes, int32_t K,
                             float* out2, void* stream) {
    SGX_REQUIRE(ade && fde && scene_start && out2 && n_scenes >= 1, "sgx_best_of_k: bad arguments");
    SGX_UNSUPPORTED(K < 1 || K > 32, "sgx_best_of_k: K=%d samples, built for 1..32", K);
    cudaStream_t st = (cudaStream_t)stream;
    SGX_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(float), st));
    sgx::best_of_k_kernel<<<(unsigned)std::min<int64_t>(sgx::blocks_for(n_scenes * 32, 256), 1184), 256, 0, st>>>(
        ade, fde, scene_start, (int)n_scenes, K, out2);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}
