// Generic fp32 GEMM with element strides: C[m,n] (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn].
// Serves the parameter-gradient reductions (tall-skinny, K = batch -> split-K with atomics) and the
// standalone dense GCN/GAT layers.  Not on the inference hot path.
#include "sgx_common.cuh"

namespace sgx {

constexpr int GT = 64;   // C tile (GT x GT)
constexpr int GK = 16;   // K tile

__global__ void __launch_bounds__(256)
gemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbk,
            int64_t sbn, float* __restrict__ C, int64_t ldc, int M, int N, int64_t K, int64_t k_per_split,
            int accumulate, int relu, int use_atomic, const float* __restrict__ bias_m,
            const float* __restrict__ bias_n) {
    __shared__ float As[GK][GT + 1];
    __shared__ float Bs[GK][GT + 1];
    const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
    const int64_t kb = (int64_t)blockIdx.z * k_per_split;
    const int64_t ke = min(K, kb + k_per_split);
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int64_t k0 = kb; k0 < ke; k0 += GK) {
        // 256 threads load GK*GT = 1024 elements of each operand (4 per thread)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int e = threadIdx.x + r * 256;
            // A tile: choose the faster-varying index along the smaller stride for coalescing
            int mm, kk;
            if (sak <= sam) { kk = e % GK; mm = e / GK; } else { mm = e % GT; kk = e / GT; }
            int64_t k = k0 + kk;
            float v = 0.f;
            if (m0 + mm < M && k < ke) v = A[(int64_t)(m0 + mm) * sam + k * sak];
            As[kk][mm] = v;
            int nn, kk2;
            if (sbn <= sbk) { nn = e % GT; kk2 = e / GT; } else { kk2 = e % GK; nn = e / GK; }
            int64_t k2 = k0 + kk2;
            float w = 0.f;
            if (n0 + nn < N && k2 < ke) w = B[k2 * sbk + (int64_t)(n0 + nn) * sbn];
            Bs[kk2][nn] = w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float* c = C + (int64_t)m * ldc + n;
            if (use_atomic) {
                atomicAdd(c, acc[i][j]);
            } else {
                float v = acc[i][j] + (accumulate ? *c : 0.f) + (bias_m ? bias_m[m] : 0.f) + (bias_n ? bias_n[n] : 0.f);
                *c = relu ? fmaxf(v, 0.f) : v;
            }
        }
    }
}

__global__ void zero_strided_kernel(float* C, int64_t ldc, int M, int N) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)M * N) C[(i / N) * ldc + (i % N)] = 0.f;
}

int gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
         int64_t M, int64_t N, int64_t K, int accumulate, int relu, cudaStream_t st, const float* bias_m, const float* bias_n) {
    SGX_REQUIRE(A && B && C, "gemm: null pointer");
    SGX_REQUIRE(M >= 0 && N >= 0 && K >= 0 && M < (1 << 30) && N < (1 << 30), "gemm: bad shape");
    if (M == 0 || N == 0) return SGX_OK;
    int64_t tiles = ((M + GT - 1) / GT) * ((N + GT - 1) / GT);
    int64_t splits = 1;
    if (K > 4096 && tiles < 296 && !relu && !bias_m && !bias_n) {
        // K slabs of >= 512 rows, up to ~8 CTAs per SM: the reductions this serves are [<= 128 x <= 72] outputs over
        // K = 10^4 .. 10^6 rows (2048-row slabs on 4 CTAs per SM left a 1.3-wave tail: G step 12.0 -> 10.3 ms)
        splits = std::min<int64_t>((K + 511) / 512, std::max<int64_t>(1, 1184 / tiles));
    }
    int64_t kps = (K + splits - 1) / splits;
    kps = (kps + GK - 1) / GK * GK;
    if (kps == 0) kps = GK;
    splits = std::max<int64_t>(1, (K + kps - 1) / kps);
    int use_atomic = splits > 1;
    if (use_atomic && !accumulate) {
        zero_strided_kernel<<<blocks_for(M * N, 256), 256, 0, st>>>(C, ldc, (int)M, (int)N);
        SGX_LAUNCH_CHECK();
    }
    dim3 grid((unsigned)((N + GT - 1) / GT), (unsigned)((M + GT - 1) / GT), (unsigned)splits);
    gemm_kernel<<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, (int)M, (int)N, K, kps, accumulate, relu,
                                      use_atomic, bias_m, bias_n);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

}  // namespace sgx

extern "C" int sgx_gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                        int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t accumulate, int32_t relu,
                        void* stream) {
    return sgx::gemm(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, accumulate, relu, (cudaStream_t)stream);
}
