// Generic fp32 GEMM with element strides: C[m,n] (+)= sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn].
// Serves the parameter-gradient reductions (tall-skinny, K = batch -> split-K with atomics) and the
// standalone dense GCN/GAT layers.  Not on the inference hot path.
#include "sgx_common.cuh"
#include "sgx_warp_mma.cuh"

namespace sgx {

constexpr int GK = 32;   // K tile
constexpr int GAS = GK + 4;   // row stride of the A tile [m][k]: conflict-free A fragments (4 g + t)

// BM x BN C tile per CTA (64 x 64, or 128 x 32 for the skinny N <= 32 products: a 64-wide tile would spend half of its
// MMAs on zero padding), 8 warps as (BM/32) x (BN/16), each a 32 x 16 sub-tile = 2 x 2 mma.sync m16n8k8 fragments; every
// operand is split hi + lo (3xTF32: lo.hi + hi.lo + hi.hi in fp32, ~7e-7 relative -- fp32-grade, sgx_warp_mma.cuh).
// The CUDA-core version of this kernel spent its time on shared-memory operand reads (8 LDS per 16 FMA): 8 TFLOP/s on
// the [batch x 512] x [512 x 32] products of the pooling backward, i.e. 0.5 - 0.7 ms per 500 MB operand.
template <int BM, int BN>
__global__ void __launch_bounds__(256, 2)
gemm_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbk,
            int64_t sbn, float* __restrict__ C, int64_t ldc, int M, int N, int64_t K, int64_t k_per_split,
            int accumulate, int relu, int use_atomic, const float* __restrict__ bias_m,
            const float* __restrict__ bias_n) {
    constexpr int GBS = BN + 8;   // row stride of the B tile [k][n]: conflict-free B fragments (8 t + g)
    constexpr int NA = BM / 8, NB = BN / 8;               // elements per thread and K tile
    __shared__ float As[BM][GAS];
    __shared__ float Bs[GK][GBS];
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int64_t kb = (int64_t)blockIdx.z * k_per_split;
    const int64_t ke = min(K, kb + k_per_split);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp / (BN / 16)) * 32, wn = (warp % (BN / 16)) * 16;
    float acc[2][2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.f;

    // A thread moves 8 elements of each operand per K tile: element r sits a constant stride behind element 0, so the
    // addresses are one pointer + r * stride (the index arithmetic of the first version was half of the instructions).
    // The faster-varying index runs along the operand's smaller stride so that a warp's loads coalesce.
    const bool a_kfast = sak <= sam, b_nfast = sbn <= sbk;
    const int a_m = a_kfast ? (threadIdx.x >> 5) : (threadIdx.x % BM), a_k = a_kfast ? (threadIdx.x & 31) : (threadIdx.x / BM);
    const int a_dm = a_kfast ? 8 : 0, a_dk = a_kfast ? 0 : 256 / BM;          // per-element step in (m, k)
    const int b_n = b_nfast ? (threadIdx.x % BN) : (threadIdx.x >> 5), b_k = b_nfast ? (threadIdx.x / BN) : (threadIdx.x & 31);
    const int b_dn = b_nfast ? 0 : 8, b_dk = b_nfast ? 256 / BN : 0;
    const float* pa = A + (int64_t)(m0 + a_m) * sam + (kb + a_k) * sak;
    const float* pb = B + (kb + b_k) * sbk + (int64_t)(n0 + b_n) * sbn;
    const int64_t a_step = a_dm * sam + a_dk * sak, b_step = b_dn * sbn + b_dk * sbk;
    float* sa = &As[a_m][a_k];
    float* sb = &Bs[b_k][b_n];
    const int sa_step = a_dm * GAS + a_dk, sb_step = b_dk * GBS + b_dn;
    float av[NA], bv[NB];
    auto fetch = [&](int64_t k0) {
        const bool interior = (k0 + GK <= ke) && (m0 + BM <= M) && (n0 + BN <= N);
        if (interior) {
#pragma unroll
            for (int r = 0; r < NA; ++r) av[r] = pa[r * a_step];
#pragma unroll
            for (int r = 0; r < NB; ++r) bv[r] = pb[r * b_step];
        } else {
#pragma unroll
            for (int r = 0; r < NA; ++r)
                av[r] = (m0 + a_m + r * a_dm < M && k0 + a_k + r * a_dk < ke) ? pa[r * a_step] : 0.f;
#pragma unroll
            for (int r = 0; r < NB; ++r)
                bv[r] = (n0 + b_n + r * b_dn < N && k0 + b_k + r * b_dk < ke) ? pb[r * b_step] : 0.f;
        }
        pa += GK * sak;
        pb += GK * sbk;
    };
    fetch(kb);
    for (int64_t k0 = kb; k0 < ke; k0 += GK) {
        __syncthreads();                                   // the previous tile's fragments have been read
#pragma unroll
        for (int r = 0; r < NA; ++r) sa[r * sa_step] = av[r];
#pragma unroll
        for (int r = 0; r < NB; ++r) sb[r * sb_step] = bv[r];
        __syncthreads();
        if (k0 + GK < ke) fetch(k0 + GK);                  // the next tile's loads are in flight during the MMAs
        // the tensor core adds into its accumulator with truncation: chains are kept to one K tile (12 MMAs) and the
        // tile sums are added with IEEE fp32 adds, otherwise the bias grows with K (3.8e-6 at K = 10^4)
        float part[2][2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) part[i][j][0] = part[i][j][1] = part[i][j][2] = part[i][j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < GK / 8; ++ks) {
            uint32_t ah[2][4], al[2][4], bh[2][2], bl[2][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float* ar = &As[wm + i * 16 + g][ks * 8 + t];
                split_tf32_rn(ar[0], ah[i][0], al[i][0]);
                split_tf32_rn(ar[8 * GAS], ah[i][1], al[i][1]);
                split_tf32_rn(ar[4], ah[i][2], al[i][2]);
                split_tf32_rn(ar[8 * GAS + 4], ah[i][3], al[i][3]);
            }
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float* br = &Bs[ks * 8 + t][wn + j * 8 + g];
                split_tf32_rn(br[0], bh[j][0], bl[j][0]);
                split_tf32_rn(br[4 * GBS], bh[j][1], bl[j][1]);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(part[i][j], al[i], bh[j][0], bh[j][1]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(part[i][j], ah[i], bl[j][0], bl[j][1]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) mma_tf32(part[i][j], ah[i], bh[j][0], bh[j][1]);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][j][q] += part[i][j][q];
    }
    const bool vec2 = !use_atomic && (ldc % 2 == 0) && ((reinterpret_cast<uintptr_t>(C) & 7) == 0);
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int m = m0 + wm + i * 16 + g + h * 8, n = n0 + wn + j * 8 + 2 * t;
                if (m >= M || n >= N) continue;
                float* c = C + (int64_t)m * ldc + n;
                if (vec2 && n + 1 < N) {                    // the fragment's two columns as one 8-byte store
                    float2 v = make_float2(acc[i][j][h * 2], acc[i][j][h * 2 + 1]);
                    if (accumulate) { const float2 o = *reinterpret_cast<const float2*>(c); v.x += o.x; v.y += o.y; }
                    if (bias_m) { v.x += bias_m[m]; v.y += bias_m[m]; }
                    if (bias_n) { v.x += bias_n[n]; v.y += bias_n[n + 1]; }
                    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
                    *reinterpret_cast<float2*>(c) = v;
                    continue;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (n + q >= N) continue;
                    const float a = acc[i][j][h * 2 + q];
                    if (use_atomic) {
                        atomicAdd(c + q, a);
                    } else {
                        const float v = a + (accumulate ? c[q] : 0.f) + (bias_m ? bias_m[m] : 0.f) + (bias_n ? bias_n[n + q] : 0.f);
                        c[q] = relu ? fmaxf(v, 0.f) : v;
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
    const bool skinny = N <= 32;                          // 128 x 32 tiles instead of 64 x 64
    const int BM = skinny ? 128 : 64, BN = skinny ? 32 : 64;
    int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
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
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)splits);
    if (skinny)
        gemm_kernel<128, 32><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, (int)M, (int)N, K, kps, accumulate, relu,
                                                   use_atomic, bias_m, bias_n);
    else
        gemm_kernel<64, 64><<<grid, 256, 0, st>>>(A, sam, sak, B, sbk, sbn, C, ldc, (int)M, (int)N, K, kps, accumulate, relu,
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
