// Two-layer context MLP of the generator / decoder / discriminator in ONE launch (inference):
//     out = ReLU(W2 ReLU(W1 [xa ; xb] + b1) + b2)
// = make_mlp([in, mid, out]) with activation relu, batch_norm 0, dropout 0 (sgan/models.py:7-20) at its call sites
// sgan/models.py:898 (mlp_decoder_context, SGAN-P wiring), :165-166 (decoder.mlp after per-step pooling) and :990
// (real_classifier).  The reference runs cat + addmm + relu + addmm + relu (5 launches, the 64-wide hidden row goes
// through HBM twice); here a warp owns 32 consecutive pedestrians: their rows are staged in shared memory, both linear
// maps are warp-level 3xTF32 tensor-core GEMMs (sgx_warp_mma.cuh: ~7e-7 relative, inside the 1e-5 contract) with the
// hidden row kept in shared memory, and the concatenation [xa ; xb] (final_encoder_h ; pool_h) is folded into the load.
// HBM traffic = the algorithmic (in + out) * 4 bytes per pedestrian.
#include "sgx_common.cuh"
#include "sgx_warp_mma.cuh"

namespace sgx {

constexpr int MLP_WARPS = 8;

template <int IN, int HID, int OUTP>
struct MlpCfg {
    static constexpr int SA = (IN > HID ? IN : HID) + 4;      // row stride: = 4 mod 32 -> conflict-free A fragments
    static constexpr int SB1 = HID + 8;                       // = 8 mod 32 -> conflict-free B fragments
    static constexpr int SB2 = OUTP == 24 ? 24 : OUTP + 8;    // 24 and (8k + 8) are conflict-free too
    static constexpr int WFLOATS = IN * SB1 + HID * SB2 + HID + OUTP;
    static constexpr int SMEM = (WFLOATS + MLP_WARPS * 32 * SA) * (int)sizeof(float);
};

template <int IN, int HID, int OUTP>
__global__ void __launch_bounds__(MLP_WARPS * 32)
mlp2_fused_kernel(const float* __restrict__ xa, int da, const float* __restrict__ xb, int db, int64_t batch,
                  const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                  const float* __restrict__ b2, int OUT, float* __restrict__ out) {
    using C = MlpCfg<IN, HID, OUTP>;
    extern __shared__ __align__(16) uint8_t raw[];
    float* sW1 = reinterpret_cast<float*>(raw);          // [IN][SB1]  = W1^T
    float* sW2 = sW1 + IN * C::SB1;                      // [HID][SB2] = W2^T, columns >= OUT zero
    float* sb1 = sW2 + HID * C::SB2;
    float* sb2 = sb1 + HID;
    float* rows = sb2 + OUTP;
    for (int e = threadIdx.x; e < IN * HID; e += blockDim.x) {
        const int n = e / IN, k = e % IN;                // coalesced read of W1[n][k]
        sW1[k * C::SB1 + n] = W1[e];
    }
    for (int e = threadIdx.x; e < HID * C::SB2; e += blockDim.x) {
        const int k = e / C::SB2, n = e % C::SB2;
        sW2[e] = n < OUT ? W2[n * HID + k] : 0.f;
    }
    for (int e = threadIdx.x; e < HID; e += blockDim.x) sb1[e] = b1[e];
    for (int e = threadIdx.x; e < OUTP; e += blockDim.x) sb2[e] = e < OUT ? b2[e] : 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float* RB = rows + warp * 32 * C::SA;
    const int64_t n_chunks = (batch + 31) / 32;
    for (int64_t chunk = (int64_t)blockIdx.x * MLP_WARPS + warp; chunk < n_chunks; chunk += (int64_t)gridDim.x * MLP_WARPS) {
        const int64_t p0 = chunk * 32;
        const int np = (int)((batch - p0) < 32 ? (batch - p0) : 32);
        // ---- [xa ; xb] rows of the chunk -> shared memory (both sources are contiguous over the chunk) ----
        {
            const int qa = da / 4;
            const float4* src = reinterpret_cast<const float4*>(xa + p0 * da);
            for (int i = lane; i < 32 * qa; i += 32) {
                const int r = i / qa, c = i % qa;
                const float4 v = r < np ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(RB + r * C::SA + 4 * c) = v;
            }
            if (db > 0) {
                const int qb = db / 4;
                const float4* srcb = reinterpret_cast<const float4*>(xb + p0 * db);
                for (int i = lane; i < 32 * qb; i += 32) {
                    const int r = i / qb, c = i % qb;
                    const float4 v = r < np ? srcb[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    *reinterpret_cast<float4*>(RB + r * C::SA + da + 4 * c) = v;
                }
            }
        }
        __syncwarp();
        // ---- hidden = ReLU(x W1^T + b1), written over the x rows (the GEMM finishes an m-tile before storing it) ----
        warp_gemm_3xtf32<IN, HID / 8, C::SA, C::SB1>(RB, sW1, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            const float u0 = sb1[col], u1 = sb1[col + 1];
            *reinterpret_cast<float2*>(RB + r * C::SA + col) = make_float2(fmaxf(c[0] + u0, 0.f), fmaxf(c[1] + u1, 0.f));
            *reinterpret_cast<float2*>(RB + (r + 8) * C::SA + col) = make_float2(fmaxf(c[2] + u0, 0.f), fmaxf(c[3] + u1, 0.f));
        });
        // ---- out = ReLU(hidden W2^T + b2) ----
        warp_gemm_3xtf32<HID, OUTP / 8, C::SA, C::SB2>(RB, sW2, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            const float u0 = sb2[col], u1 = sb2[col + 1];
            const float y0 = fmaxf(c[0] + u0, 0.f), y1 = fmaxf(c[1] + u1, 0.f);
            const float y2 = fmaxf(c[2] + u0, 0.f), y3 = fmaxf(c[3] + u1, 0.f);
            if (OUTP % 2 == 0 && col + 1 < OUT) {
                if (r < np) *reinterpret_cast<float2*>(out + (p0 + r) * OUT + col) = make_float2(y0, y1);
                if (r + 8 < np) *reinterpret_cast<float2*>(out + (p0 + r + 8) * OUT + col) = make_float2(y2, y3);
            } else if (col < OUT) {                       // odd OUT (the discriminator's single score)
                if (r < np) out[(p0 + r) * OUT + col] = y0;
                if (r + 8 < np) out[(p0 + r + 8) * OUT + col] = y2;
            }
        });
    }
}

template <int IN, int HID, int OUTP>
static int launch_mlp2(const float* xa, int da, const float* xb, int db, int64_t batch, const float* W1, const float* b1,
                       const float* W2, const float* b2, int OUT, float* out, cudaStream_t st) {
    using C = MlpCfg<IN, HID, OUTP>;
    auto kern = mlp2_fused_kernel<IN, HID, OUTP>;
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    int dev = 0, sms = 148;
    SGX_CUDA(cudaGetDevice(&dev));
    SGX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t n_chunks = (batch + 31) / 32;
    const unsigned grid = (unsigned)std::min<int64_t>((n_chunks + MLP_WARPS - 1) / MLP_WARPS, 2 * sms);
    kern<<<grid, MLP_WARPS * 32, C::SMEM, st>>>(xa, da, xb, db, batch, W1, b1, W2, b2, OUT, out);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

template <int IN, int OUTP>
int mlp2_tc_forward(const float* xa, int da, const float* xb, int db, int64_t batch, const float* W1, const float* b1,
                    const float* W2, const float* b2, int OUT, float* out, cudaStream_t st);   // sgx_mlp_tc.cu

}  // namespace sgx

using namespace sgx;

extern "C" int sgx_mlp2_supported(int32_t IN, int32_t HID, int32_t OUT) {
    const bool in_ok = IN == 32 || IN == 40 || IN == 48;
    const bool out_ok = OUT == 1 || OUT == 24 || OUT == 32;
    return (in_ok && HID == 64 && out_ok) ? 1 : 0;
}

extern "C" int sgx_mlp2_fwd(const float* xa, int32_t da, const float* xb, int32_t db, int64_t batch, const float* W1,
                            const float* b1, const float* W2, const float* b2, int32_t HID, int32_t OUT, float* out,
                            void* stream) {
    SGX_REQUIRE(xa && W1 && b1 && W2 && b2 && out && batch > 0, "sgx_mlp2_fwd: null pointer / empty batch");
    SGX_REQUIRE(da > 0 && da % 4 == 0 && db >= 0 && db % 4 == 0 && (db == 0 || xb), "sgx_mlp2_fwd: input widths must be multiples of 4");
    const int IN = da + db;
    SGX_UNSUPPORTED(!sgx_mlp2_supported(IN, HID, OUT), "sgx_mlp2_fwd is built for in 32|40|48, mid 64, out 1|24|32; got %d -> %d -> %d",
                    IN, HID, OUT);
    cudaStream_t st = (cudaStream_t)stream;
#define SGX_MLP_CASE(I, O, OP)                                                                                        \
    if (IN == I && OUT == O)                                                                                          \
        return opt_graph_tc() ? mlp2_tc_forward<I, OP>(xa, da, xb, db, batch, W1, b1, W2, b2, OUT, out, st)           \
                              : launch_mlp2<I, 64, OP>(xa, da, xb, db, batch, W1, b1, W2, b2, OUT, out, st);
    SGX_MLP_CASE(32, 24, 24) SGX_MLP_CASE(40, 24, 24) SGX_MLP_CASE(48, 24, 24)
    SGX_MLP_CASE(32, 32, 32) SGX_MLP_CASE(40, 32, 32) SGX_MLP_CASE(48, 32, 32)
    SGX_MLP_CASE(32, 1, 8) SGX_MLP_CASE(40, 1, 8) SGX_MLP_CASE(48, 1, 8)
#undef SGX_MLP_CASE
    return SGX_ERR_UNSUPPORTED;
}
