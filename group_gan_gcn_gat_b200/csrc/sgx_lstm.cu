// Fused per-pedestrian LSTM recurrences of the generator / discriminator.
//
// Reference: Encoder.forward (sgan/models.py:62-92): Linear(2,E) + single-layer LSTM over obs_len steps;
//            Decoder.forward (sgan/models.py:142-178) with pool_every_timestep = 0: per step LSTM cell ->
//            hidden2pos -> spatial_embedding of the predicted displacement -> next step.
// In both, every pedestrian's recurrence is independent of every other pedestrian, so the whole
// T-step loop runs in ONE kernel with one thread per pedestrian: h in registers, c / h_next in shared
// memory columns, W_hh (4H x H, 16 KB for H = 32) broadcast from shared memory as 128-bit loads.  The
// embedding is folded into the input weights exactly:  W_ih (We d + be) = (W_ih We) d + W_ih be.
// The reference issues one cuDNN call (plus ~6 small kernels) per decoder step; on the bench workload
// cuDNN's persistent-RNN kernel took 3.7 ms per call (13 calls = 90 % of the generator forward).
// Inference only: under autograd the modules keep using nn.LSTM (cuDNN) so training semantics are unchanged.
#include <stdlib.h>

#include "sgx_common.cuh"

namespace sgx {

__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
    // 1 - 2/(1+e^{2x}); __expf is ex2.approx (2 ulp), the subtraction is absolutely accurate to ~1e-7
    const float e = __expf(2.f * x);
    return 1.f - __fdividef(2.f, 1.f + e);
}

constexpr int PPT = 1;           // pedestrians per thread: every broadcast W_hh load feeds 2 x 4 FMAs, which moves the
                                 // kernel from shared-memory-LSU bound (ncu: 79 % LSU wavefronts, 45 % FMA) to FMA bound

template <int H, int THREADS>
struct LstmSmem {
    float whh[4 * H * H];            // [4H][H]
    float4 wxb[4 * H];               // per gate row: ((W_ih We)[r][0], (W_ih We)[r][1], W_ih be + b_ih + b_hh, 0): one LDS.128
    float c[PPT * H * THREADS];      // cell state, column per (ped slot, thread)
    float hn[PPT * H * THREADS];     // next hidden state
    float whp[2 * H + 2];            // hidden2pos (decoder only)
};

template <int H, int THREADS>
__device__ __forceinline__ void lstm_load_weights(LstmSmem<H, THREADS>& s, const float* __restrict__ We,
                                                  const float* __restrict__ be, const float* __restrict__ W_ih,
                                                  const float* __restrict__ W_hh, const float* __restrict__ b_ih,
                                                  const float* __restrict__ b_hh, int E) {
    for (int e = threadIdx.x; e < 4 * H * H; e += THREADS) s.whh[e] = W_hh[e];
    for (int r = threadIdx.x; r < 4 * H; r += THREADS) {
        float ax = 0.f, ay = 0.f, b = b_ih[r] + b_hh[r];
        for (int e = 0; e < E; ++e) {
            const float w = W_ih[r * E + e];
            ax = fmaf(w, We[2 * e], ax);
            ay = fmaf(w, We[2 * e + 1], ay);
            b = fmaf(w, be[e], b);
        }
        s.wxb[r] = make_float4(ax, ay, b, 0.f);
    }
}

// one LSTM cell update for this thread's PPT pedestrians; input of ped k is the 2-vector (dx[k], dy[k])
template <int H, int THREADS>
__device__ __forceinline__ void lstm_cell(LstmSmem<H, THREADS>& s, float (&h)[PPT][H], const float (&dx)[PPT],
                                          const float (&dy)[PPT]) {
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int u = 0; u < H; ++u) {
        float g[PPT][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = q * H + u;
            const float4 wi = s.wxb[r];
            float2 acc[PPT];       // packed FFMA2: .x accumulates even k, .y odd k
#pragma unroll
            for (int k = 0; k < PPT; ++k) acc[k] = make_float2(fmaf(wi.x, dx[k], wi.z), wi.y * dy[k]);
            const float4* w = reinterpret_cast<const float4*>(&s.whh[r * H]);
#pragma unroll
            for (int j = 0; j < H / 4; ++j) {
                const float4 v = w[j];
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    acc[k] = ffma2(make_float2(v.x, v.y), make_float2(h[k][4 * j], h[k][4 * j + 1]), acc[k]);
                    acc[k] = ffma2(make_float2(v.z, v.w), make_float2(h[k][4 * j + 2], h[k][4 * j + 3]), acc[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < PPT; ++k) g[k][q] = acc[k].x + acc[k].y;
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int col = (k * H + u) * THREADS + tid;
            const float cn = sigmoid_f(g[k][1]) * s.c[col] + sigmoid_f(g[k][0]) * tanh_f(g[k][2]);
            s.c[col] = cn;
            s.hn[col] = sigmoid_f(g[k][3]) * tanh_f(cn);
        }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k)
#pragma unroll
        for (int u = 0; u < H; ++u) h[k][u] = s.hn[(k * H + u) * THREADS + tid];
}

template <int H, int THREADS>
__global__ void __launch_bounds__(THREADS)
lstm_encoder_kernel(const float* __restrict__ obs_rel, int T, int batch, const float* __restrict__ We,
                    const float* __restrict__ be, const float* __restrict__ W_ih, const float* __restrict__ W_hh,
                    const float* __restrict__ b_ih, const float* __restrict__ b_hh, int E, float* __restrict__ h_out) {
    extern __shared__ __align__(16) uint8_t raw[];
    LstmSmem<H, THREADS>& s = *reinterpret_cast<LstmSmem<H, THREADS>*>(raw);
    lstm_load_weights<H, THREADS>(s, We, be, W_ih, W_hh, b_ih, b_hh, E);
    __syncthreads();
    const int tid = threadIdx.x;
    for (int p0 = blockIdx.x * THREADS * PPT; p0 < batch; p0 += gridDim.x * THREADS * PPT) {
        int p[PPT];
        bool live[PPT];
        float h[PPT][H];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            p[k] = p0 + k * THREADS + tid;
            live[k] = p[k] < batch;
#pragma unroll
            for (int u = 0; u < H; ++u) { h[k][u] = 0.f; s.c[(k * H + u) * THREADS + tid] = 0.f; }
        }
        for (int t = 0; t < T; ++t) {
            float dx[PPT], dy[PPT];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                float2 d = make_float2(0.f, 0.f);
                if (live[k]) d = *reinterpret_cast<const float2*>(obs_rel + ((int64_t)t * batch + p[k]) * 2);
                dx[k] = d.x; dy[k] = d.y;
            }
            lstm_cell<H, THREADS>(s, h, dx, dy);
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!live[k]) continue;
#pragma unroll
            for (int u = 0; u < H; u += 4)
                *reinterpret_cast<float4*>(h_out + (int64_t)p[k] * H + u) =
                    make_float4(h[k][u], h[k][u + 1], h[k][u + 2], h[k][u + 3]);
        }
    }
}

template <int H, int THREADS>
__global__ void __launch_bounds__(THREADS)
lstm_decoder_kernel(const float* __restrict__ h0, const float* __restrict__ c0, const float* __restrict__ last_pos_rel,
                    const float* __restrict__ z, const int32_t* __restrict__ ped_scene, int nz, int steps, int batch, const float* __restrict__ We, const float* __restrict__ be,
                    const float* __restrict__ W_ih, const float* __restrict__ W_hh, const float* __restrict__ b_ih,
                    const float* __restrict__ b_hh, const float* __restrict__ W_hp, const float* __restrict__ b_hp, int E,
                    float* __restrict__ pred_rel, float* __restrict__ h_final, float* __restrict__ c_final) {
    extern __shared__ __align__(16) uint8_t raw[];
    LstmSmem<H, THREADS>& s = *reinterpret_cast<LstmSmem<H, THREADS>*>(raw);
    lstm_load_weights<H, THREADS>(s, We, be, W_ih, W_hh, b_ih, b_hh, E);
    for (int e = threadIdx.x; e < 2 * H; e += THREADS) s.whp[e] = W_hp[e];
    if (threadIdx.x < 2) s.whp[2 * H + threadIdx.x] = b_hp[threadIdx.x];
    __syncthreads();
    const int tid = threadIdx.x;
    for (int p0 = blockIdx.x * THREADS * PPT; p0 < batch; p0 += gridDim.x * THREADS * PPT) {
        int p[PPT];
        bool live[PPT];
        float h[PPT][H], dx[PPT], dy[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            p[k] = p0 + k * THREADS + tid;
            live[k] = p[k] < batch;
#pragma unroll
            // h0 row = [h0[p][0 .. H-nz) | z[ped_scene[p]][0 .. nz)]: the add_noise concat of models.py:837-846 folded in
            const int hc = H - nz;
            const int sc = (live[k] && nz > 0) ? ped_scene[p[k]] : 0;
#pragma unroll
            for (int u = 0; u < H; ++u) {
                float v = 0.f;
                if (live[k]) v = (u < hc) ? h0[(int64_t)p[k] * hc + u] : z[(int64_t)sc * nz + (u - hc)];
                h[k][u] = v;
                s.c[(k * H + u) * THREADS + tid] = (live[k] && c0) ? c0[(int64_t)p[k] * H + u] : 0.f;
            }
            float2 d = make_float2(0.f, 0.f);
            if (live[k]) d = *reinterpret_cast<const float2*>(last_pos_rel + (int64_t)p[k] * 2);
            dx[k] = d.x; dy[k] = d.y;
        }
        for (int t = 0; t < steps; ++t) {
            lstm_cell<H, THREADS>(s, h, dx, dy);
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                float rx = s.whp[2 * H], ry = s.whp[2 * H + 1];
#pragma unroll
                for (int u = 0; u < H; ++u) {
                    rx = fmaf(s.whp[u], h[k][u], rx);
                    ry = fmaf(s.whp[H + u], h[k][u], ry);
                }
                dx[k] = rx; dy[k] = ry;
                if (live[k]) *reinterpret_cast<float2*>(pred_rel + ((int64_t)t * batch + p[k]) * 2) = make_float2(rx, ry);
            }
        }
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!live[k]) continue;
            if (h_final) {
#pragma unroll
                for (int u = 0; u < H; ++u) h_final[(int64_t)p[k] * H + u] = h[k][u];
            }
            if (c_final) {
#pragma unroll
                for (int u = 0; u < H; ++u) c_final[(int64_t)p[k] * H + u] = s.c[(k * H + u) * THREADS + tid];
            }
        }
    }
}

template <int H>
static int launch_encoder(const float* obs_rel, int T, int64_t batch, const float* We, const float* be,
                          const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, int E,
                          float* h_out, cudaStream_t st) {
    constexpr int THREADS = 128;
    auto kern = lstm_encoder_kernel<H, THREADS>;
    const int smem = (int)sizeof(LstmSmem<H, THREADS>);
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned grid = (unsigned)std::min<int64_t>((batch + THREADS * PPT - 1) / (THREADS * PPT), 148 * 4);
    kern<<<grid, THREADS, smem, st>>>(obs_rel, T, (int)batch, We, be, W_ih, W_hh, b_ih, b_hh, E, h_out);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

template <int H>
static int launch_decoder(const float* h0, const float* c0, const float* last_pos_rel, const float* z,
                          const int32_t* ped_scene, int nz, int steps, int64_t batch,
                          const float* We, const float* be, const float* W_ih, const float* W_hh, const float* b_ih,
                          const float* b_hh, const float* W_hp, const float* b_hp, int E, float* pred_rel,
                          float* h_final, float* c_final, cudaStream_t st) {
    constexpr int THREADS = 128;
    auto kern = lstm_decoder_kernel<H, THREADS>;
    const int smem = (int)sizeof(LstmSmem<H, THREADS>);
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned grid = (unsigned)std::min<int64_t>((batch + THREADS * PPT - 1) / (THREADS * PPT), 148 * 4);
    kern<<<grid, THREADS, smem, st>>>(h0, c0, last_pos_rel, z, ped_scene, nz, steps, (int)batch, We, be, W_ih, W_hh, b_ih, b_hh, W_hp,
                                      b_hp, E, pred_rel, h_final, c_final);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

}  // namespace sgx

using namespace sgx;

// tensor-core variant (sgx_lstm_tc.cu): H = 32, large batches
int64_t sgx_lstm_tc_ws_bytes();
int sgx_lstm_tc_run(bool decoder, const float* seq_in, const float* h0, const float* c0, const float* z,
                    const int32_t* ped_scene, int nz, int T, int64_t batch, const float* We, const float* be,
                    const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const float* W_hp,
                    const float* b_hp, int E, float* seq_out, float* h_out, void* ws, cudaStream_t st);

static bool use_tc(int H, int T, int64_t batch, const void* ws, int64_t ws_bytes) {
    const char* off = getenv("SGX_LSTM_TC");
    if (off && off[0] == '0') return false;
    return H == 32 && T >= 2 && batch >= 8192 && ws != nullptr && ws_bytes >= sgx_lstm_tc_ws_bytes();
}

extern "C" int64_t sgx_lstm_ws_bytes(void) { return sgx_lstm_tc_ws_bytes(); }

extern "C" int sgx_lstm_encoder_fwd(const float* obs_rel, int32_t T, int64_t batch, const float* We, const float* be,
                                    const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh,
                                    int32_t E, int32_t H, float* h_out, void* workspace, int64_t ws_bytes,
                                    void* stream) {
    SGX_REQUIRE(obs_rel && We && be && W_ih && W_hh && b_ih && b_hh && h_out, "sgx_lstm_encoder_fwd: null pointer");
    SGX_REQUIRE(T >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && E >= 1, "sgx_lstm_encoder_fwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (use_tc(H, T, batch, workspace, ws_bytes))
        return sgx_lstm_tc_run(false, obs_rel, nullptr, nullptr, nullptr, nullptr, 0, T, batch, We, be, W_ih, W_hh, b_ih,
                               b_hh, nullptr, nullptr, E, nullptr, h_out, workspace, st);
    if (H == 32) return launch_encoder<32>(obs_rel, T, batch, We, be, W_ih, W_hh, b_ih, b_hh, E, h_out, st);
    if (H == 48) return launch_encoder<48>(obs_rel, T, batch, We, be, W_ih, W_hh, b_ih, b_hh, E, h_out, st);
    if (H == 64) return launch_encoder<64>(obs_rel, T, batch, We, be, W_ih, W_hh, b_ih, b_hh, E, h_out, st);
    sgx::set_error("fused LSTM is built for h_dim in {32, 48, 64}; got %d", H);
    return SGX_ERR_UNSUPPORTED;
}

extern "C" int sgx_lstm_decoder_fwd(const float* h0, const float* c0, const float* last_pos_rel, const float* z,
                                    const int32_t* ped_scene, int32_t nz, int32_t steps, int64_t batch, const float* We, const float* be, const float* W_ih,
                                    const float* W_hh, const float* b_ih, const float* b_hh, const float* W_hp,
                                    const float* b_hp, int32_t E, int32_t H, float* pred_rel, float* h_final,
                                    float* c_final, void* workspace, int64_t ws_bytes, void* stream) {
    SGX_REQUIRE(h0 && last_pos_rel && We && be && W_ih && W_hh && b_ih && b_hh && W_hp && b_hp && pred_rel,
                "sgx_lstm_decoder_fwd: null pointer");
    SGX_REQUIRE(steps >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && E >= 1, "sgx_lstm_decoder_fwd: bad shape");
    SGX_REQUIRE(nz == 0 || (z && ped_scene && nz > 0 && nz < H), "sgx_lstm_decoder_fwd: noise needs z, ped_scene, 0 < nz < H");
    cudaStream_t st = (cudaStream_t)stream;
    if (!c_final && use_tc(H, steps, batch, workspace, ws_bytes))
        return sgx_lstm_tc_run(true, last_pos_rel, h0, c0, z, ped_scene, nz, steps, batch, We, be, W_ih, W_hh, b_ih, b_hh,
                               W_hp, b_hp, E, pred_rel, h_final, workspace, st);
    if (H == 32)
        return launch_decoder<32>(h0, c0, last_pos_rel, z, ped_scene, nz, steps, batch, We, be, W_ih, W_hh, b_ih, b_hh, W_hp, b_hp, E,
                                  pred_rel, h_final, c_final, st);
    if (H == 48)
        return launch_decoder<48>(h0, c0, last_pos_rel, z, ped_scene, nz, steps, batch, We, be, W_ih, W_hh, b_ih, b_hh, W_hp, b_hp, E,
                                  pred_rel, h_final, c_final, st);
    if (H == 64)
        return launch_decoder<64>(h0, c0, last_pos_rel, z, ped_scene, nz, steps, batch, We, be, W_ih, W_hh, b_ih, b_hh, W_hp, b_hp, E,
                                  pred_rel, h_final, c_final, st);
    sgx::set_error("fused LSTM is built for h_dim in {32, 48, 64}; got %d", H);
    return SGX_ERR_UNSUPPORTED;
}
