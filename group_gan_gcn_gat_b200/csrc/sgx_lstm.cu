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
// (Tried for small batches and dropped, measured: four lanes per pedestrian, one gate each, pre-activations exchanged by
// shuffles -- 55 % slower on a 5 k-pedestrian batch than this thread-per-pedestrian form.)
// Training: the same kernels with SAVE = true write a tape; lstm_bwd_kernel walks it backwards (one thread per
// pedestrian) and three tall-skinny GEMMs reduce the parameter gradients (sgx_lstm_*_train_fwd / sgx_lstm_bwd).
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

// Training tape (backward): everything the backward kernel and the parameter-gradient GEMMs read,
// unit-major [rows][T][batch] so that thread-per-pedestrian accesses coalesce and the merged (t, ped) index is the
// contiguous K dimension of the GEMMs.
struct LstmTape {
    float* gact;     // [4H][T][B]  i, f, g, o after their non-linearities
    float* cprev;    // [H][T][B]   cell state entering step t
    float* tanhc;    // [H][T][B]   tanh of the cell state leaving step t
    float* hprev;    // [H][T][B]   hidden state entering step t
    float* hout;     // [H+1][T][B] hidden state leaving step t, last row = 1 (bias column of the hidden2pos gradient)
    float* xaug;     // [3][T][B]   step input (dx, dy, 1)
    int T;
    int64_t B;
};

// one LSTM cell update for this thread's PPT pedestrians; input of ped k is the 2-vector (dx[k], dy[k])
template <int H, int THREADS, bool SAVE = false>
__device__ __forceinline__ void lstm_cell(LstmSmem<H, THREADS>& s, float (&h)[PPT][H], const float (&dx)[PPT],
                                          const float (&dy)[PPT], const LstmTape* tape = nullptr, int t = 0,
                                          const int* p = nullptr, const bool* live = nullptr) {
    const int tid = threadIdx.x;
    if (SAVE) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!live[k]) continue;
            const int64_t at = (int64_t)t * tape->B + p[k], plane = (int64_t)tape->T * tape->B;
#pragma unroll
            for (int u = 0; u < H; ++u) tape->hprev[u * plane + at] = h[k][u];
            tape->xaug[at] = dx[k]; tape->xaug[plane + at] = dy[k]; tape->xaug[2 * plane + at] = 1.f;
        }
    }
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
            if (SAVE) {
                const float gi = sigmoid_f(g[k][0]), gf = sigmoid_f(g[k][1]), gg = tanh_f(g[k][2]), go = sigmoid_f(g[k][3]);
                const float cp = s.c[col];
                const float cn = gf * cp + gi * gg;
                const float tc = tanh_f(cn);
                s.c[col] = cn;
                s.hn[col] = go * tc;
                if (live[k]) {
                    const int64_t at = (int64_t)t * tape->B + p[k], plane = (int64_t)tape->T * tape->B;
                    tape->gact[(0 * H + u) * plane + at] = gi;
                    tape->gact[(1 * H + u) * plane + at] = gf;
                    tape->gact[(2 * H + u) * plane + at] = gg;
                    tape->gact[(3 * H + u) * plane + at] = go;
                    tape->cprev[u * plane + at] = cp;
                    tape->tanhc[u * plane + at] = tc;
                    tape->hout[u * plane + at] = go * tc;
                }
            } else {
                const float cn = sigmoid_f(g[k][1]) * s.c[col] + sigmoid_f(g[k][0]) * tanh_f(g[k][2]);
                s.c[col] = cn;
                s.hn[col] = sigmoid_f(g[k][3]) * tanh_f(cn);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < PPT; ++k)
#pragma unroll
        for (int u = 0; u < H; ++u) h[k][u] = s.hn[(k * H + u) * THREADS + tid];
    if (SAVE) {
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (live[k]) tape->hout[H * (int64_t)tape->T * tape->B + (int64_t)t * tape->B + p[k]] = 1.f;
    }
}

template <int H, int THREADS, bool SAVE>
__global__ void __launch_bounds__(THREADS)
lstm_encoder_kernel(const float* __restrict__ obs_rel, int T, int batch, const float* __restrict__ We,
                    const float* __restrict__ be, const float* __restrict__ W_ih, const float* __restrict__ W_hh,
                    const float* __restrict__ b_ih, const float* __restrict__ b_hh, int E, float* __restrict__ h_out,
                    LstmTape tape) {
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
            lstm_cell<H, THREADS, SAVE>(s, h, dx, dy, &tape, t, p, live);
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

template <int H, int THREADS, bool SAVE>
__global__ void __launch_bounds__(THREADS)
lstm_decoder_kernel(const float* __restrict__ h0, const float* __restrict__ c0, const float* __restrict__ last_pos_rel,
                    const float* __restrict__ z, const int32_t* __restrict__ ped_scene, int nz, int steps, int batch, const float* __restrict__ We, const float* __restrict__ be,
                    const float* __restrict__ W_ih, const float* __restrict__ W_hh, const float* __restrict__ b_ih,
                    const float* __restrict__ b_hh, const float* __restrict__ W_hp, const float* __restrict__ b_hp, int E,
                    float* __restrict__ pred_rel, float* __restrict__ h_final, float* __restrict__ c_final,
                    LstmTape tape) {
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
            lstm_cell<H, THREADS, SAVE>(s, h, dx, dy, &tape, t, p, live);
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

template <int H, bool SAVE = false>
static int launch_encoder(const float* obs_rel, int T, int64_t batch, const float* We, const float* be,
                          const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, int E,
                          float* h_out, cudaStream_t st, LstmTape tape = LstmTape()) {
    constexpr int THREADS = 128;
    auto kern = lstm_encoder_kernel<H, THREADS, SAVE>;
    const int smem = (int)sizeof(LstmSmem<H, THREADS>);
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned grid = (unsigned)std::min<int64_t>((batch + THREADS * PPT - 1) / (THREADS * PPT), 148 * 4);
    kern<<<grid, THREADS, smem, st>>>(obs_rel, T, (int)batch, We, be, W_ih, W_hh, b_ih, b_hh, E, h_out, tape);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

template <int H, bool SAVE = false>
static int launch_decoder(const float* h0, const float* c0, const float* last_pos_rel, const float* z,
                          const int32_t* ped_scene, int nz, int steps, int64_t batch,
                          const float* We, const float* be, const float* W_ih, const float* W_hh, const float* b_ih,
                          const float* b_hh, const float* W_hp, const float* b_hp, int E, float* pred_rel,
                          float* h_final, float* c_final, cudaStream_t st, LstmTape tape = LstmTape()) {
    constexpr int THREADS = 128;
    auto kern = lstm_decoder_kernel<H, THREADS, SAVE>;
    const int smem = (int)sizeof(LstmSmem<H, THREADS>);
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    unsigned grid = (unsigned)std::min<int64_t>((batch + THREADS * PPT - 1) / (THREADS * PPT), 148 * 4);
    kern<<<grid, THREADS, smem, st>>>(h0, c0, last_pos_rel, z, ped_scene, nz, steps, (int)batch, We, be, W_ih, W_hh, b_ih, b_hh, W_hp,
                                      b_hp, E, pred_rel, h_final, c_final, tape);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

// ------------------------------------------------------------------------------------------------
// Backward through the whole recurrence in one kernel (thread per pedestrian, t = T-1 .. 0), reading the tape.
//   d(pre-activation gates) -> dG [4H][T][B] for the parameter-gradient GEMMs
//   decoder: the step input is hidden2pos of the previous step, so d(rel_t) = d_pred_rel[t] + Wx^T dG_{t+1} and
//            dh_t += W_hp^T d(rel_t); d(rel_t) is kept in dRel [2][T][B]; dh_0's carry is d_h0.
//   encoder: d(seq_in[t]) = Wx^T dG_t (needed when the sequence is a generator output fed to the discriminator).
// dh lives in the hn columns of shared memory, dc in the c columns, W_hh^T dG accumulates in registers.
// ------------------------------------------------------------------------------------------------
template <int H, int THREADS, bool DECODER>
__global__ void __launch_bounds__(THREADS)
lstm_bwd_kernel(LstmTape tape, int T, int batch, const float* __restrict__ We, const float* __restrict__ be,
                const float* __restrict__ W_ih, const float* __restrict__ W_hh, const float* __restrict__ b_ih,
                const float* __restrict__ b_hh, const float* __restrict__ W_hp, int E,
                const float* __restrict__ d_seq_out, const float* __restrict__ d_h_last, float* __restrict__ d_seq_in,
                float* __restrict__ d_h0, float* __restrict__ d_c0, float* __restrict__ dG, float* __restrict__ dRel) {
    extern __shared__ __align__(16) uint8_t raw[];
    LstmSmem<H, THREADS>& s = *reinterpret_cast<LstmSmem<H, THREADS>*>(raw);
    lstm_load_weights<H, THREADS>(s, We, be, W_ih, W_hh, b_ih, b_hh, E);
    if (DECODER) for (int e = threadIdx.x; e < 2 * H; e += THREADS) s.whp[e] = W_hp[e];
    __syncthreads();
    const int tid = threadIdx.x;
    const int64_t plane = (int64_t)T * batch;
    for (int p = blockIdx.x * THREADS + tid; p < batch; p += gridDim.x * THREADS) {
#pragma unroll 4
        for (int u = 0; u < H; ++u) {
            s.hn[u * THREADS + tid] = d_h_last ? d_h_last[(int64_t)p * H + u] : 0.f;      // dh
            s.c[u * THREADS + tid] = 0.f;                                                   // dc
        }
        float cx = 0.f, cy = 0.f;                        // Wx^T dG of the step after this one
        for (int t = T - 1; t >= 0; --t) {
            const int64_t at = (int64_t)t * batch + p;
            if (DECODER) {
                const float2 go = *reinterpret_cast<const float2*>(d_seq_out + at * 2);
                const float rx = go.x + cx, ry = go.y + cy;
                dRel[at] = rx;
                dRel[plane + at] = ry;
#pragma unroll 4
                for (int u = 0; u < H; ++u)
                    s.hn[u * THREADS + tid] += s.whp[u] * rx + s.whp[H + u] * ry;
            }
            float dhp[H];
#pragma unroll
            for (int j = 0; j < H; ++j) dhp[j] = 0.f;
            float dx = 0.f, dy = 0.f;
#pragma unroll 1
            for (int u = 0; u < H; ++u) {
                const float gi = tape.gact[(0 * H + u) * plane + at], gf = tape.gact[(1 * H + u) * plane + at];
                const float gg = tape.gact[(2 * H + u) * plane + at], go = tape.gact[(3 * H + u) * plane + at];
                const float cp = tape.cprev[u * plane + at], tc = tape.tanhc[u * plane + at];
                const float dh = s.hn[u * THREADS + tid];
                const float dc = s.c[u * THREADS + tid] + dh * go * (1.f - tc * tc);
                float d[4];
                d[0] = dc * gg * gi * (1.f - gi);
                d[1] = dc * cp * gf * (1.f - gf);
                d[2] = dc * gi * (1.f - gg * gg);
                d[3] = dh * tc * go * (1.f - go);
                s.c[u * THREADS + tid] = dc * gf;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = q * H + u;
                    dG[r * plane + at] = d[q];
                    const float4 wi = s.wxb[r];
                    dx = fmaf(wi.x, d[q], dx);
                    dy = fmaf(wi.y, d[q], dy);
                    const float4* w = reinterpret_cast<const float4*>(&s.whh[r * H]);
#pragma unroll
                    for (int j = 0; j < H / 4; ++j) {
                        const float4 v = w[j];
                        dhp[4 * j] = fmaf(v.x, d[q], dhp[4 * j]);
                        dhp[4 * j + 1] = fmaf(v.y, d[q], dhp[4 * j + 1]);
                        dhp[4 * j + 2] = fmaf(v.z, d[q], dhp[4 * j + 2]);
                        dhp[4 * j + 3] = fmaf(v.w, d[q], dhp[4 * j + 3]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < H; ++j) s.hn[j * THREADS + tid] = dhp[j];
            if (DECODER) { cx = dx; cy = dy; }
            else if (d_seq_in) *reinterpret_cast<float2*>(d_seq_in + at * 2) = make_float2(dx, dy);
        }
        if (d_h0) {
#pragma unroll 4
            for (int u = 0; u < H; ++u) d_h0[(int64_t)p * H + u] = s.hn[u * THREADS + tid];
        }
        if (d_c0) {
#pragma unroll 4
            for (int u = 0; u < H; ++u) d_c0[(int64_t)p * H + u] = s.c[u * THREADS + tid];
        }
    }
}

static LstmTape carve_tape(float* base, int T, int64_t B, int H) {
    LstmTape t;
    const int64_t plane = (int64_t)T * B;
    t.gact = base;
    t.cprev = t.gact + 4 * H * plane;
    t.tanhc = t.cprev + H * plane;
    t.hprev = t.tanhc + H * plane;
    t.hout = t.hprev + H * plane;
    t.xaug = t.hout + (H + 1) * plane;
    t.T = T;
    t.B = B;
    return t;
}

template <int H>
static int lstm_backward(bool decoder, LstmTape tape, int T, int64_t batch, const float* We, const float* be,
                         const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const float* W_hp,
                         int E, const float* d_seq_out, const float* d_h_last, float* d_seq_in, float* d_h0,
                         float* d_c0, float* dW_hh, float* dS, float* dW_hp_aug, float* ws, cudaStream_t st) {
    constexpr int THREADS = 128;
    const int64_t plane = (int64_t)T * batch;
    float* dG = ws;
    float* dRel = ws + 4 * H * plane;
    const int smem = (int)sizeof(LstmSmem<H, THREADS>);
    const unsigned grid = (unsigned)std::min<int64_t>((batch + THREADS - 1) / THREADS, 148 * 4);
    if (decoder) {
        auto kern = lstm_bwd_kernel<H, THREADS, true>;
        SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, THREADS, smem, st>>>(tape, T, (int)batch, We, be, W_ih, W_hh, b_ih, b_hh, W_hp, E, d_seq_out, d_h_last,
                                          d_seq_in, d_h0, d_c0, dG, dRel);
    } else {
        auto kern = lstm_bwd_kernel<H, THREADS, false>;
        SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, THREADS, smem, st>>>(tape, T, (int)batch, We, be, W_ih, W_hh, b_ih, b_hh, W_hp, E, d_seq_out, d_h_last,
                                          d_seq_in, d_h0, d_c0, dG, dRel);
    }
    SGX_LAUNCH_CHECK();
    int rc;
    // dW_hh [4H,H] = dG [4H, T*B] . hprev [H, T*B]^T ;  dS [4H,3] = dG . (dx, dy, 1)^T  (input weights, embedding, biases)
    if ((rc = gemm(dG, plane, 1, tape.hprev, 1, plane, dW_hh, H, 4 * H, H, plane, 0, 0, st))) return rc;
    if ((rc = gemm(dG, plane, 1, tape.xaug, 1, plane, dS, 3, 4 * H, 3, plane, 0, 0, st))) return rc;
    if (decoder)   // [dW_hp | db_hp] [2, H+1] = dRel [2, T*B] . (hout ; 1) [H+1, T*B]^T
        if ((rc = gemm(dRel, plane, 1, tape.hout, 1, plane, dW_hp_aug, H + 1, 2, H + 1, plane, 0, 0, st))) return rc;
    return SGX_OK;
}

}  // namespace sgx

using namespace sgx;

// tensor-core variant (sgx_lstm_tc.cu): H = 32, large batches
int64_t sgx_lstm_tc_ws_bytes();
int sgx_lstm_tc_run(bool decoder, const float* seq_in, const float* h0, const float* c0, const float* z,
                    const int32_t* ped_scene, int nz, int T, int64_t batch, const float* We, const float* be,
                    const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const float* W_hp,
                    const float* b_hp, int E, float* seq_out, float* h_out, void* ws, cudaStream_t st, bool ws_prepared);

static bool use_tc(int H, int T, int64_t batch, const void* ws, int64_t ws_bytes) {
    if (!opt_lstm_tc()) return false;      // sgx_set_option("lstm_tc", 0): CUDA-core kernels for every batch (parity tests)
    return H == 32 && T >= 2 && batch >= 8192 && ws != nullptr && ws_bytes >= sgx_lstm_tc_ws_bytes();
}

extern "C" int64_t sgx_lstm_ws_bytes(void) { return sgx_lstm_tc_ws_bytes(); }

// Fills a workspace with the tensor-core weight images of one recurrence (embedding folded into the input weights, gate
// rows pre-scaled, 3-way bf16 splits): pass it to sgx_lstm_encoder_fwd / _decoder_fwd with ws_prepared = 1 for as long as
// these weights do not change, and the per-call prep launch (4.7 us) disappears.  A no-op for h_dim != 32.
extern "C" int sgx_lstm_prep(const float* We, const float* be, const float* W_ih, const float* W_hh, const float* b_ih,
                             const float* b_hh, int32_t E, int32_t H, void* workspace, int64_t ws_bytes, void* stream) {
    SGX_REQUIRE(We && be && W_ih && W_hh && b_ih && b_hh && workspace, "sgx_lstm_prep: null pointer");
    SGX_REQUIRE(ws_bytes >= sgx_lstm_tc_ws_bytes(), "sgx_lstm_prep: workspace too small");
    if (H != 32) return SGX_OK;
    return sgx_lstm_tc_run(false, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, We, be, W_ih, W_hh, b_ih, b_hh, nullptr,
                           nullptr, E, nullptr, nullptr, workspace, (cudaStream_t)stream, false);
}

extern "C" int sgx_lstm_encoder_fwd(const float* obs_rel, int32_t T, int64_t batch, const float* We, const float* be,
                                    const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh,
                                    int32_t E, int32_t H, float* h_out, void* workspace, int64_t ws_bytes,
                                    int32_t ws_prepared, void* stream) {
    SGX_REQUIRE(obs_rel && We && be && W_ih && W_hh && b_ih && b_hh && h_out, "sgx_lstm_encoder_fwd: null pointer");
    SGX_REQUIRE(T >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && E >= 1, "sgx_lstm_encoder_fwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    if (use_tc(H, T, batch, workspace, ws_bytes))
        return sgx_lstm_tc_run(false, obs_rel, nullptr, nullptr, nullptr, nullptr, 0, T, batch, We, be, W_ih, W_hh, b_ih,
                               b_hh, nullptr, nullptr, E, nullptr, h_out, workspace, st, ws_prepared != 0);
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
                                    float* c_final, void* workspace, int64_t ws_bytes, int32_t ws_prepared, void* stream) {
    SGX_REQUIRE(h0 && last_pos_rel && We && be && W_ih && W_hh && b_ih && b_hh && W_hp && b_hp && pred_rel,
                "sgx_lstm_decoder_fwd: null pointer");
    SGX_REQUIRE(steps >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && E >= 1, "sgx_lstm_decoder_fwd: bad shape");
    SGX_REQUIRE(nz == 0 || (z && ped_scene && nz > 0 && nz < H), "sgx_lstm_decoder_fwd: noise needs z, ped_scene, 0 < nz < H");
    cudaStream_t st = (cudaStream_t)stream;
    if (!c_final && use_tc(H, steps, batch, workspace, ws_bytes))
        return sgx_lstm_tc_run(true, last_pos_rel, h0, c0, z, ped_scene, nz, steps, batch, We, be, W_ih, W_hh, b_ih, b_hh,
                               W_hp, b_hp, E, pred_rel, h_final, workspace, st, ws_prepared != 0);
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

// ---- training path (forward with tape, backward) -------------------------------------------------
extern "C" int64_t sgx_lstm_tape_floats(int32_t T, int64_t batch, int32_t H) {
    return ((int64_t)4 * H + 3 * (int64_t)H + (H + 1) + 3) * T * batch;
}
extern "C" int64_t sgx_lstm_bwd_ws_bytes(int32_t T, int64_t batch, int32_t H) {
    return ((int64_t)4 * H + 2) * T * batch * (int64_t)sizeof(float);
}

#define SGX_LSTM_DISPATCH_H(H, CALL)                                                         \
    if (H == 32) { constexpr int HH = 32; return CALL; }                                     \
    if (H == 48) { constexpr int HH = 48; return CALL; }                                     \
    if (H == 64) { constexpr int HH = 64; return CALL; }                                     \
    sgx::set_error("fused LSTM is built for h_dim in {32, 48, 64}; got %d", H);             \
    return SGX_ERR_UNSUPPORTED;

extern "C" int sgx_lstm_encoder_train_fwd(const float* seq_in, int32_t T, int64_t batch, const float* We,
                                          const float* be, const float* W_ih, const float* W_hh, const float* b_ih,
                                          const float* b_hh, int32_t E, int32_t H, float* h_out, float* tape,
                                          void* stream) {
    SGX_REQUIRE(seq_in && We && be && W_ih && W_hh && b_ih && b_hh && h_out && tape,
                "sgx_lstm_encoder_train_fwd: null pointer");
    SGX_REQUIRE(T >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && E >= 1, "sgx_lstm_encoder_train_fwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    SGX_LSTM_DISPATCH_H(H, (launch_encoder<HH, true>(seq_in, T, batch, We, be, W_ih, W_hh, b_ih, b_hh, E, h_out, st,
                                                      carve_tape(tape, T, batch, HH))))
}

extern "C" int sgx_lstm_decoder_train_fwd(const float* h0, const float* c0, const float* last_pos_rel, int32_t steps,
                                          int64_t batch, const float* We, const float* be, const float* W_ih,
                                          const float* W_hh, const float* b_ih, const float* b_hh, const float* W_hp,
                                          const float* b_hp, int32_t E, int32_t H, float* pred_rel, float* h_final,
                                          float* tape, void* stream) {
    SGX_REQUIRE(h0 && last_pos_rel && We && be && W_ih && W_hh && b_ih && b_hh && W_hp && b_hp && pred_rel && tape,
                "sgx_lstm_decoder_train_fwd: null pointer");
    SGX_REQUIRE(steps >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && E >= 1, "sgx_lstm_decoder_train_fwd: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    SGX_LSTM_DISPATCH_H(H, (launch_decoder<HH, true>(h0, c0, last_pos_rel, nullptr, nullptr, 0, steps, batch, We, be, W_ih,
                                                      W_hh, b_ih, b_hh, W_hp, b_hp, E, pred_rel, h_final, nullptr, st,
                                                      carve_tape(tape, steps, batch, HH))))
}

// decoder != 0: d_seq_out = d(pred_rel) [T,B,2] (required), d_h0 / d_c0 [B,H] receive d(h0) / d(c0), dW_hp_aug [2,H+1] = [dW_hp | db_hp].
// decoder == 0: d_seq_in [T,B,2] (nullable) receives d(seq_in); d_seq_out / W_hp / dW_hp_aug unused.
// dS [4H,3] = dG . (x, y, 1)^T: the caller forms dW_ih = dS[:, :2] We^T + dS[:, 2] be^T, dWe = W_ih^T dS[:, :2],
// dbe = W_ih^T dS[:, 2], db_ih = db_hh = dS[:, 2]  (the embedding is folded into the input weights in the kernel).
extern "C" int sgx_lstm_bwd(int32_t decoder, const float* tape, int32_t T, int64_t batch, const float* We,
                            const float* be, const float* W_ih, const float* W_hh, const float* b_ih,
                            const float* b_hh, const float* W_hp, int32_t E, int32_t H, const float* d_seq_out,
                            const float* d_h_last, float* d_seq_in, float* d_h0, float* d_c0, float* dW_hh,
                            float* dS, float* dW_hp_aug, void* workspace, int64_t ws_bytes, void* stream) {
    SGX_REQUIRE(tape && We && be && W_ih && W_hh && b_ih && b_hh && dW_hh && dS && workspace, "sgx_lstm_bwd: null pointer");
    SGX_REQUIRE(!decoder || (W_hp && d_seq_out && dW_hp_aug), "sgx_lstm_bwd: decoder needs W_hp, d_seq_out, dW_hp_aug");
    SGX_REQUIRE(T >= 1 && batch >= 1 && batch < ((int64_t)1 << 31) && E >= 1, "sgx_lstm_bwd: bad shape");
    SGX_REQUIRE(ws_bytes >= sgx_lstm_bwd_ws_bytes(T, batch, H), "sgx_lstm_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    SGX_LSTM_DISPATCH_H(H, (lstm_backward<HH>(decoder != 0, carve_tape(const_cast<float*>(tape), T, batch, HH), T, batch, We,
                                               be, W_ih, W_hh, b_ih, b_hh, W_hp, E, d_seq_out, d_h_last, d_seq_in, d_h0,
                                               d_c0, dW_hh, dS, dW_hp_aug, (float*)workspace, st)))
}
