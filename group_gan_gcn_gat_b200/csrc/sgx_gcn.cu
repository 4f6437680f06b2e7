// GCNModule forward/backward (sgan/models.py:628-712, GCN.forward 573-580), gcn_layers = 2.
//
// The reference multiplies dense N x N (row-normalised) adjacencies: A_intra has fl(1/|g|) on the
// members of the row's group, A_inter = 1/G everywhere.  A.H is therefore a *segmented mean*:
// every member of a group sees the same aggregated row, every group of a scene sees the same
// inter-group row.  The kernels below work on that structure directly (segmented SpMM):
//
//   gcn_group_kernel  (thread per ped, leaders work):  M1 = sum_{j in g} a X_j ; H1 = relu(M1 W0)
//                      M2 = sum_{j in g} a H1 ; X1 = relu(M2 W1) ; Xg = sum_{j in g} a X1
//   gcn_scene_kernel  (thread per scene):              N1 = sum_g c Xg ; K1 = relu(N1 V0)
//                      N2 = sum_g c K1 ; Y = relu(N2 V1)
//   gcn_out_kernel    (thread per ped):                out = Wo [X1_g ; a Y] + bo
//
// Sums over identical rows are performed term by term (k additions of a*v) so the arithmetic stays
// as close to the dense matmul as a different summation order allows.  Backward reuses the same three
// levels in reverse; parameter gradients are tall-skinny GEMMs over per-group / per-scene rows.
#include <stdlib.h>

#include "sgx_common.cuh"
#include "sgx_warp_mma.cuh"

namespace sgx {

template <int NI, int NO>
__device__ __forceinline__ void matvec(const float* __restrict__ sW /*[NO][NI] in smem*/, const float (&x)[NI],
                                       float (&y)[NO]) {
#pragma unroll
    for (int o = 0; o < NO; ++o) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < NI; c += 4) {
            float4 w = *reinterpret_cast<const float4*>(sW + o * NI + c);
            s = fmaf(x[c], w.x, s); s = fmaf(x[c + 1], w.y, s); s = fmaf(x[c + 2], w.z, s); s = fmaf(x[c + 3], w.w, s);
        }
        y[o] = s;
    }
}

__device__ __forceinline__ float repeat_sum(float term, int k) {
    float s = 0.f;
    for (int t = 0; t < k; ++t) s += term;
    return s;
}

// load a [R][C] row-major global matrix into smem, optionally transposed to [C][R]
__device__ __forceinline__ void load_w(float* dst, const float* __restrict__ src, int R, int C, bool transpose) {
    for (int e = threadIdx.x; e < R * C; e += blockDim.x) {
        int r = e / C, c = e % C;
        dst[transpose ? c * R + r : e] = src[e];
    }
}

template <int IN, int HID, int OUT>
__global__ void __launch_bounds__(128)
gcn_group_kernel(const float* __restrict__ x, const int32_t* __restrict__ leader, const int32_t* __restrict__ gsize,
                 const int32_t* __restrict__ ped_start, const int32_t* __restrict__ ped_end, int batch,
                 const float* __restrict__ W0, const float* __restrict__ W1, float* __restrict__ X1g,
                 float* __restrict__ Xg, float* __restrict__ M1s, float* __restrict__ M2s) {
    __shared__ __align__(16) float sW0t[HID * IN];   // [HID][IN]
    __shared__ __align__(16) float sW1[HID * OUT];   // [HID][OUT]
    load_w(sW0t, W0, IN, HID, true);
    load_w(sW1, W1, HID, OUT, false);
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const bool lead = (leader[p] == p);
    float x1[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) x1[o] = 0.f;
    if (lead) {
        const int k = gsize[p];
        const float a = __frcp_rn((float)k);
        float m1[IN];
#pragma unroll
        for (int c = 0; c < IN; ++c) m1[c] = 0.f;
        const int e = ped_end[p];
        for (int q = p; q < e; ++q) {
            if (leader[q] != p) continue;
            const float4* row = reinterpret_cast<const float4*>(x + (int64_t)q * IN);
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) {
                float4 v = row[c];
                m1[4 * c] = fmaf(a, v.x, m1[4 * c]); m1[4 * c + 1] = fmaf(a, v.y, m1[4 * c + 1]);
                m1[4 * c + 2] = fmaf(a, v.z, m1[4 * c + 2]); m1[4 * c + 3] = fmaf(a, v.w, m1[4 * c + 3]);
            }
        }
        if (M1s) {
#pragma unroll
            for (int c = 0; c < IN; ++c) M1s[(int64_t)p * IN + c] = m1[c];
        }
        for (int f = 0; f < HID; ++f) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < IN; c += 4) {
                float4 w = *reinterpret_cast<const float4*>(sW0t + f * IN + c);
                s = fmaf(m1[c], w.x, s); s = fmaf(m1[c + 1], w.y, s); s = fmaf(m1[c + 2], w.z, s); s = fmaf(m1[c + 3], w.w, s);
            }
            const float h1 = fmaxf(s, 0.f);
            const float m2 = repeat_sum(a * h1, k);
            if (M2s) M2s[(int64_t)p * HID + f] = m2;
#pragma unroll
            for (int o = 0; o < OUT; o += 4) {
                float4 w = *reinterpret_cast<const float4*>(sW1 + f * OUT + o);
                x1[o] = fmaf(m2, w.x, x1[o]); x1[o + 1] = fmaf(m2, w.y, x1[o + 1]);
                x1[o + 2] = fmaf(m2, w.z, x1[o + 2]); x1[o + 3] = fmaf(m2, w.w, x1[o + 3]);
            }
        }
#pragma unroll
        for (int o = 0; o < OUT; ++o) {
            x1[o] = fmaxf(x1[o], 0.f);
            X1g[(int64_t)p * OUT + o] = x1[o];
            Xg[(int64_t)p * OUT + o] = repeat_sum(a * x1[o], k);
        }
    } else if (M1s) {  // backward consumes these as GEMM operands: non-leader rows must be zero
#pragma unroll
        for (int c = 0; c < IN; ++c) M1s[(int64_t)p * IN + c] = 0.f;
        for (int f = 0; f < HID; ++f) M2s[(int64_t)p * HID + f] = 0.f;
    }
}

// per scene: rows of Yrow / N1s / N2s / K1s are indexed by the scene's first pedestrian
template <int HID, int OUT>
__global__ void __launch_bounds__(128)
gcn_scene_kernel(const float* __restrict__ Xg, const int32_t* __restrict__ leader,
                 const int32_t* __restrict__ scene_start, const int32_t* __restrict__ n_group, int n_scenes,
                 const float* __restrict__ V0, const float* __restrict__ V1, float* __restrict__ Yrow,
                 float* __restrict__ N1s, float* __restrict__ N2s, float* __restrict__ K1s) {
    __shared__ __align__(16) float sV0t[HID * OUT];  // [HID][OUT]
    __shared__ __align__(16) float sV1[HID * OUT];   // [HID][OUT]
    load_w(sV0t, V0, OUT, HID, true);
    load_w(sV1, V1, HID, OUT, false);
    __syncthreads();
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scenes) return;
    const int b = scene_start[s], e = scene_start[s + 1], G = n_group[s];
    const float c = __frcp_rn((float)G);
    float n1[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) n1[o] = 0.f;
    for (int q = b; q < e; ++q) {
        if (leader[q] != q) continue;
#pragma unroll
        for (int o = 0; o < OUT; ++o) n1[o] = fmaf(c, Xg[(int64_t)q * OUT + o], n1[o]);
    }
    float y[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) y[o] = 0.f;
    for (int f = 0; f < HID; ++f) {
        float sacc = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; o += 4) {
            float4 w = *reinterpret_cast<const float4*>(sV0t + f * OUT + o);
            sacc = fmaf(n1[o], w.x, sacc); sacc = fmaf(n1[o + 1], w.y, sacc);
            sacc = fmaf(n1[o + 2], w.z, sacc); sacc = fmaf(n1[o + 3], w.w, sacc);
        }
        const float k1 = fmaxf(sacc, 0.f);
        const float n2 = repeat_sum(c * k1, G);
        if (N2s) { N2s[(int64_t)b * HID + f] = n2; K1s[(int64_t)b * HID + f] = k1; }
#pragma unroll
        for (int o = 0; o < OUT; o += 4) {
            float4 w = *reinterpret_cast<const float4*>(sV1 + f * OUT + o);
            y[o] = fmaf(n2, w.x, y[o]); y[o + 1] = fmaf(n2, w.y, y[o + 1]);
            y[o + 2] = fmaf(n2, w.z, y[o + 2]); y[o + 3] = fmaf(n2, w.w, y[o + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        Yrow[(int64_t)b * OUT + o] = fmaxf(y[o], 0.f);
        if (N1s) N1s[(int64_t)b * OUT + o] = n1[o];
    }
}

// The same per-scene computation with one WARP per scene (dense crowds: gcn_scene_kernel's single thread walks N
// pedestrians and G x 72 repeat-sum terms serially -- 0.7 ms for 64 scenes of 1024).  Lanes <-> features; the scene's
// leaders are found 32 candidates at a time (ballot) and accumulated in ascending order, and every dot product keeps
// the thread kernel's term order, so the two kernels give bit-identical results.
template <int HID, int OUT>
__global__ void __launch_bounds__(128)
gcn_scene_warp_kernel(const float* __restrict__ Xg, const int32_t* __restrict__ leader,
                      const int32_t* __restrict__ scene_start, const int32_t* __restrict__ n_group, int n_scenes,
                      const float* __restrict__ V0, const float* __restrict__ V1, float* __restrict__ Yrow,
                      float* __restrict__ N1s, float* __restrict__ N2s, float* __restrict__ K1s) {
    static_assert(OUT == 16 && HID <= 96, "lanes 0..15 <-> output features, <= 3 hidden units per lane");
    __shared__ __align__(16) float sV0t[HID * OUT];  // [HID][OUT]
    __shared__ __align__(16) float sV1[HID * OUT];   // [HID][OUT]
    load_w(sV0t, V0, OUT, HID, true);
    load_w(sV1, V1, HID, OUT, false);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_scenes) return;
    const int b = scene_start[s], e = scene_start[s + 1], G = n_group[s];
    const float c = __frcp_rn((float)G);
    float n1 = 0.f;                                   // lane o < 16: N1[o]
    for (int q0 = b; q0 < e; q0 += 32) {
        const int q = q0 + lane;
        uint32_t m = __ballot_sync(0xffffffffu, q < e && leader[q] == q);
        for (; m; m &= m - 1) {
            const int qq = q0 + __ffs(m) - 1;
            if (lane < OUT) n1 = fmaf(c, Xg[(int64_t)qq * OUT + lane], n1);
        }
    }
    float n1v[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) n1v[o] = __shfl_sync(0xffffffffu, n1, o);
    float n2[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const int f = lane + 32 * u;
        if (f < HID) {
            float sacc = 0.f;
#pragma unroll
            for (int o = 0; o < OUT; o += 4) {
                float4 w = *reinterpret_cast<const float4*>(sV0t + f * OUT + o);
                sacc = fmaf(n1v[o], w.x, sacc); sacc = fmaf(n1v[o + 1], w.y, sacc);
                sacc = fmaf(n1v[o + 2], w.z, sacc); sacc = fmaf(n1v[o + 3], w.w, sacc);
            }
            const float k1 = fmaxf(sacc, 0.f);
            n2[u] = repeat_sum(c * k1, G);
            if (N2s) { N2s[(int64_t)b * HID + f] = n2[u]; K1s[(int64_t)b * HID + f] = k1; }
        }
    }
    float y = 0.f;                                    // lane o < 16: Y[o], terms in ascending f like the thread kernel
#pragma unroll
    for (int f = 0; f < HID; ++f) {
        const float v = __shfl_sync(0xffffffffu, n2[f / 32], f % 32);
        if (lane < OUT) y = fmaf(v, sV1[f * OUT + lane], y);
    }
    if (lane < OUT) {
        Yrow[(int64_t)b * OUT + lane] = fmaxf(y, 0.f);
        if (N1s) N1s[(int64_t)b * OUT + lane] = n1;
    }
}

template <int OUT, int FIN>
__global__ void __launch_bounds__(128)
gcn_out_kernel(const float* __restrict__ X1g, const float* __restrict__ Yrow, const int32_t* __restrict__ leader,
               const int32_t* __restrict__ gsize, const int32_t* __restrict__ ped_start, int batch,
               const float* __restrict__ Wo, const float* __restrict__ bo, float* __restrict__ out,
               float* __restrict__ cat_save) {
    __shared__ __align__(16) float sWo[FIN * 2 * OUT];
    __shared__ float sbo[FIN];
    load_w(sWo, Wo, FIN, 2 * OUT, false);
    for (int e = threadIdx.x; e < FIN; e += blockDim.x) sbo[e] = bo[e];
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const int l = leader[p];
    const float a = __frcp_rn((float)gsize[p]);
    float cat[2 * OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        cat[o] = X1g[(int64_t)l * OUT + o];
        cat[OUT + o] = a * Yrow[(int64_t)ped_start[p] * OUT + o];
    }
    if (cat_save) {
#pragma unroll
        for (int o = 0; o < 2 * OUT; ++o) cat_save[(int64_t)p * 2 * OUT + o] = cat[o];
    }
    if (!out) return;
    float y[FIN];
    matvec<2 * OUT, FIN>(sWo, cat, y);
#pragma unroll
    for (int o = 0; o < FIN; ++o) out[(int64_t)p * FIN + o] = y[o] + sbo[o];
}

// ---- backward -----------------------------------------------------------------------------------
// per ped: dcat = Wo^T dOut   ([dx1 | dx2])
template <int OUT, int FIN>
__global__ void __launch_bounds__(128)
gcn_bwd_out_kernel(const float* __restrict__ gout, int batch, const float* __restrict__ Wo, float* __restrict__ dcat) {
    __shared__ __align__(16) float sWot[2 * OUT * FIN];  // [2*OUT][FIN]
    load_w(sWot, Wo, FIN, 2 * OUT, true);
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    float g[FIN];
#pragma unroll
    for (int o = 0; o < FIN; ++o) g[o] = gout[(int64_t)p * FIN + o];
    float d[2 * OUT];
    matvec<FIN, 2 * OUT>(sWot, g, d);
#pragma unroll
    for (int o = 0; o < 2 * OUT; ++o) dcat[(int64_t)p * 2 * OUT + o] = d[o];
}

// per scene: dYsum = sum_p a_p dx2_p ; back through the inter GCN; writes dXg (same for every group of
// the scene) at the scene row, and the GEMM operands for dV0 / dV1.
// One WARP per scene (a thread per scene walked its N pedestrians alone: 1.2 ms for 64 scenes of 1024): lanes strided over
// the pedestrians for the d(y) sum and over the hidden units for the two small products, warp sums in between.
template <int HID, int OUT>
__global__ void __launch_bounds__(128)
gcn_bwd_scene_kernel(const float* __restrict__ dcat, const float* __restrict__ Yrow, const float* __restrict__ K1s,
                     const int32_t* __restrict__ gsize, const int32_t* __restrict__ scene_start,
                     const int32_t* __restrict__ n_group, int n_scenes, const float* __restrict__ V0,
                     const float* __restrict__ V1, float* __restrict__ dXg_row, float* __restrict__ dYm,
                     float* __restrict__ dK1m) {
    __shared__ __align__(16) float sV1[HID * OUT];  // [HID][OUT]  (dN2[f] = sum_o dYm[o] V1[f][o])
    __shared__ __align__(16) float sV0[OUT * HID];  // [OUT][HID]  (dN1[o] = sum_f dK1m[f] V0[o][f])
    load_w(sV1, V1, HID, OUT, false);
    load_w(sV0, V0, OUT, HID, false);
    __syncthreads();
    const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (s >= n_scenes) return;
    const int b = scene_start[s], e = scene_start[s + 1], G = n_group[s];
    const float c = __frcp_rn((float)G);
    auto wsum = [](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    };
    float dy[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) dy[o] = 0.f;
    for (int q = b + lane; q < e; q += 32) {
        const float a = __frcp_rn((float)gsize[q]);
        const float4* row = reinterpret_cast<const float4*>(dcat + (int64_t)q * 2 * OUT + OUT);
#pragma unroll
        for (int o = 0; o < OUT / 4; ++o) {
            const float4 v = row[o];
            dy[4 * o] = fmaf(a, v.x, dy[4 * o]); dy[4 * o + 1] = fmaf(a, v.y, dy[4 * o + 1]);
            dy[4 * o + 2] = fmaf(a, v.z, dy[4 * o + 2]); dy[4 * o + 3] = fmaf(a, v.w, dy[4 * o + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        dy[o] = wsum(dy[o]);
        if (!(Yrow[(int64_t)b * OUT + o] > 0.f)) dy[o] = 0.f;
        if (lane == o) dYm[(int64_t)b * OUT + o] = dy[o];
    }
    float dn1[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) dn1[o] = 0.f;
    for (int f = lane; f < HID; f += 32) {
        float dn2 = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) dn2 = fmaf(dy[o], sV1[f * OUT + o], dn2);
        // N2_i = sum_g c K1_g for each of the G rows i  =>  dK1 (per row g) = c * sum_i dN2_i = c * dn2
        const float dk1 = (K1s[(int64_t)b * HID + f] > 0.f) ? c * dn2 : 0.f;
        // operand for dV0 = sum_g N1^T (dK1_g * mask) = N1^T (G * dk1)
        dK1m[(int64_t)b * HID + f] = (float)G * dk1;
#pragma unroll
        for (int o = 0; o < OUT; ++o) dn1[o] = fmaf(dk1, sV0[o * HID + f], dn1[o]);
    }
    // N1_g = sum_g' c Xg_g' for each of the G rows  =>  dXg_g' = c * sum_g dN1_g = c * G * dn1
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        const float v = wsum(dn1[o]);
        if (lane == o) dXg_row[(int64_t)b * OUT + o] = c * (float)G * v;
    }
}

// per leader: gather dx1 of the members, add the GPool path, back through the intra GCN.
template <int IN, int HID, int OUT>
__global__ void __launch_bounds__(128)
gcn_bwd_group_kernel(const float* __restrict__ dcat, const float* __restrict__ dXg_row,
                     const float* __restrict__ X1g, const float* __restrict__ M1s, const int32_t* __restrict__ leader,
                     const int32_t* __restrict__ gsize, const int32_t* __restrict__ ped_start,
                     const int32_t* __restrict__ ped_end, int batch, const float* __restrict__ W0,
                     const float* __restrict__ W1, float* __restrict__ dX1m, float* __restrict__ dH1m,
                     float* __restrict__ dM1sum) {
    __shared__ __align__(16) float sW0t[HID * IN];   // [HID][IN]: recompute H1 mask and dM1 = dH1m W0^T
    __shared__ __align__(16) float sW1[HID * OUT];   // [HID][OUT]
    load_w(sW0t, W0, IN, HID, true);
    load_w(sW1, W1, HID, OUT, false);
    __syncthreads();
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    if (leader[p] != p) {
#pragma unroll
        for (int o = 0; o < OUT; ++o) dX1m[(int64_t)p * OUT + o] = 0.f;
        for (int f = 0; f < HID; ++f) dH1m[(int64_t)p * HID + f] = 0.f;
        return;
    }
    const int k = gsize[p];
    const float a = __frcp_rn((float)k);
    const int sb = ped_start[p], e = ped_end[p];
    float dx1[OUT];
#pragma unroll
    for (int o = 0; o < OUT; ++o) dx1[o] = 0.f;
    for (int q = p; q < e; ++q) {
        if (leader[q] != p) continue;
#pragma unroll
        for (int o = 0; o < OUT; ++o) dx1[o] += dcat[(int64_t)q * 2 * OUT + o];
    }
    // Xg_g = sum_{j in g} a X1_j: every member row receives a * dXg; k members
#pragma unroll
    for (int o = 0; o < OUT; ++o) {
        dx1[o] += (float)k * a * dXg_row[(int64_t)sb * OUT + o];
        if (!(X1g[(int64_t)p * OUT + o] > 0.f)) dx1[o] = 0.f;
        dX1m[(int64_t)p * OUT + o] = dx1[o];   // operand of dW1 = M2^T dX1m
    }
    float m1[IN], dm1[IN];
#pragma unroll
    for (int c = 0; c < IN; ++c) { m1[c] = M1s[(int64_t)p * IN + c]; dm1[c] = 0.f; }
    for (int f = 0; f < HID; ++f) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < IN; ++c) s = fmaf(m1[c], sW0t[f * IN + c], s);
        float dm2 = 0.f;
#pragma unroll
        for (int o = 0; o < OUT; ++o) dm2 = fmaf(dx1[o], sW1[f * OUT + o], dm2);
        // M2_i = sum_j a H1_j (k rows i, k rows j) => dH1_j = a * sum_i dM2_i = a * dm2 ; summed over j: k*a*dm2
        float dh1 = (s > 0.f) ? (float)k * a * dm2 : 0.f;
        dH1m[(int64_t)p * HID + f] = dh1;       // operand of dW0 = M1^T dH1m
#pragma unroll
        for (int c = 0; c < IN; ++c) dm1[c] = fmaf(dh1, sW0t[f * IN + c], dm1[c]);
    }
#pragma unroll
    for (int c = 0; c < IN; ++c) dM1sum[(int64_t)p * IN + c] = dm1[c];
}

// per ped: dX_p = a * dM1sum[leader]
template <int IN>
__global__ void gcn_bwd_x_kernel(const float* __restrict__ dM1sum, const int32_t* __restrict__ leader,
                                 const int32_t* __restrict__ gsize, int batch, float* __restrict__ dx) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)batch * IN) return;
    int p = (int)(idx / IN), c = (int)(idx % IN);
    dx[idx] = __frcp_rn((float)gsize[p]) * dM1sum[(int64_t)leader[p] * IN + c];
}

__global__ void colsum_kernel(const float* __restrict__ m, int64_t rows, int cols, float* __restrict__ out) {
    // out[c] = sum_r m[r][c]; one block per column chunk of 32, grid-stride over rows, atomics to finish
    int c = blockIdx.x * 32 + (threadIdx.x & 31);
    int r0 = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int stride = gridDim.y * (blockDim.x >> 5);
    float s = 0.f;
    if (c < cols)
        for (int64_t r = r0; r < rows; r += stride) s += m[r * cols + c];
    if (c < cols && s != 0.f) atomicAdd(&out[c], s);
}

struct GcnWs {
    float *X1g, *Xg, *M1s, *M2s, *Yrow, *N1s, *N2s, *K1s, *cat, *dcat, *dXg_row, *dYm, *dK1m, *dX1m, *dH1m, *dM1sum;
};

static int64_t carve_gcn(Carver& c, GcnWs& w, int64_t batch, int IN, int HID, int OUT) {
    w.X1g = c.take<float>(batch * OUT); w.Xg = c.take<float>(batch * OUT);
    w.M1s = c.take<float>(batch * IN); w.M2s = c.take<float>(batch * HID);
    w.Yrow = c.take<float>(batch * OUT); w.N1s = c.take<float>(batch * OUT);
    w.N2s = c.take<float>(batch * HID); w.K1s = c.take<float>(batch * HID);
    w.cat = c.take<float>(batch * 2 * OUT); w.dcat = c.take<float>(batch * 2 * OUT);
    w.dXg_row = c.take<float>(batch * OUT); w.dYm = c.take<float>(batch * OUT);
    w.dK1m = c.take<float>(batch * HID); w.dX1m = c.take<float>(batch * OUT);
    w.dH1m = c.take<float>(batch * HID); w.dM1sum = c.take<float>(batch * IN);
    return c.off;
}

template <int IN, int HID, int OUT, int FIN>
static int gcn_forward(const float* x, const int32_t* leader, const int32_t* gsize, const int32_t* ped_start,
                       const int32_t* ped_end, const int32_t* scene_start, const int32_t* n_group, int64_t batch,
                       int64_t S, const float* W0, const float* W1, const float* V0, const float* V1, const float* Wo,
                       const float* bo, float* out, GcnWs& w, bool save, cudaStream_t st) {
    gcn_group_kernel<IN, HID, OUT><<<blocks_for(batch, 128), 128, 0, st>>>(
        x, leader, gsize, ped_start, ped_end, (int)batch, W0, W1, w.X1g, w.Xg, save ? w.M1s : nullptr,
        save ? w.M2s : nullptr);
    SGX_LAUNCH_CHECK();
    // one warp per scene (bit-identical to the thread-per-scene gcn_scene_kernel, kept as its readable statement)
    gcn_scene_warp_kernel<HID, OUT><<<blocks_for(S, 4), 128, 0, st>>>(w.Xg, leader, scene_start, n_group, (int)S, V0, V1,
                                                                     w.Yrow, save ? w.N1s : nullptr,
                                                                     save ? w.N2s : nullptr, save ? w.K1s : nullptr);
    SGX_LAUNCH_CHECK();
    if (out || save) {
        gcn_out_kernel<OUT, FIN><<<blocks_for(batch, 128), 128, 0, st>>>(w.X1g, w.Yrow, leader, gsize, ped_start,
                                                                        (int)batch, Wo, bo, out,
                                                                        save ? w.cat : nullptr);
        SGX_LAUNCH_CHECK();
    }
    return SGX_OK;
}


// ------------------------------------------------------------------------------------------------
// Fused forward for batches whose scenes all fit a warp (N <= 32): one WARP per chunk of whole scenes, lanes <-> peds.
// Leader lanes run the intra GCN on their group's mean, the first lane of every scene runs the inter GCN on the mean
// of the scene's group states, every lane writes its output row: one launch, only x / group structure / out in HBM.
// ------------------------------------------------------------------------------------------------
constexpr int GF_WARPS = 8;
constexpr int GF_RA = 44;                                     // stride of the 40-wide x rows (16 B aligned)
constexpr int GF_SCRATCH = 32 * GF_RA + 32 * 16 * 3 + 32;     // x rows, X1 / Xg / Y rows, leader slots

template <int NI, int NO>
__device__ __forceinline__ void gemv_regs2(const float (&x)[NI], const float* __restrict__ W /*[NI][NO] smem*/,
                                           float (&y)[NO]) {
    float2 acc[NO / 2];
#pragma unroll
    for (int o = 0; o < NO / 2; ++o) acc[o] = make_float2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NI; ++c) {
        const float2 xx = make_float2(x[c], x[c]);
        const float4* w = reinterpret_cast<const float4*>(W + c * NO);
#pragma unroll
        for (int o = 0; o < NO / 4; ++o) {
            const float4 v = w[o];
            acc[2 * o] = ffma2(xx, make_float2(v.x, v.y), acc[2 * o]);
            acc[2 * o + 1] = ffma2(xx, make_float2(v.z, v.w), acc[2 * o + 1]);
        }
    }
#pragma unroll
    for (int o = 0; o < NO / 2; ++o) { y[2 * o] = acc[o].x; y[2 * o + 1] = acc[o].y; }
}

template <int IN, int HID, int OUT, int FIN>
__global__ void __launch_bounds__(GF_WARPS * 32)
gcn_fused_fwd_kernel(const float* __restrict__ x, const int32_t* __restrict__ leader, const int32_t* __restrict__ gsize,
                     const int32_t* __restrict__ ped_start, const int32_t* __restrict__ ped_end,
                     const int32_t* __restrict__ scene_start, const int32_t* __restrict__ chunk_scene, int n_chunks,
                     const float* __restrict__ W0, const float* __restrict__ W1, const float* __restrict__ V0,
                     const float* __restrict__ V1, const float* __restrict__ Wo, const float* __restrict__ bo,
                     float* __restrict__ out) {
    extern __shared__ __align__(16) uint8_t raw[];
    float* sW0 = reinterpret_cast<float*>(raw);               // [IN][HID]
    float* sW1 = sW0 + IN * HID;                              // [HID][OUT]
    float* sV0 = sW1 + HID * OUT;                             // [OUT][HID]
    float* sV1 = sV0 + OUT * HID;                             // [HID][OUT]
    float* sWo = sV1 + HID * OUT;                             // [FIN][2*OUT]
    float* sbo = sWo + FIN * 2 * OUT;                         // [FIN]
    float* bufs = sbo + ((FIN + 3) / 4) * 4;
    {
        const float* src[6] = {W0, W1, V0, V1, Wo, bo};
        float* dst[6] = {sW0, sW1, sV0, sV1, sWo, sbo};
        const int cnt[6] = {IN * HID, HID * OUT, OUT * HID, HID * OUT, FIN * 2 * OUT, FIN};
        for (int k = 0; k < 6; ++k)
            for (int e = threadIdx.x; e < cnt[k]; e += blockDim.x) dst[k][e] = src[k][e];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* A = bufs + warp * GF_SCRATCH;
    float* X1s = A + 32 * GF_RA;
    float* Xgs = X1s + 32 * 16;
    float* Ys = Xgs + 32 * 16;
    int* lead_slot = reinterpret_cast<int*>(Ys + 32 * 16);
    const int n_warps_total = gridDim.x * GF_WARPS;
    for (int chunk = blockIdx.x * GF_WARPS + warp; chunk < n_chunks; chunk += n_warps_total) {
        const int p0 = scene_start[chunk_scene[chunk]];
        const int np = scene_start[chunk_scene[chunk + 1]] - p0;
        const bool live = lane < np;
        const int p = p0 + lane;
        int b = 0, e = 0, my_lead = lane, k = 1;
        if (live) {
            b = ped_start[p] - p0; e = ped_end[p] - p0; my_lead = leader[p] - p0; k = gsize[p];
            const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)p * IN);
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) reinterpret_cast<float4*>(A + lane * GF_RA)[c] = xr[c];
        }
        lead_slot[lane] = live ? my_lead : -1;
        const float a = __frcp_rn((float)k);
        __syncwarp();
        // ---- intra GCN on the group mean (leader lanes) ----
        if (live && my_lead == lane) {
            float m1[IN];
#pragma unroll
            for (int c = 0; c < IN; ++c) m1[c] = 0.f;
            for (int q = lane; q < e; ++q) {
                if (lead_slot[q] != lane) continue;
                const float4* row = reinterpret_cast<const float4*>(A + q * GF_RA);
#pragma unroll
                for (int c = 0; c < IN / 4; ++c) {
                    const float4 v = row[c];
                    m1[4 * c] = fmaf(a, v.x, m1[4 * c]); m1[4 * c + 1] = fmaf(a, v.y, m1[4 * c + 1]);
                    m1[4 * c + 2] = fmaf(a, v.z, m1[4 * c + 2]); m1[4 * c + 3] = fmaf(a, v.w, m1[4 * c + 3]);
                }
            }
            float h1[HID];
            gemv_regs2<IN, HID>(m1, sW0, h1);
#pragma unroll
            for (int f = 0; f < HID; ++f) h1[f] = repeat_sum(a * fmaxf(h1[f], 0.f), k);      // = M2
            float x1[OUT];
            gemv_regs2<HID, OUT>(h1, sW1, x1);
#pragma unroll
            for (int o = 0; o < OUT; ++o) {
                x1[o] = fmaxf(x1[o], 0.f);
                X1s[lane * 16 + o] = x1[o];
                Xgs[lane * 16 + o] = repeat_sum(a * x1[o], k);
            }
        }
        __syncwarp();
        // ---- inter GCN on the mean of the scene's group states (first lane of every scene) ----
        if (live && lane == b) {
            int G = 0;
            for (int q = b; q < e; ++q) G += (lead_slot[q] == q) ? 1 : 0;
            const float c = __frcp_rn((float)G);
            float n1[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) n1[o] = 0.f;
            for (int q = b; q < e; ++q) {
                if (lead_slot[q] != q) continue;
#pragma unroll
                for (int o = 0; o < OUT; ++o) n1[o] = fmaf(c, Xgs[q * 16 + o], n1[o]);
            }
            float k1[HID];
            gemv_regs2<OUT, HID>(n1, sV0, k1);
#pragma unroll
            for (int f = 0; f < HID; ++f) k1[f] = repeat_sum(c * fmaxf(k1[f], 0.f), G);      // = N2
            float y[OUT];
            gemv_regs2<HID, OUT>(k1, sV1, y);
#pragma unroll
            for (int o = 0; o < OUT; ++o) Ys[lane * 16 + o] = fmaxf(y[o], 0.f);
        }
        __syncwarp();
        // ---- output Linear on [X1 of my group ; Y of my scene / |group|] ----
        if (live) {
            float cat[2 * OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) { cat[o] = X1s[my_lead * 16 + o]; cat[OUT + o] = a * Ys[b * 16 + o]; }
            float* orow = out + (int64_t)p * FIN;
#pragma unroll
            for (int o4 = 0; o4 < FIN / 4; ++o4) {
                float yv[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int o = 4 * o4 + kk;
                    float acc = sbo[o];
                    const float4* wr = reinterpret_cast<const float4*>(sWo + o * 2 * OUT);
#pragma unroll
                    for (int c = 0; c < 2 * OUT / 4; ++c) {
                        const float4 v = wr[c];
                        acc = fmaf(cat[4 * c], v.x, acc); acc = fmaf(cat[4 * c + 1], v.y, acc);
                        acc = fmaf(cat[4 * c + 2], v.z, acc); acc = fmaf(cat[4 * c + 3], v.w, acc);
                    }
                    yv[kk] = acc;
                }
                reinterpret_cast<float4*>(orow)[o4] = make_float4(yv[0], yv[1], yv[2], yv[3]);
            }
        }
        __syncwarp();
    }
}


// ------------------------------------------------------------------------------------------------
// The same single-launch forward with the five linear maps as warp-level tensor-core GEMMs over the chunk's 32 slots
// (sgx_warp_mma.cuh, 3xTF32 = fp32-level accuracy).  In the GEMV form only the leader lanes (~45 %) and the first lane
// of every scene (~27 %) worked during the two GCNs; as a 32-row GEMM the slots that are not leaders / scene heads just
// carry rows nobody reads.  Row-wise steps (group mean, ReLU, the k-fold repeat sums that reproduce the reference's
// summation, the scene mean) stay lane-per-slot and use lane masks for their neighbour sets.
// ------------------------------------------------------------------------------------------------

// row[f] <- sum of `times` copies of scale * relu(row[f]) (the reference sums the identical rows of a group / scene one
// by one), F floats.  ONE loop over `times` with the whole row in registers: a loop per element made 72 divergent
// loops per row (branch stalls + 87 M instructions per launch).
template <int F>
__device__ __forceinline__ void relu_scale_repeat_row(float* __restrict__ row, float scale, int times) {
    float term[F], acc[F];
#pragma unroll
    for (int c = 0; c < F / 4; ++c) {
        const float4 v = reinterpret_cast<const float4*>(row)[c];
        term[4 * c] = scale * fmaxf(v.x, 0.f); term[4 * c + 1] = scale * fmaxf(v.y, 0.f);
        term[4 * c + 2] = scale * fmaxf(v.z, 0.f); term[4 * c + 3] = scale * fmaxf(v.w, 0.f);
    }
#pragma unroll
    for (int f = 0; f < F; ++f) acc[f] = 0.f;
    for (int t = 0; t < times; ++t) {
#pragma unroll
        for (int f = 0; f < F; ++f) acc[f] += term[f];
    }
#pragma unroll
    for (int c = 0; c < F / 4; ++c)
        reinterpret_cast<float4*>(row)[c] = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
}

constexpr int GM_WARPS = 10;
constexpr int GM_RS = 76;                                     // 72-wide row buffer (group mean of x, hidden rows, cat)
constexpr int GM_RX = 44;                                     // x rows
constexpr int GM_RN = 20;                                     // 16-wide rows
constexpr int GM_SCRATCH = 32 * GM_RS + 32 * GM_RN + 32 * 16 * 3;     // the x rows alias X1s/Xgs/Ys (dead before those are written)
static_assert(32 * GM_RX <= 32 * 16 * 3, "x rows must fit the aliased region");
constexpr int GM_SW_WIDE = 72;                                // W0 / V0 [K][72]: 72 % 32 = 8 -> conflict-free B fragments
constexpr int GM_SW_NARROW = 24;                              // W1 / V1 [72][16] padded to 24
template <int FIN> struct GmWo { static constexpr int STRIDE = (FIN % 32 == 0) ? FIN + 8 : FIN; };

template <int IN, int HID, int OUT, int FIN>
__global__ void __launch_bounds__(GM_WARPS * 32)
gcn_fused_mma_kernel(const float* __restrict__ x, const int32_t* __restrict__ leader, const int32_t* __restrict__ gsize,
                     const int32_t* __restrict__ ped_start, const int32_t* __restrict__ ped_end,
                     const int32_t* __restrict__ scene_start, const int32_t* __restrict__ chunk_scene, int n_chunks,
                     const float* __restrict__ W0, const float* __restrict__ W1, const float* __restrict__ V0,
                     const float* __restrict__ V1, const float* __restrict__ Wo, const float* __restrict__ bo,
                     float* __restrict__ out) {
    static_assert(HID == 72 && OUT == 16 && IN % 8 == 0 && FIN % 8 == 0, "built for hidden 72, out 16");
    constexpr int SWO = GmWo<FIN>::STRIDE;
    extern __shared__ __align__(16) uint8_t raw[];
    float* sW0 = reinterpret_cast<float*>(raw);               // [IN][72]
    float* sW1 = sW0 + IN * GM_SW_WIDE;                       // [72][24]
    float* sV0 = sW1 + HID * GM_SW_NARROW;                    // [16][72]
    float* sV1 = sV0 + OUT * GM_SW_WIDE;                      // [72][24]
    float* sWoT = sV1 + HID * GM_SW_NARROW;                   // [32][SWO] = Wo^T
    float* sbo = sWoT + 2 * OUT * SWO;                        // [FIN]
    float* bufs = sbo + FIN;
    for (int e = threadIdx.x; e < IN * HID; e += blockDim.x) sW0[e] = W0[e];
    for (int e = threadIdx.x; e < OUT * HID; e += blockDim.x) sV0[e] = V0[e];
    for (int e = threadIdx.x; e < HID * GM_SW_NARROW; e += blockDim.x) {
        const int k = e / GM_SW_NARROW, n = e % GM_SW_NARROW;
        sW1[e] = n < OUT ? W1[k * OUT + n] : 0.f;
        sV1[e] = n < OUT ? V1[k * OUT + n] : 0.f;
    }
    for (int e = threadIdx.x; e < 2 * OUT * SWO; e += blockDim.x) {
        const int k = e / SWO, n = e % SWO;
        sWoT[e] = n < FIN ? Wo[n * 2 * OUT + k] : 0.f;
    }
    for (int e = threadIdx.x; e < FIN; e += blockDim.x) sbo[e] = bo[e];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    float* RB = bufs + warp * GM_SCRATCH;                     // [32][GM_RS]
    float* SB = RB + 32 * GM_RS;                              // [32][GM_RN]
    float* X1s = SB + 32 * GM_RN;                             // [32][16] ReLU'd intra output of the group (at its leader)
    float* XB = X1s;                                          // [32][GM_RX] x rows, only until the group means are taken
    float* Xgs = X1s + 32 * 16;                               // [32][16] pooled group state
    float* Ys = Xgs + 32 * 16;                                // [32][16] inter output of the scene (at its first slot)
    auto to_rb = [&](int mt, int nt, const float (&c)[4]) {
        const int r = mt * 16 + g;
        *reinterpret_cast<float2*>(RB + r * GM_RS + nt * 8 + 2 * t) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(RB + (r + 8) * GM_RS + nt * 8 + 2 * t) = make_float2(c[2], c[3]);
    };
    auto to_sb = [&](int mt, int nt, const float (&c)[4]) {
        const int r = mt * 16 + g;
        *reinterpret_cast<float2*>(SB + r * GM_RN + nt * 8 + 2 * t) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(SB + (r + 8) * GM_RN + nt * 8 + 2 * t) = make_float2(c[2], c[3]);
    };
    const int n_warps_total = gridDim.x * GM_WARPS;
    for (int chunk = blockIdx.x * GM_WARPS + warp; chunk < n_chunks; chunk += n_warps_total) {
        const int p0 = scene_start[chunk_scene[chunk]];
        const int np = scene_start[chunk_scene[chunk + 1]] - p0;
        const bool live = lane < np;
        const int p = p0 + lane;
        int b = 0, e = 0, my_lead = lane, k = 1;
        {
            float4 xv[IN / 4];
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) xv[c] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                b = ped_start[p] - p0; e = ped_end[p] - p0; my_lead = leader[p] - p0; k = gsize[p];
                const float4* xr = reinterpret_cast<const float4*>(x + (int64_t)p * IN);
#pragma unroll
                for (int c = 0; c < IN / 4; ++c) xv[c] = xr[c];
            }
#pragma unroll
            for (int c = 0; c < IN / 4; ++c) reinterpret_cast<float4*>(XB + lane * GM_RX)[c] = xv[c];
        }
        const bool is_lead = live && my_lead == lane;
        const bool is_head = live && lane == b;
        const float a = __frcp_rn((float)k);
        const uint32_t group_mask = __match_any_sync(0xffffffffu, live ? my_lead : 32 + lane);
        const uint32_t scene_mask = (e >= 32 ? 0xffffffffu : ((1u << e) - 1u)) & ~((1u << b) - 1u);
        const uint32_t leader_mask = __ballot_sync(0xffffffffu, is_lead) & scene_mask;
        __syncwarp();
        // ---- group mean of x at the leader slots -> RB rows ----
        {
            float m1[IN];
#pragma unroll
            for (int c = 0; c < IN; ++c) m1[c] = 0.f;
            if (is_lead) {
                for (uint32_t mm = group_mask; mm; mm &= mm - 1) {
                    const float4* row = reinterpret_cast<const float4*>(XB + (__ffs(mm) - 1) * GM_RX);
#pragma unroll
                    for (int c = 0; c < IN / 4; ++c) {
                        const float4 v = row[c];
                        m1[4 * c] = fmaf(a, v.x, m1[4 * c]); m1[4 * c + 1] = fmaf(a, v.y, m1[4 * c + 1]);
                        m1[4 * c + 2] = fmaf(a, v.z, m1[4 * c + 2]); m1[4 * c + 3] = fmaf(a, v.w, m1[4 * c + 3]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < IN / 4; ++c)
                reinterpret_cast<float4*>(RB + lane * GM_RS)[c] = make_float4(m1[4 * c], m1[4 * c + 1], m1[4 * c + 2], m1[4 * c + 3]);
        }
        __syncwarp();
        // ---- intra GCN: H1 = M1 W0 (in place), M2 = k-fold sum of relu(H1)/k, X1 = relu(M2 W1) ----
        warp_gemm_3xtf32<IN, HID / 8, GM_RS, GM_SW_WIDE>(RB, sW0, lane, to_rb);
        {
            relu_scale_repeat_row<HID>(RB + lane * GM_RS, a, k);
        }
        __syncwarp();
        warp_gemm_3xtf32<HID, OUT / 8, GM_RS, GM_SW_NARROW>(RB, sW1, lane, to_sb);
        {
            const float4* row = reinterpret_cast<const float4*>(SB + lane * GM_RN);
#pragma unroll
            for (int c = 0; c < OUT / 4; ++c) {
                float4 v = row[c];
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
                reinterpret_cast<float4*>(X1s + lane * 16)[c] = v;
                reinterpret_cast<float4*>(Xgs + lane * 16)[c] = make_float4(repeat_sum(a * v.x, k), repeat_sum(a * v.y, k),
                                                                            repeat_sum(a * v.z, k), repeat_sum(a * v.w, k));
            }
        }
        __syncwarp();
        // ---- inter GCN at the scene heads: N1 = mean of the scene's group states ----
        const int G = __popc(leader_mask);
        const float cg = __frcp_rn((float)(G > 0 ? G : 1));
        {
            float n1[OUT];
#pragma unroll
            for (int o = 0; o < OUT; ++o) n1[o] = 0.f;
            if (is_head) {
                for (uint32_t mm = leader_mask; mm; mm &= mm - 1) {
                    const int q = __ffs(mm) - 1;
#pragma unroll
                    for (int o = 0; o < OUT; ++o) n1[o] = fmaf(cg, Xgs[q * 16 + o], n1[o]);
                }
            }
#pragma unroll
            for (int c = 0; c < OUT / 4; ++c)
                reinterpret_cast<float4*>(SB + lane * GM_RN)[c] = make_float4(n1[4 * c], n1[4 * c + 1], n1[4 * c + 2], n1[4 * c + 3]);
        }
        __syncwarp();
        warp_gemm_3xtf32<OUT, HID / 8, GM_RN, GM_SW_WIDE>(SB, sV0, lane, to_rb);
        {
            relu_scale_repeat_row<HID>(RB + lane * GM_RS, cg, is_head ? G : 1);   // rows of the other slots are never read
        }
        __syncwarp();
        warp_gemm_3xtf32<HID, OUT / 8, GM_RS, GM_SW_NARROW>(RB, sV1, lane, to_sb);
        {
            const float4* row = reinterpret_cast<const float4*>(SB + lane * GM_RN);
#pragma unroll
            for (int c = 0; c < OUT / 4; ++c) {
                const float4 v = row[c];
                reinterpret_cast<float4*>(Ys + lane * 16)[c] =
                    make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
            }
        }
        __syncwarp();
        // ---- cat = [X1 of my group ; Y of my scene / |group|] -> RB rows, out = cat Wo^T + bo ----
#pragma unroll
        for (int c = 0; c < OUT / 4; ++c) {
            const float4 u = reinterpret_cast<const float4*>(X1s + my_lead * 16)[c];
            const float4 v = reinterpret_cast<const float4*>(Ys + b * 16)[c];
            reinterpret_cast<float4*>(RB + lane * GM_RS)[c] = u;
            reinterpret_cast<float4*>(RB + lane * GM_RS + OUT)[c] = make_float4(a * v.x, a * v.y, a * v.z, a * v.w);
        }
        __syncwarp();
        warp_gemm_3xtf32<2 * OUT, FIN / 8, GM_RS, SWO>(RB, sWoT, lane, [&](int mt, int nt, const float (&c)[4]) {
            const int r = mt * 16 + g, col = nt * 8 + 2 * t;
            const float b0 = sbo[col], b1 = sbo[col + 1];
            if (r < np) *reinterpret_cast<float2*>(out + (int64_t)(p0 + r) * FIN + col) = make_float2(c[0] + b0, c[1] + b1);
            if (r + 8 < np)
                *reinterpret_cast<float2*>(out + (int64_t)(p0 + r + 8) * FIN + col) = make_float2(c[2] + b0, c[3] + b1);
        });
    }
}

template <int IN, int FIN>
int gcn_fused_tc_forward(const float* x, const int32_t* leader, const int32_t* gsize, const float* labels, const int32_t* ps, const int32_t* pe,
                         const int32_t* scene_start, const int32_t* chunk_scene, int n_chunks, const float* W0,
                         const float* W1, const float* V0, const float* V1, const float* Wo, const float* bo, float* out,
                         cudaStream_t st, const void* prep = nullptr, void* prep_out = nullptr);   // sgx_gcn_tc.cu
template <int IN, int FIN>
int64_t gcn_tc_prep_bytes();

template <int IN, int FIN>
static int gcn_fused_launch(const float* x, const int32_t* leader, const int32_t* gsize, const int32_t* ps,
                            const int32_t* pe, const int32_t* scene_start, const int32_t* chunk_scene, int n_chunks,
                            const float* W0, const float* W1, const float* V0, const float* V1, const float* Wo,
                            const float* bo, float* out, cudaStream_t st) {
    if (opt_graph_tc())                    // linear maps on tcgen05 (sgx_gcn_tc.cu)
        return gcn_fused_tc_forward<IN, FIN>(x, leader, gsize, nullptr, ps, pe, scene_start, chunk_scene, n_chunks, W0, W1, V0, V1,
                                             Wo, bo, out, st);
#ifdef SGX_AB_VARIANTS
    const bool mma = opt_gcn_mma();        // A/B builds only: sgx_set_option("gcn_mma", 0) selects the CUDA-core GEMV kernel
#else
    constexpr bool mma = true;
#endif
    if (mma) {                             // linear maps on the tensor cores
        auto kern_m = gcn_fused_mma_kernel<IN, 72, 16, FIN>;
        const int wf = IN * GM_SW_WIDE + 2 * 72 * GM_SW_NARROW + 16 * GM_SW_WIDE + 32 * GmWo<FIN>::STRIDE + FIN;
        const int smem_m = (wf + GM_WARPS * GM_SCRATCH) * (int)sizeof(float);
        SGX_CUDA(cudaFuncSetAttribute(kern_m, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_m));
        const int grid_m = std::min((n_chunks + GM_WARPS - 1) / GM_WARPS, 148);
        kern_m<<<grid_m, GM_WARPS * 32, smem_m, st>>>(x, leader, gsize, ps, pe, scene_start, chunk_scene, n_chunks, W0, W1,
                                                      V0, V1, Wo, bo, out);
        SGX_LAUNCH_CHECK();
        return SGX_OK;
    }
#ifdef SGX_AB_VARIANTS
    auto kern = gcn_fused_fwd_kernel<IN, 72, 16, FIN>;
    const int wfloats = IN * 72 + 72 * 16 * 3 + FIN * 32 + ((FIN + 3) / 4) * 4;
    const int smem = (wfloats + GF_WARPS * GF_SCRATCH) * (int)sizeof(float);
    SGX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = std::min((n_chunks + GF_WARPS - 1) / GF_WARPS, 148);
    kern<<<grid, GF_WARPS * 32, smem, st>>>(x, leader, gsize, ps, pe, scene_start, chunk_scene, n_chunks, W0, W1, V0, V1,
                                            Wo, bo, out);
    SGX_LAUNCH_CHECK();
#endif
    return SGX_OK;
}

}  // namespace sgx

using namespace sgx;

extern "C" int64_t sgx_gcn_module_ws_bytes(int64_t batch, int64_t n_scenes, int32_t IN, int32_t HID, int32_t OUT,
                                           int32_t FIN) {
    (void)n_scenes; (void)FIN;
    Carver c(nullptr);
    GcnWs w;
    return carve_gcn(c, w, batch, IN, HID, OUT);
}

#define GCN_DISPATCH(CALL)                                                                        \
    if (IN == 40 && HID == 72 && OUT == 16 && FIN == 24) { constexpr int I = 40, F = 24; CALL; }      \
    else if (IN == 32 && HID == 72 && OUT == 16 && FIN == 24) { constexpr int I = 32, F = 24; CALL; } \
    else if (IN == 40 && HID == 72 && OUT == 16 && FIN == 32) { constexpr int I = 40, F = 32; CALL; } \
    else if (IN == 32 && HID == 72 && OUT == 16 && FIN == 32) { constexpr int I = 32, F = 32; CALL; } \
    else {                                                                                        \
        sgx::set_error("GCNModule dims (in=%d hid=%d out=%d final=%d) have no kernel instance; "      \
                       "built: in {32,40}, hid 72, out 16, final {24,32}", IN, HID, OUT, FIN);         \
        return SGX_ERR_UNSUPPORTED;                                                               \
    }

extern "C" int sgx_gcn_module_fwd(const float* x, const int32_t* leader, const int32_t* group_size,
                                  const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                                  const int32_t* n_group, int64_t batch, int64_t n_scenes, const float* W0,
                                  const float* W1, const float* V0, const float* V1, const float* Wo, const float* bo,
                                  int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, float* out, void* workspace,
                                  int64_t ws_bytes, void* stream) {
    SGX_REQUIRE(x && leader && group_size && ped_start && ped_end && scene_start && n_group && W0 && W1 && V0 && V1 &&
                    Wo && bo && out && workspace, "sgx_gcn_module_fwd: null pointer");
    SGX_REQUIRE(batch > 0 && n_scenes > 0, "sgx_gcn_module_fwd: empty batch");
    SGX_REQUIRE(ws_bytes >= sgx_gcn_module_ws_bytes(batch, n_scenes, IN, HID, OUT, FIN), "workspace too small");
    Carver c(workspace);
    GcnWs w;
    carve_gcn(c, w, batch, IN, HID, OUT);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = SGX_OK;
    GCN_DISPATCH((rc = gcn_forward<I, 72, 16, F>(x, leader, group_size, ped_start, ped_end, scene_start, n_group, batch,
                                                 n_scenes, W0, W1, V0, V1, Wo, bo, out, w, false, st)));
    return rc;
}

template <int IN, int HID, int OUT, int FIN>
static int gcn_backward(const float* x, const float* gout, const int32_t* leader, const int32_t* gsize,
                        const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                        const int32_t* n_group, int64_t batch, int64_t S, const float* W0, const float* W1,
                        const float* V0, const float* V1, const float* Wo, const float* bo, float* gx, float* gW0,
                        float* gW1, float* gV0, float* gV1, float* gWo, float* gbo, GcnWs& w, cudaStream_t st) {
    // per-scene operands live at the row of the scene's first ped; all other rows must be zero for the GEMMs
    SGX_CUDA(cudaMemsetAsync(w.N1s, 0, (size_t)batch * OUT * 4, st));
    SGX_CUDA(cudaMemsetAsync(w.N2s, 0, (size_t)batch * HID * 4, st));
    SGX_CUDA(cudaMemsetAsync(w.dYm, 0, (size_t)batch * OUT * 4, st));
    SGX_CUDA(cudaMemsetAsync(w.dK1m, 0, (size_t)batch * HID * 4, st));
    int rc = gcn_forward<IN, HID, OUT, FIN>(x, leader, gsize, ped_start, ped_end, scene_start, n_group, batch, S, W0, W1,
                                            V0, V1, Wo, bo, nullptr, w, true, st);
    if (rc) return rc;
    gcn_bwd_out_kernel<OUT, FIN><<<blocks_for(batch, 128), 128, 0, st>>>(gout, (int)batch, Wo, w.dcat);
    SGX_LAUNCH_CHECK();
    // dWo = gout^T cat ; dbo = colsum(gout)
    if ((rc = gemm(gout, 1, FIN, w.cat, 2 * OUT, 1, gWo, 2 * OUT, FIN, 2 * OUT, batch, 0, 0, st))) return rc;
    SGX_CUDA(cudaMemsetAsync(gbo, 0, FIN * 4, st));
    colsum_kernel<<<dim3((FIN + 31) / 32, 64), 256, 0, st>>>(gout, batch, FIN, gbo);
    SGX_LAUNCH_CHECK();
    gcn_bwd_scene_kernel<HID, OUT><<<blocks_for(S * 32, 128), 128, 0, st>>>(w.dcat, w.Yrow, w.K1s, gsize, scene_start,
                                                                      n_group, (int)S, V0, V1, w.dXg_row, w.dYm, w.dK1m);
    SGX_LAUNCH_CHECK();
    if ((rc = gemm(w.N2s, 1, HID, w.dYm, OUT, 1, gV1, OUT, HID, OUT, batch, 0, 0, st))) return rc;
    if ((rc = gemm(w.N1s, 1, OUT, w.dK1m, HID, 1, gV0, HID, OUT, HID, batch, 0, 0, st))) return rc;
    gcn_bwd_group_kernel<IN, HID, OUT><<<blocks_for(batch, 128), 128, 0, st>>>(
        w.dcat, w.dXg_row, w.X1g, w.M1s, leader, gsize, ped_start, ped_end, (int)batch, W0, W1, w.dX1m, w.dH1m,
        w.dM1sum);
    SGX_LAUNCH_CHECK();
    if ((rc = gemm(w.M2s, 1, HID, w.dX1m, OUT, 1, gW1, OUT, HID, OUT, batch, 0, 0, st))) return rc;
    if ((rc = gemm(w.M1s, 1, IN, w.dH1m, HID, 1, gW0, HID, IN, HID, batch, 0, 0, st))) return rc;
    gcn_bwd_x_kernel<IN><<<blocks_for(batch * IN, 256), 256, 0, st>>>(w.dM1sum, leader, gsize, (int)batch, gx);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

extern "C" int sgx_gcn_module_bwd(const float* x, const float* grad_out, const int32_t* leader,
                                  const int32_t* group_size, const int32_t* ped_start, const int32_t* ped_end,
                                  const int32_t* scene_start, const int32_t* n_group, int64_t batch, int64_t n_scenes,
                                  const float* W0, const float* W1, const float* V0, const float* V1, const float* Wo,
                                  const float* bo, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, float* grad_x,
                                  float* grad_W0, float* grad_W1, float* grad_V0, float* grad_V1, float* grad_Wo,
                                  float* grad_bo, void* workspace, int64_t ws_bytes, void* stream) {
    SGX_REQUIRE(x && grad_out && leader && group_size && ped_start && ped_end && scene_start && n_group && W0 && W1 &&
                    V0 && V1 && Wo && bo && grad_x && grad_W0 && grad_W1 && grad_V0 && grad_V1 && grad_Wo && grad_bo &&
                    workspace, "sgx_gcn_module_bwd: null pointer");
    SGX_REQUIRE(batch > 0 && n_scenes > 0, "sgx_gcn_module_bwd: empty batch");
    SGX_REQUIRE(ws_bytes >= sgx_gcn_module_ws_bytes(batch, n_scenes, IN, HID, OUT, FIN), "workspace too small");
    Carver c(workspace);
    GcnWs w;
    carve_gcn(c, w, batch, IN, HID, OUT);
    cudaStream_t st = (cudaStream_t)stream;
    int rc = SGX_OK;
    GCN_DISPATCH((rc = gcn_backward<I, 72, 16, F>(x, grad_out, leader, group_size, ped_start, ped_end, scene_start,
                                                  n_group, batch, n_scenes, W0, W1, V0, V1, Wo, bo, grad_x, grad_W0,
                                                  grad_W1, grad_V0, grad_V1, grad_Wo, grad_bo, w, st)));
    return rc;
}

// Fused single-launch forward for batches whose scenes all have <= 32 pedestrians (chunk_scene / n_chunks from
// sgx_schedule_chunks with cap = 32).  Same dims as sgx_gcn_module_fwd with hid = 72, out = 16, final % 4 == 0.
extern "C" int sgx_gcn_module_fused_fwd(const float* x, const int32_t* leader, const int32_t* group_size,
                                        const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                                        const int32_t* chunk_scene, int64_t n_chunks, const float* W0, const float* W1,
                                        const float* V0, const float* V1, const float* Wo, const float* bo, int32_t IN,
                                        int32_t HID, int32_t OUT, int32_t FIN, float* out, void* stream) {
    SGX_REQUIRE(x && leader && group_size && ped_start && ped_end && scene_start && chunk_scene && W0 && W1 && V0 && V1 &&
                    Wo && bo && out, "sgx_gcn_module_fused_fwd: null pointer");
    SGX_REQUIRE(n_chunks > 0 && n_chunks < ((int64_t)1 << 31), "sgx_gcn_module_fused_fwd: bad chunk count");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = SGX_OK;
    GCN_DISPATCH((rc = gcn_fused_launch<I, F>(x, leader, group_size, ped_start, ped_end, scene_start, chunk_scene,
                                              (int)n_chunks, W0, W1, V0, V1, Wo, bo, out, st)));
    return rc;
}

// The same forward with the group structure derived inside the kernel from the datasets_group labels (no sgx_group_ids
// pass, no leader / size / n_group arrays): scenes <= 32 pedestrians, tcgen05 kernel.  prep (nullable): the weight images
// of sgx_gcn_module_tc_prep for THESE weights; without it the kernel builds them itself (~8 us per launch).
extern "C" int sgx_gcn_module_fused_fwd_labels(const float* x, const float* labels, const int32_t* ped_start,
                                               const int32_t* ped_end, const int32_t* scene_start,
                                               const int32_t* chunk_scene, int64_t n_chunks, const float* W0,
                                               const float* W1, const float* V0, const float* V1, const float* Wo,
                                               const float* bo, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN,
                                               const void* prep, float* out, void* stream) {
    SGX_REQUIRE(x && labels && ped_start && ped_end && scene_start && chunk_scene && W0 && W1 && V0 && V1 && Wo && bo && out,
                "sgx_gcn_module_fused_fwd_labels: null pointer");
    SGX_REQUIRE(n_chunks > 0 && n_chunks < ((int64_t)1 << 31), "sgx_gcn_module_fused_fwd_labels: bad chunk count");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = SGX_OK;
    GCN_DISPATCH((rc = sgx::gcn_fused_tc_forward<I, F>(x, nullptr, nullptr, labels, ped_start, ped_end, scene_start, chunk_scene,
                                                       (int)n_chunks, W0, W1, V0, V1, Wo, bo, out, st, prep, nullptr)));
    return rc;
}

extern "C" int64_t sgx_gcn_module_tc_prep_bytes(int32_t IN, int32_t HID, int32_t OUT, int32_t FIN) {
    int64_t n = -1;
    if (HID == 72 && OUT == 16) {
        if (IN == 40 && FIN == 24) n = sgx::gcn_tc_prep_bytes<40, 24>();
        else if (IN == 32 && FIN == 24) n = sgx::gcn_tc_prep_bytes<32, 24>();
        else if (IN == 40 && FIN == 32) n = sgx::gcn_tc_prep_bytes<40, 32>();
        else if (IN == 32 && FIN == 32) n = sgx::gcn_tc_prep_bytes<32, 32>();
    }
    return n;
}

extern "C" int sgx_gcn_module_tc_prep(const float* W0, const float* W1, const float* V0, const float* V1, const float* Wo,
                                      const float* bo, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, void* prep,
                                      void* stream) {
    SGX_REQUIRE(W0 && W1 && V0 && V1 && Wo && bo && prep, "sgx_gcn_module_tc_prep: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = SGX_OK;
    GCN_DISPATCH((rc = sgx::gcn_fused_tc_forward<I, F>(nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                                                       W0, W1, V0, V1, Wo, bo, nullptr, st, nullptr, prep)));
    return rc;
}
