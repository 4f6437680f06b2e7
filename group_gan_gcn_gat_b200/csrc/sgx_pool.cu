// PoolHiddenNet forward/backward, fp32 CUDA-core path (the 1e-5 parity path).
//
// Reference: sgan/models.py:497-549.  For every ordered pair (i,j) of a scene the reference
// materialises [We (P_j-P_i)+be ; h_j] (N^2 x 48), runs Linear(48,512)+ReLU, Linear(512,B)+ReLU
// and max-reduces over j.  Here nothing of size N^2 x 512 ever exists:
//
//   layer 1 is split exactly (hard part 3):
//       W1 [We d + be ; h_j] + b1 = (W1e We) d + (W1h h_j + W1e be + b1) = Aeff d + C_j
//   stage A (per ped)  : C^T[k][j] for all 512 hidden units   (one small GEMM, K = H)
//   stage B (per pair) : z_k = Aeff[k].d + C^T[k][j];  y_b += W2[b][k] * relu(z_k), k = 0..511,
//                        with the 512-wide hidden vector living only in registers;
//                        pairs are laid out flat (sorted by (i,j)), lanes <-> consecutive pairs,
//                        segmented warp max over pairs that share i, then one 64-bit atomicMax
//                        of (value_bits << 32 | j) per (segment, channel): post-ReLU values are
//                        >= 0 so unsigned compare of the bit pattern is exact, and the low word
//                        carries the argmax for the sparse backward.
//   stage F            : unpack to out fp32 / argmax int32.
#include "sgx_common.cuh"

namespace sgx {

constexpr int HID = SGX_POOL_HIDDEN;

// Aeff[k] = W1[k,:E] . We[:,c]   c0[k] = b1[k] + W1[k,:E] . be
__global__ void pool_prep_kernel(const float* __restrict__ We, const float* __restrict__ be,
                                 const float* __restrict__ W1, const float* __restrict__ b1, int E, int H,
                                 float2* __restrict__ Aeff, float* __restrict__ c0) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= HID) return;
    const float* w = W1 + (int64_t)k * (E + H);
    float ax = 0.f, ay = 0.f, c = b1[k];
    for (int e = 0; e < E; ++e) {
        float v = w[e];
        ax = fmaf(v, We[2 * e], ax);
        ay = fmaf(v, We[2 * e + 1], ay);
        c = fmaf(v, be[e], c);
    }
    Aeff[k] = make_float2(ax, ay);
    c0[k] = c;
}

__device__ __forceinline__ int find_ped(const int64_t* __restrict__ pair_off, int lo, int hi, int64_t q) {
    // largest i in [lo,hi] with pair_off[i] <= q
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (pair_off[mid] <= q) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// BC output channels per CTA column (blockIdx.y selects the channel block), R pairs per thread.
template <int BC, int R, int THREADS>
__global__ void __launch_bounds__(THREADS)
pool_pair_kernel(const float* __restrict__ Ct, int64_t ldc, const float* __restrict__ pos,
                 const int32_t* __restrict__ ped_start, const int64_t* __restrict__ pair_off,
                 const int32_t* __restrict__ tile_first, int64_t n_tiles128, int batch, int64_t n_pairs,
                 const float2* __restrict__ Aeff, const float* __restrict__ W2, const float* __restrict__ b2, int B,
                 unsigned long long* __restrict__ packed) {
    __shared__ float2 sA[HID];
    __shared__ __align__(16) float sW[HID * BC];  // [k][c]
    const int cb = blockIdx.y * BC;
    for (int k = threadIdx.x; k < HID; k += THREADS) sA[k] = Aeff[k];
    for (int e = threadIdx.x; e < HID * BC; e += THREADS) {
        int c = e / HID, k = e % HID;  // coalesced read of W2[cb+c][k]
        sW[k * BC + c] = W2[(int64_t)(cb + c) * HID + k];
    }
    __syncthreads();

    const int64_t base = (int64_t)blockIdx.x * (THREADS * R);
    int ped_i[R], ped_j[R];
    float dx[R], dy[R];
    const float* cptr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int64_t q = base + (int64_t)r * THREADS + threadIdx.x;
        ped_i[r] = -1;
        ped_j[r] = 0;
        dx[r] = dy[r] = 0.f;
        cptr[r] = Ct;
        if (q < n_pairs) {
            int64_t t = q >> 7;
            int lo = tile_first[t];
            int hi = (t + 1 < n_tiles128) ? tile_first[t + 1] : batch - 1;
            int i = find_ped(pair_off, lo, hi, q);
            int j = ped_start[i] + (int)(q - pair_off[i]);
            ped_i[r] = i;
            ped_j[r] = j;
            dx[r] = pos[2 * j] - pos[2 * i];          // P_j - P_i first, never A.P_j - A.P_i
            dy[r] = pos[2 * j + 1] - pos[2 * i + 1];
            cptr[r] = Ct + j;
        }
    }
    float acc[R][BC];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int c = 0; c < BC; ++c) acc[r][c] = 0.f;

#pragma unroll 2
    for (int k = 0; k < HID; ++k) {
        const float2 a = sA[k];
        float w[BC];
#pragma unroll
        for (int c4 = 0; c4 < BC / 4; ++c4) {
            float4 v = *reinterpret_cast<const float4*>(&sW[k * BC + 4 * c4]);
            w[4 * c4] = v.x; w[4 * c4 + 1] = v.y; w[4 * c4 + 2] = v.z; w[4 * c4 + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float z = fmaf(a.x, dx[r], fmaf(a.y, dy[r], __ldg(cptr[r] + (int64_t)k * ldc)));
            asm("max.NaN.f32 %0, %0, 0f00000000;" : "+f"(z));     // ReLU that keeps a NaN (FMNMX.NAN), like torch.relu
#pragma unroll
            for (int c = 0; c < BC; ++c) acc[r][c] = fmaf(w[c], z, acc[r][c]);
        }
    }

    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int key = ped_i[r];
        const int key_prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (key_prev != key);
        bool same[5];
#pragma unroll
        for (int sft = 0; sft < 5; ++sft) {
            const int okey = __shfl_down_sync(0xffffffffu, key, 1 << sft);
            same[sft] = (lane + (1 << sft) < 32) && (okey == key);
        }
#pragma unroll
        for (int c = 0; c < BC; ++c) {
            // a NaN keeps its (canonical) bit pattern: it orders above every finite value, like torch.max
            const float y = acc[r][c] + b2[cb + c];
            const unsigned ybits = (y != y) ? 0x7fc00000u : (__float_as_uint(fmaxf(y, 0.f)) & 0x7fffffffu);
            unsigned long long pk = ((unsigned long long)ybits << 32) | (unsigned)ped_j[r];
#pragma unroll
            for (int sft = 0; sft < 5; ++sft) {
                const unsigned long long other = __shfl_down_sync(0xffffffffu, pk, 1 << sft);
                if (same[sft] && other > pk) pk = other;
            }
            if (head && key >= 0) atomicMax(&packed[(int64_t)key * B + cb + c], pk);
        }
    }
}

__global__ void pool_unpack_kernel(const unsigned long long* __restrict__ packed, int64_t n, float* __restrict__ out,
                                   int32_t* __restrict__ argmax) {
    pdl_trigger();
    pdl_wait();
    // four keys per thread (n = batch * bottleneck_dim, a multiple of 8): two 16-byte loads, one 16-byte store each
    const int64_t i = 4 * ((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n) return;
    const ulonglong2 p0 = *reinterpret_cast<const ulonglong2*>(packed + i);
    const ulonglong2 p1 = *reinterpret_cast<const ulonglong2*>(packed + i + 2);
    *reinterpret_cast<float4*>(out + i) =
        make_float4(__uint_as_float((unsigned)(p0.x >> 32)), __uint_as_float((unsigned)(p0.y >> 32)),
                    __uint_as_float((unsigned)(p1.x >> 32)), __uint_as_float((unsigned)(p1.y >> 32)));
    if (argmax)
        *reinterpret_cast<int4*>(argmax + i) = make_int4((int32_t)(unsigned)(p0.x & 0xffffffffu), (int32_t)(unsigned)(p0.y & 0xffffffffu),
                                                         (int32_t)(unsigned)(p1.x & 0xffffffffu), (int32_t)(unsigned)(p1.y & 0xffffffffu));
}

// ------------------------------------------------------------------------------------------------
// Backward.  Only the (i, b) -> j* = argmax pairs carry gradient (hard part 4).
// One warp per pedestrian i; lanes own 16 hidden units each.  Channels of i that share the same j*
// are merged before the scatter into dC[j*].  Per-CTA partial sums for dW2 / dAeff / db2 live in
// shared memory and are flushed once.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pool_bwd_event_kernel(const float* __restrict__ Cr /*[batch][512]*/, const float* __restrict__ pos,
                      const float* __restrict__ out, const int32_t* __restrict__ argmax,
                      const float* __restrict__ gout, int batch, int B, const float2* __restrict__ Aeff,
                      const float* __restrict__ W2, float* __restrict__ dC /*[batch][512]*/,
                      float* __restrict__ dW2, float* __restrict__ db2, float* __restrict__ dAeff /*[512][2]*/,
                      float* __restrict__ dpos) {
    extern __shared__ float smem[];
    float* sdW2 = smem;                 // [B][512]
    float* sdA = sdW2 + (size_t)B * HID;  // [512][2]
    float* sdb2 = sdA + 2 * HID;          // [B]
    for (int e = threadIdx.x; e < B * HID + 2 * HID + B; e += blockDim.x) smem[e] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = blockIdx.x * nwarp + warp; i < batch; i += gridDim.x * nwarp) {
        const float pix = pos[2 * i], piy = pos[2 * i + 1];
        for (int b0 = 0; b0 < B; ++b0) {
            float g0 = gout[(int64_t)i * B + b0];
            bool act0 = (out[(int64_t)i * B + b0] > 0.f) && (g0 != 0.f);
            if (!act0) continue;
            const int j = argmax[(int64_t)i * B + b0];
            bool first = true;  // is b0 the first active channel with this j?
            for (int b = 0; b < b0; ++b) {
                if (argmax[(int64_t)i * B + b] == j && out[(int64_t)i * B + b] > 0.f &&
                    gout[(int64_t)i * B + b] != 0.f) { first = false; break; }
            }
            if (!first) continue;
            const float ddx = pos[2 * j] - pix, ddy = pos[2 * j + 1] - piy;
            float z[16], dz[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                int k = u * 32 + lane;
                float2 a = Aeff[k];
                z[u] = fmaf(a.x, ddx, fmaf(a.y, ddy, Cr[(int64_t)j * HID + k]));
                dz[u] = 0.f;
            }
            for (int b = b0; b < B; ++b) {
                float g = gout[(int64_t)i * B + b];
                if (argmax[(int64_t)i * B + b] != j || !(out[(int64_t)i * B + b] > 0.f) || g == 0.f) continue;
                if (lane == 0) atomicAdd(&sdb2[b], g);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    int k = u * 32 + lane;
                    float r = fmaxf(z[u], 0.f);
                    if (r > 0.f) {
                        atomicAdd(&sdW2[b * HID + k], g * r);
                        dz[u] = fmaf(g, W2[(int64_t)b * HID + k], dz[u]);
                    }
                }
            }
            float sx = 0.f, sy = 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                int k = u * 32 + lane;
                if (dz[u] != 0.f) {
                    atomicAdd(&dC[(int64_t)j * HID + k], dz[u]);
                    atomicAdd(&sdA[2 * k], dz[u] * ddx);
                    atomicAdd(&sdA[2 * k + 1], dz[u] * ddy);
                    float2 a = Aeff[k];
                    sx = fmaf(dz[u], a.x, sx);
                    sy = fmaf(dz[u], a.y, sy);
                }
            }
            sx = warp_sum(sx);
            sy = warp_sum(sy);
            if (lane == 0 && (sx != 0.f || sy != 0.f)) {
                atomicAdd(&dpos[2 * j], sx);
                atomicAdd(&dpos[2 * j + 1], sy);
                atomicAdd(&dpos[2 * i], -sx);
                atomicAdd(&dpos[2 * i + 1], -sy);
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < B * HID; e += blockDim.x)
        if (sdW2[e] != 0.f) atomicAdd(&dW2[e], sdW2[e]);
    for (int e = threadIdx.x; e < 2 * HID; e += blockDim.x)
        if (sdA[e] != 0.f) atomicAdd(&dAeff[e], sdA[e]);
    for (int e = threadIdx.x; e < B; e += blockDim.x)
        if (sdb2[e] != 0.f) atomicAdd(&db2[e], sdb2[e]);
}

// ------------------------------------------------------------------------------------------------
// Backward events without global atomics (round 2).  pool_bwd_event_kernel above scatters every event's 512 hidden-unit
// gradients into dC[j*] with atomicAdd: 10^9 L2 atomics per call at 245 k pedestrians, 3.0 ms of the 5.2 ms backward.
// Here a CTA owns whole scenes -- block b = the scenes whose first pedestrian lies in [16 b, 16 b + 16) -- and thread k
// owns hidden unit k: the block's rows of Cr are staged in shared memory once, the events of the block's pedestrians
// (argmax pairs never leave a scene) accumulate into a shared dC tile that only thread k touches in column k, and the
// tile is written once with plain stores.  dW2 / dAeff / dc0 live in registers for the whole kernel (dc0 = the column
// sums of dC: no separate pass over the 500 MB matrix).  (Tried and dropped: recomputing the block's Cr rows in the kernel
// as a 3xTF32 warp GEMM from h instead of reading the GEMM's output -- correct, but the weight fragments do not fit in
// registers next to the per-channel accumulators and streaming them from L1 made the kernel 1.47 -> 2.53 ms, more than
// the 0.45 ms GEMM it removes.)  A block spanning more than PB_ROWS rows (a scene > 33) zeroes its
// own rows and falls back to atomics.
// ------------------------------------------------------------------------------------------------
constexpr int PB_IB = 16;            // pedestrians whose scenes start in one block
constexpr int PB_ROWS = 48;          // rows of a block held in shared memory (15 + largest scene <= 48)

template <int B, bool DPOS>
__global__ void __launch_bounds__(HID, 1)
pool_bwd_block_kernel(const float* __restrict__ Cr /*[batch][512]*/, const float* __restrict__ pos,
                      const float* __restrict__ out, const int32_t* __restrict__ argmax,
                      const float* __restrict__ gout, const int32_t* __restrict__ ped_start,
                      const int32_t* __restrict__ ped_end, int batch, const float2* __restrict__ Aeff,
                      const float* __restrict__ W2, float* __restrict__ dC /*[batch][512]*/, float* __restrict__ dW2,
                      float* __restrict__ db2, float* __restrict__ dAeff /*[512][2]*/, float* __restrict__ dc0,
                      float* __restrict__ dpos) {
    extern __shared__ __align__(16) float pb_smem[];
    float* sCr = pb_smem;                                 // [PB_ROWS][512]
    float* sdC = sCr + PB_ROWS * HID;                     // [PB_ROWS][512]
    float2* spos = reinterpret_cast<float2*>(sdC + PB_ROWS * HID);   // [PB_ROWS]
    float* sdpos = reinterpret_cast<float*>(spos + PB_ROWS);         // [PB_ROWS][2]
    float* ev_g = sdpos + 2 * PB_ROWS;                    // [PB_ROWS * B]
    int* ev_j = reinterpret_cast<int*>(ev_g + PB_ROWS * B);          // [PB_ROWS * B]  row of j* inside the block, -1: no gradient
    float* s_db2 = reinterpret_cast<float*>(ev_j + PB_ROWS * B);     // [B]
    const int k = threadIdx.x, lane = threadIdx.x & 31;
    const float2 a = Aeff[k];
    float w2[B], dw2[B];
#pragma unroll
    for (int b = 0; b < B; ++b) { w2[b] = W2[(int64_t)b * HID + k]; dw2[b] = 0.f; }
    float dax = 0.f, day = 0.f, dc0k = 0.f;
    if (k < B) s_db2[k] = 0.f;
    const int n_blocks = (batch + PB_IB - 1) / PB_IB;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        // rows of this block: from the first scene start >= 16 blk to the first scene start >= 16 (blk + 1)
        const int i0 = blk * PB_IB, i1 = i0 + PB_IB;
        const int lo = (ped_start[i0] == i0) ? i0 : ped_end[i0];
        const int hi = (i1 >= batch) ? batch : ((ped_start[i1] == i1) ? i1 : ped_end[i1]);
        const int rows = hi - lo;
        if (rows <= 0) continue;
        if (rows > PB_ROWS) {
            // ---- a scene too large for the tile: zero the block's rows, then the atomic scatter of the general kernel ----
            for (int r = 0; r < rows; ++r) dC[(int64_t)(lo + r) * HID + k] = 0.f;
            __syncthreads();
            for (int i = lo; i < hi; ++i) {
                const float2 pi = *reinterpret_cast<const float2*>(pos + 2 * (int64_t)i);
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const float g = gout[(int64_t)i * B + b];
                    if (!(out[(int64_t)i * B + b] > 0.f) || g == 0.f) continue;
                    const int j = argmax[(int64_t)i * B + b];
                    const float2 pj = *reinterpret_cast<const float2*>(pos + 2 * (int64_t)j);
                    const float ddx = pj.x - pi.x, ddy = pj.y - pi.y;
                    const float z = fmaf(a.x, ddx, fmaf(a.y, ddy, Cr[(int64_t)j * HID + k]));
                    if (k == 0) atomicAdd(&s_db2[b], g);
                    float dz = 0.f;
                    if (z > 0.f) {
                        dw2[b] = fmaf(g, z, dw2[b]);
                        dz = g * w2[b];
                        atomicAdd(&dC[(int64_t)j * HID + k], dz);
                        dc0k += dz;
                        dax = fmaf(dz, ddx, dax);
                        day = fmaf(dz, ddy, day);
                    }
                    if (DPOS) {
                        const float sx = warp_sum(dz * a.x), sy = warp_sum(dz * a.y);
                        if (lane == 0 && (sx != 0.f || sy != 0.f)) {
                            atomicAdd(&dpos[2 * j], sx); atomicAdd(&dpos[2 * j + 1], sy);
                            atomicAdd(&dpos[2 * i], -sx); atomicAdd(&dpos[2 * i + 1], -sy);
                        }
                    }
                }
            }
            __syncthreads();
            continue;
        }
        // ---- stage the block: Cr rows, positions, the event list; zero the dC tile ----
        __syncthreads();                                  // the previous block's tile has been written out
        for (int r = 0; r < rows; ++r) {
            sCr[r * HID + k] = Cr[(int64_t)(lo + r) * HID + k];
            sdC[r * HID + k] = 0.f;
        }
        if (k < rows) {
            spos[k] = *reinterpret_cast<const float2*>(pos + 2 * (int64_t)(lo + k));
            if (DPOS) { sdpos[2 * k] = 0.f; sdpos[2 * k + 1] = 0.f; }
        }
        for (int idx = k; idx < rows * B; idx += HID) {
            const int64_t src = (int64_t)lo * B + idx;    // rows are contiguous: [lo*B, hi*B)
            const float g = gout[src];
            const bool act = (out[src] > 0.f) && (g != 0.f);
            ev_g[idx] = g;
            ev_j[idx] = act ? (argmax[src] - lo) : -1;
            if (act) atomicAdd(&s_db2[idx % B], g);
        }
        __syncthreads();
        // ---- events: channel b outermost so that the per-channel registers are indexed statically.  (Loading four rows'
        //      metadata and Cr values up front did not help, 1.68 -> 1.86 ms: the loop is issue bound -- 16 warps x ~20
        //      instructions per event -- not latency bound.) ----
#pragma unroll
        for (int b = 0; b < B; ++b) {
            for (int r = 0; r < rows; ++r) {
                const int jr = ev_j[r * B + b];
                if (jr < 0) continue;
                const float g = ev_g[r * B + b];
                const float2 pi = spos[r], pj = spos[jr];
                const float ddx = pj.x - pi.x, ddy = pj.y - pi.y;
                const float z = fmaf(a.x, ddx, fmaf(a.y, ddy, sCr[jr * HID + k]));
                float dz = 0.f;
                if (z > 0.f) {
                    dw2[b] = fmaf(g, z, dw2[b]);
                    dz = g * w2[b];
                    sdC[jr * HID + k] += dz;
                    dax = fmaf(dz, ddx, dax);
                    day = fmaf(dz, ddy, day);
                }
                if (DPOS) {
                    const float sx = warp_sum(dz * a.x), sy = warp_sum(dz * a.y);
                    if (lane == 0 && (sx != 0.f || sy != 0.f)) {
                        atomicAdd(&sdpos[2 * jr], sx); atomicAdd(&sdpos[2 * jr + 1], sy);
                        atomicAdd(&sdpos[2 * r], -sx); atomicAdd(&sdpos[2 * r + 1], -sy);
                    }
                }
            }
        }
        // ---- the tile, once ----
        for (int r = 0; r < rows; ++r) {
            const float v = sdC[r * HID + k];
            dC[(int64_t)(lo + r) * HID + k] = v;
            dc0k += v;
        }
        if (DPOS) {
            __syncthreads();
            if (k < 2 * rows) dpos[2 * (int64_t)lo + k] = sdpos[k];
        }
    }
    __syncthreads();
#pragma unroll
    for (int b = 0; b < B; ++b)
        if (dw2[b] != 0.f) atomicAdd(&dW2[(int64_t)b * HID + k], dw2[b]);
    if (dax != 0.f) atomicAdd(&dAeff[2 * k], dax);
    if (day != 0.f) atomicAdd(&dAeff[2 * k + 1], day);
    if (dc0k != 0.f) atomicAdd(&dc0[k], dc0k);
    if (k < B && s_db2[k] != 0.f) atomicAdd(&db2[k], s_db2[k]);
}

// dc0[k] = sum_p dC[p][k]
__global__ void colsum512_kernel(const float* __restrict__ dC, int batch, float* __restrict__ dc0) {
    int k = threadIdx.x + (blockIdx.x % 2) * 256;
    int chunk = blockIdx.x / 2, nchunk = gridDim.x / 2;
    float s = 0.f;
    for (int p = chunk; p < batch; p += nchunk) s += dC[(int64_t)p * HID + k];
    atomicAdd(&dc0[k], s);
}

// chain rule through Aeff = W1e We and c0 = b1 + W1e be  (one thread per hidden unit k, then per e)
__global__ void pool_bwd_chain_kernel(const float* __restrict__ We, const float* __restrict__ be,
                                      const float* __restrict__ W1, const float* __restrict__ dAeff,
                                      const float* __restrict__ dc0, int E, int H, float* __restrict__ dW1,
                                      float* __restrict__ db1, float* __restrict__ dWe, float* __restrict__ dbe) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < HID) {
        int k = t;
        float ax = dAeff[2 * k], ay = dAeff[2 * k + 1], c = dc0[k];
        for (int e = 0; e < E; ++e)
            dW1[(int64_t)k * (E + H) + e] = fmaf(ax, We[2 * e], fmaf(ay, We[2 * e + 1], c * be[e]));
        db1[k] = c;
    } else if (t < HID + E) {
        int e = t - HID;
        float wx = 0.f, wy = 0.f, b = 0.f;
        for (int k = 0; k < HID; ++k) {
            float w = W1[(int64_t)k * (E + H) + e];
            wx = fmaf(w, dAeff[2 * k], wx);
            wy = fmaf(w, dAeff[2 * k + 1], wy);
            b = fmaf(w, dc0[k], b);
        }
        dWe[2 * e] = wx;
        dWe[2 * e + 1] = wy;
        dbe[e] = b;
    }
}

static int check_dims(int E, int H, int B) {
    SGX_REQUIRE(E >= 1 && E <= 256 && H >= 1 && H <= 1024, "pool: unsupported embedding/hidden dims E=%d H=%d", E, H);
    SGX_UNSUPPORTED(B < 8 || B % 8 != 0 || B > 1024, "pool: bottleneck_dim=%d must be a multiple of 8 in [8,1024]", B);
    return SGX_OK;
}

static int64_t ldc_for(int64_t batch) { return align_up(batch, 32); }

}  // namespace sgx

using namespace sgx;

// implemented in sgx_pool_tc.cu (bf16 operands) and sgx_pool_tc32.cu (fp16 hi/lo operand splits)
int64_t sgx_pool_bf16_prep_bytes(int E, int H, int B);
int64_t sgx_pool_bf16_ws_bytes(int64_t batch, int E, int H, int B);
int sgx_pool_bf16_prep(const float* We, const float* be, const float* W1, const float* b1, const float* W2, int E, int H,
                       int B, void* prep, cudaStream_t st);
int sgx_pool_fwd_bf16(const float* h, const float* pos, const int32_t* ped_start, const int64_t* pair_off,
                      const int32_t* tile_first, int64_t batch, int64_t n_pairs, const void* prep, const float* b2, int E,
                      int H, int B, unsigned long long* packed, void* ws, cudaStream_t st);
bool sgx_pool_tc32_supported(int E, int H, int B);
int64_t sgx_pool_tc32_prep_bytes(int E, int H, int B);
int64_t sgx_pool_tc32_ws_bytes(int64_t batch, int H);
int sgx_pool_tc32_prep(const float* We, const float* be, const float* W1, const float* b1, const float* W2, int E, int H,
                       int B, void* prep, cudaStream_t st);
int sgx_pool_fwd_tc32(const float* h, const float* pos, const int32_t* ped_start, const int64_t* pair_off,
                      const int32_t* tile_first, int64_t batch, int64_t n_pairs, const void* prep, const float* b2, int E,
                      int H, int B, unsigned long long* packed, void* ws, cudaStream_t st, bool prep_in_this_call);

// Prepared weights: everything that depends on the parameters only (folded first layer, operand images of the
// tensor-core kernels).  Built once per weight version by sgx_pool_prep; sgx_pool_fwd builds it into its workspace.
extern "C" int64_t sgx_pool_prep_bytes(int32_t E, int32_t H, int32_t B, int32_t precision) {
    if (precision == SGX_PRECISION_BF16) return sgx_pool_bf16_prep_bytes(E, H, B);
    if (precision == SGX_PRECISION_TC32) return sgx_pool_tc32_prep_bytes(E, H, B);
    return align_up(HID * 8, 256) + align_up(HID * 4, 256);
}

extern "C" int sgx_pool_prep(const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                             const float* b2, int32_t E, int32_t H, int32_t B, int32_t precision, void* prep,
                             int64_t prep_bytes, void* stream) {
    (void)b2;
    SGX_REQUIRE(We && be && W1 && b1 && W2 && prep, "sgx_pool_prep: null pointer");
    int rc = check_dims(E, H, B);
    if (rc) return rc;
    SGX_REQUIRE(precision >= SGX_PRECISION_FP32 && precision <= SGX_PRECISION_TC32, "sgx_pool_prep: unknown precision %d",
                precision);
    SGX_REQUIRE(prep_bytes >= sgx_pool_prep_bytes(E, H, B, precision), "sgx_pool_prep: buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (precision == SGX_PRECISION_BF16) return sgx_pool_bf16_prep(We, be, W1, b1, W2, E, H, B, prep, st);
    if (precision == SGX_PRECISION_TC32) return sgx_pool_tc32_prep(We, be, W1, b1, W2, E, H, B, prep, st);
    Carver c(prep);
    float2* Aeff = c.take<float2>(HID);
    float* c0 = c.take<float>(HID);
    pool_prep_kernel<<<2, 256, 0, st>>>(We, be, W1, b1, E, H, Aeff, c0);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

// workspace: [packed | 256 B of per-call statistics (zeroed with packed) | precision-specific | room for prepared weights]
static int64_t pool_call_ws(int64_t batch, int E, int H, int B, int precision) {
    int64_t n = align_up(batch * B * 8, 256) + 256;
    if (precision == SGX_PRECISION_BF16) return n + sgx_pool_bf16_ws_bytes(batch, E, H, B);
    if (precision == SGX_PRECISION_TC32) return n + sgx_pool_tc32_ws_bytes(batch, H) - 256;
    return n + align_up(ldc_for(batch) * HID * 4, 256);
}

extern "C" int64_t sgx_pool_ws_bytes(int64_t batch, int32_t E, int32_t H, int32_t B, int32_t precision) {
    return pool_call_ws(batch, E, H, B, precision) + sgx_pool_prep_bytes(E, H, B, precision);
}

extern "C" int sgx_pool_fwd_prepped(const float* h, const float* pos, const int32_t* ped_start, const int32_t* ped_end,
                                    const int64_t* pair_off, const int32_t* tile_first, int64_t batch, int64_t n_pairs,
                                    const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                                    const float* b2, int32_t E, int32_t H, int32_t B, int32_t precision,
                                    const void* prep, float* out, int32_t* argmax, void* workspace, int64_t ws_bytes,
                                    void* stream) {
    (void)ped_end;
    SGX_REQUIRE(h && pos && ped_start && pair_off && tile_first && We && be && W1 && b1 && W2 && b2 && out && workspace,
                "sgx_pool_fwd: null pointer");
    SGX_REQUIRE(batch > 0 && n_pairs >= batch, "sgx_pool_fwd: bad batch/n_pairs");
    SGX_REQUIRE(((uintptr_t)out & 15u) == 0 && ((uintptr_t)argmax & 15u) == 0 && ((uintptr_t)workspace & 15u) == 0 &&
                    ((uintptr_t)h & 15u) == 0,
                "sgx_pool_fwd: h, out, argmax and workspace must be 16-byte aligned");
    int rc = check_dims(E, H, B);
    if (rc) return rc;
    SGX_REQUIRE(precision >= SGX_PRECISION_FP32 && precision <= SGX_PRECISION_TC32, "sgx_pool_fwd: unknown precision %d",
                precision);
    SGX_REQUIRE(ws_bytes >= pool_call_ws(batch, E, H, B, precision) + (prep ? 0 : sgx_pool_prep_bytes(E, H, B, precision)),
                "sgx_pool_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    Carver ws(workspace);
    unsigned long long* packed = ws.take<unsigned long long>(batch * B);
    unsigned* stat = ws.take<unsigned>(64);
    SGX_CUDA(cudaMemsetAsync(packed, 0, (size_t)((char*)stat - (char*)packed) + 256, st));
    const bool prep_in_this_call = (prep == nullptr);
    if (!prep) {   // weights not prepared by the caller: build the images behind the per-call regions
        void* own = ws.base + pool_call_ws(batch, E, H, B, precision);
        rc = sgx_pool_prep(We, be, W1, b1, W2, b2, E, H, B, precision, own, sgx_pool_prep_bytes(E, H, B, precision), stream);
        if (rc) return rc;
        prep = own;
    }
    const int64_t n_tiles128 = (n_pairs + 127) / 128;
    if (precision == SGX_PRECISION_BF16) {
        rc = sgx_pool_fwd_bf16(h, pos, ped_start, pair_off, tile_first, batch, n_pairs, prep, b2, E, H, B, packed,
                               ws.base + ws.off, st);
        if (rc) return rc;
    } else if (precision == SGX_PRECISION_TC32) {
        rc = sgx_pool_fwd_tc32(h, pos, ped_start, pair_off, tile_first, batch, n_pairs, prep, b2, E, H, B, packed, stat, st,
                               prep_in_this_call);
        if (rc) return rc;
    } else {
        const int64_t ldc = ldc_for(batch);
        float* Ct = ws.take<float>(ldc * HID);
        Carver pc(const_cast<void*>(prep));
        const float2* Aeff = pc.take<float2>(HID);
        const float* c0 = pc.take<float>(HID);
        // C^T[k][p] = c0[k] + sum_h W1[k][E+h] * h[p][h]
        rc = gemm(W1 + E, E + H, 1, h, 1, H, Ct, ldc, HID, batch, H, 0, 0, st, c0);
        if (rc) return rc;
        cudaEvent_t ev0, ev1;
        profile_events(&ev0, &ev1);
        if (ev0 && ev1) SGX_CUDA(cudaEventRecord(ev0, st));
        if (B % 16 == 0) {
            constexpr int BC = 16, R = 4, T = 256;
            dim3 grid((unsigned)((n_pairs + T * R - 1) / (T * R)), B / BC);
            pool_pair_kernel<BC, R, T><<<grid, T, 0, st>>>(Ct, ldc, pos, ped_start, pair_off, tile_first, n_tiles128,
                                                           (int)batch, n_pairs, Aeff, W2, b2, B, packed);
        } else {
            constexpr int BC = 8, R = 8, T = 256;
            dim3 grid((unsigned)((n_pairs + T * R - 1) / (T * R)), B / BC);
            pool_pair_kernel<BC, R, T><<<grid, T, 0, st>>>(Ct, ldc, pos, ped_start, pair_off, tile_first, n_tiles128,
                                                           (int)batch, n_pairs, Aeff, W2, b2, B, packed);
        }
        SGX_LAUNCH_CHECK();
        if (ev0 && ev1) SGX_CUDA(cudaEventRecord(ev1, st));
    }
    SGX_CUDA(launch_pdl(pool_unpack_kernel, dim3(blocks_for(batch * B / 4, 256)), dim3(256), 0, st, true, packed, batch * B, out,
                        argmax));
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

extern "C" int sgx_pool_fwd(const float* h, const float* pos, const int32_t* ped_start, const int32_t* ped_end,
                            const int64_t* pair_off, const int32_t* tile_first, int64_t batch, int64_t n_pairs,
                            const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                            const float* b2, int32_t E, int32_t H, int32_t B, int32_t precision, float* out,
                            int32_t* argmax, void* workspace, int64_t ws_bytes, void* stream) {
    return sgx_pool_fwd_prepped(h, pos, ped_start, ped_end, pair_off, tile_first, batch, n_pairs, We, be, W1, b1, W2, b2,
                                E, H, B, precision, nullptr, out, argmax, workspace, ws_bytes, stream);
}

/* 1 when precision TC32 has a kernel for these dims on the current device */
extern "C" int sgx_pool_tc32_available(int32_t E, int32_t H, int32_t B) {
    return (sgx_has_tcgen05() && sgx_pool_tc32_supported(E, H, B)) ? 1 : 0;
}

extern "C" int64_t sgx_pool_bwd_ws_bytes(int64_t batch, int32_t E, int32_t H, int32_t B) {
    (void)E; (void)H; (void)B;
    // Cr, dC [batch][512]; Aeff, dAeff [512][2]; c0, dc0 [512]
    return 2 * align_up(batch * HID * 4, 256) + 2 * align_up(HID * 8, 256) + 2 * align_up(HID * 4, 256);
}

static int pool_bwd_impl(const float* h, const float* pos, const float* out, const int32_t* argmax,
                         const float* grad_out, const int32_t* ped_start, const int32_t* ped_end, int64_t batch,
                         const float* We, const float* be, const float* W1, const float* b1, const float* W2, int32_t E,
                         int32_t H, int32_t B, float* grad_h, float* grad_pos, float* grad_We, float* grad_be,
                         float* grad_W1, float* grad_b1, float* grad_W2, float* grad_b2, void* workspace, int64_t ws_bytes,
                         void* stream) {
    SGX_REQUIRE(h && pos && out && argmax && grad_out && We && be && W1 && b1 && W2 && grad_h && grad_We &&
                    grad_be && grad_W1 && grad_b1 && grad_W2 && grad_b2 && workspace,
                "sgx_pool_bwd: null pointer");
    int rc = check_dims(E, H, B);
    if (rc) return rc;
    SGX_REQUIRE(ws_bytes >= sgx_pool_bwd_ws_bytes(batch, E, H, B), "sgx_pool_bwd: workspace too small");
    SGX_REQUIRE(batch > 0 && batch < ((int64_t)1 << 31), "sgx_pool_bwd: bad batch");
    cudaStream_t st = (cudaStream_t)stream;
    Carver ws(workspace);
    float* Cr = ws.take<float>(batch * HID);
    float* dC = ws.take<float>(batch * HID);
    float2* Aeff = ws.take<float2>(HID);
    float* dAeff = ws.take<float>(2 * HID);
    float* c0 = ws.take<float>(HID);
    float* dc0 = ws.take<float>(HID);
    // the scene-owned kernel (no global atomics) needs the scene bounds and has instances for the shipped bottlenecks
    const bool blocks = ped_start && ped_end && (B == 8 || B == 48);
    if (!blocks) SGX_CUDA(cudaMemsetAsync(dC, 0, (size_t)batch * HID * 4, st));
    SGX_CUDA(cudaMemsetAsync(dAeff, 0, 2 * HID * 4, st));
    SGX_CUDA(cudaMemsetAsync(dc0, 0, HID * 4, st));
    SGX_CUDA(cudaMemsetAsync(grad_W2, 0, (size_t)B * HID * 4, st));
    SGX_CUDA(cudaMemsetAsync(grad_b2, 0, (size_t)B * 4, st));
    if (grad_pos) SGX_CUDA(cudaMemsetAsync(grad_pos, 0, (size_t)batch * 2 * 4, st));
    pool_prep_kernel<<<2, 256, 0, st>>>(We, be, W1, b1, E, H, Aeff, c0);
    SGX_LAUNCH_CHECK();
    // Cr[p][k] = c0[k] + sum_h h[p][h] W1[k][E+h]   (row-major: a warp reads one ped's 512 values coalesced)
    rc = gemm(h, H, 1, W1 + E, 1, E + H, Cr, HID, batch, HID, H, 0, 0, st, nullptr, c0);
    if (rc) return rc;
    if (blocks) {
        const size_t smem = ((size_t)2 * PB_ROWS * HID + 4 * PB_ROWS + 2 * (size_t)PB_ROWS * B + B) * sizeof(float);
        const int grid = (int)std::min<int64_t>((batch + PB_IB - 1) / PB_IB, 148);
#define SGX_PB_LAUNCH(BB, DP)                                                                                              \
    do {                                                                                                                   \
        SGX_CUDA(cudaFuncSetAttribute(pool_bwd_block_kernel<BB, DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        pool_bwd_block_kernel<BB, DP><<<grid, HID, smem, st>>>(Cr, pos, out, argmax, grad_out, ped_start, ped_end,          \
                                                               (int)batch, Aeff, W2, dC, grad_W2, grad_b2, dAeff, dc0,      \
                                                               grad_pos);                                                   \
    } while (0)
        if (B == 8) { if (grad_pos) SGX_PB_LAUNCH(8, true); else SGX_PB_LAUNCH(8, false); }
        else { if (grad_pos) SGX_PB_LAUNCH(48, true); else SGX_PB_LAUNCH(48, false); }
#undef SGX_PB_LAUNCH
        SGX_LAUNCH_CHECK();
    } else {
        SGX_REQUIRE(grad_pos != nullptr, "sgx_pool_bwd: grad_pos is optional only with scene bounds and bottleneck 8 / 48");
        const size_t smem = ((size_t)B * HID + 2 * HID + B) * sizeof(float);
        SGX_UNSUPPORTED(smem > 200 * 1024, "sgx_pool_bwd: bottleneck_dim=%d too large for the backward kernel (max 96)", B);
        SGX_CUDA(cudaFuncSetAttribute(pool_bwd_event_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = (int)std::min<int64_t>((batch + 7) / 8, 148 * 2);
        pool_bwd_event_kernel<<<grid, 256, smem, st>>>(Cr, pos, out, argmax, grad_out, (int)batch, B, Aeff, W2, dC, grad_W2,
                                                      grad_b2, dAeff, grad_pos);
        SGX_LAUNCH_CHECK();
        int nchunk = (int)std::min<int64_t>(batch, 128);
        colsum512_kernel<<<2 * nchunk, 256, 0, st>>>(dC, (int)batch, dc0);
        SGX_LAUNCH_CHECK();
    }
    // dW1[:, E:] = dC^T h     (K = batch: split-K inside gemm)
    rc = gemm(dC, 1, HID, h, H, 1, grad_W1 + E, E + H, HID, H, batch, 0, 0, st);
    if (rc) return rc;
    // dh = dC W1[:, E:]
    rc = gemm(dC, HID, 1, W1 + E, E + H, 1, grad_h, H, batch, H, HID, 0, 0, st);
    if (rc) return rc;
    pool_bwd_chain_kernel<<<blocks_for(HID + E, 256), 256, 0, st>>>(We, be, W1, dAeff, dc0, E, H, grad_W1, grad_b1,
                                                                     grad_We, grad_be);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

extern "C" int sgx_pool_bwd(const float* h, const float* pos, const float* out, const int32_t* argmax,
                            const float* grad_out, int64_t batch, const float* We, const float* be, const float* W1,
                            const float* b1, const float* W2, const float* b2, int32_t E, int32_t H, int32_t B,
                            float* grad_h, float* grad_pos, float* grad_We, float* grad_be, float* grad_W1,
                            float* grad_b1, float* grad_W2, float* grad_b2, void* workspace, int64_t ws_bytes,
                            void* stream) {
    (void)b2;
    SGX_REQUIRE(grad_pos != nullptr, "sgx_pool_bwd: null pointer");
    return pool_bwd_impl(h, pos, out, argmax, grad_out, nullptr, nullptr, batch, We, be, W1, b1, W2, E, H, B, grad_h, grad_pos,
                         grad_We, grad_be, grad_W1, grad_b1, grad_W2, grad_b2, workspace, ws_bytes, stream);
}

// The same backward with the scene bounds of every pedestrian (ped_start / ped_end from sgx_schedule_fill): the events
// are accumulated scene by scene in shared memory instead of with global atomics.  grad_pos may be null (positions that
// need no gradient: the generator's pooling on observed positions) -- the per-event position reduction is then skipped.
extern "C" int sgx_pool_bwd_scenes(const float* h, const float* pos, const float* out, const int32_t* argmax,
                                   const float* grad_out, const int32_t* ped_start, const int32_t* ped_end, int64_t batch,
                                   const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                                   const float* b2, int32_t E, int32_t H, int32_t B, float* grad_h, float* grad_pos,
                                   float* grad_We, float* grad_be, float* grad_W1, float* grad_b1, float* grad_W2,
                                   float* grad_b2, void* workspace, int64_t ws_bytes, void* stream) {
    (void)b2;
    SGX_REQUIRE(ped_start && ped_end, "sgx_pool_bwd_scenes: null pointer");
    return pool_bwd_impl(h, pos, out, argmax, grad_out, ped_start, ped_end, batch, We, be, W1, b1, W2, E, H, B, grad_h,
                         grad_pos, grad_We, grad_be, grad_W1, grad_b1, grad_W2, grad_b2, workspace, ws_bytes, stream);
}
