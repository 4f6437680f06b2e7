// Warp-level fp32-accurate GEMM on the tensor cores for the per-scene-chunk graph kernels (sgx_gat.cu, sgx_gcn.cu):
// mma.sync m16n8k8 TF32 with every operand split hi + lo (hi = top 19 bits), hi*hi + lo*hi + hi*lo accumulated in fp32
// (~7e-7 relative).  One warp multiplies its chunk's 32 rows; rows are independent, so rows of dead lanes may hold
// anything.
#pragma once
#include <stdint.h>

namespace sgx {

__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(v) & 0xFFFFE000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
// round-to-nearest variant: hi and lo are both rounded (cvt.rna), so the residual error is unbiased and 4x smaller
// (2^-22 per operand).  The truncating split above leaves a one-sided 2^-20 error that adds up linearly over very long
// reductions (K = 10^4 .. 10^6 in the split-K parameter-gradient GEMMs).
__device__ __forceinline__ void split_tf32_rn(float v, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    const float r = v - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// C [16 MT x 8 NT] = A [16 MT x K] . B [K x 8 NT] for one warp (MT = 2: the 32 slots of a warp chunk); A rows in shared memory (stride SA floats), B row-major
// (stride SB).  store(mt, nt, acc) receives the m16n8 accumulator fragment: acc[0..1] = row mt*16 + lane/4, columns
// nt*8 + 2 (lane%4) + {0,1}; acc[2..3] = the same columns of row + 8.  A __syncwarp precedes the stores of an m-tile,
// so C may overwrite the A rows of that m-tile.  (mt0, mts): the m-tiles mt0, mt0 + mts, ... only -- two warps that share
// a chunk's buffers take one m-tile each (the callers put a barrier of the pair around the call).
template <int K, int NT, int SA, int SB, int MT = 2, class Store>
__device__ __forceinline__ void warp_gemm_3xtf32(const float* __restrict__ A, const float* __restrict__ B, int lane,
                                                 Store&& store, int mt0 = 0, int mts = 1) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int mt = mt0; mt < MT; mt += mts) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll 1
        for (int ks = 0; ks < K / 8; ++ks) {
            const float* ar = A + (mt * 16 + g) * SA + ks * 8 + t;
            uint32_t ah[4], al[4];
            split_tf32(ar[0], ah[0], al[0]);
            split_tf32(ar[8 * SA], ah[1], al[1]);
            split_tf32(ar[4], ah[2], al[2]);
            split_tf32(ar[8 * SA + 4], ah[3], al[3]);
            const float* br = B + (ks * 8 + t) * SB + g;
            // n-tiles in groups: the three products of one n-tile accumulate into the same fragment, so issuing them
            // back to back serialises on the MMA latency; interleaving NG n-tiles puts NG independent MMAs between.
            constexpr int NG = (NT % 5 == 0) ? 5 : (NT % 4 == 0) ? 4 : (NT % 3 == 0) ? 3 : (NT % 2 == 0) ? 2 : 1;
#pragma unroll
            for (int n0 = 0; n0 < NT; n0 += NG) {
                uint32_t bh0[NG], bl0[NG], bh1[NG], bl1[NG];
#pragma unroll
                for (int j = 0; j < NG; ++j) {
                    split_tf32(br[(n0 + j) * 8], bh0[j], bl0[j]);
                    split_tf32(br[4 * SB + (n0 + j) * 8], bh1[j], bl1[j]);
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], al, bh0[j], bh1[j]);
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, bl0[j], bl1[j]);
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, bh0[j], bh1[j]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) store(mt, nt, acc[nt]);
    }
    __syncwarp();
}


// C [32 x 8 NT] = A [32 x K] . Bt  with Bt(k, n) = B[n * SB + k]: the right operand is a row-major block whose ROWS are
// indexed by the output column (dX = dY W^T read straight from the forward's [K_in][N_out] weight block).  Same
// fragment conventions as warp_gemm_3xtf32.  B fragment loads are 2-way bank conflicted for SB = 24 / 88 (g and g + 4).
template <int K, int NT, int SA, int SB, int MT = 2, class Store>
__device__ __forceinline__ void warp_gemm_3xtf32_bt(const float* __restrict__ A, const float* __restrict__ B, int lane,
                                                    Store&& store, int mt0 = 0, int mts = 1) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll 1
    for (int mt = mt0; mt < MT; mt += mts) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll 1
        for (int ks = 0; ks < K / 8; ++ks) {
            const float* ar = A + (mt * 16 + g) * SA + ks * 8 + t;
            uint32_t ah[4], al[4];
            split_tf32(ar[0], ah[0], al[0]);
            split_tf32(ar[8 * SA], ah[1], al[1]);
            split_tf32(ar[4], ah[2], al[2]);
            split_tf32(ar[8 * SA + 4], ah[3], al[3]);
            const float* br = B + g * SB + ks * 8 + t;
            constexpr int NG = (NT % 5 == 0) ? 5 : (NT % 4 == 0) ? 4 : (NT % 3 == 0) ? 3 : (NT % 2 == 0) ? 2 : 1;
#pragma unroll
            for (int n0 = 0; n0 < NT; n0 += NG) {
                uint32_t bh0[NG], bl0[NG], bh1[NG], bl1[NG];
#pragma unroll
                for (int j = 0; j < NG; ++j) {
                    split_tf32(br[(n0 + j) * 8 * SB], bh0[j], bl0[j]);
                    split_tf32(br[(n0 + j) * 8 * SB + 4], bh1[j], bl1[j]);
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], al, bh0[j], bh1[j]);
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, bl0[j], bl1[j]);
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, bh0[j], bh1[j]);
            }
        }
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) store(mt, nt, acc[nt]);
    }
    __syncwarp();
}

// C [M x 8 NT] = A^T . B over the warp's 32 rows (K = 32): A [32][SA] (columns 0..M-1 used), B [32][SB] -- the
// per-chunk parameter gradient dW = input^T dY.  store(m0, nt, acc): acc[0..1] = rows m0 + lane/4, columns
// nt*8 + 2 (lane%4) + {0,1}; acc[2..3] = row + 8.  Rows >= M of the last m-tile are computed from clamped columns and
// must be dropped by the caller.  Every one of the 32 rows of A and B must be finite (zero for dead lanes).
template <int M, int NT, int SA, int SB, class Store>
__device__ __forceinline__ void warp_gemm_3xtf32_at(const float* __restrict__ A, const float* __restrict__ B, int lane,
                                                    Store&& store, int mt0 = 0, int mts = 1) {
    const int g = lane >> 2, t = lane & 3;
    constexpr int MT = (M + 15) / 16;
#pragma unroll 1
    for (int mt = mt0; mt < MT; mt += mts) {
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
        const int m_lo = (mt * 16 + g) < M ? (mt * 16 + g) : (M - 1), m_hi = (mt * 16 + g + 8) < M ? (mt * 16 + g + 8) : (M - 1);
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
            const float* ar = A + (ks * 8 + t) * SA;
            uint32_t ah[4], al[4];
            split_tf32(ar[m_lo], ah[0], al[0]);
            split_tf32(ar[m_hi], ah[1], al[1]);
            split_tf32(ar[4 * SA + m_lo], ah[2], al[2]);
            split_tf32(ar[4 * SA + m_hi], ah[3], al[3]);
            const float* br = B + (ks * 8 + t) * SB + g;
            constexpr int NG = (NT % 3 == 0) ? 3 : (NT % 2 == 0) ? 2 : 1;
#pragma unroll
            for (int n0 = 0; n0 < NT; n0 += NG) {
                uint32_t bh0[NG], bl0[NG], bh1[NG], bl1[NG];
#pragma unroll
                for (int j = 0; j < NG; ++j) {
                    split_tf32(br[(n0 + j) * 8], bh0[j], bl0[j]);
                    split_tf32(br[4 * SB + (n0 + j) * 8], bh1[j], bl1[j]);
                }
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], al, bh0[j], bh1[j]);
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, bl0[j], bl1[j]);
#pragma unroll
                for (int j = 0; j < NG; ++j) mma_tf32(acc[n0 + j], ah, bh0[j], bh1[j]);
            }
        }
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) store(mt * 16, nt, acc[nt]);
    }
}

// The same product shared by a PAIR of warps along the output columns: role 0 takes the first ceil(NT / 2) n-tiles, role 1
// the rest (every m-tile each) -- an even split whatever M is, where splitting the m-tiles of M = 16 / 40 / 72 is not.
template <int M, int NT, int SA, int SB, class Store>
__device__ __forceinline__ void warp_gemm_3xtf32_at_pair(const float* __restrict__ A, const float* __restrict__ B, int lane,
                                                         Store&& store, int role) {
    constexpr int N0 = (NT + 1) / 2, N1 = NT - N0;
    if (role == 0) {
        warp_gemm_3xtf32_at<M, N0, SA, SB>(A, B, lane, store);
    } else if constexpr (N1 > 0) {
        warp_gemm_3xtf32_at<M, N1, SA, SB>(A, B + N0 * 8, lane,
                                           [&](int m0, int nt, const float (&c)[4]) { store(m0, nt + N0, c); });
    }
}

}  // namespace sgx
