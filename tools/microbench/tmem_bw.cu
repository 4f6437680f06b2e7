// TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) throughput per SM on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a --cudart shared -O3 -o /tmp/tmem_bw tmem_bw.cu && /tmp/tmem_bw
// (dynamic cudart, binary kept out of the repo tree)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(uint32_t a, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
        "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(a));
}
__device__ __forceinline__ void ld32_pack(uint32_t a, uint32_t (&v)[32]) {   // 64 columns of 16-bit data
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
        "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(a));
}
__device__ __forceinline__ void st32(uint32_t a, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
        "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
        "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
        "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]),
        "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}

// mode 0: ld x32, wait every load; 1: ld x32, wait every 4 loads; 2: st x32; 3: ld pack16
__global__ void bw_kernel(int mode, int iters, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = slot + ((uint32_t)((warp & 3) << 5) << 16);
    uint32_t acc = 0;
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                ld32(tm + ((it * 4 + b) * 32) % 512, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                acc ^= v[0] ^ v[31];
            }
        } else if (mode == 1) {
            uint32_t w0[32], w1[32], w2[32], w3[32];
            ld32(tm + 0 + (it & 3) * 128 % 512, w0);
            ld32(tm + 32 + (it & 3) * 128 % 512, w1);
            ld32(tm + 64 + (it & 3) * 128 % 512, w2);
            ld32(tm + 96 + (it & 3) * 128 % 512, w3);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            acc ^= w0[0] ^ w1[5] ^ w2[7] ^ w3[31];
        } else if (mode == 2) {
#pragma unroll
            for (int b = 0; b < 4; ++b) st32(tm + ((it * 4 + b) * 32) % 512, v);
            asm volatile("tcgen05.wait::st.sync.aligned;");
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                ld32_pack(tm + ((it * 4 + b) * 64) % 512, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;");
                acc ^= v[0] ^ v[31];
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot));
}

int main() {
    long long* cyc;
    uint32_t* sink;
    cudaMalloc(&cyc, 148 * 8);
    cudaMalloc(&sink, 148 * 1024 * 4);
    const int iters = 2000;
    const char* names[] = {"ld.32x32b.x32 wait each", "ld.32x32b.x32 wait/4", "st.32x32b.x32", "ld.x32.pack16 wait each"};
    for (int mode = 0; mode < 4; ++mode)
        for (int nw : {1, 4, 8, 16}) {
            bw_kernel<<<148, nw * 32>>>(mode, iters, cyc, sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double bytes = (double)iters * 4 * 32 * 32 * 4 * nw;   // per SM (register bytes moved)
            printf("%-26s warps=%2d  cycles=%lld  B/cycle/SM=%.1f\n", names[mode], nw, h[0], bytes / h[0]);
        }
    return 0;
}
