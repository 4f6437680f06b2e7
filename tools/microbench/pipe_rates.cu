// Issue-rate microbenchmark for the instructions of the tc32 pooling epilogue (sm_100a):
//   nvcc -gencode arch=compute_100a,code=sm_100a --cudart shared -O3 -o /tmp/pipe_rates tools/microbench/pipe_rates.cu
// (build OUTSIDE the repo tree or delete the binary afterwards; link cudart dynamically.)
// Prints warp-instructions per clock per SM for 4 / 8 / 16 resident warps per SM, 8 independent chains per thread.
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>

template <int OP>
__global__ void rate_kernel(float* out, long long* cyc, int iters) {
    float x[8];
    uint32_t u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 0.001f + i; u[i] = threadIdx.x * 77u + i; }
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) {          // F2FP.RELU.F16.F32.PACK_AB.RZ
                asm volatile("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(__uint_as_float(u[i])));
            } else if (OP == 1) {   // F2FP.BF16 relu (the bf16 kernel's convert)
                asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(__uint_as_float(u[i])));
            } else if (OP == 2) {   // FHFMA (mixed f16 x f16 + f32)
                asm volatile("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %1;\n\tfma.rn.f32.f16 %0, a, b, %0;\n\t}" : "+f"(x[i]) : "r"(u[i]));
            } else if (OP == 3) {   // HADD2.F32 (f16 -> f32)
                asm volatile("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %1;\n\tcvt.f32.f16 %0, a;\n\t}" : "=f"(x[i]) : "r"(u[i] + (uint32_t)__float_as_uint(x[i])));
            } else if (OP == 4) {   // FADD
                asm volatile("add.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(1.0001f));
            } else if (OP == 5) {   // LOP3
                asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            } else if (OP == 6) {   // PRMT
                asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            } else if (OP == 7) {   // FMNMX
                asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(__uint_as_float(u[i])));
            } else if (OP == 8) {   // FFMA
                asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(x[i]) : "f"(1.0001f));
            } else if (OP == 9) {   // F2FP.F16 rn (no relu)
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(__uint_as_float(u[i])));
            }
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(u[i]);
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 148 * 8);
    const int iters = 4096;
    printf("%-34s", name);
    for (int warps : {4, 8, 16}) {
        rate_kernel<OP><<<148, warps * 32>>>(out, cyc, iters);
        rate_kernel<OP><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
        printf("  %2d warps: %.3f winst/clk/SM", warps, (double)warps * iters * 8 / avg);
    }
    printf("\n");
}

int main() {
    run<0>("F2FP.RELU.F16.F32.PACK_AB.RZ");
    run<9>("F2FP.F16.F32.PACK_AB");
    run<1>("F2FP.RELU.BF16.F32.PACK_AB");
    run<2>("FHFMA");
    run<3>("HADD2.F32 (cvt.f32.f16)");
    run<4>("FADD");
    run<8>("FFMA");
    run<5>("LOP3");
    run<6>("PRMT");
    run<7>("FMNMX");
    return 0;
}
