// Shared helpers for libsgx_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <algorithm>

#include "../../include/sgx.h"

namespace sgx {

void set_error(const char* fmt, ...);
void count_launch();
// optional events recorded around the dominant pooling kernel (bench instrumentation); null when unset
void profile_events(cudaEvent_t* start, cudaEvent_t* stop);
// process-wide switches (sgx_set_option; resolved once by the host side, never getenv on a call path)
bool opt_lstm_tc();
bool opt_graph_tc();      // tcgen05 GATEncoder / GCNModule forwards (default 1); 0: the mma.sync kernels (parity tests)
bool opt_pdl();           // programmatic dependent launch between the kernels of the SGAN-P chain (default 1)
#ifdef SGX_AB_VARIANTS
bool opt_gat_mma();
bool opt_gcn_mma();
#endif

#define SGX_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            sgx::set_error(__VA_ARGS__);       \
            return SGX_ERR_INVALID;            \
        }                                      \
    } while (0)

#define SGX_UNSUPPORTED(cond, ...)             \
    do {                                       \
        if (cond) {                            \
            sgx::set_error(__VA_ARGS__);       \
            return SGX_ERR_UNSUPPORTED;        \
        }                                      \
    } while (0)

#define SGX_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            sgx::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__,   \
                           __LINE__);                                                           \
            return SGX_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

// every kernel launch of the library is followed by this: counts launches (sgx_launch_count) and surfaces errors
#define SGX_LAUNCH_CHECK()              \
    do {                                \
        sgx::count_launch();            \
        SGX_CUDA(cudaGetLastError());   \
    } while (0)

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
static inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

// Programmatic dependent launch (PDL).  The kernels of one generator forward (encoder recurrence -> pooling statistics ->
// h image -> pooling -> unpack -> context MLP -> decoder recurrence) are queued back to back on one stream; each of
// them calls pdl_trigger() first thing and pdl_wait() before its first access to global memory another kernel of the
// chain may have written (or may still read), so the CTAs of kernel n+1 are resident, with barriers initialised and
// TMEM requested, while the last CTAs of kernel n drain.  griddepcontrol.wait returns when the prerequisite grid has
// COMPLETED and its writes are visible, and every kernel of the chain waits before it exits, so completion is
// transitive along the chain.  A kernel launched without the attribute, or after something that is not a kernel
// (memset, event record), is ordered as usual and both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (pdl && opt_pdl()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// carve typed regions out of a caller-provided workspace
struct Carver {
    char* base;
    int64_t off;
    explicit Carver(void* p) : base((char*)p), off(0) {}
    template <typename T>
    T* take(int64_t n) {
        T* r = (T*)(base + off);
        off += align_up(n * (int64_t)sizeof(T), 256);
        return r;
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Blackwell packed fp32 FMA (SASS FFMA2): d = a * b + c on both halves; halves the FMA issue slots of the
// CUDA-core GEMV loops.  The two halves accumulate even / odd k separately and are added at the end.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long*>(&a), rb = *reinterpret_cast<unsigned long long*>(&b),
                       rc = *reinterpret_cast<unsigned long long*>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// generic fp32 GEMM used inside the library (same as the exported sgx_gemm)
// bias_m / bias_n (nullable): per-row / per-column bias added to C (disables split-K)
int gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
         int64_t M, int64_t N, int64_t K, int accumulate, int relu, cudaStream_t st, const float* bias_m = nullptr,
         const float* bias_n = nullptr);

}  // namespace sgx
