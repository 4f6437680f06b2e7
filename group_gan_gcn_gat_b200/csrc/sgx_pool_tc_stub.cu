// temporary stub until the tcgen05 kernel lands
#include "sgx_common.cuh"
int sgx_pool_fwd_bf16(const float*, const float*, const int32_t*, const int64_t*, const int32_t*, int64_t, int64_t,
                      const float*, const float*, const float*, const float*, const float*, const float*, int, int, int,
                      unsigned long long*, void*, int64_t, cudaStream_t) {
    sgx::set_error("bf16 pooling kernel not built");
    return SGX_ERR_UNSUPPORTED;
}
int64_t sgx_pool_bf16_ws_bytes(int64_t, int, int, int) { return 0; }
extern "C" int sgx_has_tcgen05(void) { return 0; }
