// Scene schedule built ON THE DEVICE (a11: the ragged segmentation the reference walks with
// `for (start, end) in seq_start_end` + two .item() calls per scene, sgan/models.py:507-510, 256-262, 639-644).
//
// sgx_schedule_build (sgx_api.cu) fills the same arrays on the host in one pass over the pedestrians; at 65 k scenes /
// 206 k pedestrians that pass plus the upload of its 7 MB of index arrays was 0.75 ms of host time per minibatch, in
// front of the first pooling launch of an end-to-end evaluation step.  Here the host only validates seq_start_end and
// takes its totals (sgx_schedule_stats, one pass over the SCENES), uploads the 16 bytes per scene, and four small
// launches derive everything else:
//   scan A / B / C : pair_base[s] = sum_{s' < s} N_s'^2 (exclusive scan over scenes: 1024 scenes per block, block sums
//                    scanned by one block), scene_start[s]
//   fill           : thread per pedestrian p -- binary search of its scene over the scene starts -- ped_start, ped_end,
//                    ped_scene, pair_off[p] = pair_base[s] + (p - start_s) N_s;  thread per 128-pair tile t -- binary
//                    search of the scene holding pair 128 t over pair_base -- tile_first[t] = start_s + (128 t - pair_base[s]) / N_s
// Integer work, bit-identical to the host pass (tests/test_gpu_schedule.py).
// seq_start_end may be handed over as a DEVICE pointer or as a pointer into PINNED host memory (cudaHostAlloc: mapped
// into the device address space under unified addressing): scan A reads it once, over PCIe in the second case, and
// leaves a device copy for the other launches.  No copy-engine transfer is involved, so the schedule of minibatch k
// never queues behind the H2D prefetch of minibatch k+1 (a 1 MB cudaMemcpyAsync issued after the 50 MB prefetch stalled
// the compute stream for 0.75 ms per step in the pipelined evaluation loop).  sgx_fetch_pinned is the same idea for the
// small chunk lists of the graph kernels.
#include "sgx_common.cuh"

namespace sgx {
namespace sched {

constexpr int SCAN_THREADS = 256, SCAN_PER_THREAD = 4, SCAN_BLOCK = SCAN_THREADS * SCAN_PER_THREAD;

// inclusive scan of one int64 per thread over the block (SCAN_THREADS threads); *total = sum over the block
__device__ __forceinline__ long long block_inclusive_scan(long long v, long long* s_warp, long long* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const long long u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
    }
    if (lane == 31) s_warp[warp] = v;
    __syncthreads();
    long long base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        const long long x = s_warp[w];
        if (w < warp) base += x;
        tot += x;
    }
    *total = tot;
    return v + base;
}

// WRITE = false: block_sum[b] = sum of N^2 over the block's scenes.  WRITE = true: block_sum holds the EXCLUSIVE scan of
// those sums; writes pair_base[s] and scene_start[s] (and the closing entries [S]).
// (WRITE = false additionally copies the scenes it reads to sse_copy: the device copy the later launches use)
template <bool WRITE>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(const int64_t* __restrict__ sse, int64_t S, long long* __restrict__ block_sum, long long* __restrict__ pair_base,
            int32_t* __restrict__ scene_start, int64_t batch, int64_t n_pairs, int64_t* __restrict__ sse_copy) {
    __shared__ long long s_warp[SCAN_THREADS / 32];
    const int64_t s0 = ((int64_t)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_PER_THREAD;
    long long c[SCAN_PER_THREAD], tot = 0;
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u) {
        const int64_t s = s0 + u;
        c[u] = 0;
        if (s < S) {
            const longlong2 ab = *reinterpret_cast<const longlong2*>(sse + 2 * s);
            const long long n = ab.y - ab.x;
            c[u] = n * n;
            if (WRITE) scene_start[s] = (int32_t)ab.x;
            else *reinterpret_cast<longlong2*>(sse_copy + 2 * s) = ab;
        }
        tot += c[u];
    }
    long long block_total;
    const long long incl = block_inclusive_scan(tot, s_warp, &block_total);
    if (!WRITE) {
        if (threadIdx.x == 0) block_sum[blockIdx.x] = block_total;
        return;
    }
    long long excl = block_sum[blockIdx.x] + incl - tot;
#pragma unroll
    for (int u = 0; u < SCAN_PER_THREAD; ++u) {
        const int64_t s = s0 + u;
        if (s < S) pair_base[s] = excl;
        excl += c[u];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        pair_base[S] = n_pairs;
        scene_start[S] = (int32_t)batch;
    }
}

// exclusive scan of block_sum[0 .. n) in place, one block
__global__ void __launch_bounds__(SCAN_THREADS)
scan_blocks_kernel(long long* __restrict__ block_sum, int64_t n) {
    __shared__ long long s_warp[SCAN_THREADS / 32];
    long long carry = 0;
    for (int64_t b0 = 0; b0 < n; b0 += SCAN_THREADS) {
        const int64_t i = b0 + threadIdx.x;
        const long long v = i < n ? block_sum[i] : 0;
        long long total;
        const long long incl = block_inclusive_scan(v, s_warp, &total);
        if (i < n) block_sum[i] = carry + incl - v;
        carry += total;
        __syncthreads();                       // s_warp is reused by the next trip
    }
}

__global__ void __launch_bounds__(256)
fill_kernel(const int64_t* __restrict__ sse, const long long* __restrict__ pair_base, int64_t S, int64_t batch,
            int64_t n_tiles, int64_t n_pairs, int32_t* __restrict__ ped_start, int32_t* __restrict__ ped_end,
            int32_t* __restrict__ ped_scene, int64_t* __restrict__ pair_off, int32_t* __restrict__ tile_first) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < batch) {
        // scene of pedestrian t: the largest s with start_s <= t (scenes tile [0, batch) without gaps: validated on the host)
        int64_t lo = 0, hi = S - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (sse[2 * mid] <= t) lo = mid; else hi = mid - 1;
        }
        const longlong2 ab = *reinterpret_cast<const longlong2*>(sse + 2 * lo);
        ped_start[t] = (int32_t)ab.x;
        ped_end[t] = (int32_t)ab.y;
        ped_scene[t] = (int32_t)lo;
        pair_off[t] = pair_base[lo] + (t - ab.x) * (ab.y - ab.x);
        if (t == batch - 1) pair_off[batch] = n_pairs;
    } else if (t - batch < n_tiles) {
        // pedestrian owning the first pair of 128-pair tile tt: scene = the largest s with pair_base[s] <= 128 tt
        const int64_t tt = t - batch;
        const long long q = 128 * tt;
        int64_t lo = 0, hi = S - 1;
        while (lo < hi) {
            const int64_t mid = (lo + hi + 1) >> 1;
            if (pair_base[mid] <= q) lo = mid; else hi = mid - 1;
        }
        const longlong2 ab = *reinterpret_cast<const longlong2*>(sse + 2 * lo);
        tile_first[tt] = (int32_t)(ab.x + (q - pair_base[lo]) / (ab.y - ab.x));
    }
}

// 16-byte words from mapped pinned host memory (or anywhere) to device memory, by the SMs instead of a copy engine
__global__ void __launch_bounds__(256)
fetch_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace sched
}  // namespace sgx

using namespace sgx;

extern "C" int64_t sgx_schedule_device_ws_bytes(int64_t S) {
    if (S < 1) return 256;
    const int64_t n_blocks = (S + sched::SCAN_BLOCK - 1) / sched::SCAN_BLOCK;
    return align_up((S + 1) * 8, 256) + align_up(n_blocks * 8, 256) + align_up(S * 16, 256);
}

extern "C" int sgx_fetch_pinned(void* d_dst, const void* src, int64_t nbytes, void* stream) {
    SGX_REQUIRE(d_dst && src && nbytes >= 0 && nbytes % 16 == 0 && ((uintptr_t)d_dst & 15u) == 0 && ((uintptr_t)src & 15u) == 0,
                "sgx_fetch_pinned: pointers and size must be multiples of 16 bytes");
    if (nbytes == 0) return SGX_OK;
    sched::fetch_kernel<<<(unsigned)std::min<int64_t>(blocks_for(nbytes / 16, 256), 592), 256, 0, (cudaStream_t)stream>>>(
        (const uint4*)src, (uint4*)d_dst, nbytes / 16);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}

extern "C" int sgx_schedule_build_device(const int64_t* sse_src, int64_t S, int64_t batch, int64_t n_pairs, int64_t n_tiles,
                                         int32_t* scene_start, int32_t* ped_start, int32_t* ped_end, int64_t* pair_off,
                                         int32_t* tile_first, int32_t* ped_scene, void* workspace, int64_t ws_bytes,
                                         void* stream) {
    SGX_REQUIRE(sse_src && scene_start && ped_start && ped_end && pair_off && tile_first && ped_scene && workspace,
                "sgx_schedule_build_device: null pointer");
    SGX_REQUIRE(S >= 1 && batch >= S && batch < ((int64_t)1 << 31) && n_pairs >= batch && n_tiles == (n_pairs + 127) / 128,
                "sgx_schedule_build_device: totals do not describe a schedule (take them from sgx_schedule_stats)");
    SGX_REQUIRE(ws_bytes >= sgx_schedule_device_ws_bytes(S), "sgx_schedule_build_device: workspace too small");
    SGX_REQUIRE(((uintptr_t)sse_src & 15u) == 0, "sgx_schedule_build_device: seq_start_end must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    Carver ws(workspace);
    long long* pair_base = ws.take<long long>(S + 1);
    const int64_t n_blocks = (S + sched::SCAN_BLOCK - 1) / sched::SCAN_BLOCK;
    long long* block_sum = ws.take<long long>(n_blocks);
    int64_t* d_sse = ws.take<int64_t>(2 * S);
    sched::scan_kernel<false><<<(unsigned)n_blocks, sched::SCAN_THREADS, 0, st>>>(sse_src, S, block_sum, pair_base, scene_start,
                                                                                 batch, n_pairs, d_sse);
    SGX_LAUNCH_CHECK();
    sched::scan_blocks_kernel<<<1, sched::SCAN_THREADS, 0, st>>>(block_sum, n_blocks);
    SGX_LAUNCH_CHECK();
    sched::scan_kernel<true><<<(unsigned)n_blocks, sched::SCAN_THREADS, 0, st>>>(d_sse, S, block_sum, pair_base, scene_start,
                                                                                batch, n_pairs, nullptr);
    SGX_LAUNCH_CHECK();
    sched::fill_kernel<<<blocks_for(batch + n_tiles, 256), 256, 0, st>>>(d_sse, pair_base, S, batch, n_tiles, n_pairs, ped_start,
                                                                         ped_end, ped_scene, pair_off, tile_first);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}
