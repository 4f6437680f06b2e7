// Error string, version and the host-side scene schedule.
//
// The schedule replaces the per-scene python loop headers of the reference
// (sgan/models.py:507-510, 256-262, 639-644): seq_start_end is read ONCE per minibatch on the
// host and turned into flat per-ped arrays that every kernel indexes without a device sync.
#include <algorithm>
#include <atomic>
#include <numeric>
#include <queue>
#include <string>
#include <vector>

#include "sgx_common.cuh"

namespace sgx {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static thread_local cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
void profile_events(cudaEvent_t* start, cudaEvent_t* stop) { *start = g_ev_start; *stop = g_ev_stop; }
}  // namespace sgx

namespace sgx {
// Process-wide switches, set ONCE by the host side at load time (group_gan_gcn_gat_b200/_lib.py resolves the SGX_*
// environment variables) or by a test through sgx_set_option -- no getenv on any call path.
static std::atomic<int> g_opt_lstm_tc{1};
bool opt_lstm_tc() { return g_opt_lstm_tc.load(std::memory_order_relaxed) != 0; }
static std::atomic<int> g_opt_graph_tc{1};
bool opt_graph_tc() { return g_opt_graph_tc.load(std::memory_order_relaxed) != 0; }
static std::atomic<int> g_opt_pdl{1};
bool opt_pdl() { return g_opt_pdl.load(std::memory_order_relaxed) != 0; }
#ifdef SGX_AB_VARIANTS
static std::atomic<int> g_opt_gat_mma{1}, g_opt_gcn_mma{1};
bool opt_gat_mma() { return g_opt_gat_mma.load(std::memory_order_relaxed) != 0; }
bool opt_gcn_mma() { return g_opt_gcn_mma.load(std::memory_order_relaxed) != 0; }
#endif
}  // namespace sgx

extern "C" int sgx_set_option(const char* name, int32_t value) {
    SGX_REQUIRE(name != nullptr, "sgx_set_option: null name");
    const std::string n(name);
    if (n == "lstm_tc") { sgx::g_opt_lstm_tc.store(value); return SGX_OK; }
    if (n == "graph_tc") { sgx::g_opt_graph_tc.store(value); return SGX_OK; }
    if (n == "pdl") { sgx::g_opt_pdl.store(value); return SGX_OK; }
#ifdef SGX_AB_VARIANTS
    if (n == "gat_mma") { sgx::g_opt_gat_mma.store(value); return SGX_OK; }
    if (n == "gcn_mma") { sgx::g_opt_gcn_mma.store(value); return SGX_OK; }
#endif
    sgx::set_error("sgx_set_option: unknown option '%s' in this build", name);
    return SGX_ERR_UNSUPPORTED;
}

extern "C" long long sgx_launch_count(void) { return sgx::g_launches.load(); }
extern "C" int sgx_profile_events(void* ev_start, void* ev_stop) {
    sgx::g_ev_start = (cudaEvent_t)ev_start;
    sgx::g_ev_stop = (cudaEvent_t)ev_stop;
    return SGX_OK;
}

extern "C" const char* sgx_last_error(void) { return sgx::g_err; }
extern "C" int sgx_version(void) { return 100; }

static int validate(const int64_t* sse, int64_t S, int64_t* batch_out, int64_t* max_n, int64_t* pairs) {
    SGX_REQUIRE(sse != nullptr && S >= 1, "seq_start_end must hold at least one scene");
    int64_t cur = 0, mx = 0, p = 0;
    for (int64_t s = 0; s < S; ++s) {
        int64_t a = sse[2 * s], b = sse[2 * s + 1];
        SGX_REQUIRE(a == cur, "scene %lld starts at %lld, expected %lld: scenes must tile [0,batch) contiguously",
                    (long long)s, (long long)a, (long long)cur);
        SGX_REQUIRE(b > a, "scene %lld is empty (start=%lld end=%lld)", (long long)s, (long long)a, (long long)b);
        int64_t n = b - a;
        mx = std::max(mx, n);
        p += n * n;
        cur = b;
    }
    SGX_REQUIRE(cur < (int64_t)1 << 31, "batch too large for int32 indexing");
    *batch_out = cur;
    *max_n = mx;
    *pairs = p;
    return SGX_OK;
}

extern "C" int sgx_schedule_stats(const int64_t* sse, int64_t S, int64_t* stats) {
    SGX_REQUIRE(stats != nullptr, "null stats");
    int64_t batch, mx, pairs;
    int rc = validate(sse, S, &batch, &mx, &pairs);
    if (rc) return rc;
    stats[0] = batch;
    stats[1] = mx;
    stats[2] = pairs;
    stats[3] = (pairs + 127) / 128;
    stats[4] = S;
    return SGX_OK;
}

extern "C" int sgx_schedule_fill(const int64_t* sse, int64_t S, int32_t* scene_start, int32_t* ped_start,
                                 int32_t* ped_end, int64_t* pair_off, int32_t* tile_first) {
    int64_t batch, mx, pairs;
    int rc = validate(sse, S, &batch, &mx, &pairs);
    if (rc) return rc;
    SGX_REQUIRE(scene_start && ped_start && ped_end && pair_off && tile_first, "null output array");
    int64_t acc = 0, next_tile = 0;
    for (int64_t s = 0; s < S; ++s) {
        int64_t a = sse[2 * s], b = sse[2 * s + 1], n = b - a;
        scene_start[s] = (int32_t)a;
        for (int64_t i = a; i < b; ++i) {
            ped_start[i] = (int32_t)a;
            ped_end[i] = (int32_t)b;
            pair_off[i] = acc;
            // every 128-pair tile whose first pair lies in ped i's row block
            while (next_tile * 128 < acc + n) tile_first[next_tile++] = (int32_t)i;
            acc += n;
        }
    }
    scene_start[S] = (int32_t)batch;
    pair_off[batch] = acc;
    return SGX_OK;
}

// sgx_schedule_fill + the per-pedestrian scene index + (when every scene fits `chunk_cap`) the chunk boundaries of
// sgx_schedule_chunks, in ONE pass over the scenes: what the host side builds per minibatch.
extern "C" int sgx_schedule_build(const int64_t* sse, int64_t S, int32_t* scene_start, int32_t* ped_start,
                                  int32_t* ped_end, int64_t* pair_off, int32_t* tile_first, int32_t* ped_scene,
                                  int32_t chunk_cap, int32_t* chunk_scene, int64_t* n_chunks) {
    int64_t batch, mx, pairs;
    int rc = validate(sse, S, &batch, &mx, &pairs);
    if (rc) return rc;
    SGX_REQUIRE(scene_start && ped_start && ped_end && pair_off && tile_first && ped_scene, "null output array");
    const bool chunks = chunk_scene != nullptr && n_chunks != nullptr && chunk_cap >= 1 && mx <= chunk_cap;
    int64_t acc = 0, next_tile = 0, c = 0, fill = 0;
    if (chunks) chunk_scene[0] = 0;
    for (int64_t s = 0; s < S; ++s) {
        const int64_t a = sse[2 * s], b = sse[2 * s + 1], n = b - a;
        scene_start[s] = (int32_t)a;
        if (chunks) {
            if (fill + n > chunk_cap) { chunk_scene[++c] = (int32_t)s; fill = 0; }
            fill += n;
        }
        for (int64_t i = a; i < b; ++i) {
            ped_start[i] = (int32_t)a;
            ped_end[i] = (int32_t)b;
            ped_scene[i] = (int32_t)s;
            pair_off[i] = acc;
            while (next_tile * 128 < acc + n) tile_first[next_tile++] = (int32_t)i;
            acc += n;
        }
    }
    scene_start[S] = (int32_t)batch;
    pair_off[batch] = acc;
    if (chunks) chunk_scene[++c] = (int32_t)S;
    if (n_chunks) *n_chunks = chunks ? c : 0;
    return SGX_OK;
}

// Longest-processing-time-first partition of scenes over ranks, cost = N^2 (pairwise pooling).
extern "C" int sgx_schedule_partition(const int64_t* sse, int64_t S, int32_t world, int32_t* rank_of_scene,
                                      int64_t* rank_cost) {
    int64_t batch, mx, pairs;
    int rc = validate(sse, S, &batch, &mx, &pairs);
    if (rc) return rc;
    SGX_REQUIRE(world >= 1 && rank_of_scene != nullptr, "bad world size / null output");
    std::vector<int64_t> order(S);
    std::iota(order.begin(), order.end(), 0);
    auto cost = [&](int64_t s) { int64_t n = sse[2 * s + 1] - sse[2 * s]; return n * n; };
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return cost(x) > cost(y); });
    typedef std::pair<int64_t, int32_t> Load;  // (cost so far, rank): min-heap, ties -> lower rank
    std::priority_queue<Load, std::vector<Load>, std::greater<Load>> heap;
    for (int32_t r = 0; r < world; ++r) heap.push(Load(0, r));
    std::vector<int64_t> tot(world, 0);
    for (int64_t k = 0; k < S; ++k) {
        Load l = heap.top();
        heap.pop();
        rank_of_scene[order[k]] = l.second;
        l.first += cost(order[k]);
        tot[l.second] = l.first;
        heap.push(l);
    }
    if (rank_cost)
        for (int32_t r = 0; r < world; ++r) rank_cost[r] = tot[r];
    return SGX_OK;
}

// Greedy packing of consecutive whole scenes into chunks of at most `cap` pedestrians (one warp / CTA team per
// chunk in the fused GAT kernel).  h_chunk_scene [S+1]: scene index where each chunk starts, closed by S.
// Returns the number of chunks in *h_n_chunks, or SGX_ERR_UNSUPPORTED if a scene exceeds `cap`.
extern "C" int sgx_schedule_chunks(const int64_t* sse, int64_t S, int32_t cap, int32_t* chunk_scene,
                                   int64_t* n_chunks) {
    int64_t batch, mx, pairs;
    int rc = validate(sse, S, &batch, &mx, &pairs);
    if (rc) return rc;
    SGX_REQUIRE(chunk_scene && n_chunks && cap >= 1, "sgx_schedule_chunks: bad arguments");
    SGX_UNSUPPORTED(mx > cap, "largest scene has %lld pedestrians, chunk capacity is %d", (long long)mx, cap);
    int64_t c = 0, fill = 0;
    chunk_scene[0] = 0;
    for (int64_t s = 0; s < S; ++s) {
        int64_t n = sse[2 * s + 1] - sse[2 * s];
        if (fill + n > cap) { chunk_scene[++c] = (int32_t)s; fill = 0; }
        fill += n;
    }
    chunk_scene[++c] = (int32_t)S;
    *n_chunks = c;
    return SGX_OK;
}
