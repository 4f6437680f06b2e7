// Group structure from the datasets_group labels -- integer/bit exact.
//
// Reference: sgan/models.py:263-267 (M_intra, A_intra), 271-278 (unique rows + reverse => groups
// ordered by ascending smallest member) and the identical copy at 654-680 (GCNModule).
// The reference builds N x N boolean matrices per scene; here every pedestrian gets
//   leader     = smallest global index j of its scene with the same non-zero label (itself if label == 0)
//   group_size = number of members,  group_id = rank of its leader among the scene's leaders
// which is all the GAT/GCN kernels need.  sgx_group_dense re-materialises the reference's dense
// matrices for the parity tests.
#include "sgx_common.cuh"

namespace sgx {

__global__ void group_leader_kernel(const float* __restrict__ labels, const int32_t* __restrict__ ped_start,
                                    const int32_t* __restrict__ ped_end, int batch, int32_t* __restrict__ leader,
                                    int32_t* __restrict__ group_size) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const int s = ped_start[p], e = ped_end[p];
    const float lab = labels[p];
    int lead = p, cnt = 1;
    if (lab != 0.f) {  // float compare, exactly like (A_g == B_g) & (A_g != 0)
        cnt = 0;
        bool found = false;
        for (int q = s; q < e; ++q) {
            bool same = (labels[q] == lab);
            cnt += same ? 1 : 0;
            if (same && !found) { lead = q; found = true; }
        }
        if (cnt == 0) { cnt = 1; lead = p; }  // NaN label: only the diagonal survives
    }
    leader[p] = lead;
    group_size[p] = cnt;
}

__global__ void group_id_kernel(const int32_t* __restrict__ leader, const int32_t* __restrict__ ped_start, int batch,
                                int32_t* __restrict__ group_id) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= batch) return;
    const int s = ped_start[p], l = leader[p];
    int g = 0;
    for (int q = s; q < l; ++q) g += (leader[q] == q) ? 1 : 0;
    group_id[p] = g;
}

__global__ void group_count_kernel(const int32_t* __restrict__ leader, const int32_t* __restrict__ scene_start,
                                   int n_scenes, int32_t* __restrict__ n_group) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_scenes) return;
    int g = 0;
    for (int q = scene_start[s]; q < scene_start[s + 1]; ++q) g += (leader[q] == q) ? 1 : 0;
    n_group[s] = g;
}

__global__ void group_dense_kernel(const float* __restrict__ labels, const int32_t* __restrict__ group_size,
                                   const int32_t* __restrict__ group_id, int start, int n, uint8_t* M, float* A,
                                   uint8_t* R, float* Rn) {
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)n * n) return;
    int i = (int)(idx / n), j = (int)(idx % n);
    float li = labels[start + i], lj = labels[start + j];
    bool m = ((li == lj) && (li != 0.f)) || (i == j);
    if (M) M[idx] = m ? 1 : 0;
    // normalize(): rowsum (int64) -> float -> pow(-1); bool * float
    if (A) A[idx] = m ? __frcp_rn((float)group_size[start + i]) : 0.f;
    // row g of R = members of group g (groups ordered by smallest member); only rows < n_group are written
    int g = i;  // reuse the n x n index space: row g, column j
    bool r = (group_id[start + j] == g);
    if (r) {
        if (R) R[idx] = 1;
        if (Rn) Rn[idx] = __frcp_rn((float)group_size[start + j]);
    } else {
        if (R) R[idx] = 0;
        if (Rn) Rn[idx] = 0.f;
    }
}

}  // namespace sgx

extern "C" int sgx_group_ids(const float* labels, const int32_t* ped_start, const int32_t* ped_end,
                             const int32_t* scene_start, int64_t batch, int64_t n_scenes, int32_t* leader,
                             int32_t* group_size, int32_t* group_id, int32_t* n_group, void* stream) {
    using namespace sgx;
    SGX_REQUIRE(labels && ped_start && ped_end && leader && group_size, "sgx_group_ids: null pointer");
    SGX_REQUIRE(batch > 0 && batch < ((int64_t)1 << 31), "sgx_group_ids: bad batch");
    cudaStream_t st = (cudaStream_t)stream;
    group_leader_kernel<<<blocks_for(batch, 128), 128, 0, st>>>(labels, ped_start, ped_end, (int)batch, leader,
                                                               group_size);
    SGX_LAUNCH_CHECK();
    if (group_id) {
        group_id_kernel<<<blocks_for(batch, 128), 128, 0, st>>>(leader, ped_start, (int)batch, group_id);
        SGX_LAUNCH_CHECK();
    }
    if (n_group) {
        SGX_REQUIRE(scene_start && n_scenes > 0, "sgx_group_ids: n_group needs scene_start");
        group_count_kernel<<<blocks_for(n_scenes, 128), 128, 0, st>>>(leader, scene_start, (int)n_scenes, n_group);
        SGX_LAUNCH_CHECK();
    }
    return SGX_OK;
}

extern "C" int sgx_group_dense(const float* labels, const int32_t* leader, const int32_t* group_size,
                               const int32_t* group_id, int64_t start, int64_t end, uint8_t* M, float* A, uint8_t* R,
                               float* Rn, void* stream) {
    using namespace sgx;
    (void)leader;
    SGX_REQUIRE(labels && group_size && group_id && end > start, "sgx_group_dense: bad arguments");
    int64_t n = end - start;
    SGX_REQUIRE(n <= 16384, "sgx_group_dense: scene too large for a dense dump");
    group_dense_kernel<<<blocks_for(n * n, 256), 256, 0, (cudaStream_t)stream>>>(labels, group_size, group_id,
                                                                                (int)start, (int)n, M, A, R, Rn);
    SGX_LAUNCH_CHECK();
    return SGX_OK;
}
