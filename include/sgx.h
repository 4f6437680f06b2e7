/* sgx.h -- C ABI of libsgx_b200.so: the sgan social-interaction hot path on B200 (sm_100a).
 *
 * The reference (peaceminusones/Group-GAN-GCN-GAT) has no FFI: its boundary for this path is the
 * nn.Module API of sgan/models.py.  Each entry point below replaces the body of one reference
 * method; the Python modules in group_gan_gcn_gat_b200/ keep the reference signatures and
 * state_dict names and call these through ctypes + torch.library.
 *
 * Conventions
 *   - plain C symbols, POD arguments; every pointer is a raw CUDA device pointer unless the
 *     parameter name starts with `h_` (host memory).  No torch types.
 *   - the CALLER owns and allocates every buffer (outputs and workspace; sizes via *_ws_bytes),
 *     the library never allocates device memory and never synchronises: all work is queued on
 *     `stream` (a cudaStream_t passed as void*).
 *   - fp32 tensors are dense row-major; "batch" = total pedestrians of the minibatch,
 *     scenes are the ragged segments seq_start_end[s] = (start, end) of sgan/models.py:507-510.
 *   - return value: 0 = ok, <0 = error (see SGX_ERR_*); sgx_last_error() gives the text.
 *     No C++ exception crosses the ABI.
 *   - re-entrant per stream; no global state except the last-error string (thread local).
 */
#ifndef SGX_H_
#define SGX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SGX_OK 0
#define SGX_ERR_INVALID (-1)     /* bad shape / size / null pointer / non-contiguous scenes */
#define SGX_ERR_UNSUPPORTED (-2) /* valid request this build has no kernel for               */
#define SGX_ERR_CUDA (-3)        /* a CUDA runtime call failed                               */

#define SGX_PRECISION_FP32 0     /* CUDA-core path, 1e-5 relative parity                     */
#define SGX_PRECISION_BF16 1     /* tcgen05/TMEM path, bf16 operands, 2e-2 parity            */
#define SGX_PRECISION_TC32 2     /* tcgen05/TMEM path, fp16 hi+lo operand splits with fp32    */
                                 /* accumulation: 1e-5 relative parity (the fp32 contract)   */

#define SGX_POOL_HIDDEN 512      /* hard-coded mid width of mlp_pre_pool, sgan/models.py:473 */

const char* sgx_last_error(void);
int sgx_version(void);
/* 1 when the current CUDA device can run the tcgen05 kernels of this build (compute capability 10.x) */
int sgx_has_tcgen05(void);
/* Process-wide switches, meant to be set once at load time (the Python host side resolves the SGX_* environment
 * variables there; nothing in the library reads the environment).  "lstm_tc" (default 1): tcgen05 recurrence kernels
 * for large inference batches, 0 = CUDA-core kernels for every batch.  "graph_tc" (default 1): tcgen05 kernels for the
 * single-launch GATEncoder / GCNModule forwards, 0 = the warp-level mma.sync kernels (parity tests).  Builds with -DSGX_AB_VARIANTS also know
 * "gat_mma" / "gcn_mma" (0 = the CUDA-core GEMV single-launch kernels kept for A/B timing). */
int sgx_set_option(const char* name, int32_t value);
/* number of kernels this library has launched in this process (bench.py's gpu_launches evidence) */
long long sgx_launch_count(void);
/* bench instrumentation: when both are non-null CUDA events (cudaEvent_t), sgx_pool_fwd records them on its
 * stream directly around its dominant kernel (the pair kernel); pass nulls to switch off.  Thread local. */
int sgx_profile_events(void* ev_start, void* ev_stop);

/* ---------------------------------------------------------------------------------------------
 * Scene schedule (host).  Replaces the per-scene `.item()` loop headers of
 * sgan/models.py:507-510, 256-262, 639-644 and the seq_start_end layout of
 * sgan/data/trajectories_GCN.py:19-22,36.  Built once per minibatch, reused by every call.
 *
 * sgx_schedule_stats : validates h_seq_start_end (int64 [S,2], scenes must tile [0,batch)
 *                      contiguously, every scene >= 1 ped) and returns
 *                      h_stats[0]=batch, [1]=max scene size, [2]=sum N^2 (ordered pairs),
 *                      [3]=number of 128-pair tiles, [4]=S.
 * sgx_schedule_fill  : fills the host arrays (caller uploads them):
 *      h_scene_start int32 [S+1]
 *      h_ped_start   int32 [batch]   start of the ped's scene
 *      h_ped_end     int32 [batch]   end of the ped's scene
 *      h_pair_off    int64 [batch+1] prefix sum of scene sizes over peds: ordered pair (i,j)
 *                                    has flat index pair_off[i] + (j - ped_start[i])
 *      h_tile_first  int32 [n_tiles] ped owning the first pair of every 128-pair tile
 * sgx_schedule_partition : LPT split of scenes over `world` ranks by cost N^2;
 *      h_rank_of_scene int32 [S].  Deterministic: ties broken by scene index.
 */
int sgx_schedule_stats(const int64_t* h_seq_start_end, int64_t n_scenes, int64_t* h_stats);
int sgx_schedule_fill(const int64_t* h_seq_start_end, int64_t n_scenes, int32_t* h_scene_start,
                      int32_t* h_ped_start, int32_t* h_ped_end, int64_t* h_pair_off, int32_t* h_tile_first);
/* sgx_schedule_fill + h_ped_scene int32 [batch] (scene index of every pedestrian: the per-scene -> per-ped repeat of
 * add_noise, sgan/models.py:837-846, and of evaluate_helper, scripts/evaluate_model.py:58-69) + the chunk boundaries of
 * sgx_schedule_chunks(cap = chunk_cap) in one pass; *h_n_chunks = 0 (chunk_scene untouched) when a scene exceeds
 * chunk_cap or h_chunk_scene is null. */
int sgx_schedule_build(const int64_t* h_seq_start_end, int64_t n_scenes, int32_t* h_scene_start,
                       int32_t* h_ped_start, int32_t* h_ped_end, int64_t* h_pair_off, int32_t* h_tile_first,
                       int32_t* h_ped_scene, int32_t chunk_cap, int32_t* h_chunk_scene, int64_t* h_n_chunks);
/* The same arrays as sgx_schedule_build (without the chunk list), derived ON THE DEVICE from seq_start_end (int64 [S,2],
 * 16-byte aligned: a device pointer OR a pointer into pinned host memory, which the first launch reads over PCIe -- no
 * copy-engine transfer that could queue behind the prefetch of the next minibatch) and the totals of sgx_schedule_stats
 * (batch = stats[0], n_pairs = stats[2], n_tiles = stats[3]): the host validates and hands over 16 bytes per scene instead
 * of filling and uploading ~34 bytes per pedestrian.  Four small launches on `stream`, no synchronisation; bit-identical
 * to the host pass.  All output pointers are device pointers; workspace: sgx_schedule_device_ws_bytes(S). */
int64_t sgx_schedule_device_ws_bytes(int64_t n_scenes);
int sgx_schedule_build_device(const int64_t* seq_start_end_dev_or_pinned, int64_t n_scenes, int64_t batch, int64_t n_pairs,
                              int64_t n_tiles, int32_t* d_scene_start, int32_t* d_ped_start, int32_t* d_ped_end,
                              int64_t* d_pair_off, int32_t* d_tile_first, int32_t* d_ped_scene, void* workspace,
                              int64_t ws_bytes, void* stream);
/* nbytes (multiple of 16) from pinned host memory (or device memory) to d_dst by a kernel instead of a copy engine: for
 * the small index lists of a minibatch that must not wait behind the H2D prefetch of the next one. */
int sgx_fetch_pinned(void* d_dst, const void* src_pinned, int64_t nbytes, void* stream);
int sgx_schedule_partition(const int64_t* h_seq_start_end, int64_t n_scenes, int32_t world,
                           int32_t* h_rank_of_scene, int64_t* h_rank_cost);
/* Greedy packing of consecutive whole scenes into chunks of <= cap pedestrians (fused GAT kernel: one warp per
 * chunk).  h_chunk_scene int32 [S+1] receives the first scene of every chunk, closed by S; *h_n_chunks the count.
 * SGX_ERR_UNSUPPORTED when a scene is larger than cap. */
int sgx_schedule_chunks(const int64_t* h_seq_start_end, int64_t n_scenes, int32_t cap, int32_t* h_chunk_scene,
                        int64_t* h_n_chunks);

/* ---------------------------------------------------------------------------------------------
 * Group structure.  Replaces sgan/models.py:263-278 (GATEncoder) == 654-680 (GCNModule):
 * M_intra = (g_i == g_j & g_i != 0) | eye, its row normalisation, unique-rows + reverse.
 *   labels      fp32 [batch]  group label of the last observed frame (float compare, 0 = none)
 *   leader      int32 [batch] global index of the smallest member of the ped's group
 *   group_size  int32 [batch]
 *   group_id    int32 [batch] scene-local id, groups numbered by ascending smallest member
 *   n_group     int32 [S]
 * sgx_group_dense writes, for ONE scene, the dense matrices the reference materialises
 * (for bit-exact parity tests): M uint8 [N,N], A fp32 [N,N] (= M * fl(1/rowsum)),
 * R uint8 [N,N] and Rn fp32 [N,N] whose first n_group rows are R_intra / its row normalisation
 * (row g = members of group g) and whose other rows are zero.  Any pointer may be null.
 */
int sgx_group_ids(const float* labels, const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                  int64_t batch, int64_t n_scenes, int32_t* leader, int32_t* group_size, int32_t* group_id,
                  int32_t* n_group, void* stream);
int sgx_group_dense(const float* labels, const int32_t* leader, const int32_t* group_size, const int32_t* group_id,
                    int64_t start, int64_t end, uint8_t* M, float* A, uint8_t* R, float* Rn, void* stream);

/* ---------------------------------------------------------------------------------------------
 * PoolHiddenNet.forward (sgan/models.py:497-549):
 *   out[i,b] = max_j ReLU(b2[b] + W2[b,:] . ReLU(b1 + W1 [We (P_j - P_i) + be ; h_j]))   j in scene(i)
 *   h [batch,H]  pos [batch,2]  We [E,2] be [E]  W1 [512,E+H] b1 [512]  W2 [B,512] b2 [B]
 *   out fp32 [batch,B]   argmax int32 [batch,B] (global index j attaining the max; ties -> larger j)
 * precision: SGX_PRECISION_FP32 (CUDA cores; any E, H, B multiple of 8),
 *            SGX_PRECISION_TC32 (tcgen05, fp32-grade; (H,B) = (32,8): sgx_pool_tc32_available) or
 *            SGX_PRECISION_BF16 (tcgen05; (H,B) in {(32,8),(48,48)}) ... see DESIGN.md.
 * workspace: sgx_pool_ws_bytes(batch, E, H, B, precision).  h, out, argmax and workspace: 16-byte aligned (vector
 *   loads / stores; every torch allocation is).
 * Prepared weights: everything that depends only on the parameters (the folded first layer
 *   Aeff = W1[:, :E] We, c = W1[:, :E] be + b1, and the operand images of the tensor-core kernels) can be built
 *   once per weight version with sgx_pool_prep into a caller-owned buffer of sgx_pool_prep_bytes() and passed
 *   to sgx_pool_fwd_prepped (prep = NULL: built per call inside the workspace, which is what sgx_pool_fwd does).
 * A NaN in any pair's output makes out[i,b] NaN (like torch.max).
 * Backward (argmax-sparse, fp32): grads of sum(out*grad_out) w.r.t. every input; gradient
 * buffers are OVERWRITTEN (not accumulated).  Reference: autograd through models.py:516-541.
 */
int64_t sgx_pool_ws_bytes(int64_t batch, int32_t E, int32_t H, int32_t B, int32_t precision);
int sgx_pool_fwd(const float* h, const float* pos, const int32_t* ped_start, const int32_t* ped_end,
                 const int64_t* pair_off, const int32_t* tile_first, int64_t batch, int64_t n_pairs,
                 const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                 const float* b2, int32_t E, int32_t H, int32_t B, int32_t precision, float* out,
                 int32_t* argmax, void* workspace, int64_t ws_bytes, void* stream);
int64_t sgx_pool_prep_bytes(int32_t E, int32_t H, int32_t B, int32_t precision);
int sgx_pool_prep(const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                  const float* b2, int32_t E, int32_t H, int32_t B, int32_t precision, void* prep,
                  int64_t prep_bytes, void* stream);
int sgx_pool_fwd_prepped(const float* h, const float* pos, const int32_t* ped_start, const int32_t* ped_end,
                         const int64_t* pair_off, const int32_t* tile_first, int64_t batch, int64_t n_pairs,
                         const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                         const float* b2, int32_t E, int32_t H, int32_t B, int32_t precision, const void* prep,
                         float* out, int32_t* argmax, void* workspace, int64_t ws_bytes, void* stream);
int sgx_pool_tc32_available(int32_t E, int32_t H, int32_t B);
int64_t sgx_pool_bwd_ws_bytes(int64_t batch, int32_t E, int32_t H, int32_t B);
int sgx_pool_bwd(const float* h, const float* pos, const float* out, const int32_t* argmax, const float* grad_out,
                 int64_t batch, const float* We, const float* be, const float* W1, const float* b1,
                 const float* W2, const float* b2, int32_t E, int32_t H, int32_t B, float* grad_h,
                 float* grad_pos, float* grad_We, float* grad_be, float* grad_W1, float* grad_b1,
                 float* grad_W2, float* grad_b2, void* workspace, int64_t ws_bytes, void* stream);
/* The same backward given the scene bounds of every pedestrian (ped_start / ped_end, as for sgx_pool_fwd): the argmax
 * pairs never leave a scene, so the hidden-layer gradients are accumulated scene by scene in shared memory and written
 * once instead of scattered with global atomics (bottleneck 8 or 48; other dims take the path of sgx_pool_bwd).
 * grad_pos may be NULL when the positions need no gradient (PoolHiddenNet on observed positions). */
int sgx_pool_bwd_scenes(const float* h, const float* pos, const float* out, const int32_t* argmax,
                        const float* grad_out, const int32_t* ped_start, const int32_t* ped_end, int64_t batch,
                        const float* We, const float* be, const float* W1, const float* b1, const float* W2,
                        const float* b2, int32_t E, int32_t H, int32_t B, float* grad_h, float* grad_pos,
                        float* grad_We, float* grad_be, float* grad_W1, float* grad_b1, float* grad_W2,
                        float* grad_b2, void* workspace, int64_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * make_mlp([in, mid, out], activation relu, batch_norm 0, dropout 0) forward (sgan/models.py:7-20) in one launch,
 * inference only:   out = ReLU(W2 ReLU(W1 [xa ; xb] + b1) + b2)
 * at its call sites sgan/models.py:898 (mlp_decoder_context(cat(final_encoder_h, pool_h)), the SGAN-P wiring),
 * :165-166 (decoder.mlp(cat(decoder_h, pool_h))) and :990 (real_classifier).  xa [batch,da], xb [batch,db] or NULL with
 * db = 0 (the concatenation is folded into the load), W1 [HID,da+db] b1 [HID] W2 [OUT,HID] b2 [OUT] (nn.Linear layout),
 * out [batch,OUT].  Built for da+db in {32,40,48}, HID 64, OUT in {1,24,32}: sgx_mlp2_supported.
 */
int sgx_mlp2_supported(int32_t IN, int32_t HID, int32_t OUT);
int sgx_mlp2_fwd(const float* xa, int32_t da, const float* xb, int32_t db, int64_t batch, const float* W1,
                 const float* b1, const float* W2, const float* b2, int32_t HID, int32_t OUT, float* out,
                 void* stream);

/* ---------------------------------------------------------------------------------------------
 * GCNModule.forward (sgan/models.py:628-712) with gcn_layers = 2:
 *   x [batch,IN]  W0 [IN,HID] W1 [HID,OUT] (intra)  V0 [OUT,HID] V1 [HID,OUT] (inter)
 *   Wo [FIN,2*OUT] bo [FIN]  ->  out [batch,FIN]
 * Needs the group structure of sgx_group_ids.  Gradient buffers are overwritten.
 */
int64_t sgx_gcn_module_ws_bytes(int64_t batch, int64_t n_scenes, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN);
int sgx_gcn_module_fwd(const float* x, const int32_t* leader, const int32_t* group_size, const int32_t* ped_start,
                       const int32_t* ped_end, const int32_t* scene_start, const int32_t* n_group, int64_t batch,
                       int64_t n_scenes, const float* W0, const float* W1, const float* V0, const float* V1,
                       const float* Wo, const float* bo, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN,
                       float* out, void* workspace, int64_t ws_bytes, void* stream);
/* Same forward in ONE launch for batches whose scenes all have <= 32 pedestrians (chunks from sgx_schedule_chunks). */
int sgx_gcn_module_fused_fwd(const float* x, const int32_t* leader, const int32_t* group_size,
                             const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                             const int32_t* chunk_scene, int64_t n_chunks, const float* W0, const float* W1,
                             const float* V0, const float* V1, const float* Wo, const float* bo, int32_t IN,
                             int32_t HID, int32_t OUT, int32_t FIN, float* out, void* stream);
/* ... with the group structure derived in the kernel from the datasets_group labels [batch] (what sgx_group_ids would
 * compute, sgan/models.py:654-680, bit for bit): no leader / size arrays, no extra pass.  tcgen05 kernel.
 * prep: NULL, or the fp16 hi/lo weight images built by sgx_gcn_module_tc_prep for exactly these weights
 * (sgx_gcn_module_tc_prep_bytes bytes, device memory; -1 for dims without a kernel instance) -- the host side caches it
 * per weight version; without it every launch rebuilds the images (~8 us). */
int sgx_gcn_module_fused_fwd_labels(const float* x, const float* labels, const int32_t* ped_start,
                                    const int32_t* ped_end, const int32_t* scene_start, const int32_t* chunk_scene,
                                    int64_t n_chunks, const float* W0, const float* W1, const float* V0,
                                    const float* V1, const float* Wo, const float* bo, int32_t IN, int32_t HID,
                                    int32_t OUT, int32_t FIN, const void* prep, float* out, void* stream);
int64_t sgx_gcn_module_tc_prep_bytes(int32_t IN, int32_t HID, int32_t OUT, int32_t FIN);
int sgx_gcn_module_tc_prep(const float* W0, const float* W1, const float* V0, const float* V1, const float* Wo,
                           const float* bo, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, void* prep, void* stream);
int sgx_gcn_module_bwd(const float* x, const float* grad_out, const int32_t* leader, const int32_t* group_size,
                       const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                       const int32_t* n_group, int64_t batch, int64_t n_scenes, const float* W0, const float* W1,
                       const float* V0, const float* V1, const float* Wo, const float* bo, int32_t IN, int32_t HID,
                       int32_t OUT, int32_t FIN, float* grad_x, float* grad_W0, float* grad_W1, float* grad_V0,
                       float* grad_V1, float* grad_Wo, float* grad_bo, void* workspace, int64_t ws_bytes,
                       void* stream);

/* GCNModule backward in one launch (+ a 30 KB reduction) for the batches sgx_gcn_module_fused_fwd covers (every scene <= 32
 * peds; input 32|40, hidden 72, out 16, final 24|32): the forward is recomputed per chunk inside the kernel, nothing but x,
 * grad_out and grad_x touches HBM.  Gradient buffers are overwritten.  workspace: sgx_gcn_module_fused_bwd_ws_bytes(). */
int64_t sgx_gcn_module_fused_bwd_ws_bytes(void);
int sgx_gcn_module_fused_bwd(const float* x, const float* grad_out, const int32_t* leader, const int32_t* group_size,
                             const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                             const int32_t* chunk_scene, int64_t n_chunks, const float* W0, const float* W1,
                             const float* V0, const float* V1, const float* Wo, const float* bo, int32_t IN, int32_t HID,
                             int32_t OUT, int32_t FIN, float* grad_x, float* grad_W0, float* grad_W1, float* grad_V0,
                             float* grad_V1, float* grad_Wo, float* grad_bo, void* workspace, int64_t ws_bytes,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * GATEncoder.forward (sgan/models.py:254-294), dropout = 0:
 *   x [batch,IN]; every GAT is  n_heads x GraphAttentionLayer(F_in -> HID, concat) then
 *   out_att (n_heads*HID -> OUT), ELU, log_softmax over the OUT features (models.py:231-237).
 *   intra: Wi [n_heads,IN,HID]  ai [n_heads,2*HID]  Wio [n_heads*HID,OUT]  aio [2*OUT]
 *   inter: We [n_heads,OUT,HID] ae [n_heads,2*HID]  Weo [n_heads*HID,OUT]  aeo [2*OUT]
 *   Wo [FIN,2*OUT] bo [FIN] -> out [batch,FIN].  alpha = LeakyReLU slope.
 * The reference hard-codes IN=40, HID=72, OUT=16, FIN=24 (models.py:242-244).
 */
int64_t sgx_gat_encoder_ws_bytes(int64_t batch, int64_t n_scenes, int32_t n_heads, int32_t IN, int32_t HID,
                                 int32_t OUT, int32_t FIN);
int sgx_gat_encoder_fwd(const float* x, const int32_t* leader, const int32_t* group_size, const int32_t* ped_start,
                        const int32_t* ped_end, int64_t batch, int64_t n_scenes, const float* Wi, const float* ai,
                        const float* Wio, const float* aio, const float* We, const float* ae, const float* Weo,
                        const float* aeo, const float* Wo, const float* bo, float alpha, int32_t n_heads, int32_t IN,
                        int32_t HID, int32_t OUT, int32_t FIN, float* out, void* workspace, int64_t ws_bytes,
                        void* stream);
/* Same forward in ONE launch for batches whose scenes all have <= chunk_cap pedestrians, chunk_cap = 32 or 64
 * (chunk_scene / n_chunks from sgx_schedule_chunks with that cap; every ETH/UCY window has <= 57): no intermediate
 * leaves the SM.  n_heads = 1, dims 40/72/16/24 only. */
int sgx_gat_encoder_fused_fwd(const float* x, const int32_t* leader, const int32_t* group_size,
                              const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                              const int32_t* chunk_scene, int64_t n_chunks, int32_t chunk_cap, const float* Wi,
                              const float* ai, const float* Wio, const float* aio, const float* We, const float* ae,
                              const float* Weo, const float* aeo, const float* Wo, const float* bo, float alpha,
                              int32_t n_heads, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, float* out,
                              void* stream);
/* ... with the group structure derived in the kernel from the datasets_group labels [batch] (what sgx_group_ids would
 * compute, sgan/models.py:263-267, bit for bit); scenes <= 32 pedestrians.  tcgen05 kernel.
 * prep: NULL, or the fp16 hi/lo weight images built by sgx_gat_encoder_tc_prep for exactly these weights
 * (sgx_gat_encoder_tc_prep_bytes bytes, device memory) -- cached per weight version by the host side. */
int sgx_gat_encoder_fused_fwd_labels(const float* x, const float* labels, const int32_t* ped_start,
                                     const int32_t* ped_end, const int32_t* scene_start, const int32_t* chunk_scene,
                                     int64_t n_chunks, const float* Wi, const float* ai, const float* Wio,
                                     const float* aio, const float* We, const float* ae, const float* Weo,
                                     const float* aeo, const float* Wo, const float* bo, float alpha, int32_t n_heads,
                                     int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, const void* prep, float* out,
                                     void* stream);
int64_t sgx_gat_encoder_tc_prep_bytes(void);
int sgx_gat_encoder_tc_prep(const float* Wi, const float* ai, const float* Wio, const float* aio, const float* We,
                            const float* ae, const float* Weo, const float* aeo, const float* Wo, const float* bo,
                            int32_t n_heads, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, void* prep, void* stream);
int sgx_gat_encoder_bwd(const float* x, const float* grad_out, const int32_t* leader, const int32_t* group_size,
                        const int32_t* ped_start, const int32_t* ped_end, int64_t batch, int64_t n_scenes,
                        const float* Wi, const float* ai, const float* Wio, const float* aio, const float* We,
                        const float* ae, const float* Weo, const float* aeo, const float* Wo, const float* bo,
                        float alpha, int32_t n_heads, int32_t IN, int32_t HID, int32_t OUT, int32_t FIN,
                        float* grad_x, float* grad_Wi, float* grad_ai, float* grad_Wio, float* grad_aio,
                        float* grad_We, float* grad_ae, float* grad_Weo, float* grad_aeo, float* grad_Wo,
                        float* grad_bo, void* workspace, int64_t ws_bytes, void* stream);

/* Dense crowds (BASELINE.json configs[3]: scenes of 64-1024 pedestrians): the same forward / backward with the attention
 * layers on warp-per-row scene kernels (a CTA per 32 rows of one scene, candidates and their scores staged in shared
 * memory, the inter-level Wh tiles staged by the whole CTA) when every scene has 65 .. 2048 pedestrians; otherwise
 * identical to sgx_gat_encoder_fwd / _bwd.  scene_start int32 [n_scenes+1]; max_scene = largest scene of the batch. */
int sgx_gat_encoder_fwd_dense(const float* x, const int32_t* leader, const int32_t* group_size,
                              const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                              int64_t batch, int64_t n_scenes, int32_t max_scene, const float* Wi, const float* ai,
                              const float* Wio, const float* aio, const float* We, const float* ae, const float* Weo,
                              const float* aeo, const float* Wo, const float* bo, float alpha, int32_t n_heads,
                              int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, float* out, void* workspace,
                              int64_t ws_bytes, void* stream);
int sgx_gat_encoder_bwd_dense(const float* x, const float* grad_out, const int32_t* leader, const int32_t* group_size,
                              const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                              int64_t batch, int64_t n_scenes, int32_t max_scene, const float* Wi, const float* ai,
                              const float* Wio, const float* aio, const float* We, const float* ae, const float* Weo,
                              const float* aeo, const float* Wo, const float* bo, float alpha, int32_t n_heads,
                              int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, float* grad_x, float* grad_Wi,
                              float* grad_ai, float* grad_Wio, float* grad_aio, float* grad_We, float* grad_ae,
                              float* grad_Weo, float* grad_aeo, float* grad_Wo, float* grad_bo, void* workspace,
                              int64_t ws_bytes, void* stream);

/* GATEncoder backward in one launch (+ a 30 KB reduction) for the batches sgx_gat_encoder_fused_fwd with chunk_cap 32
 * covers (every scene <= 32 peds, n_heads 1, dims 40/72/16/24): the forward is recomputed per chunk inside the kernel,
 * nothing but x, grad_out and grad_x touches HBM.  Gradient buffers are overwritten.  workspace:
 * sgx_gat_encoder_fused_bwd_ws_bytes() (per-CTA gradient blocks). */
int64_t sgx_gat_encoder_fused_bwd_ws_bytes(void);
int sgx_gat_encoder_fused_bwd(const float* x, const float* grad_out, const int32_t* leader, const int32_t* group_size,
                              const int32_t* ped_start, const int32_t* ped_end, const int32_t* scene_start,
                              const int32_t* chunk_scene, int64_t n_chunks, const float* Wi, const float* ai,
                              const float* Wio, const float* aio, const float* We, const float* ae, const float* Weo,
                              const float* aeo, const float* Wo, const float* bo, float alpha, int32_t n_heads,
                              int32_t IN, int32_t HID, int32_t OUT, int32_t FIN, float* grad_x, float* grad_Wi,
                              float* grad_ai, float* grad_Wio, float* grad_aio, float* grad_We, float* grad_ae,
                              float* grad_Weo, float* grad_aeo, float* grad_Wo, float* grad_bo, void* workspace,
                              int64_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused per-pedestrian LSTM recurrences (inference), SURVEY.md 8f row f1.
 * Encoder.forward (sgan/models.py:62-92): obs_rel [T,batch,2] -> Linear(2,E) -> LSTM(E,H) from zero state ->
 *   h_out [batch,H].  We [E,2] be [E]  W_ih [4H,E] W_hh [4H,H] b_ih,b_hh [4H] (PyTorch gate order i,f,g,o).
 * Decoder.forward (sgan/models.py:142-178) without per-step pooling: h0 [batch,H], c0 [batch,H] or null (= 0),
 *   first input = embedding of last_pos_rel [batch,2]; per step LSTM cell -> rel = W_hp h + b_hp -> next input =
 *   embedding of rel.  pred_rel [steps,batch,2]; h_final / c_final [batch,H] optional (null to skip).
 *   steps = 1 gives the single cell update used when the caller pools between steps (models.py:162-168).
 *   nz > 0 folds add_noise (models.py:837-846, 'global' mix) in: h0 is then [batch,H-nz] and the last nz features
 *   of pedestrian p are z[ped_scene[p]] (z [S,nz], ped_scene int32 [batch]); pass z = ped_scene = null, nz = 0 otherwise.
 * Built for H in {32, 48, 64}.  With a workspace of sgx_lstm_ws_bytes() and H = 32, batch >= 8192, steps >= 2 the
 * recurrence runs on the tensor cores (tcgen05, 3-way bf16 operand splits = fp32-level accuracy); workspace may be
 * null (CUDA-core kernel).
 */
int64_t sgx_lstm_ws_bytes(void);
int sgx_lstm_encoder_fwd(const float* obs_rel, int32_t T, int64_t batch, const float* We, const float* be,
                         const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, int32_t E,
                         int32_t H, float* h_out, void* workspace, int64_t ws_bytes, int32_t ws_prepared, void* stream);
int sgx_lstm_decoder_fwd(const float* h0, const float* c0, const float* last_pos_rel, const float* z,
                         const int32_t* ped_scene, int32_t nz, int32_t steps, int64_t batch,
                         const float* We, const float* be, const float* W_ih, const float* W_hh, const float* b_ih,
                         const float* b_hh, const float* W_hp, const float* b_hp, int32_t E, int32_t H,
                         float* pred_rel, float* h_final, float* c_final, void* workspace, int64_t ws_bytes,
                         int32_t ws_prepared, void* stream);
/* ws_prepared = 1: `workspace` already holds the tensor-core weight images of exactly these weights, written by
 * sgx_lstm_prep (a workspace the caller keeps per recurrence and weight version); 0: the call builds them itself. */
int sgx_lstm_prep(const float* We, const float* be, const float* W_ih, const float* W_hh, const float* b_ih,
                  const float* b_hh, int32_t E, int32_t H, void* workspace, int64_t ws_bytes, void* stream);

/* Training path of the same recurrences (what autograd does through nn.LSTM + the step loop of sgan/models.py:157-175,
 * replaced by one forward kernel that writes a tape and one backward kernel + 2-3 reductions):
 *   tape: sgx_lstm_tape_floats(T, batch, H) floats, caller-owned, kept between forward and backward.
 *   sgx_lstm_encoder_train_fwd / sgx_lstm_decoder_train_fwd: as the inference entry points (zero initial state for the
 *     encoder; decoder without noise fold-in -- pass the concatenated [batch,H] state), plus the tape.
 *   sgx_lstm_bwd: decoder != 0: d_seq_out = d(pred_rel) [T,batch,2], d_h_last = d(h_final) [batch,H] or null;
 *       outputs d_h0 / d_c0 [batch,H] (null to skip), dW_hp_aug [2,H+1] = [dW_hp | db_hp].
 *     decoder == 0: d_h_last = d(final_h) [batch,H]; d_seq_in [T,batch,2] receives d(seq_in) (null to skip).
 *     Both: dW_hh [4H,H]; dS [4H,3] = sum_t dG_t (x_t, y_t, 1)^T, from which the caller forms (the embedding is folded
 *     into the input weights inside the kernels)  dW_ih = dS[:, :2] We^T + dS[:, 2] be^T,  dWe = W_ih^T dS[:, :2],
 *     dbe = W_ih^T dS[:, 2],  db_ih = db_hh = dS[:, 2].   workspace: sgx_lstm_bwd_ws_bytes(T, batch, H).
 */
int64_t sgx_lstm_tape_floats(int32_t T, int64_t batch, int32_t H);
int64_t sgx_lstm_bwd_ws_bytes(int32_t T, int64_t batch, int32_t H);
int sgx_lstm_encoder_train_fwd(const float* seq_in, int32_t T, int64_t batch, const float* We, const float* be,
                               const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, int32_t E,
                               int32_t H, float* h_out, float* tape, void* stream);
int sgx_lstm_decoder_train_fwd(const float* h0, const float* c0, const float* last_pos_rel, int32_t steps,
                               int64_t batch, const float* We, const float* be, const float* W_ih, const float* W_hh,
                               const float* b_ih, const float* b_hh, const float* W_hp, const float* b_hp, int32_t E,
                               int32_t H, float* pred_rel, float* h_final, float* tape, void* stream);
int sgx_lstm_bwd(int32_t decoder, const float* tape, int32_t T, int64_t batch, const float* We, const float* be,
                 const float* W_ih, const float* W_hh, const float* b_ih, const float* b_hh, const float* W_hp,
                 int32_t E, int32_t H, const float* d_seq_out, const float* d_h_last, float* d_seq_in, float* d_h0,
                 float* d_c0, float* dW_hh, float* dS, float* dW_hp_aug, void* workspace, int64_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Displacement metrics + best-of-K reduction: relative_to_abs (sgan/utils.py:83-96) +
 * displacement_error / final_displacement_error mode='raw' (sgan/losses.py:74-119) for sample k, written to column k
 * of ade / fde [batch,K]; then evaluate_helper (scripts/evaluate_model.py:58-69): out2[0] = sum over scenes of
 * min_k sum_{peds} ade[p][k], out2[1] the same for fde.  pred_rel, gt [T,batch,2]; start_pos [batch,2]; K <= 32.
 */
int sgx_displacement_errors(const float* pred_rel, const float* start_pos, const float* gt, int32_t T, int64_t batch,
                            float* ade, float* fde, int32_t K, int32_t k, void* stream);
int sgx_best_of_k(const float* ade, const float* fde, const int32_t* scene_start, int64_t n_scenes, int32_t K,
                  float* out2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Standalone dense-adjacency layers (API parity for GraphAttentionLayer.forward(h, adj),
 * sgan/models.py:198-210, and GCN.forward(A, X), models.py:573-580): the masked-softmax rows below
 * plus sgx_gemm for every product (Wh = h W, att Wh, (A H) W ...).  n x n dense `adj`.
 */
/* att[i,:] = softmax_j(adj[i,j] > 0 ? LeakyReLU(st[i,0] + st[j,1]) : -9e15)   st [n,2], adj/att [n,n] */
int sgx_dense_att_fwd(const float* st, const float* adj, int64_t n, float alpha, float* att, void* stream);
/* datt_inout: in = dL/d(att), out = dL/d(pre-activation) (0 where adj <= 0);  ds[i] = sum_j of the output row */
int sgx_dense_att_bwd(const float* st, const float* adj, const float* att, int64_t n, float alpha,
                      float* datt_inout, float* ds, void* stream);
/* C[M,N] (+)= op(A) op(B), fp32, generic strides (element strides); used by the dense layers and
 * by the parameter-gradient reductions.  accumulate != 0 adds into C. */
int sgx_gemm(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
             int64_t ldc, int64_t M, int64_t N, int64_t K, int32_t accumulate, int32_t relu, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SGX_H_ */
