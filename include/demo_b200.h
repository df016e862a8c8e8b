/*
 * libdemo_b200 -- C ABI of the B200-native DeMo retrieval hot path.
 *
 * The reference (maxingan2412/DeMo2) has no FFI: its boundary for this path is the Python
 * function surface imported by engine/processor.py:7 and layers/make_loss.py:9.  The Python
 * shim `demo2_b200` re-exports that surface and calls the entry points below through ctypes;
 * every entry point cites the reference interface it replaces.  See INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (demo_last_error() has the message);
 *   - all pointers are DEVICE pointers unless the name ends in _host; matrices are row-major
 *     fp32 with an explicit leading dimension in elements; labels are int32;
 *   - the caller owns every buffer including the workspace (query the size first); the
 *     library never allocates or frees device memory and keeps no pointer after returning;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous
 *     unless documented otherwise;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef DEMO_B200_H_
#define DEMO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEMO_API __attribute__((visibility("default")))

enum {
  DEMO_OK = 0,
  DEMO_ERR_INVALID = -1,
  DEMO_ERR_CUDA = -2,
  DEMO_ERR_WORKSPACE = -3,
  DEMO_ERR_UNSUPPORTED = -4,
  DEMO_ERR_CAPACITY = -5
};

/* distance flavour (low 2 bits of `flags`) */
enum {
  DEMO_DIST_SQ = 0,       /* |q|^2+|g|^2-2qg          utils/metrics.py:395-401 euclidean_distance   */
  DEMO_DIST_SQRT = 1,     /* sqrt(clamp(.,1e-12))     layers/triplet_loss.py:16-31 euclidean_dist    */
  DEMO_DIST_COS_SIM = 2,  /* qg/(|q||g|)              cosine_similarity (north_star; absent upstream) */
  DEMO_DIST_COS_DIST = 3, /* (1-qg/(|q||g|))/2        layers/triplet_loss.py:34-48 cosine_dist       */
  DEMO_FLAG_L2NORM = 0x10,       /* F.normalize rows first   utils/metrics.py:345 */
  DEMO_FLAG_TRIPLET_NORM = 0x20, /* x/(|x|+1e-12) first      layers/triplet_loss.py:5-13 */
  DEMO_FLAG_SIMT = 0x40,         /* FFMA cross-check kernel instead of the tcgen05 GEMM */
  DEMO_FLAG_HOST_INPUT = 0x80    /* demo_eval_prepare: x is pinned host memory, read in place over PCIe */
};

DEMO_API const char* demo_last_error(void);
DEMO_API int demo_version(void);
/* 1 when a CUDA device with compute capability 10.x is current, else 0 (host-only query). */
DEMO_API int demo_device_ok(void);

/* ---- distance matrix ------------------------------------------------------------------
 * Replaces euclidean_distance(qf, gf) (utils/metrics.py:395-401), euclidean_dist / cosine_dist
 * (layers/triplet_loss.py:16-48) and the all-pairs matrix of re_ranking (utils/reranking.py:36-41).
 * out[Q][G] (ld = ldo).  rowmax (optional, [Q]) receives max_g out[q][g].
 * qn_out / gn_out (optional, [Q][d] / [G][d] contiguous) receive the normalised rows. */
DEMO_API size_t demo_sqdist_workspace_bytes(int Q, int G, int d, int flags);
DEMO_API int demo_sqdist_f32(const float* q, const float* g, int Q, int G, int d, int64_t ldq,
                             int64_t ldg, float* out, int64_t ldo, int flags, float* rowmax,
                             float* qn_out, float* gn_out, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ---- CMC / mAP by rank counts ------------------------------------------------------------
 * Replaces eval_func (utils/metrics.py:110-169) and the distance + eval part of
 * R1_mAP_eval.compute (utils/metrics.py:341-369).  For every valid positive p of query q:
 * r_p = 1 + #{valid g before p}, c_p = 1 + #{positive g before p}, "before" = lexicographic on
 * (distance, gallery index); AP = mean_p c_p/r_p, CMC[k] = [min_p r_p <= k+1].  The counts are
 * additive over gallery shards (multi-GPU: all-gather records, all-reduce counts).
 *
 * Stages (each an entry point so that a host can put collectives between them):
 *   demo_eval_plan         labels -> pid-sorted permutations, record CSR, band work list
 *   demo_eval_records      prepare operands; distances/global index/junk flag of every
 *                          same-identity (query, gallery) pair  ("records", CSR by sorted query)
 *   demo_build_thresholds  per query: valid positives sorted by (d, gidx) + #junk before each
 *   demo_eval_count        += #{local gallery items before each threshold}  (fused GEMM epilogue)
 *   demo_cmc_map_finalize  counts -> per-query AP / first rank -> cmc[max_rank], mAP, #valid
 * demo_eval_features chains them on one GPU; demo_eval_matrix does the same for a materialised
 * distance matrix (one streaming pass, 4 B per pair).                                         */
DEMO_API size_t demo_plan_bytes(int Q, int G);
/* The gallery is sorted by (pid not asked for by any query, pid): the "queried" rows -- the only
 * ones records / thresholds come from -- are the first info[3] rows of the sorted order.
 * info_host (optional, host int64[4]) = {T records, max same-pid count, band units, #queried
 * gallery rows}; when given the call synchronises `stream`. */
DEMO_API int demo_eval_plan(const int* q_pid, const int* g_pid, int Q, int G, void* plan, size_t plan_bytes,
                            int64_t* info_host, void* stream);
DEMO_API int demo_plan_pointers(const void* plan, size_t plan_bytes, int Q, int G, const int** q_perm,
                                const int** g_perm, const int** rec_ofs, const int** g_lo);
/* device pointer of info[4] (same numbers as info_host) for hosts that enqueue the plan without
 * synchronising and read the numbers later, together with other data */
DEMO_API int demo_plan_info(const void* plan, size_t plan_bytes, int Q, int G, const int** info);
DEMO_API size_t demo_eval_workspace_bytes(int Q, int G, int d, int64_t T);
/* max_cnt > 63 (info[1] of the plan / the merged maximum) adds a 256-row distance slab: query
 * blocks holding a row with more than 63 thresholds are then counted from ONE stored GEMM pass
 * (4 B per pair of HBM traffic) instead of one full GEMM per 63 thresholds. */
DEMO_API size_t demo_eval_workspace_bytes_ex(int Q, int G, int d, int64_t T, int max_cnt);
DEMO_API size_t demo_eval_matrix_workspace_bytes(int Q, int G, int64_t T);
/* Staged form of demo_eval_records for streamed / pipelined evaluation
 * (R1_mAP_eval.compute with host-resident features, utils/metrics.py:341-369):
 *   demo_eval_prepare     normalise + fp16 hi/lo split of the pid-sorted rows [row0, row0+nrows) of
 *                         the queries (which = 0) or the gallery (which = 1).  x only has to be
 *                         DEVICE-ACCESSIBLE: pinned host memory is pulled over PCIe by the kernel
 *                         itself, in sorted order (queried rows first), slab by slab.
 *   demo_eval_extract     records from the prepared queries + queried gallery rows; g_index
 *                         (optional, [G]) = global gallery index per local row (tie-break key).
 *                         [q_row0, q_row0+q_nrows) (pid-sorted queries, 128-aligned start; 0, Q =
 *                         all): only the records of that query group -- its gallery rows (a prefix
 *                         of the sorted gallery) must be prepared, the later ones may still be in
 *                         flight.
 *   demo_eval_count_range counts the pid-sorted queries [q_row0, q_row0+q_nrows) (start a multiple
 *                         of 256 rows, 1024 with the slab path) against the sorted gallery rows
 *                         [g_row0, g_row0+g_nrows): the (query block x gallery block) rectangles of
 *                         successive calls must tile Q x G exactly once;
 *                         reserve_sms > 0: the persistent GEMM grid leaves that many SMs free for
 *                         the kernel that pulls in the next slab meanwhile.                      */
DEMO_API int demo_eval_prepare(const float* x, int n, int d, int64_t ld, int flags, int which, int row0, int nrows,
                               const void* plan, size_t plan_bytes, int Q, int G, int64_t T, void* ws,
                               size_t ws_bytes, float* xn_out, void* stream);
DEMO_API int demo_eval_extract(int Q, int G, int d, const int* q_cam, const int* g_cam, int g_index_base,
                               const int* g_index, const void* plan, size_t plan_bytes, int64_t T, void* ws,
                               size_t ws_bytes, float* rec_dist, int* rec_gidx, int* rec_junk, int q_row0,
                               int q_nrows, void* stream);
DEMO_API int demo_eval_count_range(int Q, int G, int d, int64_t T_local, void* ws, size_t ws_bytes,
                                   const int* thr_ofs, const int* thr_cnt, const float* thr_val,
                                   const int* thr_gidx, unsigned* counts, int max_cnt, int chunk_tiles,
                                   int g_row0, int g_nrows, int q_row0, int q_nrows, int reserve_sms,
                                   void* stream);
DEMO_API int demo_eval_records(const float* q, const float* g, int Q, int G, int d, int64_t ldq, int64_t ldg,
                               int flags, const int* q_cam, const int* g_cam, int g_index_base,
                               const void* plan, size_t plan_bytes, int64_t T, void* ws, size_t ws_bytes,
                               float* rec_dist, int* rec_gidx, int* rec_junk, float* qn_out, float* gn_out,
                               void* stream);
DEMO_API int demo_build_thresholds(const int* rec_ofs, const float* rec_dist, const int* rec_gidx,
                                   const int* rec_junk, int Q, int* thr_cnt, float* thr_val, int* thr_gidx,
                                   int* thr_junk, void* stream);
DEMO_API int demo_eval_count(int Q, int G, int d, int64_t T_local, void* ws, size_t ws_bytes, const int* thr_ofs,
                             const int* thr_cnt, const float* thr_val, const int* thr_gidx, unsigned* counts,
                             int max_cnt, int chunk_tiles, void* stream);
DEMO_API int demo_cmc_map_finalize(const int* thr_ofs, const int* thr_cnt, const int* thr_junk,
                                   const unsigned* counts, const int* q_perm, int Q, int max_rank,
                                   float* cmc_out, double* map_out, int* num_valid_out, double* ap_out,
                                   int* first_out, void* scratch, void* stream);
DEMO_API int demo_eval_features(const float* q, const float* g, int Q, int G, int d, int64_t ldq, int64_t ldg,
                                int flags, const int* q_cam, const int* g_cam, const void* plan,
                                size_t plan_bytes, int64_t T, int max_cnt, int max_rank, void* ws,
                                size_t ws_bytes, float* cmc_out, double* map_out, int* num_valid_out,
                                double* ap_out, int* first_out, float* qn_out, float* gn_out, void* stream);
DEMO_API int demo_eval_matrix(const float* distmat, int Q, int G, int64_t ld, const int* q_cam, const int* g_cam,
                              const void* plan, size_t plan_bytes, int64_t T, int max_cnt, int max_rank,
                              void* ws, size_t ws_bytes, float* cmc_out, double* map_out, int* num_valid_out,
                              double* ap_out, int* first_out, void* stream);
DEMO_API int demo_eval_ws_pointers(void* ws, size_t ws_bytes, int Q, int G, int d, int64_t T, float** cmc,
                                   double** map, int** nvalid, double** ap, int** first, unsigned** counts,
                                   int** thr_cnt, float** thr_val, int** thr_gidx, int** thr_junk,
                                   float** rec_dist, int** rec_gidx, int** rec_junk);

/* ---- k-reciprocal re-ranking and top-k ----------------------------------------------------
 * demo_rerank replaces re_ranking(probFea, galFea, k1, k2, lambda_value, local_distmat=None,
 * only_local=False) (utils/reranking.py:29-100): feat = cat(probFea, galFea) [N][d],
 * out [Q][N-Q] fp32.  float16 stores / adds follow the reference.
 * demo_rerank_matrix starts from the all-pairs matrix X (the reference's original_dist before
 * :46), which also serves the distance-matrix form re_ranking(q_g, q_q, g_g, k1, k2, lambda).
 * demo_topk_rows: k smallest per row ascending by (value, column), k <= 256
 * (np.argsort(...)[:, :k] at utils/reranking.py:48 and utils/metrics.py:279). */
DEMO_API size_t demo_rerank_workspace_bytes(int N, int Q, int d, int k1, int k2);
DEMO_API int demo_rerank(const float* feat, int N, int Q, int d, int64_t ld, int flags, int k1, int k2,
                         double lambda_value, const float* local_distmat, int64_t ld_local, int only_local,
                         float* out, int64_t ldo, float* feat_n_out, void* ws, size_t ws_bytes, void* stream);
DEMO_API int demo_rerank_matrix(const float* X, int64_t ldx, int N, int Q, int k1, int k2, double lambda_value,
                                float* out, int64_t ldo, void* ws, size_t ws_bytes, void* stream);
DEMO_API int demo_topk_rows(const float* mat, int rows, int cols, int64_t ld, int k, int* idx_out,
                            float* val_out, void* stream);

/* Row-sharded re-ranking (multi-GPU): rank r owns the contiguous rows
 * [row0, row0+nrows) of the N x N problem, features are replicated, and the HOST all-gathers the
 * neighbour lists (after _topk), the sparse V rows (after _krecip) and the expanded rows (after
 * _expand) -- full-size caller-owned arrays rank_all [N][K] int32, v_idx [N][cap] int32,
 * v_val [N][cap] fp16, v_cnt [N], q_idx / q_val [N][capq], q_cnt [N]; K / cap / capq from
 * demo_rerank_dims.  Every stage is bit-identical to the corresponding part of demo_rerank. */
DEMO_API int demo_rerank_dims(int N, int k1, int k2, int* K, int* cap, int* capq);
DEMO_API size_t demo_rerank_shard_workspace_bytes(int N, int Q, int d, int k1, int k2, int rows_cap);
DEMO_API int demo_rerank_shard_topk(const float* feat, int N, int Q, int d, int64_t ld, int flags, int k1, int k2,
                                    int row0, int nrows, int rows_cap, int* rank_rows, float* feat_n_out,
                                    void* ws, size_t ws_bytes, void* stream);
DEMO_API int demo_rerank_shard_krecip(int N, int Q, int d, int k1, int k2, int row0, int nrows, int rows_cap,
                                      const int* rank_all, int* v_idx, void* v_val, int* v_cnt, void* ws,
                                      size_t ws_bytes, void* stream);
DEMO_API int demo_rerank_shard_expand(int N, int k1, int k2, int row0, int nrows, const int* rank_all,
                                      const int* v_idx, const void* v_val, const int* v_cnt, int* q_idx,
                                      void* q_val, int* q_cnt, void* stream);
DEMO_API int demo_rerank_shard_jaccard(int N, int Q, int d, int k1, int k2, double lambda_value, int row0,
                                       int nrows, int rows_cap, const int* f_idx, const void* f_val,
                                       const int* f_cnt, float* out_rows, int64_t ldo, void* ws, size_t ws_bytes,
                                       void* stream);

/* ---- batch-hard triplet mining -------------------------------------------------------------
 * demo_triplet_hard_fwd replaces the core of TripletLoss.__call__ (layers/triplet_loss.py:124-125):
 * euclidean_dist(x, x) + hard_example_mining(dist_mat, labels, return_inds=True) fused into the
 * distance GEMM epilogue (the N x N matrix is never written).  Positives include the anchor
 * itself (as in the reference); ties go to the lowest index; npos (optional) = #same-label
 * samples per anchor so the host can reproduce the reference's equal-count requirement (:79).
 * demo_triplet_hard_bwd: d loss / d x from the upstream gradients of dist_ap / dist_an (only the
 * 2N selected pairs carry gradient).  demo_hard_example_mining: the same selection on a
 * caller-provided distance matrix (layers/triplet_loss.py:51-104). */
DEMO_API size_t demo_triplet_workspace_bytes(int N, int d);
DEMO_API int demo_triplet_hard_fwd(const float* x, int N, int d, int64_t ld, const int* labels, float* dist_ap,
                                   float* dist_an, int64_t* p_idx, int64_t* n_idx, int* npos, void* ws,
                                   size_t ws_bytes, void* stream);
DEMO_API int demo_triplet_hard_bwd(const float* x, int N, int d, int64_t ld, const int64_t* p_idx,
                                   const int64_t* n_idx, const float* dist_ap, const float* dist_an,
                                   const float* g_ap, const float* g_an, float* grad_x, int64_t ldg,
                                   void* stream);
DEMO_API int demo_hard_example_mining(const float* dist_mat, int N, int64_t ld, const int* labels,
                                      float* dist_ap, float* dist_an, int64_t* p_idx, int64_t* n_idx,
                                      int* npos, void* stream);

/* Training-size batches (N <= demo_triplet_loss_max_batch() = 256): the WHOLE TripletLoss forward
 * of layers/triplet_loss.py:121-135 -- euclidean_dist, hard_example_mining, the (1 +- hard_factor)
 * scaling, SoftMarginLoss (margin < 0 / NaN = the reference's margin=None) or MarginRankingLoss and
 * its mean -- in ONE launch for B <= 8 feature matrices sharing the labels (the per-modality calls
 * of layers/make_loss.py:47-52), and the whole backward in another.
 *   xs      HOST array of B device pointers, each [N][ld] fp32
 *   labels  device int32 (label_is_i64 = 0) or int64 (1, the reference's LongTensor)
 *   loss [B]; dist_ap / dist_an [B][N] as the reference returns them (already scaled); p_idx /
 *   n_idx [B][N] (n_idx = -1: no negative); status [B] (optional): bit 0 = anchors with different
 *   numbers of positives (the reference's view(N, -1) at :79 raises), bit 1 = anchor without negative
 *   ws      demo_triplet_loss_workspace_bytes() bytes, zero-initialised ONCE by the caller
 *   bwd     g_loss [B] (optional), g_ap_ext / g_an_ext [B][N] (optional upstream gradients of the
 *           returned distances) -> grad [B][N][ldg]                                              */
DEMO_API size_t demo_triplet_loss_workspace_bytes(void);
DEMO_API int demo_triplet_loss_max_batch(void);
DEMO_API int demo_triplet_loss_fwd(const float* const* xs, int B, int N, int d, int64_t ld, const void* labels,
                                   int label_is_i64, float margin, float hard_factor, float* loss, float* dist_ap,
                                   float* dist_an, int64_t* p_idx, int64_t* n_idx, int* status, void* ws,
                                   size_t ws_bytes, void* stream);
DEMO_API int demo_triplet_loss_bwd(const float* const* xs, int B, int N, int d, int64_t ld, float margin,
                                   float hard_factor, const float* dist_ap, const float* dist_an,
                                   const int64_t* p_idx, const int64_t* n_idx, const float* g_loss,
                                   const float* g_ap_ext, const float* g_an_ext, float* grad, int64_t ldg,
                                   void* stream);

/* ---- communicator (multi-GPU evaluation) ----------------------------------
 * The reference evaluates on rank 0 only (engine/processor.py:145-156); here the gallery is
 * sharded over one process per GPU and the two exchanges of an evaluation -- all-gather of the
 * same-identity records, all-reduce(sum) of the rank counts -- run on the compute stream through
 * an NCCL communicator owned by the library (libnccl is resolved with dlopen at init time; the
 * copy already loaded by PyTorch is preferred).  One communicator per process.
 *   demo_comm_unique_id   rank 0: 128-byte id, distributed to the other ranks by the host
 *   demo_comm_init        collective; the current CUDA device is the rank's device
 *   demo_comm_check       ncclCommGetAsyncError -> error code                                  */
DEMO_API int demo_comm_available(void);
DEMO_API int demo_comm_nccl_version(void);
DEMO_API int demo_comm_unique_id(void* id_out_host);
DEMO_API int demo_comm_init(int rank, int world, const void* unique_id_host);
DEMO_API int demo_comm_destroy(void);
DEMO_API int demo_comm_info(int* rank, int* world);
DEMO_API int demo_comm_check(void);
DEMO_API int demo_comm_all_gather(const void* send, void* recv, size_t bytes_per_rank, void* stream);
DEMO_API int demo_comm_all_reduce_sum_u32(void* buf, size_t count, void* stream);
DEMO_API int demo_comm_broadcast(void* buf, size_t bytes, int root, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEMO_B200_H_ */
