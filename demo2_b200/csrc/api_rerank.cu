// extern "C" entry points for re-ranking and top-k (include/demo_b200.h).
#include "gemm_epilogues.cuh"
#include "gemm2_sm100.cuh"

#include <cstdlib>
#include "rerank.cuh"

using namespace demo;

namespace {

__global__ void keys_to_float_kernel2(const unsigned* keys, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = key_float(keys[i]);
}

struct RrWs {
  PrepView a;
  float* E;             // [N][N]
  unsigned* rowmax_key; // [N]
  float* rowmax;        // [N]
  RerankWs r;
};

// Leading dimension of the all-pairs matrix: rows start on 128-byte lines, so the store epilogue
// can use its line-coalesced float4 path and a mirrored warp store never straddles five sectors
// (N = 10 290 as the pitch: 41 M sector writes for 13 M sectors of data).
inline long long e_pitch(int N) { return round_up(static_cast<long long>(N > 0 ? N : 1), 32ll); }

size_t carve_rr(Carver& c, int N, int Q, int d, int k1, int k2, RrWs* w) {
  RrWs t;
  const size_t n = N > 0 ? N : 1;
  prep_carve(c, N, d > 0 ? d : 8, &t.a);
  t.E = c.take<float>(n * static_cast<size_t>(e_pitch(N)));
  t.rowmax_key = c.take<unsigned>(n);
  t.rowmax = c.take<float>(n);
  rerank_carve(c, N, Q, k1, k2, &t.r);
  if (w) *w = t;
  return c.off;
}

// One launch of the store GEMM over A rows x B rows of the SAME feature set (global offsets a0 /
// b0), restricted to the tiles that hold an element with global row <= global column.  The value
// of a pair is the GEMM result with the lower index on the A side (EpiStore::Params), whatever
// the tiling or sharding: the full matrix is produced from its upper triangle (half the MMA work)
// and a row shard from two launches, bit-identically.
int launch_sym_store(const PrepView& a, int a0, const PrepView& b, int b0, EpiStore::Params ep, cudaStream_t stream) {
  ep.a_norm = a.norm;
  ep.a_inv = a.inv_scale;
  ep.b_norm = b.norm;
  ep.b_inv = b.inv_scale;
  ep.M = a.rows;
  ep.mode = DIST_SQ;
  ep.a_global0 = a0;
  ep.b_global0 = b0;
  GemmOperands ops;
  if (prefer_pair_kernel(a.rows, b.rows)) {
    DEMO_TRY(make_gemm2_operands(a, b, &ops));
    Schedule s = make_dense_schedule2(a.rows, b.rows);
    s.tri = 1;
    s.tri_a0 = a0;
    s.tri_b0 = b0;
    // whole square problem whose prepared rows stay in L2 (no need for the n-grouped raster) and
    // few tiles per worker: folded enumeration of the upper triangle, every unit non-empty
    static const bool no_fold = getenv("DEMO_NO_FOLD") != nullptr;   // A/B experiments
    if (!no_fold && a0 == b0 && a.rows == b.rows && a.hi == b.hi &&
        static_cast<size_t>(a.rows) * a.pitch * 4 <= (size_t(72) << 20))
      s = make_folded_schedule2(a.rows);
    return launch_sqdist_gemm2<EpiStore>(ops, s, s.num_units, ep, stream);
  }
  DEMO_TRY(make_gemm_operands(a, b, &ops));
  Schedule s = make_dense_schedule(a.rows, b.rows);
  s.tri = 1;
  s.tri_a0 = a0;
  s.tri_b0 = b0;
  return launch_sqdist_gemm<EpiStore>(ops, s, s.num_units, ep, stream);
}

PrepView row_range(const PrepView& v, int row0, int nrows) {
  PrepView s = v;
  s.hi = v.hi + static_cast<size_t>(row0) * 2 * v.pitch;
  s.lo = s.hi + 32;
  s.norm = v.norm + row0;
  s.inv_scale = v.inv_scale + row0;
  s.rows = nrows;
  return s;
}

// Row-sharded re-ranking: workspace of one rank owning up to rows_cap rows.
struct RrShardWs {
  PrepView a;           // all N rows prepared (features are replicated)
  float* E;             // [rows_cap][N]  local rows of the all-pairs matrix
  unsigned* rowmax_key; // [rows_cap]
  float* rowmax;        // [rows_cap]
  RerankWs r;           // inverted index / scan scratch (the full-size sparse arrays are the caller's)
};

size_t carve_rr_shard(Carver& c, int N, int Q, int d, int k1, int k2, int rows_cap, RrShardWs* w) {
  RrShardWs t;
  const size_t n = N > 0 ? N : 1, rc = rows_cap > 0 ? rows_cap : 1;
  prep_carve(c, N, d > 0 ? d : 8, &t.a);
  t.E = c.take<float>(rc * static_cast<size_t>(e_pitch(N)));
  t.rowmax_key = c.take<unsigned>(rc);
  t.rowmax = c.take<float>(rc);
  rerank_carve(c, N, Q, k1, k2, &t.r);
  if (w) *w = t;
  return c.off;
}

int get_rr_shard(void* ws, size_t ws_bytes, int N, int Q, int d, int k1, int k2, int rows_cap, RrShardWs* w) {
  Carver c(ws, ws_bytes);
  carve_rr_shard(c, N, Q, d, k1, k2, rows_cap, w);
  if (!ws || !c.ok()) {
    set_error("rerank shard: workspace missing or too small (%zu < %zu)", ws_bytes, c.off);
    return DEMO_ERR_WORKSPACE;
  }
  return DEMO_OK;
}

}  // namespace

extern "C" {

// ---- row-sharded re-ranking (multi-GPU; the host all-gathers between the stages) ----
int demo_rerank_dims(int N, int k1, int k2, int* K, int* cap, int* capq) {
  if (K) *K = rerank_k(k1, k2);
  if (cap) *cap = rerank_cap(k1);
  if (capq) *capq = rerank_capq(N, k1, k2);
  return DEMO_OK;
}

size_t demo_rerank_shard_workspace_bytes(int N, int Q, int d, int k1, int k2, int rows_cap) {
  Carver c(nullptr, ~size_t(0));
  return round_up(carve_rr_shard(c, N, Q, d, k1, k2, rows_cap, nullptr), size_t(1024));
}

// Stage 1: rows [row0, row0+nrows) of the all-pairs squared-distance matrix (kept in the
// workspace with their row maxima) and their K = max(k1+1, k2) nearest neighbours by
// (E[i][j] / rowmax_i, j)  ->  rank_rows [nrows][K]   (utils/reranking.py:38-48).
int demo_rerank_shard_topk(const float* feat, int N, int Q, int d, int64_t ld, int flags, int k1, int k2, int row0,
                           int nrows, int rows_cap, int* rank_rows, float* feat_n_out, void* ws, size_t ws_bytes,
                           void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(feat && N > 1 && Q >= 1 && Q < N && d > 0 && ld >= d, "rerank shard: bad arguments (N=%d, Q=%d)", N, Q);
  DEMO_REQUIRE(row0 >= 0 && nrows >= 0 && nrows <= rows_cap && row0 + nrows <= N, "rerank shard: bad row range");
  RrShardWs w;
  DEMO_TRY(get_rr_shard(ws, ws_bytes, N, Q, d, k1, k2, rows_cap, &w));
  const int nm = (flags & DEMO_FLAG_L2NORM) ? PREP_NORM_F_NORMALIZE : PREP_NORM_NONE;
  DEMO_TRY(launch_prep_rows(feat, N, d, ld, nm, nullptr, w.a, feat_n_out, d, stream));
  if (nrows == 0) return DEMO_OK;
  DEMO_REQUIRE(rank_rows, "rerank shard: null output");
  const PrepView rows = row_range(w.a, row0, nrows);   // the local rows
  DEMO_CHECK_CUDA(cudaMemsetAsync(w.rowmax_key, 0, sizeof(unsigned) * nrows, stream));
  {
    // columns j >= i: local rows on the A side, stored in place
    EpiStore::Params ep;
    ep.out = w.E;
    ep.ldo = e_pitch(N);
    ep.rowmax_key = w.rowmax_key;
    ep.sym_mask = 1;
    DEMO_TRY(launch_sym_store(rows, row0, w.a, 0, ep, stream));
  }
  {
    // columns j < i: the pair's value is defined with the lower index (j) on the A side; the
    // local rows are the B operand and every result is stored at the transposed position
    EpiStore::Params ep;
    ep.out = w.E;
    ep.ldo = e_pitch(N);
    ep.rowmax_key = nullptr;
    ep.store_normal = 0;
    ep.sym_mirror = 1;
    ep.out_t = w.E;
    ep.ldo_t = e_pitch(N);
    ep.rowmax_key_t = w.rowmax_key;
    DEMO_TRY(launch_sym_store(row_range(w.a, 0, row0 + nrows), 0, rows, row0, ep, stream));
  }
  keys_to_float_kernel2<<<ceil_div(nrows, 256), 256, 0, stream>>>(w.rowmax_key, w.rowmax, nrows);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return launch_topk_rows(w.E, e_pitch(N), nrows, N, w.rowmax, rerank_k(k1, k2), rank_rows, nullptr, stream);
}

// Stage 2: V rows of the local rows from the gathered neighbour lists (:51-71); writes rows
// [row0, row0+nrows) of the full-size arrays v_idx [.][cap], v_val (fp16), v_cnt.
int demo_rerank_shard_krecip(int N, int Q, int d, int k1, int k2, int row0, int nrows, int rows_cap,
                             const int* rank_all, int* v_idx, void* v_val, int* v_cnt, void* ws, size_t ws_bytes,
                             void* stream_) {
  RrShardWs w;
  DEMO_TRY(get_rr_shard(ws, ws_bytes, N, Q, d, k1, k2, rows_cap, &w));
  DEMO_REQUIRE(rank_all && v_idx && v_val && v_cnt, "rerank shard: null pointer");
  return launch_krecip_rows(w.E, e_pitch(N), w.rowmax, rank_all, N, k1, k2, row0, nrows, v_idx, static_cast<__half*>(v_val),
                            v_cnt, w.r.rh_idx, w.r.rh_cnt, static_cast<cudaStream_t>(stream_));
}

// Stage 3: local query expansion of the local rows from the gathered V rows (:73-78).
int demo_rerank_shard_expand(int N, int k1, int k2, int row0, int nrows, const int* rank_all, const int* v_idx,
                             const void* v_val, const int* v_cnt, int* q_idx, void* q_val, int* q_cnt,
                             void* stream_) {
  DEMO_REQUIRE(rank_all && v_idx && v_val && v_cnt && q_idx && q_val && q_cnt, "rerank shard: null pointer");
  return launch_expand_rows(rank_all, N, k1, k2, row0, nrows, v_idx, static_cast<const __half*>(v_val), v_cnt, q_idx,
                            static_cast<__half*>(q_val), q_cnt, static_cast<cudaStream_t>(stream_));
}

// Stage 4: inverted index of the gathered final V (f_* = q_* arrays, or v_* when k2 == 1), then
// Jaccard + blend for the local query rows -> out_rows [nq_local][N-Q], nq_local =
// max(0, min(row0+nrows, Q) - row0)   (:80-99).
int demo_rerank_shard_jaccard(int N, int Q, int d, int k1, int k2, double lambda_value, int row0, int nrows,
                              int rows_cap, const int* f_idx, const void* f_val, const int* f_cnt, float* out_rows,
                              int64_t ldo, void* ws, size_t ws_bytes, void* stream_) {
  RrShardWs w;
  DEMO_TRY(get_rr_shard(ws, ws_bytes, N, Q, d, k1, k2, rows_cap, &w));
  DEMO_REQUIRE(f_idx && f_val && f_cnt, "rerank shard: null pointer");
  int nq_local = (row0 + nrows < Q ? row0 + nrows : Q) - row0;
  if (nq_local < 0) nq_local = 0;
  DEMO_REQUIRE(nq_local == 0 || (out_rows && ldo >= N - Q), "rerank shard: bad output");
  const int f_cap = k2 != 1 ? rerank_capq(N, k1, k2) : rerank_cap(k1);
  return launch_jaccard_rows(w.E, e_pitch(N), w.rowmax, N, Q, k1, k2, lambda_value, row0, nq_local, f_idx,
                             static_cast<const __half*>(f_val), f_cnt, f_cap, w.r, out_rows, ldo,
                             static_cast<cudaStream_t>(stream_));
}

size_t demo_rerank_workspace_bytes(int N, int Q, int d, int k1, int k2) {
  Carver c(nullptr, ~size_t(0));
  return round_up(carve_rr(c, N, Q, d, k1, k2, nullptr), size_t(1024));
}

// re_ranking(probFea, galFea, k1, k2, lambda_value, local_distmat=None, only_local=False)
// (utils/reranking.py:29-100).  feat = cat(probFea, galFea) [N][d]; out [Q][N-Q].
int demo_rerank(const float* feat, int N, int Q, int d, int64_t ld, int flags, int k1, int k2, double lambda_value,
                const float* local_distmat, int64_t ld_local, int only_local, float* out, int64_t ldo,
                float* feat_n_out, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(out && ws && N > 1 && Q >= 1 && Q < N, "rerank: bad arguments (N=%d, Q=%d)", N, Q);
  DEMO_REQUIRE(!only_local || local_distmat, "rerank: only_local needs local_distmat");
  Carver c(ws, ws_bytes);
  RrWs w;
  carve_rr(c, N, Q, d, k1, k2, &w);
  if (!c.ok()) {
    set_error("rerank: workspace too small (%zu < %zu)", ws_bytes, c.off);
    return DEMO_ERR_WORKSPACE;
  }
  if (!only_local) {
    DEMO_REQUIRE(feat && d > 0 && ld >= d, "rerank: bad features");
    const int nm = (flags & DEMO_FLAG_L2NORM) ? PREP_NORM_F_NORMALIZE : PREP_NORM_NONE;
    DEMO_TRY(launch_prep_rows(feat, N, d, ld, nm, nullptr, w.a, feat_n_out, d, stream));
    const bool fused_max = local_distmat == nullptr;
    if (fused_max) DEMO_CHECK_CUDA(cudaMemsetAsync(w.rowmax_key, 0, sizeof(unsigned) * N, stream));
    // upper triangle only; every result is also stored at its transposed position
    EpiStore::Params ep;
    ep.out = w.E;
    ep.ldo = e_pitch(N);
    ep.rowmax_key = fused_max ? w.rowmax_key : nullptr;
    ep.sym_mask = 1;
    ep.sym_mirror = 1;
    ep.out_t = w.E;
    ep.ldo_t = e_pitch(N);
    ep.rowmax_key_t = fused_max ? w.rowmax_key : nullptr;
    DEMO_TRY(launch_sym_store(w.a, 0, w.a, 0, ep, stream));
    if (fused_max) {
      keys_to_float_kernel2<<<ceil_div(N, 256), 256, 0, stream>>>(w.rowmax_key, w.rowmax, N);
    } else {
      // X = D + local; E = X^T (see rerank.cu header).  Our GEMM output plays the role of D^T.
      DEMO_TRY(launch_transpose_add(local_distmat, ld_local, w.E, e_pitch(N), N, true, stream));
      DEMO_TRY(launch_rowmax(w.E, e_pitch(N), N, w.rowmax, stream));
    }
  } else {
    DEMO_TRY(launch_transpose_add(local_distmat, ld_local, w.E, e_pitch(N), N, false, stream));
    DEMO_TRY(launch_rowmax(w.E, e_pitch(N), N, w.rowmax, stream));
  }
  DEMO_CHECK_CUDA(cudaGetLastError());
  return run_rerank_stages(w.E, e_pitch(N), w.rowmax, N, Q, k1, k2, lambda_value, w.r, out, ldo, stream);
}

// Same, starting from the reference's all-pairs matrix X = `original_dist` (before :46), e.g.
// [[q_q, q_g], [q_g^T, g_g]] for the distance-matrix form re_ranking(q_g, q_q, g_g, ...).
int demo_rerank_matrix(const float* X, int64_t ldx, int N, int Q, int k1, int k2, double lambda_value, float* out,
                       int64_t ldo, void* ws, size_t ws_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(X && out && ws && N > 1 && Q >= 1 && Q < N && ldx >= N, "rerank_matrix: bad arguments");
  Carver c(ws, ws_bytes);
  RrWs w;
  carve_rr(c, N, Q, 8, k1, k2, &w);
  if (!c.ok()) {
    set_error("rerank_matrix: workspace too small (%zu < %zu)", ws_bytes, c.off);
    return DEMO_ERR_WORKSPACE;
  }
  DEMO_TRY(launch_transpose_add(X, ldx, w.E, e_pitch(N), N, false, stream));
  DEMO_TRY(launch_rowmax(w.E, e_pitch(N), N, w.rowmax, stream));
  return run_rerank_stages(w.E, e_pitch(N), w.rowmax, N, Q, k1, k2, lambda_value, w.r, out, ldo, stream);
}

// k smallest entries of every row, ascending by (value, column index); k <= 256.
// Replaces the np.argsort(...)[:, :k] uses of the path (utils/reranking.py:48, metrics.py:279).
int demo_topk_rows(const float* mat, int rows, int cols, int64_t ld, int k, int* idx_out, float* val_out,
                   void* stream_) {
  DEMO_REQUIRE(mat && idx_out && ld >= cols, "topk_rows: bad arguments");
  return launch_topk_rows(mat, ld, rows, cols, nullptr, k, idx_out, val_out, static_cast<cudaStream_t>(stream_));
}

}  // extern "C"
