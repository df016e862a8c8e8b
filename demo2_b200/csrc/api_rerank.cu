// extern "C" entry points for re-ranking and top-k (include/demo_b200.h).
#include "gemm_epilogues.cuh"
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

size_t carve_rr(Carver& c, int N, int Q, int d, int k1, int k2, RrWs* w) {
  RrWs t;
  const size_t n = N > 0 ? N : 1;
  prep_carve(c, N, d > 0 ? d : 8, &t.a);
  t.E = c.take<float>(n * n);
  t.rowmax_key = c.take<unsigned>(n);
  t.rowmax = c.take<float>(n);
  rerank_carve(c, N, Q, k1, k2, &t.r);
  if (w) *w = t;
  return c.off;
}

}  // namespace

extern "C" {

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
    GemmOperands ops;
    DEMO_TRY(make_gemm_operands(w.a, w.a, &ops));
    const bool fused_max = local_distmat == nullptr;
    if (fused_max) DEMO_CHECK_CUDA(cudaMemsetAsync(w.rowmax_key, 0, sizeof(unsigned) * N, stream));
    EpiStore::Params ep;
    ep.a_norm = w.a.norm;
    ep.a_inv = w.a.inv_scale;
    ep.b_norm = w.a.norm;
    ep.b_inv = w.a.inv_scale;
    ep.out = w.E;
    ep.ldo = N;
    ep.M = N;
    ep.mode = DIST_SQ;
    ep.rowmax_key = fused_max ? w.rowmax_key : nullptr;
    const Schedule s = make_dense_schedule(N, N);
    DEMO_TRY(launch_sqdist_gemm<EpiStore>(ops, s, s.num_units, ep, stream));
    if (fused_max) {
      keys_to_float_kernel2<<<ceil_div(N, 256), 256, 0, stream>>>(w.rowmax_key, w.rowmax, N);
    } else {
      // X = D + local; E = X^T (see rerank.cu header).  Our GEMM output plays the role of D^T.
      DEMO_TRY(launch_transpose_add(local_distmat, ld_local, w.E, N, N, true, stream));
      DEMO_TRY(launch_rowmax(w.E, N, N, w.rowmax, stream));
    }
  } else {
    DEMO_TRY(launch_transpose_add(local_distmat, ld_local, w.E, N, N, false, stream));
    DEMO_TRY(launch_rowmax(w.E, N, N, w.rowmax, stream));
  }
  DEMO_CHECK_CUDA(cudaGetLastError());
  return run_rerank_stages(w.E, N, w.rowmax, N, Q, k1, k2, lambda_value, w.r, out, ldo, stream);
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
  DEMO_TRY(launch_transpose_add(X, ldx, w.E, N, N, false, stream));
  DEMO_TRY(launch_rowmax(w.E, N, N, w.rowmax, stream));
  return run_rerank_stages(w.E, N, w.rowmax, N, Q, k1, k2, lambda_value, w.r, out, ldo, stream);
}

// k smallest entries of every row, ascending by (value, column index); k <= 256.
// Replaces the np.argsort(...)[:, :k] uses of the path (utils/reranking.py:48, metrics.py:279).
int demo_topk_rows(const float* mat, int rows, int cols, int64_t ld, int k, int* idx_out, float* val_out,
                   void* stream_) {
  DEMO_REQUIRE(mat && idx_out && ld >= cols, "topk_rows: bad arguments");
  return launch_topk_rows(mat, ld, rows, cols, nullptr, k, idx_out, val_out, static_cast<cudaStream_t>(stream_));
}

}  // extern "C"
