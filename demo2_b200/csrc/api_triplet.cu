// Batch-hard triplet mining (layers/triplet_loss.py:16-135): fused distance + mining forward on
// the tcgen05 GEMM (EpiMine), standalone mining on a given matrix, and the sparse backward.
#include "gemm_epilogues.cuh"

using namespace demo;

namespace {

__global__ void init_best_kernel(unsigned long long* best_pos, unsigned long long* best_neg, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    best_pos[i] = 0ull;
    best_neg[i] = ~0ull;
  }
}

// (key << 32 | code) -> distance, index; also counts the positives (same-label columns) per row
__global__ void decode_best_kernel(const unsigned long long* __restrict__ best_pos,
                                   const unsigned long long* __restrict__ best_neg, const int* __restrict__ labels,
                                   int n, float* __restrict__ dist_ap, float* __restrict__ dist_an,
                                   long long* __restrict__ p_idx, long long* __restrict__ n_idx, int* __restrict__ npos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long bp = best_pos[i], bn = best_neg[i];
  dist_ap[i] = key_float(static_cast<unsigned>(bp >> 32));
  p_idx[i] = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned>(bp & 0xFFFFFFFFu));
  if (bn == ~0ull) {  // no negative in the batch
    dist_an[i] = INFINITY;
    n_idx[i] = -1;
  } else {
    dist_an[i] = key_float(static_cast<unsigned>(bn >> 32));
    n_idx[i] = static_cast<long long>(bn & 0xFFFFFFFFu);
  }
  if (npos) {
    int c = 0;
    const int lab = labels[i];
    for (int j = 0; j < n; ++j) c += labels[j] == lab;
    npos[i] = c;
  }
}

// hard_example_mining(dist_mat, labels) on a given N x N matrix: one warp per row
__global__ void mining_matrix_kernel(const float* __restrict__ dm, long long ld, const int* __restrict__ labels, int n,
                                     float* __restrict__ dist_ap, float* __restrict__ dist_an,
                                     long long* __restrict__ p_idx, long long* __restrict__ n_idx,
                                     int* __restrict__ npos) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int lab = labels[i];
  unsigned long long bp = 0ull, bn = ~0ull;
  int c = 0;
  for (int j = lane; j < n; j += 32) {
    const unsigned key = float_key(dm[(long long)i * ld + j] + 0.f);
    if (labels[j] == lab) {
      const unsigned long long v = (static_cast<unsigned long long>(key) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(j));
      bp = v > bp ? v : bp;
      ++c;
    } else {
      const unsigned long long v = (static_cast<unsigned long long>(key) << 32) | static_cast<unsigned>(j);
      bn = v < bn ? v : bn;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long op = __shfl_xor_sync(0xffffffffu, bp, o), on = __shfl_xor_sync(0xffffffffu, bn, o);
    bp = op > bp ? op : bp;
    bn = on < bn ? on : bn;
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if (lane == 0) {
    dist_ap[i] = key_float(static_cast<unsigned>(bp >> 32));
    p_idx[i] = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned>(bp & 0xFFFFFFFFu));
    dist_an[i] = bn == ~0ull ? INFINITY : key_float(static_cast<unsigned>(bn >> 32));
    n_idx[i] = bn == ~0ull ? -1 : static_cast<long long>(bn & 0xFFFFFFFFu);
    if (npos) npos[i] = c;
  }
}

constexpr int kTripletMaxN = 12800;   // 200 KB of per-anchor coefficients in the backward kernel

// grad_x[r] = sum over anchors a of the pair terms that touch row r (deterministic gather form):
//   pair (a, b) with upstream g and distance D contributes  +g (x_a - x_b)/D to row a, - the same to row b;
//   zero when the clamp was active (D^2 <= 1e-12).
__global__ void triplet_bwd_kernel(const float* __restrict__ x, int n, int d, long long ld,
                                   const long long* __restrict__ p_idx, const long long* __restrict__ n_idx,
                                   const float* __restrict__ dist_ap, const float* __restrict__ dist_an,
                                   const float* __restrict__ g_ap, const float* __restrict__ g_an,
                                   float* __restrict__ grad, long long ldg) {
  extern __shared__ float s_coef[];  // [2][n] pair coefficients g / D (0 when clamped), then indices
  int* s_idx = reinterpret_cast<int*>(s_coef + 2 * n);
  const int r = blockIdx.x;
  for (int a = threadIdx.x; a < n; a += blockDim.x) {
    const float dp = dist_ap[a], dn = dist_an[a];
    s_coef[a] = (dp * dp > 1e-12f) ? g_ap[a] / dp : 0.f;
    s_coef[n + a] = (n_idx[a] >= 0 && dn * dn > 1e-12f) ? g_an[a] / dn : 0.f;
    s_idx[a] = static_cast<int>(p_idx[a]);
    s_idx[n + a] = static_cast<int>(n_idx[a]);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float xr = x[(long long)r * ld + k];
    float acc = 0.f;
    for (int a = 0; a < n; ++a) {
      const int p = s_idx[a], q = s_idx[n + a];
      const float cp = s_coef[a], cn = s_coef[n + a];
      if (a == r) {
        if (cp != 0.f) acc += cp * (xr - x[(long long)p * ld + k]);
        if (cn != 0.f) acc += cn * (xr - x[(long long)q * ld + k]);
      }
      if (p == r && cp != 0.f) acc -= cp * (x[(long long)a * ld + k] - xr);
      if (q == r && cn != 0.f) acc -= cn * (x[(long long)a * ld + k] - xr);
    }
    grad[(long long)r * ldg + k] = acc;
  }
}

struct TripletWs {
  PrepView a;
  unsigned long long* best_pos;
  unsigned long long* best_neg;
};

size_t carve_triplet(Carver& c, int N, int d, TripletWs* w) {
  TripletWs t;
  prep_carve(c, N, d, &t.a);
  t.best_pos = c.take<unsigned long long>(N > 0 ? N : 1);
  t.best_neg = c.take<unsigned long long>(N > 0 ? N : 1);
  if (w) *w = t;
  return c.off;
}

}  // namespace

extern "C" {

size_t demo_triplet_workspace_bytes(int N, int d) {
  Carver c(nullptr, ~size_t(0));
  return round_up(carve_triplet(c, N, d, nullptr), size_t(1024));
}

// TripletLoss forward core (layers/triplet_loss.py:124-125): dist_mat = euclidean_dist(x, x) and
// hard_example_mining(dist_mat, labels, return_inds=True) in one fused GEMM; the N x N matrix is
// never written.  p_idx / n_idx are int64 like the reference's LongTensors.
int demo_triplet_hard_fwd(const float* x, int N, int d, int64_t ld, const int* labels, float* dist_ap,
                          float* dist_an, int64_t* p_idx, int64_t* n_idx, int* npos, void* ws, size_t ws_bytes,
                          void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(x && labels && dist_ap && dist_an && p_idx && n_idx && ws && N > 0 && d > 0 && ld >= d,
               "triplet_hard_fwd: bad arguments");
  // the backward kernel stages 16 B per anchor in shared memory: refuse here what it could not take
  DEMO_REQUIRE(N <= kTripletMaxN, "triplet_hard_fwd: batch too large (N=%d, max %d)", N, kTripletMaxN);
  Carver c(ws, ws_bytes);
  TripletWs w;
  carve_triplet(c, N, d, &w);
  if (!c.ok()) {
    set_error("triplet_hard_fwd: workspace too small (%zu < %zu)", ws_bytes, c.off);
    return DEMO_ERR_WORKSPACE;
  }
  DEMO_TRY(launch_prep_rows(x, N, d, ld, PREP_NORM_NONE, nullptr, w.a, nullptr, 0, stream));
  init_best_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(w.best_pos, w.best_neg, N);
  GemmOperands ops;
  DEMO_TRY(make_gemm_operands(w.a, w.a, &ops));
  EpiMine::Params ep;
  ep.a_norm = w.a.norm;
  ep.a_inv = w.a.inv_scale;
  ep.b_norm = w.a.norm;
  ep.b_inv = w.a.inv_scale;
  ep.a_lab = labels;
  ep.b_lab = labels;
  ep.best_pos = w.best_pos;
  ep.best_neg = w.best_neg;
  ep.M = N;
  const Schedule s = make_dense_schedule(N, N);
  DEMO_TRY(launch_sqdist_gemm<EpiMine>(ops, s, s.num_units, ep, stream));
  decode_best_kernel<<<ceil_div(N, 256), 256, 0, stream>>>(w.best_pos, w.best_neg, labels, N, dist_ap, dist_an,
                                                           reinterpret_cast<long long*>(p_idx),
                                                           reinterpret_cast<long long*>(n_idx), npos);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

int demo_triplet_hard_bwd(const float* x, int N, int d, int64_t ld, const int64_t* p_idx, const int64_t* n_idx,
                          const float* dist_ap, const float* dist_an, const float* g_ap, const float* g_an,
                          float* grad_x, int64_t ldg, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(x && p_idx && n_idx && dist_ap && dist_an && g_ap && g_an && grad_x && N > 0 && d > 0,
               "triplet_hard_bwd: bad arguments");
  const size_t smem = static_cast<size_t>(N) * 16;
  DEMO_REQUIRE(N <= kTripletMaxN, "triplet_hard_bwd: batch too large (N=%d, max %d)", N, kTripletMaxN);
  static PerDeviceInt configured;
  DEMO_CHECK_CUDA(ensure_dynamic_smem(configured, triplet_bwd_kernel, kTripletMaxN * 16));
  triplet_bwd_kernel<<<N, 256, smem, stream>>>(x, N, d, ld, reinterpret_cast<const long long*>(p_idx),
                                               reinterpret_cast<const long long*>(n_idx), dist_ap, dist_an, g_ap,
                                               g_an, grad_x, ldg);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

// hard_example_mining(dist_mat, labels, return_inds=True) on a caller-provided matrix
// (layers/triplet_loss.py:51-104).
int demo_hard_example_mining(const float* dist_mat, int N, int64_t ld, const int* labels, float* dist_ap,
                             float* dist_an, int64_t* p_idx, int64_t* n_idx, int* npos, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(dist_mat && labels && dist_ap && dist_an && p_idx && n_idx && N > 0 && ld >= N,
               "hard_example_mining: bad arguments");
  mining_matrix_kernel<<<ceil_div(N * 32, 256), 256, 0, stream>>>(dist_mat, ld, labels, N, dist_ap, dist_an,
                                                                  reinterpret_cast<long long*>(p_idx),
                                                                  reinterpret_cast<long long*>(n_idx), npos);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // extern "C"
