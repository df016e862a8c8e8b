// Label planning, record / threshold bookkeeping, the streaming rank-count over a materialised
// matrix, and CMC/mAP finalisation.  See rank.cuh for the pipeline.
#include "rank.cuh"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "gemm_sm100.cuh"

namespace demo {

constexpr int kBandTiles = 8;  // tiles per extract work unit

// ---------------------------------------------------------------------------------------
// plan
// ---------------------------------------------------------------------------------------
int plan_band_cap(int Q, int G) {
  const long long mb = ceil_div(Q > 0 ? Q : 1, kBM);
  const long long per = ceil_div<long long>(G > 0 ? G : 1, kBandTiles * kBN);
  long long cap = mb * per;
  // a block's band never exceeds G rows, and all bands together cover at most G + (mb-1)*maxgroup rows;
  // mb*per is the simple safe bound.
  if (cap > (1ll << 26)) cap = 1ll << 26;
  return static_cast<int>(cap);
}

static size_t cub_tmp_bytes_for(int Q, int G) {
  size_t a = 0, b = 0, c = 0;
  int* np = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, a, np, np, np, np, G > 0 ? G : 1);
  cub::DeviceRadixSort::SortPairs(nullptr, b, np, np, np, np, Q > 0 ? Q : 1);
  cub::DeviceScan::ExclusiveSum(nullptr, c, np, np, Q + 1);
  size_t m = a > b ? a : b;
  return (m > c ? m : c) + 256;
}

size_t plan_carve(Carver& c, int Q, int G, PlanView* v) {
  PlanView p;
  p.Q = Q;
  p.G = G;
  const size_t q1 = Q > 0 ? Q : 1, g1 = G > 0 ? G : 1;
  p.q_perm = c.take<int>(q1);
  p.q_pid_sorted = c.take<int>(q1);
  p.g_perm = c.take<int>(g1);
  p.g_pid_sorted = c.take<int>(g1);
  p.g_lo = c.take<int>(q1);
  p.rec_ofs = c.take<int>(q1 + 1);
  p.band_cap = plan_band_cap(Q, G);
  p.band_list = c.take<int4>(p.band_cap);
  p.band_count = c.take<int>(4);
  p.info = c.take<int>(4);
  p.iota = c.take<int>(q1 > g1 ? q1 : g1);
  p.cnt = c.take<int>(q1 + 1);
  p.cub_tmp_bytes = cub_tmp_bytes_for(Q, G);
  p.cub_tmp = c.take<char>(p.cub_tmp_bytes);
  if (v) *v = p;
  return c.off;
}

namespace {

__global__ void iota_kernel(int* a, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = i;
}

// per sorted query: [lower_bound, upper_bound) of its pid in the sorted gallery pids
__global__ void ranges_kernel(const int* __restrict__ q_pid_sorted, const int* __restrict__ g_pid_sorted,
                              int Q, int G, int* __restrict__ g_lo, int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > Q) return;
  if (i == Q) {
    cnt[Q] = 0;
    return;
  }
  const int pid = q_pid_sorted[i];
  int lo = 0, hi = G;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (g_pid_sorted[mid] < pid) lo = mid + 1; else hi = mid;
  }
  const int first = lo;
  hi = G;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (g_pid_sorted[mid] <= pid) lo = mid + 1; else hi = mid;
  }
  g_lo[i] = first;
  cnt[i] = lo - first;
}

// one block: band work list for the extract GEMM + summary numbers
__global__ void __launch_bounds__(1024)
band_list_kernel(const int* __restrict__ g_lo, const int* __restrict__ cnt, const int* __restrict__ rec_ofs,
                 int Q, int4* __restrict__ list, int cap, int* __restrict__ band_count, int* __restrict__ info) {
  __shared__ int s_scan[1024];
  __shared__ int s_max[1024];
  const int t = threadIdx.x;
  const int m_blocks = ceil_div(Q, kBM);
  const int per = ceil_div(m_blocks, 1024);
  const int b0 = t * per, b1 = min(m_blocks, b0 + per);
  constexpr int kUnitRows = kBandTiles * kBN;
  int local = 0;
  for (int b = b0; b < b1; ++b) {
    const int first = b * kBM, last = min(Q, first + kBM) - 1;
    const int rows = g_lo[last] + cnt[last] - g_lo[first];
    local += ceil_div(rows, kUnitRows);
  }
  int mx = 0;
  for (int i = t; i < Q; i += 1024) mx = max(mx, cnt[i]);
  s_scan[t] = local;
  s_max[t] = mx;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // Hillis-Steele inclusive scan + max reduce
    const int v = t >= o ? s_scan[t - o] : 0;
    const int m2 = t >= o ? s_max[t - o] : 0;
    __syncthreads();
    s_scan[t] += v;
    s_max[t] = max(s_max[t], m2);
    __syncthreads();
  }
  int pos = s_scan[t] - local;
  for (int b = b0; b < b1; ++b) {
    const int first = b * kBM, last = min(Q, first + kBM) - 1;
    const int lo = g_lo[first];
    const int rows = g_lo[last] + cnt[last] - lo;
    for (int r = 0; r < rows; r += kUnitRows) {
      if (pos < cap) list[pos] = make_int4(b, lo + r, min(kUnitRows, rows - r), 0);
      ++pos;
    }
  }
  if (t == 1023) {
    const int total = s_scan[1023];
    *band_count = min(total, cap);
    info[0] = rec_ofs[Q];
    info[1] = s_max[1023];
    info[2] = total;
    info[3] = 0;
  }
}

}  // namespace

int run_plan(const int* q_pid, const int* g_pid, const PlanView& p, cudaStream_t stream) {
  const int Q = p.Q, G = p.G;
  DEMO_REQUIRE(Q > 0 && G > 0, "plan: empty query or gallery set (Q=%d, G=%d)", Q, G);
  size_t tmp = p.cub_tmp_bytes;
  const int n = Q > G ? Q : G;
  iota_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(p.iota, n);
  DEMO_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(p.cub_tmp, tmp, g_pid, p.g_pid_sorted, p.iota, p.g_perm, G,
                                                  0, 32, stream));
  tmp = p.cub_tmp_bytes;
  DEMO_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(p.cub_tmp, tmp, q_pid, p.q_pid_sorted, p.iota, p.q_perm, Q,
                                                  0, 32, stream));
  ranges_kernel<<<ceil_div(Q + 1, 256), 256, 0, stream>>>(p.q_pid_sorted, p.g_pid_sorted, Q, G, p.g_lo, p.cnt);
  tmp = p.cub_tmp_bytes;
  DEMO_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(p.cub_tmp, tmp, p.cnt, p.rec_ofs, Q + 1, stream));
  band_list_kernel<<<1, 1024, 0, stream>>>(p.g_lo, p.cnt, p.rec_ofs, Q, p.band_list, p.band_cap, p.band_count,
                                           p.info);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

// ---------------------------------------------------------------------------------------
// records
// ---------------------------------------------------------------------------------------
namespace {

__global__ void fill_records_kernel(const int* __restrict__ rec_ofs, const int* __restrict__ g_lo,
                                    const int* __restrict__ q_perm, const int* __restrict__ g_perm,
                                    const int* __restrict__ q_cam, const int* __restrict__ g_cam, int Q,
                                    int g_index_base, int* __restrict__ rec_gidx, int* __restrict__ rec_junk) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= Q) return;
  const int s0 = rec_ofs[i], n = rec_ofs[i + 1] - s0, lo = g_lo[i];
  const int cam = q_cam[q_perm[i]];
  for (int s = lane; s < n; s += 32) {
    const int orig = g_perm[lo + s];
    rec_gidx[s0 + s] = g_index_base + orig;
    rec_junk[s0 + s] = g_cam[orig] == cam ? 1 : 0;
  }
}

__global__ void gather_records_kernel(const int* __restrict__ rec_ofs, const int* __restrict__ g_lo,
                                      const int* __restrict__ q_perm, const int* __restrict__ g_perm,
                                      const float* __restrict__ distmat, long long ld, int Q,
                                      float* __restrict__ rec_dist) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= Q) return;
  const int s0 = rec_ofs[i], n = rec_ofs[i + 1] - s0, lo = g_lo[i];
  const float* row = distmat + static_cast<long long>(q_perm[i]) * ld;
  for (int s = lane; s < n; s += 32) rec_dist[s0 + s] = row[g_perm[lo + s]];
}

__device__ __forceinline__ bool lex_before(float da, int ga, float db, int gb) {
  return da < db || (da == db && ga < gb);
}

// One warp per query: rank-by-counting of its same-pid records.
__global__ void build_thresholds_kernel(const int* __restrict__ rec_ofs, const float* __restrict__ rec_dist,
                                        const int* __restrict__ rec_gidx, const int* __restrict__ rec_junk,
                                        int Q, int* __restrict__ thr_cnt, float* __restrict__ thr_val,
                                        int* __restrict__ thr_gidx, int* __restrict__ thr_junk) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= Q) return;
  const int s0 = rec_ofs[i], n = rec_ofs[i + 1] - s0;
  int npos = 0;
  for (int a = lane; a < n; a += 32) {
    if (rec_junk[s0 + a]) continue;
    const float da = rec_dist[s0 + a];
    const int ga = rec_gidx[s0 + a];
    int pos_before = 0, junk_before = 0;
    for (int b = 0; b < n; ++b) {
      const bool before = lex_before(rec_dist[s0 + b], rec_gidx[s0 + b], da, ga);
      const int jb = rec_junk[s0 + b];
      pos_before += (before && !jb) ? 1 : 0;
      junk_before += (before && jb) ? 1 : 0;
    }
    thr_val[s0 + pos_before] = da;
    thr_gidx[s0 + pos_before] = ga;
    thr_junk[s0 + pos_before] = junk_before;
    ++npos;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) npos += __shfl_xor_sync(0xffffffffu, npos, o);
  if (lane == 0) thr_cnt[i] = npos;
}

}  // namespace

int launch_fill_records(const PlanView& p, const int* q_cam, const int* g_cam, int g_index_base,
                        int* rec_gidx, int* rec_junk, cudaStream_t stream) {
  fill_records_kernel<<<ceil_div(p.Q * 32, 256), 256, 0, stream>>>(p.rec_ofs, p.g_lo, p.q_perm, p.g_perm, q_cam,
                                                                   g_cam, p.Q, g_index_base, rec_gidx, rec_junk);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

int launch_gather_records(const PlanView& p, const float* distmat, long long ld, float* rec_dist,
                          cudaStream_t stream) {
  gather_records_kernel<<<ceil_div(p.Q * 32, 256), 256, 0, stream>>>(p.rec_ofs, p.g_lo, p.q_perm, p.g_perm,
                                                                     distmat, ld, p.Q, rec_dist);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

int launch_build_thresholds(const int* rec_ofs, const float* rec_dist, const int* rec_gidx,
                            const int* rec_junk, int Q, int* thr_cnt, float* thr_val, int* thr_gidx,
                            int* thr_junk, cudaStream_t stream) {
  if (Q <= 0) return DEMO_OK;
  build_thresholds_kernel<<<ceil_div(Q * 32, 256), 256, 0, stream>>>(rec_ofs, rec_dist, rec_gidx, rec_junk, Q,
                                                                     thr_cnt, thr_val, thr_gidx, thr_junk);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

// ---------------------------------------------------------------------------------------
// count over a materialised matrix: one block per query row, 4 B per (q, g) pair read once.
// Thread-private histogram columns in shared memory (no atomics), thresholds of the row
// (window of kCmWin) staged in shared memory and searched by bisection.
// ---------------------------------------------------------------------------------------
namespace {

constexpr int kCmThreads = 256;
constexpr int kCmWin = 2048;  // thresholds handled per pass over the row

__global__ void __launch_bounds__(kCmThreads)
count_matrix_kernel(const float* __restrict__ distmat, long long ld, int G, int g_index_base,
                    const int* __restrict__ q_perm, const int* __restrict__ thr_ofs,
                    const int* __restrict__ thr_cnt, const float* __restrict__ thr_val,
                    const int* __restrict__ thr_gidx, unsigned* __restrict__ counts, int window) {
  __shared__ float s_thr[kCmWin];
  __shared__ int s_tg[kCmWin];
  __shared__ unsigned s_hist[kCmWin + 8];
  __shared__ unsigned s_warp[kCmThreads / 32];
  const int i = blockIdx.x;
  const int t = threadIdx.x;
  const int tbase = thr_ofs[i] + window * kCmWin;
  const int nthr = max(0, min(kCmWin, thr_cnt[i] - window * kCmWin));
  if (nthr == 0) return;
  for (int k = t; k < nthr; k += kCmThreads) {
    s_thr[k] = thr_val[tbase + k];
    s_tg[k] = thr_gidx[tbase + k];
  }
  for (int k = t; k < kCmWin + 8; k += kCmThreads) s_hist[k] = 0u;
  __syncthreads();
  const float tmax = s_thr[nthr - 1];
  const float* row = distmat + static_cast<long long>(q_perm[i]) * ld;
  for (int g0 = 0; g0 < G; g0 += kCmThreads * 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int g = g0 + u * kCmThreads + t;
      v[u] = g < G ? __ldg(row + g) : INFINITY;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d = v[u];
      if (!(d <= tmax)) continue;  // after every threshold of the window (also skips padding)
      int lo = 0, hi = nthr;       // pos = #{thr <= d}
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_thr[mid] <= d) lo = mid + 1; else hi = mid;
      }
      int pos = lo;
      if (pos > 0 && s_thr[pos - 1] == d) {
        const int g = g_index_base + g0 + u * kCmThreads + t;
        while (pos > 0 && s_thr[pos - 1] == d && s_tg[pos - 1] > g) --pos;
      }
      atomicAdd(&s_hist[pos], 1u);
    }
  }
  __syncthreads();
  // counts[k] += sum_{b <= k} hist[b]: block-wide inclusive scan, 8 bins per thread
  unsigned c[8], sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    c[j] = s_hist[t * 8 + j];
    sum += c[j];
  }
  unsigned incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += x;
  }
  if ((t & 31) == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  unsigned run = incl - sum;
  for (int w = 0; w < (t >> 5); ++w) run += s_warp[w];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    run += c[j];
    const int k = t * 8 + j;
    if (k < nthr && run) atomicAdd(counts + tbase + k, run);
  }
}

}  // namespace

int launch_count_matrix(const float* distmat, long long ld, int G, int g_index_base, const int* q_perm,
                        const int* thr_ofs, const int* thr_cnt, const float* thr_val, const int* thr_gidx,
                        unsigned* counts, int Q, int max_cnt, cudaStream_t stream) {
  if (Q <= 0 || G <= 0) return DEMO_OK;
  static_assert(kCmWin == kCmThreads * 8, "scan layout");
  const int windows = ceil_div(max_cnt > 0 ? max_cnt : 1, kCmWin);
  for (int w = 0; w < windows; ++w) {
    count_matrix_kernel<<<Q, kCmThreads, 0, stream>>>(distmat, ld, G, g_index_base, q_perm, thr_ofs, thr_cnt,
                                                      thr_val, thr_gidx, counts, w);
    DEMO_CHECK_CUDA(cudaGetLastError());
  }
  return DEMO_OK;
}

// ---------------------------------------------------------------------------------------
// finalize
// ---------------------------------------------------------------------------------------
namespace {

// per query (one warp): AP (float64) and rank of the first correct match
__global__ void query_ap_kernel(const int* __restrict__ thr_ofs, const int* __restrict__ thr_cnt,
                                const int* __restrict__ thr_junk, const unsigned* __restrict__ counts,
                                const int* __restrict__ q_perm, int Q, double* __restrict__ ap_out,
                                int* __restrict__ first_out) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= Q) return;
  const int s0 = thr_ofs[i], n = thr_cnt[i];
  double acc = 0.0;
  for (int j = lane; j < n; j += 32) {
    const int r = 1 + static_cast<int>(counts[s0 + j]) - thr_junk[s0 + j];
    acc += static_cast<double>(j + 1) / static_cast<double>(r);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) {
    const int q = q_perm[i];
    ap_out[q] = n > 0 ? acc / static_cast<double>(n) : -1.0;  // -1: skipped (identity absent from the gallery)
    first_out[q] = n > 0 ? 1 + static_cast<int>(counts[s0]) - thr_junk[s0] : 0;  // 0: skipped
  }
}

// one block: deterministic reduction in original query order
__global__ void __launch_bounds__(1024)
reduce_metrics_kernel(const double* __restrict__ ap, const int* __restrict__ first, int Q, int max_rank,
                      float* __restrict__ cmc_out, double* __restrict__ map_out, int* __restrict__ nvalid_out,
                      unsigned* __restrict__ hist /* [max_rank] zeroed scratch */) {
  __shared__ double s_sum[1024];
  __shared__ int s_cnt[1024];
  const int t = threadIdx.x;
  const int per = ceil_div(Q, 1024);
  double s = 0.0;
  int c = 0;
  for (int q = t * per; q < min(Q, (t + 1) * per); ++q) {
    if (first[q] > 0) {
      s += ap[q];
      ++c;
      if (first[q] <= max_rank) atomicAdd(hist + first[q] - 1, 1u);
    }
  }
  s_sum[t] = s;
  s_cnt[t] = c;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (t < o) {
      s_sum[t] += s_sum[t + o];
      s_cnt[t] += s_cnt[t + o];
    }
    __syncthreads();
  }
  if (t == 0) {
    const int nv = s_cnt[0];
    *nvalid_out = nv;
    *map_out = nv > 0 ? s_sum[0] / static_cast<double>(nv) : 0.0;
    unsigned run = 0;
    for (int k = 0; k < max_rank; ++k) {
      run += hist[k];
      // float32 sum of 0/1 rows divided by the float count (utils/metrics.py:165-166)
      cmc_out[k] = nv > 0 ? static_cast<float>(run) / static_cast<float>(nv) : 0.f;
    }
  }
}

}  // namespace

int launch_finalize(const int* thr_ofs, const int* thr_cnt, const int* thr_junk, const unsigned* counts,
                    const int* q_perm, int Q, int max_rank, float* cmc_out, double* map_out,
                    int* num_valid_out, double* ap_out, int* first_out, double* scratch,
                    cudaStream_t stream) {
  DEMO_REQUIRE(max_rank > 0 && max_rank <= 4096, "finalize: max_rank out of range (%d)", max_rank);
  query_ap_kernel<<<ceil_div(Q * 32, 256), 256, 0, stream>>>(thr_ofs, thr_cnt, thr_junk, counts, q_perm, Q, ap_out,
                                                        first_out);
  DEMO_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned) * max_rank, stream));
  reduce_metrics_kernel<<<1, 1024, 0, stream>>>(ap_out, first_out, Q, max_rank, cmc_out, map_out, num_valid_out,
                                                reinterpret_cast<unsigned*>(scratch));
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // namespace demo
