// Label planning, record / threshold bookkeeping, the streaming rank-count over a materialised
// matrix, and CMC/mAP finalisation.  See rank.cuh for the pipeline.
#include "rank.cuh"

#include <cstdlib>

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
  unsigned long long* kp = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, a, kp, kp, np, np, G > 0 ? G : 1);
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
  p.gkey = c.take<unsigned long long>(g1);
  p.gkey_sorted = c.take<unsigned long long>(g1);
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

// Gallery sort key: rows whose pid is asked for by some query come first ("queried" rows: every
// record / threshold lives there, so a streamed gallery can deliver them first and the count
// GEMM can start while the rest is still on its way), then by pid (signed order).
constexpr unsigned long long kNotQueried = 1ull << 32;
__global__ void gallery_key_kernel(const int* __restrict__ g_pid, int G, const int* __restrict__ q_pid_sorted, int Q,
                                   unsigned long long* __restrict__ key) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G) return;
  const int pid = g_pid[i];
  int lo = 0, hi = Q;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (q_pid_sorted[mid] < pid) lo = mid + 1; else hi = mid;
  }
  const bool queried = lo < Q && q_pid_sorted[lo] == pid;
  key[i] = (queried ? 0ull : kNotQueried) | static_cast<unsigned long long>(static_cast<unsigned>(pid) ^ 0x80000000u);
}

// sorted keys -> sorted pids; the one thread that sees the class boundary writes P = #queried rows
__global__ void gallery_unkey_kernel(const unsigned long long* __restrict__ key_sorted, int G,
                                     int* __restrict__ g_pid_sorted, int* __restrict__ n_queried) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G) return;
  const unsigned long long k = key_sorted[i];
  g_pid_sorted[i] = static_cast<int>(static_cast<unsigned>(k) ^ 0x80000000u);
  const bool mine = k < kNotQueried;
  if (i == 0 && !mine) *n_queried = 0;
  if (mine && (i + 1 == G || key_sorted[i + 1] >= kNotQueried)) *n_queried = i + 1;
}

// per sorted query: [lower_bound, upper_bound) of its pid in the queried part of the sorted gallery
__global__ void ranges_kernel(const int* __restrict__ q_pid_sorted, const int* __restrict__ g_pid_sorted,
                              int Q, const int* __restrict__ n_queried, int* __restrict__ g_lo,
                              int* __restrict__ cnt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > Q) return;
  if (i == Q) {
    cnt[Q] = 0;
    return;
  }
  const int G = *n_queried;
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
                 int Q, int4* __restrict__ list, int cap, int* __restrict__ band_count, int* __restrict__ info,
                 const int* __restrict__ n_queried) {
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
    info[3] = *n_queried;
  }
}

}  // namespace

int run_plan(const int* q_pid, const int* g_pid, const PlanView& p, cudaStream_t stream) {
  const int Q = p.Q, G = p.G;
  DEMO_REQUIRE(Q > 0 && G > 0, "plan: empty query or gallery set (Q=%d, G=%d)", Q, G);
  size_t tmp = p.cub_tmp_bytes;
  const int n = Q > G ? Q : G;
  iota_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(p.iota, n);
  DEMO_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(p.cub_tmp, tmp, q_pid, p.q_pid_sorted, p.iota, p.q_perm, Q,
                                                  0, 32, stream));
  // gallery: stable sort by (not queried, pid); p.band_count[1] holds P = #queried rows
  int* n_queried = p.band_count + 1;
  gallery_key_kernel<<<ceil_div(G, 256), 256, 0, stream>>>(g_pid, G, p.q_pid_sorted, Q, p.gkey);
  tmp = p.cub_tmp_bytes;
  DEMO_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(p.cub_tmp, tmp, p.gkey, p.gkey_sorted, p.iota, p.g_perm, G,
                                                  0, 33, stream));
  gallery_unkey_kernel<<<ceil_div(G, 256), 256, 0, stream>>>(p.gkey_sorted, G, p.g_pid_sorted, n_queried);
  ranges_kernel<<<ceil_div(Q + 1, 256), 256, 0, stream>>>(p.q_pid_sorted, p.g_pid_sorted, Q, n_queried, p.g_lo,
                                                          p.cnt);
  tmp = p.cub_tmp_bytes;
  DEMO_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(p.cub_tmp, tmp, p.cnt, p.rec_ofs, Q + 1, stream));
  band_list_kernel<<<1, 1024, 0, stream>>>(p.g_lo, p.cnt, p.rec_ofs, Q, p.band_list, p.band_cap, p.band_count,
                                           p.info, n_queried);
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
                                    int g_index_base, const int* __restrict__ g_index,
                                    int* __restrict__ rec_gidx, int* __restrict__ rec_junk) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= Q) return;
  const int s0 = rec_ofs[i], n = rec_ofs[i + 1] - s0, lo = g_lo[i];
  const int cam = q_cam[q_perm[i]];
  for (int s = lane; s < n; s += 32) {
    const int orig = g_perm[lo + s];
    rec_gidx[s0 + s] = g_index ? g_index[orig] : g_index_base + orig;
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

// One warp per query: rank-by-counting of its same-pid records.  Up to kThrStage records are staged
// in shared memory first (the O(n^2) comparisons then read warp-wide broadcasts instead of
// global memory); longer lists fall back to global reads.
constexpr int kThrWarps = 4;
constexpr int kThrStage = 256;
constexpr int kThrLong = 64;          // lists longer than this: one block per query
constexpr int kThrLongThreads = 128;
constexpr int kThrLongStage = 1024;

__global__ void __launch_bounds__(kThrWarps * 32)
build_thresholds_kernel(const int* __restrict__ rec_ofs, const float* __restrict__ rec_dist,
                        const int* __restrict__ rec_gidx, const int* __restrict__ rec_junk,
                        int Q, int* __restrict__ thr_cnt, float* __restrict__ thr_val,
                        int* __restrict__ thr_gidx, int* __restrict__ thr_junk) {
  __shared__ float s_d[kThrWarps][kThrStage];
  __shared__ int s_g[kThrWarps][kThrStage];
  __shared__ int s_j[kThrWarps][kThrStage];
  const int w = threadIdx.x >> 5;
  const int i = blockIdx.x * kThrWarps + w;
  const int lane = threadIdx.x & 31;
  if (i >= Q) return;
  const int s0 = rec_ofs[i], n = rec_ofs[i + 1] - s0;
  if (n > kThrLong) return;                    // long lists: build_thresholds_long_kernel
  const bool staged = n <= kThrStage;
  if (staged) {
    for (int a = lane; a < n; a += 32) {
      s_d[w][a] = rec_dist[s0 + a];
      s_g[w][a] = rec_gidx[s0 + a];
      s_j[w][a] = rec_junk[s0 + a];
    }
    __syncwarp();
  }
  const float* d = staged ? s_d[w] : rec_dist + s0;
  const int* g = staged ? s_g[w] : rec_gidx + s0;
  const int* jk = staged ? s_j[w] : rec_junk + s0;
  int npos = 0;
  for (int a = lane; a < n; a += 32) {
    if (jk[a]) continue;
    const float da = d[a];
    const int ga = g[a];
    int pos_before = 0, junk_before = 0;
    for (int b = 0; b < n; ++b) {
      const bool before = lex_before(d[b], g[b], da, ga);
      const int jb = jk[b];
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

// Long lists (identities with hundreds of gallery images): the same rank-by-counting with one
// BLOCK per query -- the O(n^2) comparisons of a 171-record list kept a single warp busy for
// 1 000 dependent iterations per lane, and with one warp per query only ~12 warps per SM were in
// flight (68 us for RGBNT100's 1 715 queries; 4 warps per query: latency hidden).
__global__ void __launch_bounds__(kThrLongThreads)
build_thresholds_long_kernel(const int* __restrict__ rec_ofs, const float* __restrict__ rec_dist,
                             const int* __restrict__ rec_gidx, const int* __restrict__ rec_junk,
                             int Q, int* __restrict__ thr_cnt, float* __restrict__ thr_val,
                             int* __restrict__ thr_gidx, int* __restrict__ thr_junk) {
  __shared__ float s_d[kThrLongStage];
  __shared__ int s_g[kThrLongStage];
  __shared__ int s_j[kThrLongStage];
  __shared__ int s_npos[kThrLongThreads / 32];
  const int i = blockIdx.x, t = threadIdx.x;
  const int s0 = rec_ofs[i], n = rec_ofs[i + 1] - s0;
  if (n <= kThrLong) return;                   // build_thresholds_kernel
  const bool staged = n <= kThrLongStage;
  if (staged) {
    for (int a = t; a < n; a += kThrLongThreads) {
      s_d[a] = rec_dist[s0 + a];
      s_g[a] = rec_gidx[s0 + a];
      s_j[a] = rec_junk[s0 + a];
    }
  }
  __syncthreads();
  const float* d = staged ? s_d : rec_dist + s0;
  const int* g = staged ? s_g : rec_gidx + s0;
  const int* jk = staged ? s_j : rec_junk + s0;
  int npos = 0;
  for (int a = t; a < n; a += kThrLongThreads) {
    if (jk[a]) continue;
    const float da = d[a];
    const int ga = g[a];
    int pos_before = 0, junk_before = 0;
    for (int b = 0; b < n; ++b) {
      const bool before = lex_before(d[b], g[b], da, ga);
      const int jb = jk[b];
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
  if ((t & 31) == 0) s_npos[t >> 5] = npos;
  __syncthreads();
  if (t == 0) {
    int tot = 0;
    for (int w = 0; w < kThrLongThreads / 32; ++w) tot += s_npos[w];
    thr_cnt[i] = tot;
  }
}

}  // namespace

int launch_fill_records(const PlanView& p, const int* q_cam, const int* g_cam, int g_index_base,
                        const int* g_index, int* rec_gidx, int* rec_junk, cudaStream_t stream) {
  fill_records_kernel<<<ceil_div(p.Q * 32, 256), 256, 0, stream>>>(p.rec_ofs, p.g_lo, p.q_perm, p.g_perm, q_cam,
                                                                   g_cam, p.Q, g_index_base, g_index, rec_gidx,
                                                                   rec_junk);
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
  build_thresholds_kernel<<<ceil_div(Q, kThrWarps), kThrWarps * 32, 0, stream>>>(rec_ofs, rec_dist, rec_gidx, rec_junk, Q,
                                                                     thr_cnt, thr_val, thr_gidx, thr_junk);
  // every query belongs to exactly one of the two kernels (the blocks of the other return at once)
  build_thresholds_long_kernel<<<Q, kThrLongThreads, 0, stream>>>(rec_ofs, rec_dist, rec_gidx, rec_junk, Q, thr_cnt,
                                                                  thr_val, thr_gidx, thr_junk);
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
constexpr int kCmSmallG = 12288;   // rows of up to this many columns: bisection kernel for every row (measured: profiles/count_small_rows_r2b.txt)

__global__ void __launch_bounds__(kCmThreads)
count_matrix_kernel(const float* __restrict__ distmat, long long ld, int G, int g_index_base,
                    const int* __restrict__ q_perm, const int* __restrict__ thr_ofs,
                    const int* __restrict__ thr_cnt, const float* __restrict__ thr_val,
                    const int* __restrict__ thr_gidx, unsigned* __restrict__ counts, int window, int take_all,
                    CountRows rows) {
  __shared__ float s_thr[kCmWin];
  __shared__ int s_tg[kCmWin];
  __shared__ unsigned s_hist[kCmWin + 8];
  __shared__ unsigned s_warp[kCmThreads / 32];
  const int i = rows.row0 + blockIdx.x;   // sorted query
  const int t = threadIdx.x;
  if (rows.blk_flag && rows.blk_flag[i >> 8] == 0) return;   // block handled by the fused GEMM epilogue
  const int tbase = thr_ofs[i] + window * kCmWin;
  const int nthr = max(0, min(kCmWin, thr_cnt[i] - window * kCmWin));
  if (nthr == 0) return;
  // rows with <= 255 finite thresholds belong to count_matrix255_kernel
  if (!take_all && thr_cnt[i] <= 255 && isfinite(thr_val[thr_ofs[i] + thr_cnt[i] - 1])) return;
  for (int k = t; k < nthr; k += kCmThreads) {
    s_thr[k] = thr_val[tbase + k];
    s_tg[k] = thr_gidx[tbase + k];
  }
  for (int k = t; k < kCmWin + 8; k += kCmThreads) s_hist[k] = 0u;
  __syncthreads();
  const float tmax = s_thr[nthr - 1];
  const float* row = distmat + static_cast<long long>(q_perm ? q_perm[i] : static_cast<int>(blockIdx.x)) * ld;
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
        const int gc = g0 + u * kCmThreads + t;
        const int g = rows.col_gidx ? __ldg(rows.col_gidx + gc) : g_index_base + gc;
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

// Rows with up to 255 finite thresholds (everything from the 17 positives per query of the
// benchmark to RGBNT100's ~171 images per id).  A bisection over the thresholds costs one dependent
// shared-memory probe per level at random addresses (the 63-threshold kernel this replaces:
// 3 register + 3 table levels, 41 warp instructions per 32 elements, 1.39 ms for 4096 x 262 144;
// 255 thresholds would need five table levels).  Instead the value range [t_0, t_last] of the row's
// thresholds is cut into kC8Fine equal bins by ONE monotone arithmetic map
// f(x) = min(F + 1, u32(fma(x, s, o))): thresholds and elements go through the same expression,
// so f(thr) < f(x) implies thr < x and f(thr) > f(x) implies thr > x -- only the thresholds that
// share the element's bin are compared explicitly, and an element can only tie with a threshold
// of its own bin.  Common case, branch-free: one u16 table entry (first threshold of the bin | how
// many), one probe of that threshold, a private u8 histogram column (no atomics; flushed into u32
// totals every 240 elements per thread, before a counter can wrap).  Bins with two or more
// thresholds and bit-ties take a slow path (bisection inside the bin + the (distance, index)
// rule).  4 B per pair read once; 26-29 warp instructions per 32 elements: 1.23 ms (3.5 TB/s) at
// ~20 and 1.67 ms (2.6 TB/s) at ~170 thresholds per row for 4096 x 262 144.
constexpr int kC8Threads = 256;
constexpr int kC8Fine = 4096;
constexpr int kC8Bins = 256;
#ifndef DEMO_C8_BATCH8
#define DEMO_C8_BATCH8 0
#endif
constexpr bool kC8Batch8 = DEMO_C8_BATCH8 != 0;   // 8 look-ups in flight before the (ordered) histogram updates
constexpr int kC8FlushIters = 30;   // 8 elements per thread and iteration: 240 (+ head / tail / remainder <= 10) <= 255

__global__ void __launch_bounds__(kC8Threads, 4)
count_matrix255_kernel(const float* __restrict__ distmat, long long ld, int G, int g_index_base,
                       const int* __restrict__ q_perm, const int* __restrict__ thr_ofs,
                       const int* __restrict__ thr_cnt, const float* __restrict__ thr_val,
                       const int* __restrict__ thr_gidx, unsigned* __restrict__ counts, int hist_rows,
                       CountRows rows) {
  extern __shared__ __align__(16) unsigned char s_hist8[];   // [hist_rows][256] u8, column = thread (>= 4 rows: scratch)
  __shared__ float s_thr[kC8Bins];                           // padded with NaN (never <= or == an element, +inf included)
  __shared__ int s_tg[kC8Bins];
  __shared__ unsigned short s_fine[kC8Fine + 2];             // first threshold of the bin | count << 8
  __shared__ unsigned s_tot[kC8Bins];
  __shared__ unsigned s_warp[kC8Threads / 32];
  const int i = rows.row0 + blockIdx.x, t = threadIdx.x;     // i: sorted query
  if (rows.blk_flag && rows.blk_flag[i >> 8] == 0) return;   // block handled by the fused GEMM epilogue
  const int nthr = thr_cnt[i];
  if (nthr <= 0 || nthr > kC8Bins - 1 || nthr >= hist_rows) return;   // longer rows: count_matrix_kernel
  const int tbase = thr_ofs[i];
  if (!isfinite(__ldg(thr_val + tbase + nthr - 1))) return;
  s_thr[t] = t < nthr ? __ldg(thr_val + tbase + t) : __int_as_float(0x7fc00000);
  s_tg[t] = t < nthr ? __ldg(thr_gidx + tbase + t) : -1;
  s_tot[t] = 0u;
  __syncthreads();
  const float tmin = s_thr[0], tmax = s_thr[nthr - 1];
  // bin of x: min(F + 1, u32(fma(x, scale, off))) -- ONE rounding, monotone in x; the unsigned
  // conversion saturates everything below t_0 to bin 0, bin F + 1 lies beyond every threshold
  // (t_last lands in bin F - 1 or F).  Equal thresholds (or a single one): a tiny positive range,
  // so that the elements above it still leave the thresholds' bin.
  const float range = fmaxf(tmax - tmin, fmaxf(fabsf(tmin) * 9.5e-7f, 1e-30f));
  const float scale = static_cast<float>(kC8Fine) / range, off = -tmin * scale;
  auto fine_of = [&](float x) -> unsigned {
    return min(__float2uint_rz(fmaf(x, scale, off)), static_cast<unsigned>(kC8Fine + 1));
  };
  // s_fine[j] = (index of the first threshold whose bin is >= j) | (thresholds in bin j) << 8.
  // The thresholds' own bins (scratch: the head of the histogram area) are non-decreasing, so
  // every bin finds its first threshold by bisection -- 8 bins per thread, no serial fill of the
  // gaps between distant thresholds.
  int* tbin = reinterpret_cast<int*>(s_hist8);
  if (t < nthr) tbin[t] = static_cast<int>(fine_of(s_thr[t]));
  __syncthreads();
  {
    // thread t owns the consecutive bins [9t, 9t + 9) of the F + 2: one bisection, then a walk
    constexpr int kPer = (kC8Fine + 2 + kC8Threads - 1) / kC8Threads;
    const int j0 = t * kPer, j1 = min(j0 + kPer, kC8Fine + 2);
    int lo = 0, hi = nthr;                     // #{k : tbin[k] < j0}
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (tbin[mid] < j0) lo = mid + 1; else hi = mid;
    }
    int f = lo;
#pragma unroll 1
    for (int j = j0; j < j1; ++j) {
      int f1 = f;
      while (f1 < nthr && tbin[f1] <= j) ++f1;
      s_fine[j] = static_cast<unsigned short>(f | ((f1 - f) << 8));
      f = f1;
    }
  }
  __syncthreads();
  {
    uint4* h = reinterpret_cast<uint4*>(s_hist8);
    for (int k = t; k < hist_rows * kC8Threads / 16; k += kC8Threads) h[k] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();

  // private column: lanes of a warp on distinct banks, the four warps of a group share words
  const uint32_t col = 4u * (t & 31) + ((t >> 5) & 3) + 128u * (t >> 7);
  const uint32_t hist0 = smem_u32(s_hist8) + col;
  const float* row = distmat + static_cast<long long>(q_perm ? q_perm[i] : static_cast<int>(blockIdx.x)) * ld;

  // branch-free part: slot of the element, or "needs the slow path" (bit 31)
  auto fast_pos = [&](float d) -> unsigned {
    const unsigned e = s_fine[fine_of(d)];
    const unsigned base = e & 0xffu;
    // first threshold of the bin; when the bin is empty: the first one of a LATER bin, which is > d,
    // or the NaN padding (s_thr has 256 slots, nthr <= 255; an element may be +inf) -- so no test of
    // the count is needed and pos never exceeds nthr
    const float tv = s_thr[base];
    unsigned pos = base + (tv <= d ? 1u : 0u);
    pos |= (e > 0x1ffu || tv == d) ? 0x80000000u : 0u;     // several thresholds in the bin, or a bit-tie
    return pos;
  };
  // bins with several thresholds, bit-ties: bisection inside the bin + the (distance, index) rule
  auto slow_pos = [&](float d, int g) -> unsigned {
    const unsigned e = s_fine[fine_of(d)];
    int lo = static_cast<int>(e & 0xffu), hi = lo + static_cast<int>(e >> 8);
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (s_thr[mid] <= d) lo = mid + 1; else hi = mid;
    }
    int pos = lo;
    if (pos > 0 && s_thr[pos - 1] == d) {
      const int gi = rows.col_gidx ? __ldg(rows.col_gidx + g) : g_index_base + g;
      while (pos > 0 && s_thr[pos - 1] == d && s_tg[pos - 1] > gi) --pos;
    }
    return static_cast<unsigned>(pos);
  };
  auto bump = [&](unsigned pos) {
    const uint32_t h = hist0 + pos * kC8Threads;
    unsigned c;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(c) : "r"(h));
    c += 1u;
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(h), "r"(c));
  };
  auto place = [&](float d, int g) {
    unsigned pos = fast_pos(d);
    if (pos & 0x80000000u) pos = slow_pos(d, g);
    bump(pos);
  };
  // Elements beyond the row's last threshold only ever reach bin nthr, which nothing reads (the
  // final scan adds bins <= k for k < nthr): a vector of four such elements is dropped after three
  // fminf and one compare.  With a working retrieval model the positives -- the thresholds -- sit at
  // the head of the ranking and almost every vector goes this way; on adversarial data (thresholds
  // spread over the whole value range) the test costs ~4 % more instructions.
  auto beyond = [&](const float4& x) -> bool {
    return !(fminf(fminf(x.x, x.y), fminf(x.z, x.w)) <= tmax);
  };
  auto place8 = [&](const float4& x, const float4& y, int g0, int g1) {
    const float d[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    unsigned pos[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) pos[k] = fast_pos(d[k]);
    if ((pos[0] | pos[1] | pos[2] | pos[3] | pos[4] | pos[5] | pos[6] | pos[7]) & 0x80000000u) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (pos[k] & 0x80000000u) pos[k] = slow_pos(d[k], (k < 4 ? g0 : g1) + (k & 3));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) bump(pos[k]);
  };
  auto place4 = [&](const float4& x, int g0) {
    if (beyond(x)) return;
    const float d[4] = {x.x, x.y, x.z, x.w};
    unsigned pos[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) pos[k] = fast_pos(d[k]);
    if ((pos[0] | pos[1] | pos[2] | pos[3]) & 0x80000000u) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (pos[k] & 0x80000000u) pos[k] = slow_pos(d[k], g0 + k);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) bump(pos[k]);
  };
  auto flush = [&]() {                         // private columns -> u32 totals, columns back to zero
    __syncthreads();
    if (t <= nthr) {                           // one bin (a 256-byte row) per thread, 16-byte chunks rotated by
      uint4* rowp = reinterpret_cast<uint4*>(s_hist8 + t * kC8Threads);   // the thread index: conflict-free
      unsigned sum = 0;
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int k = (c + t) & 15;
        const uint4 x = rowp[k];
        sum += __dp4a(x.x, 0x01010101u, 0u) + __dp4a(x.y, 0x01010101u, 0u) + __dp4a(x.z, 0x01010101u, 0u) +
               __dp4a(x.w, 0x01010101u, 0u);
        rowp[k] = make_uint4(0u, 0u, 0u, 0u);
      }
      s_tot[t] += sum;
    }
    __syncthreads();
  };

  const int mis = static_cast<int>((reinterpret_cast<uintptr_t>(row) >> 2) & 3u);
  const int head = min(G, (4 - mis) & 3);
  if (t < head) place(__ldg(row + t), t);
  const int nvec = (G - head) >> 2;
  const float4* row4 = reinterpret_cast<const float4*>(row + head);
  // full iterations (every thread has both vectors): the flush cadence is uniform over the block
  const int n_full = nvec / (2 * kC8Threads);
  // the vectors of iteration k + 1 are requested before those of iteration k are placed: with one
  // pair of 16-byte loads per thread in flight, 3-4 resident blocks kept ~30 KB per SM on the way
  // -- about what ~1.5 us of loaded DRAM latency allows at the 3.1-3.5 TB/s this kernel measured
  int v = t, iters = 0;
  float4 nx = make_float4(0.f, 0.f, 0.f, 0.f), ny = nx;
  if (n_full > 0) {
    nx = __ldcs(row4 + v);
    ny = __ldcs(row4 + v + kC8Threads);
  }
  for (int k = 0; k < n_full; ++k, v += 2 * kC8Threads) {
    const float4 x = nx, y = ny;
    if (k + 1 < n_full) {
      nx = __ldcs(row4 + v + 2 * kC8Threads);
      ny = __ldcs(row4 + v + 3 * kC8Threads);
    }
    const int g0 = head + 4 * v;
    if (kC8Batch8) {
      place8(x, y, g0, g0 + 4 * kC8Threads);
    } else {
      place4(x, g0);
      place4(y, g0 + 4 * kC8Threads);
    }
    if (++iters == kC8FlushIters) {
      flush();
      iters = 0;
    }
  }
  for (; v < nvec; v += kC8Threads) {          // remainder: at most two vectors per thread
    const float4 x = __ldcs(row4 + v);
    place4(x, head + 4 * v);
  }
  const int tail0 = head + 4 * nvec;
  if (tail0 + t < G) place(__ldg(row + tail0 + t), tail0 + t);
  flush();
  // counts[k] += sum_{b <= k} tot[b]: one bin per thread, block-wide inclusive scan
  const unsigned mine = t <= nthr ? s_tot[t] : 0u;
  unsigned incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += x;
  }
  if ((t & 31) == 31) s_warp[t >> 5] = incl;
  __syncthreads();
  for (int w = 0; w < (t >> 5); ++w) incl += s_warp[w];
  if (t < nthr && incl) atomicAdd(counts + tbase + t, incl);
}

}  // namespace

int launch_count_matrix(const float* distmat, long long ld, int G, int g_index_base, const int* q_perm,
                        const int* thr_ofs, const int* thr_cnt, const float* thr_val, const int* thr_gidx,
                        unsigned* counts, int Q, int max_cnt, cudaStream_t stream, const CountRows* rows_) {
  if (Q <= 0 || G <= 0) return DEMO_OK;
  static_assert(kCmWin == kCmThreads * 8, "scan layout");
  CountRows rows;
  if (rows_) rows = *rows_;
  // Short rows: the arithmetic-bin kernel's per-row set-up (4096-bin table, 16-48 KB of private
  // histogram columns) costs more than it saves; below kSmallG columns every row goes to the
  // bisection kernel (per-element early-out beyond the last threshold, shared-memory atomics).
  static const int small_g = getenv("DEMO_CM_SMALL_G") ? atoi(getenv("DEMO_CM_SMALL_G")) : kCmSmallG;
  const bool all_generic = G <= small_g;
  // rows with up to 255 finite thresholds
  if (!all_generic) {
    // histogram rows: one per slot of the longest row this kernel takes (more resident blocks for short lists)
    const int top = max_cnt < kC8Bins - 1 ? (max_cnt > 0 ? max_cnt : 1) : kC8Bins - 1;
    const int hist_rows = (top + 1 + 31) / 32 * 32 < 64 ? 64 : (top + 1 + 31) / 32 * 32;   // >= 4 rows: setup scratch
    const int smem8 = hist_rows * kC8Threads;
    static PerDeviceInt configured8;
    DEMO_CHECK_CUDA(ensure_dynamic_smem(configured8, count_matrix255_kernel, kC8Bins * kC8Threads));
    count_matrix255_kernel<<<Q, kC8Threads, smem8, stream>>>(distmat, ld, G, g_index_base, q_perm, thr_ofs, thr_cnt,
                                                             thr_val, thr_gidx, counts, hist_rows, rows);
    DEMO_CHECK_CUDA(cudaGetLastError());
  }
  // everything else (longer rows, non-finite thresholds): generic windows
  const int windows = ceil_div(max_cnt > 0 ? max_cnt : 1, kCmWin);
  for (int w = 0; w < windows; ++w) {
    count_matrix_kernel<<<Q, kCmThreads, 0, stream>>>(distmat, ld, G, g_index_base, q_perm, thr_ofs, thr_cnt,
                                                      thr_val, thr_gidx, counts, w, all_generic ? 1 : 0, rows);
    DEMO_CHECK_CUDA(cudaGetLastError());
  }
  return DEMO_OK;
}

namespace {
// one CTA per slab of `bps` 256-row query blocks
__global__ void block_flags_kernel(const int* __restrict__ thr_cnt, int Q, int win, int bps,
                                   unsigned char* __restrict__ flag, unsigned char* __restrict__ unflag,
                                   unsigned char* __restrict__ slab_any) {
  const int nb = ceil_div(Q, 256);
  int slab = 0;
  for (int b = blockIdx.x * bps; b < min(nb, (blockIdx.x + 1) * bps); ++b) {
    int any = 0;
    for (int i = b * 256 + threadIdx.x; i < min(Q, (b + 1) * 256); i += blockDim.x) any |= thr_cnt[i] > win ? 1 : 0;
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) {
      flag[b] = any ? 1 : 0;
      unflag[b] = any ? 0 : 1;
    }
    slab |= any;
  }
  if (threadIdx.x == 0) slab_any[blockIdx.x] = slab ? 1 : 0;
}
}  // namespace

// flag[b] = 1 (unflag[b] = 0) when a row of the 256-row query block b has more than `win`
// thresholds; slab_any[s] = 1 when one of the `bps` blocks of slab s is flagged
int launch_block_flags(const int* thr_cnt, int Q, int win, int bps, unsigned char* flag, unsigned char* unflag,
                       unsigned char* slab_any, cudaStream_t stream) {
  if (Q <= 0) return DEMO_OK;
  block_flags_kernel<<<ceil_div(ceil_div(Q, 256), bps), 256, 0, stream>>>(thr_cnt, Q, win, bps, flag, unflag, slab_any);
  DEMO_CHECK_CUDA(cudaGetLastError());
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
