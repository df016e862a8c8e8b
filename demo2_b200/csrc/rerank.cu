// k-reciprocal re-ranking (utils/reranking.py:29-100) as sparse, row-local CUDA kernels.
//
// Notation.  E is the N x N matrix the ranking works on, N = Q + G.  With E[i][j] = X[j][i] where
// X is the reference's `original_dist` before normalisation (:41-45), the reference's
// `np.transpose(X / np.max(X, axis=0))` (:46) is  od[i][j] = E[i][j] / rowmax_i(E), so every
// stage below is row-local in E.  (For the feature path E is simply our all-pairs GEMM output,
// whose row maxima come fused from the GEMM epilogue.)
//
//   topk        rank[i][0..K)     K = max(k1+1, k2) smallest od[i][.] by (value, index)    (:48)
//   krecip      V row i (sorted unique indices, fp16 softmax weights)                      (:51-71)
//   expand      V_qe[i] = fp16(mean of V[rank[i][0..k2)]) -- fp32 sequential sum, /k2      (:73-78)
//   invert      column lists of V_qe                                                        (:80-82)
//   jaccard     per query: fp16 sequential accumulation over ascending columns, blend      (:84-99)
//
// Rounding contract: float32 exp, numpy pairwise float32 sum,
// fp16 stores, numpy's "compute in float32, round once" half arithmetic.
#include "rerank.cuh"

#include <cmath>

#include <cub/device/device_scan.cuh>

namespace demo {

namespace {

// numpy half arithmetic: convert to float32, operate, round to half once
__device__ __forceinline__ __half np_hadd(__half a, __half b) { return __float2half_rn(__half2float(a) + __half2float(b)); }
__device__ __forceinline__ __half np_hsub(__half a, __half b) { return __float2half_rn(__half2float(a) - __half2float(b)); }
__device__ __forceinline__ __half np_hmul(__half a, __half b) { return __float2half_rn(__half2float(a) * __half2float(b)); }
__device__ __forceinline__ __half np_hdiv(__half a, __half b) { return __float2half_rn(__half2float(a) / __half2float(b)); }

// numpy's pairwise float32 summation (numpy/core/src/umath/loops_utils.h.src pairwise_sum)
__device__ float np_pairwise_sum(const float* a, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  if (n <= 128) {
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
}

// ---------------------------------------------------------------------------------------
// row max / transpose helpers (local_distmat paths)
// ---------------------------------------------------------------------------------------
__global__ void rowmax_kernel(const float* __restrict__ E, long long ld, int N, float* __restrict__ rowmax) {
  const int i = blockIdx.x;
  float m = -INFINITY;
  for (int j = threadIdx.x; j < N; j += blockDim.x) m = fmaxf(m, E[(long long)i * ld + j]);
  __shared__ float s[256];
  s[threadIdx.x] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) s[threadIdx.x] = fmaxf(s[threadIdx.x], s[threadIdx.x + o]);
    __syncthreads();
  }
  if (threadIdx.x == 0) rowmax[i] = s[0];
}

// E[i][j] (+)= X[j][i]
__global__ void transpose_add_kernel(const float* __restrict__ X, long long ldx, float* __restrict__ E, long long lde,
                                     int N, int accumulate) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int j = by + r, i = bx + threadIdx.x;  // read X[j][i]
    tile[r][threadIdx.x] = (j < N && i < N) ? X[(long long)j * ldx + i] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = bx + r, j = by + threadIdx.x;
    if (i < N && j < N) {
      float* dst = E + (long long)i * lde + j;
      const float v = tile[threadIdx.x][r];
      *dst = accumulate ? *dst + v : v;
    }
  }
}

// ---------------------------------------------------------------------------------------
// top-k smallest per row by (value, index): radix select on ordered keys + bitonic sort of
// the k winners.  One block per row; the row is read from HBM once and kept in shared memory.
// ---------------------------------------------------------------------------------------
constexpr int kTopkThreads = 256;

constexpr int kTopkCap = 1024;  // candidate buffer of the fast path

// block-wide bitonic sort of s[0..n2) ascending (n2 a power of two <= 1024), 256 threads
__device__ void bitonic_sort_u64(unsigned long long* s, int n2) {
  for (int size = 2; size <= n2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n2; i += kTopkThreads) {
        const int partner = i ^ stride;
        if (partner > i) {
          const bool up = (i & size) == 0;
          const unsigned long long a = s[i], b = s[partner];
          if ((a > b) == up) {
            s[i] = b;
            s[partner] = a;
          }
        }
      }
      __syncthreads();
    }
  }
}

// Streams nvec vectors of a row through f(vector, vector index): kBatch independent loads per thread
// are issued before the first one is consumed.  (Round 2: `#pragma unroll 8` over float2 left a
// remainder of single dependent loads -- a 10 290-column row cost ~12 DRAM latencies per block,
// ncu: 19 of 30 stall cycles per issue on the long scoreboard at 95 % occupancy.)
template <typename Vec, int kBatch, typename F>
__device__ __forceinline__ void stream_row(const float* __restrict__ src, int nvec, int t, F&& f) {
  const Vec* s = reinterpret_cast<const Vec*>(src);
  for (int base = 0; base < nvec; base += kTopkThreads * kBatch) {
    Vec v[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int j = base + u * kTopkThreads + t;
      if (j < nvec) v[u] = __ldg(s + j);
    }
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int j = base + u * kTopkThreads + t;
      if (j < nvec) f(v[u], j);
    }
  }
}

// Fast path: the k-th smallest of the per-thread (or per-4-thread) minima is an upper bound T
// of the k-th smallest entry of the row; the few entries <= T are collected and ranked by
// (key, index).  Selection runs on the raw entries; the division by row_div (monotone) is
// applied to the candidates only, after widening T over the entries whose quotient ties with
// T's.  Fallback (more than kTopkCap candidates, i.e. massive ties): 4 x 8-bit radix select.
// Measured and dropped (round 2): the whole row in registers (one batch of up to 12 independent
// 16-byte loads per thread, bound from the whole row, filter from the registers, no second pass):
// 0.163 instead of 0.107 ms for 10 290 x 10 290 at k = 21 -- 4 resident blocks per SM instead of 6
// hide the selection / ranking phases worse than the saved L2 pass returns.  Same for a bound from
// a 4096-entry prefix on these rows (0.119 ms: more candidates to rank).
template <bool kCached>
__global__ void __launch_bounds__(kTopkThreads, kCached ? 8 : 6)
topk_rows_kernel(const float* __restrict__ mat, long long ld, int cols, const float* __restrict__ row_div, int k,
                 int* __restrict__ idx_out, float* __restrict__ val_out, int prefix) {
  extern __shared__ unsigned s_dyn[];
  __shared__ unsigned s_hist[256];
  __shared__ unsigned s_prefix, s_need, s_nless, s_neq, s_ncand, s_bound;
  __shared__ unsigned long long s_cand[kTopkCap];
  unsigned* s_raw = s_dyn;  // [cols] raw ordered keys when cached
  const int row = blockIdx.x;
  const int t = threadIdx.x;
  const float* src = mat + (long long)row * ld;
  const float div = row_div ? row_div[row] : 1.f;
  const bool has_div = row_div != nullptr;
  auto raw_at = [&](int j) -> unsigned { return kCached ? s_raw[j] : float_key(__ldg(src + j) + 0.f); };
  auto final_key = [&](unsigned raw) -> unsigned {  // ordering key: entry / row_div (canonical zero)
    return has_div ? float_key(key_float(raw) / div + 0.f) : raw;
  };
  auto key_at = [&](int j) -> unsigned { return final_key(raw_at(j)); };
  unsigned tmin = 0xFFFFFFFFu;
  // 16- or 8-byte vector loads when the row start allows it
  const int vec = kCached ? 1 : ((reinterpret_cast<uintptr_t>(src) & 15u) == 0 ? 4 : (reinterpret_cast<uintptr_t>(src) & 7u) == 0 ? 2 : 1);
  const int cols2 = vec == 4 ? (cols >> 2) * 2 : vec == 2 ? (cols >> 1) : 0;   // float2 units covered by vector loads
  // The bound only needs a subset of the row (the k-th smallest of ANY >= k entries bounds the
  // k-th smallest of all): un-cached rows derive it from a prefix, so that most of the row is
  // read from HBM once (the filter pass) instead of twice.
  const int pre = kCached ? cols : min(cols, prefix);
  const int pre2 = vec == 4 ? (pre >> 2) * 2 : vec == 2 ? (pre >> 1) : 0;
  if (vec > 1) {
    // minimum in the float domain (fminf skips NaN; + 0.f makes -0 the canonical +0), ONE key at the end
    float fmin_ = INFINITY;
    bool any = false;
    if (vec == 4)
      stream_row<float4, 4>(src, pre2 >> 1, t, [&](const float4& v, int) {
        fmin_ = fminf(fmin_, fminf(fminf(v.x + 0.f, v.y + 0.f), fminf(v.z + 0.f, v.w + 0.f)));
        any = true;
      });
    else
      stream_row<float2, 8>(src, pre2, t, [&](const float2& v, int) {
        fmin_ = fminf(fmin_, fminf(v.x + 0.f, v.y + 0.f));
        any = true;
      });
    if (any) tmin = min(tmin, float_key(fmin_));
  }
  for (int j = 2 * pre2 + t; j < pre; j += kTopkThreads) {
    const unsigned raw = float_key(__ldg(src + j) + 0.f);
    if (kCached) s_raw[j] = raw;
    tmin = min(tmin, raw);
  }
  unsigned* s_min = s_hist;  // reuse: [256]
  s_min[t] = tmin;
  if (t == 0) s_ncand = 0;
  __syncthreads();
  // bound = k-th smallest of M group minima (M = 64 groups of 4 threads when k <= 64, else 256)
  const int M = k <= 64 ? 64 : 256;
  unsigned mine = 0xFFFFFFFFu;
  if (t < M) mine = M == 64 ? min(min(s_min[4 * t], s_min[4 * t + 1]), min(s_min[4 * t + 2], s_min[4 * t + 3])) : s_min[t];
  __syncthreads();
  if (t < M) s_min[t] = mine;
  __syncthreads();
  if (t < M) {
    int rank = 0;
    for (int u = 0; u < M; ++u) {
      const unsigned o = s_min[u];
      rank += (o < mine || (o == mine && u < t)) ? 1 : 0;
    }
    if (rank == min(k, M) - 1) {
      unsigned bound = mine;
      if (has_div && bound < 0xFF000000u) {
        // widen over raw values whose quotient equals the bound's quotient
        const unsigned qb = final_key(bound);
        for (int it = 0; it < 8 && final_key(bound + 1u) == qb; ++it) ++bound;
      }
      s_bound = bound;
    }
  }
  __syncthreads();
  const unsigned bound = s_bound;
  auto offer = [&](unsigned raw, int j) {
    if (raw <= bound) {
      const unsigned p = atomicAdd(&s_ncand, 1u);
      if (p < kTopkCap) s_cand[p] = (static_cast<unsigned long long>(final_key(raw)) << 32) | static_cast<unsigned>(j);
    }
  };
  if (vec > 1) {
    // filter in the float domain: key(x) <= key(b) <=> x <= b for canonical non-NaN values (a NaN
    // bound only arises when fewer than k entries are not NaN: then every comparison must pass)
    const float bound_f = key_float(bound);
    const bool all = bound_f != bound_f;
    if (vec == 4)
      stream_row<float4, 4>(src, cols2 >> 1, t, [&](const float4& v, int j) {
        const float a = v.x + 0.f, b = v.y + 0.f, c = v.z + 0.f, e = v.w + 0.f;
        if (all || fminf(fminf(a, b), fminf(c, e)) <= bound_f) {
          offer(float_key(a), 4 * j);
          offer(float_key(b), 4 * j + 1);
          offer(float_key(c), 4 * j + 2);
          offer(float_key(e), 4 * j + 3);
        }
      });
    else
      stream_row<float2, 8>(src, cols2, t, [&](const float2& v, int j) {
        const float a = v.x + 0.f, b = v.y + 0.f;
        if (all || a <= bound_f || b <= bound_f) {
          offer(float_key(a), 2 * j);
          offer(float_key(b), 2 * j + 1);
        }
      });
  }
  for (int j = 2 * cols2 + t; j < cols; j += kTopkThreads) offer(raw_at(j), j);
  __syncthreads();
  const unsigned ncand = s_ncand;
  if (ncand <= 256) {
    // rank by counting: candidate i goes to position #{candidates smaller}
    unsigned long long me = ~0ull;
    int pos = 0;
    if (t < static_cast<int>(ncand)) {
      me = s_cand[t];
      for (unsigned u = 0; u < ncand; ++u) pos += s_cand[u] < me ? 1 : 0;
    }
    __syncthreads();
    if (t < static_cast<int>(ncand)) s_cand[pos] = me;
    __syncthreads();
  } else
  if (ncand <= kTopkCap) {
    int n2 = 32;
    while (n2 < static_cast<int>(ncand)) n2 <<= 1;
    for (int i = ncand + t; i < n2; i += kTopkThreads) s_cand[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(s_cand, n2);
  } else {
    // ---- fallback: radix select ----
    if (t == 0) {
      s_prefix = 0;
      s_need = k;
      s_nless = 0;
      s_neq = 0;
    }
    __syncthreads();
    for (int pass = 3; pass >= 0; --pass) {
      s_hist[t] = 0;
      __syncthreads();
      const unsigned prefix = s_prefix;
      const unsigned hi_mask = pass == 3 ? 0u : (0xFFFFFFFFu << ((pass + 1) * 8));
      for (int j = t; j < cols; j += kTopkThreads) {
        const unsigned key = key_at(j);
        if ((key & hi_mask) == (prefix & hi_mask)) atomicAdd(&s_hist[(key >> (pass * 8)) & 255u], 1u);
      }
      __syncthreads();
      if (t == 0) {
        unsigned need = s_need, cum = 0;
        int bin = 0;
        for (; bin < 256; ++bin) {
          if (cum + s_hist[bin] >= need) break;
          cum += s_hist[bin];
        }
        s_need = need - cum;
        s_prefix = prefix | (static_cast<unsigned>(bin) << (pass * 8));
      }
      __syncthreads();
    }
    const unsigned kth = s_prefix;
    const unsigned take_eq = s_need;  // how many keys == kth belong to the top-k (lowest columns)
    for (int j = t; j < cols; j += kTopkThreads) {
      const unsigned key = key_at(j);
      if (key < kth) {
        const unsigned p = atomicAdd(&s_nless, 1u);
        s_cand[p] = (static_cast<unsigned long long>(key) << 32) | static_cast<unsigned>(j);
      }
    }
    __syncthreads();
    if (t == 0) {
      unsigned got = 0;
      for (int j = 0; j < cols && got < take_eq; ++j)
        if (key_at(j) == kth) s_cand[(k - take_eq) + got++] = (static_cast<unsigned long long>(kth) << 32) | static_cast<unsigned>(j);
    }
    __syncthreads();
    int n2 = 32;
    while (n2 < k) n2 <<= 1;
    for (int i = k + t; i < n2; i += kTopkThreads) s_cand[i] = ~0ull;
    __syncthreads();
    bitonic_sort_u64(s_cand, n2);
  }
  if (t < k) {
    const unsigned long long c = s_cand[t];
    idx_out[(long long)row * k + t] = static_cast<int>(c & 0xFFFFFFFFu);
    if (val_out) val_out[(long long)row * k + t] = key_float(static_cast<unsigned>(c >> 32));
  }
}

// ---------------------------------------------------------------------------------------
// bitmap helpers (block-wide)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void bm_set(unsigned* bm, int x) { atomicOr(&bm[x >> 5], 1u << (x & 31)); }
__device__ __forceinline__ bool bm_get(const unsigned* bm, int x) { return (bm[x >> 5] >> (x & 31)) & 1u; }

// Enumerate the set bits of bm[0..words) in ascending order into out[] (at most cap), returns
// the total count.  All threads of the block must call (blockDim.x a multiple of 32).  s_scan: >= 32 unsigned.
__device__ int bm_enumerate(const unsigned* bm, int words, int* out, int cap, unsigned* s_scan,
                            unsigned* word_prefix = nullptr) {
  const int t = threadIdx.x, nt = blockDim.x;
  const int per = ceil_div(words, nt);
  const int w0 = min(words, t * per), w1 = min(words, w0 + per);
  unsigned c = 0;
  for (int w = w0; w < w1; ++w) c += __popc(bm[w]);
  // exclusive prefix over the threads: shuffle scan per warp + the totals of the warps before
  // (a serial scan by thread 0 -- 128 dependent shared-memory updates -- was 12 % of the expansion
  // kernel's instructions and ~2.7 us of every block's life)
  const int lane = t & 31, wid = t >> 5, nw = (nt + 31) >> 5;
  unsigned incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned x = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += x;
  }
  if (lane == 31) s_scan[wid] = incl;
  __syncthreads();
  unsigned pos = incl - c, total_u = 0;
  for (int w2 = 0; w2 < nw; ++w2) {
    const unsigned x = s_scan[w2];
    pos += w2 < wid ? x : 0u;
    total_u += x;
  }
  for (int w = w0; w < w1; ++w) {
    unsigned bits = bm[w];
    if (word_prefix) word_prefix[w] = pos;  // number of set bits before word w (rank queries)
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      if (pos < static_cast<unsigned>(cap)) out[pos] = w * 32 + b;
      ++pos;
    }
  }
  __syncthreads();   // s_scan may be reused by the caller
  return static_cast<int>(total_u);
}

// ---------------------------------------------------------------------------------------
// k-reciprocal neighbours + expansion + softmax weights (one block per row)
// ---------------------------------------------------------------------------------------
constexpr int kKrThreads = 128;
constexpr int kMaxK = 128;  // k1 + 1 <= kMaxK

// R(c, kh) for EVERY row c, once: the expansion step of every row i asks for the half-size
// reciprocal sets of its ~k1 candidates, and a row is a candidate of ~k1 other rows -- computing
// them per (i, c) repeated each reciprocity test ~k1 times.  One warp per row; the list keeps the
// forward-neighbour order (cf[np.nonzero(cb == cand)], :63-65).
__global__ void __launch_bounds__(128)
recip_half_kernel(const int* __restrict__ rank, int K, int N, int kh, int* __restrict__ rh_idx,
                  int* __restrict__ rh_cnt) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (c >= N) return;
  const int* rc = rank + (long long)c * K;
  int n = 0;
  for (int base = 0; base < kh; base += 32) {
    const int m = base + lane;
    bool rec = false;
    int x = -1;
    if (m < kh) {
      x = rc[m];
      const int* rx = rank + (long long)x * K;
      for (int mm = 0; mm < kh; ++mm) rec |= rx[mm] == c;
    }
    const unsigned ball = __ballot_sync(0xffffffffu, rec);
    if (rec) rh_idx[(long long)c * kh + n + __popc(ball & ((1u << lane) - 1u))] = x;
    n += __popc(ball);
  }
  if (lane == 0) rh_cnt[c] = n;
}

__global__ void __launch_bounds__(kKrThreads, 16)
krecip_kernel(const float* __restrict__ E, long long lde, const float* __restrict__ rowmax,
              const int* __restrict__ rank, int K, int N, int k1, int kh, int cap, int* __restrict__ v_idx,
              __half* __restrict__ v_val, int* __restrict__ v_cnt, int row0, const int* __restrict__ rh_idx,
              const int* __restrict__ rh_cnt) {
  extern __shared__ unsigned s_dyn[];
  const int words = ceil_div(N, 32);
  unsigned* bm_kri = s_dyn;            // reciprocal set R(i, k1)
  unsigned* bm_exp = s_dyn + words;    // expanded set
  int* s_list = reinterpret_cast<int*>(s_dyn + 2 * words);  // [cap]
  float* s_w = reinterpret_cast<float*>(s_list + cap);       // [cap]
  __shared__ int s_fwd[kMaxK];
  __shared__ int s_kri[kMaxK];
  __shared__ int s_nkri;
  __shared__ unsigned s_scan[kKrThreads + 1];
  __shared__ float s_sum;
  // block li works on global row i = row0 + li; E / rowmax hold only the caller's rows (local
  // index), the neighbour lists and the V rows are indexed globally (row-sharded re-ranking)
  const int li = blockIdx.x, i = li + row0, t = threadIdx.x;
  const int K1 = k1 + 1;
  for (int w = t; w < 2 * words; w += kKrThreads) s_dyn[w] = 0;
  if (t < K1) s_fwd[t] = rank[(long long)i * K + t];
  if (t == 0) s_nkri = 0;
  __syncthreads();
  // R(i, k1) = { f in fwd : i in rank[f][0..k1] }
  if (t < K1) {
    const int f = s_fwd[t];
    const int* rf = rank + (long long)f * K;
    bool found = false;
    for (int m = 0; m < K1; ++m) found |= rf[m] == i;
    if (found) {
      s_kri[atomicAdd(&s_nkri, 1)] = f;
      bm_set(bm_kri, f);
      bm_set(bm_exp, f);
    }
  }
  __syncthreads();
  const int nkri = s_nkri;
  // expansion: candidates c in R(i); R(c, k1/2); append when |R(c) & R(i)| > 2/3 |R(c)|
  const int warp = t >> 5, lane = t & 31;
  for (int ci = warp; ci < nkri; ci += kKrThreads / 32) {
    const int c = s_kri[ci];
    const int len = rh_cnt[c];                       // |R(c, k1/2)| (precomputed, recip_half_kernel)
    const int* rc = rh_idx + (long long)c * kh;
    int inter = 0;
    for (int base = 0; base < len; base += 32) {
      const int m = base + lane;
      const bool in = m < len && bm_get(bm_kri, rc[m]);
      inter += __popc(__ballot_sync(0xffffffffu, in));
    }
    if (static_cast<double>(inter) > (2.0 / 3.0) * static_cast<double>(len))
      for (int m = lane; m < len; m += 32) bm_set(bm_exp, rc[m]);
  }
  __syncthreads();
  const int n = bm_enumerate(bm_exp, words, s_list, cap, s_scan);  // np.unique: sorted ascending
  const int nn = min(n, cap);
  const float div = rowmax[li];
  const float* erow = E + (long long)li * lde;
  for (int p = t; p < nn; p += kKrThreads) s_w[p] = expf(-(erow[s_list[p]] / div));
  __syncthreads();
  if (t == 0) s_sum = np_pairwise_sum(s_w, nn);
  __syncthreads();
  const float sum = s_sum;
  for (int p = t; p < nn; p += kKrThreads) {
    v_idx[(long long)i * cap + p] = s_list[p];
    v_val[(long long)i * cap + p] = __float2half_rn(s_w[p] / sum);
  }
  if (t == 0) v_cnt[i] = nn;
}

// ---------------------------------------------------------------------------------------
// local query expansion: V_qe[i] = fp16(mean_m V[rank[i][m]]), m < k2 (one block per row)
// ---------------------------------------------------------------------------------------
constexpr int kQeThreads = 128;
constexpr int kQeAcc = 2048;    // output slots accumulated per pass of the expansion kernel
constexpr int kQeGroup = 8;     // neighbour rows whose entries are fetched together

// The union of the k2 neighbour rows is built as a bitmap (= sorted unique columns); the weights are
// then SCATTERED into their output slot, found by a rank query on the bitmap (per-word prefix +
// popcount), one neighbour row after the other -- the float32 sums run in neighbour order exactly
// like np.mean over the fp16 rows, with O(entries) work instead of a binary search per
// (output column, neighbour).
__global__ void __launch_bounds__(kQeThreads)
expand_kernel(const int* __restrict__ rank, int K, int N, int k2, int cap, const int* __restrict__ v_idx,
              const __half* __restrict__ v_val, const int* __restrict__ v_cnt, int capq,
              int* __restrict__ q_idx, __half* __restrict__ q_val, int* __restrict__ q_cnt, int row0, int acc_slots,
              int* __restrict__ col_cnt, int count_from) {
  extern __shared__ unsigned s_dyn[];
  const int words = ceil_div(N, 32);
  unsigned* bm = s_dyn;
  unsigned* wp = s_dyn + words;                                   // [words] set bits before each word
  // fp32 accumulators for kQeAcc output slots at a time (the union of k2 rows has a few hundred
  // entries; its worst-case bound capq = min(N, k2 * cap) used to size TWO shared arrays -- 82 KB
  // at k1 = 50, k2 = 15, N = 10 290, two blocks per SM); the column list goes straight to q_idx
  float* s_acc = reinterpret_cast<float*>(s_dyn + 2 * words);     // [acc_slots]
  __shared__ unsigned s_scan[kQeThreads + 1];
  __shared__ int s_nb[64], s_cnt[64];
  const int i = blockIdx.x + row0, t = threadIdx.x;
  for (int w = t; w < words; w += kQeThreads) bm[w] = 0;
  if (t < k2) {
    s_nb[t] = rank[(long long)i * K + t];
    s_cnt[t] = v_cnt[s_nb[t]];
  }
  __syncthreads();
  for (int m = 0; m < k2; ++m) {
    const int r = s_nb[m];
    const int cnt = s_cnt[m];
    for (int p = t; p < cnt; p += kQeThreads)
      if (__half2float(v_val[(long long)r * cap + p]) != 0.f) bm_set(bm, v_idx[(long long)r * cap + p]);
  }
  __syncthreads();
  const int n = bm_enumerate(bm, words, q_idx + (long long)i * capq, capq, s_scan, wp);
  const int nn = min(n, capq);
  const float k2f = static_cast<float>(k2);
  for (int c0 = 0; c0 < nn; c0 += acc_slots) {       // one pass unless the union is unusually large
    const int cn = min(acc_slots, nn - c0);
    for (int p = t; p < cn; p += kQeThreads) s_acc[p] = 0.f;
    __syncthreads();
    // sequential float32 sum in neighbour order (np.mean over fp16 rows): one barrier per neighbour
    // row.  The first entry per thread of kQeGroup rows is fetched BEFORE their barriers, so that a
    // barrier interval holds shared-memory work only (a V row rarely has more than 128 entries).
    auto add = [&](float v, int c) {
      const unsigned pos = wp[c >> 5] + __popc(bm[c >> 5] & ((1u << (c & 31)) - 1u)) - static_cast<unsigned>(c0);
      if (pos < static_cast<unsigned>(cn)) s_acc[pos] += v;   // columns of one row are distinct: no conflict
    };
    for (int m0 = 0; m0 < k2; m0 += kQeGroup) {
      float v8[kQeGroup];
      int c8[kQeGroup];
#pragma unroll
      for (int u = 0; u < kQeGroup; ++u) {
        const int m = m0 + u;
        v8[u] = 0.f;
        c8[u] = 0;
        if (m < k2 && t < s_cnt[m]) {
          const long long e = (long long)s_nb[m] * cap + t;
          v8[u] = __half2float(v_val[e]);
          c8[u] = v_idx[e];
        }
      }
#pragma unroll
      for (int u = 0; u < kQeGroup; ++u) {
        const int m = m0 + u;
        if (m >= k2) break;
        if (v8[u] != 0.f) add(v8[u], c8[u]);
        const int cnt = s_cnt[m];
        for (int p = t + kQeThreads; p < cnt; p += kQeThreads) {   // longer rows
          const float v = __half2float(v_val[(long long)s_nb[m] * cap + p]);
          if (v != 0.f) add(v, v_idx[(long long)s_nb[m] * cap + p]);
        }
        __syncthreads();
      }
    }
    // col_cnt (one-GPU flow): the list lengths of the inverted index over the gallery rows are
    // counted here, where the entries are produced, instead of by a pass of their own
    const bool counted = col_cnt != nullptr && i >= count_from;
    for (int p = t; p < cn; p += kQeThreads) {
      const __half h = __float2half_rn(s_acc[p] / k2f);
      q_val[(long long)i * capq + c0 + p] = h;
      if (counted && __half2float(h) != 0.f) atomicAdd(&col_cnt[q_idx[(long long)i * capq + c0 + p]], 1);
    }
    __syncthreads();
  }
  if (t == 0) q_cnt[i] = nn;
}

// ---------------------------------------------------------------------------------------
// inverted index of the GALLERY rows of the final V (temp_min of a query row is only read at the
// gallery columns, :94-95; the query rows were 17 % of the entries at RGBNT100 scale and half of
// them at RGBNT201 scale).  An entry is ONE word -- gallery row in the upper half, fp16 weight in
// the lower half -- or, for galleries of more than 65 536 rows, a (row, weight) pair of words.
// The order inside a column list is irrelevant: its rows are distinct.
// ---------------------------------------------------------------------------------------
struct JcWide { unsigned row, val; };
__device__ __forceinline__ unsigned jc_make(unsigned, int g, __half v) {
  return (static_cast<unsigned>(g) << 16) | __half_as_ushort(v);
}
__device__ __forceinline__ JcWide jc_make(JcWide, int g, __half v) {
  return JcWide{static_cast<unsigned>(g), __half_as_ushort(v)};
}
__device__ __forceinline__ unsigned jc_row(unsigned e) { return e >> 16; }
__device__ __forceinline__ __half jc_val(unsigned e) { return __ushort_as_half(static_cast<unsigned short>(e & 0xffffu)); }
__device__ __forceinline__ unsigned jc_row(JcWide e) { return e.row; }
__device__ __forceinline__ __half jc_val(JcWide e) { return __ushort_as_half(static_cast<unsigned short>(e.val)); }
__device__ __forceinline__ unsigned jc_load(const unsigned* p) { return __ldg(p); }
__device__ __forceinline__ JcWide jc_load(const JcWide* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  return JcWide{u.x, u.y};
}

__global__ void inv_count_kernel(const int* __restrict__ idx, const __half* __restrict__ val, const int* __restrict__ cnt,
                                 int cap, int Q, int* __restrict__ col_cnt) {
  const int i = blockIdx.x + Q;
  const int c = cnt[i];
  for (int p = threadIdx.x; p < c; p += blockDim.x)
    if (__half2float(val[(long long)i * cap + p]) != 0.f) atomicAdd(&col_cnt[idx[(long long)i * cap + p]], 1);
}
template <typename Ent>
__global__ void inv_fill_kernel(const int* __restrict__ idx, const __half* __restrict__ val, const int* __restrict__ cnt,
                                int cap, int Q, const int* __restrict__ inv_ofs, int* __restrict__ cursor,
                                Ent* __restrict__ inv_ent) {
  const int g = blockIdx.x, i = g + Q;
  const int c = cnt[i];
  for (int p = threadIdx.x; p < c; p += blockDim.x) {
    const __half v = val[(long long)i * cap + p];
    if (__half2float(v) != 0.f) {
      const int col = idx[(long long)i * cap + p];
      inv_ent[inv_ofs[col] + atomicAdd(&cursor[col], 1)] = jc_make(Ent{}, g, v);
    }
  }
}

// The Jaccard term of the blend, fp16((1 - t / (2 - t)) * fp16(1 - lambda)) with numpy's half
// arithmetic, is a function of the 16-bit temp_min value alone: tabulated once per call for
// t in [0, 2) (bit patterns 0 .. 0x3fff; temp_min is a sum of minima of weights that add up to ~1).
// At k1 = 20 / k2 = 6 half of the Jaccard kernel's instructions were the blend epilogue's two IEEE
// divisions and five half roundings per output element; the table leaves one division.
constexpr int kJacTab = 0x4000;
__device__ __forceinline__ float jaccard_term(__half tm, __half w) {
  const __half one = __float2half_rn(1.f), two = __float2half_rn(2.f);
  const __half jac = np_hsub(one, np_hdiv(tm, np_hsub(two, tm)));           // 1 - tmin / (2 - tmin)
  return __half2float(np_hmul(jac, w));
}
// out-of-table values (t >= 2: not reachable from normalised weights) -- kept out of line so that the
// epilogue's unrolled body stays small
__device__ __noinline__ float jaccard_term_slow(__half tm, __half w) { return jaccard_term(tm, w); }
__global__ void jaccard_table_kernel(float one_minus_lambda_h, float* __restrict__ tab) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < kJacTab) tab[b] = jaccard_term(__ushort_as_half(static_cast<unsigned short>(b)), __float2half_rn(one_minus_lambda_h));
}

// ---------------------------------------------------------------------------------------
// Jaccard distance + blend (one block per query row).
// temp_min[t] receives, for ascending columns j of V[i], fp16(temp_min[t] + min(V[i,j], V[t,j]))
// (:89-93), so the columns are applied one after another with a block barrier in between (a row
// may be hit by different threads in consecutive columns).  The column metadata is staged in
// shared memory per chunk and the entries of 8 columns (kAhead per thread and column: 1 for lists
// of up to 256 gallery rows, 3 for the ~500-entry lists of k1 = 50 / k2 = 15) are fetched
// together before they are applied in order.
// Measured alternative (round 2, reverted): every warp owns the rows of one residue class and
// walks the columns alone with __syncwarp only -- 1.10 ms instead of 1.07 ms at k1 = 50 / k2 = 15
// and 0.24 instead of 0.115 ms at 20 / 6: the kernel is bound by shared-memory wavefronts (a
// read-modify-write of 32 random fp16 slots costs ~7 of them), not by the barriers, and the
// per-class lists fill the lanes worse.  Also measured and reverted: the gallery rows of a query cut
// into S ranges with their own inverted lists and one block per (query, range) -- 0.76 ms (S = 1),
// 0.92 ms (S = 2), 1.32 ms (S = 4) at 50 / 15: the time follows the number of (block, column)
// steps, i.e. the per-column barrier and metadata reads, not the entries per column.
// ---------------------------------------------------------------------------------------
constexpr int kJcThreads = 256;
constexpr int kJcChunk = 512;   // columns staged per round
constexpr int kJcGroup = 8;     // columns whose entries are fetched together

template <typename Ent, int kAhead, int kMinBlocks>
__global__ void __launch_bounds__(kJcThreads, kMinBlocks)
jaccard_kernel(const float* __restrict__ E, long long lde, const float* __restrict__ rowmax, int N, int Q,
               const int* __restrict__ idx, const __half* __restrict__ val, const int* __restrict__ cnt, int cap,
               const int* __restrict__ inv_ofs, const Ent* __restrict__ inv_ent,
               float one_minus_lambda_h, float lambda_f, __half* __restrict__ scratch, float* __restrict__ out,
               long long ldo, int row0, const float* __restrict__ jac_tab) {
  extern __shared__ __half s_tmin[];
  __shared__ int s_beg[kJcChunk], s_len[kJcChunk];
  __shared__ __half s_v[kJcChunk];
  const int li = blockIdx.x, i = li + row0, t = threadIdx.x;   // E / rowmax / scratch / out: local rows
  const int G = N - Q;
  __half* tmin = scratch ? scratch + (long long)li * G : s_tmin;   // gallery rows only
  for (int j = t; j < G; j += kJcThreads) tmin[j] = __float2half_rn(0.f);
  const int c = cnt[i];
  for (int p0 = 0; p0 < c; p0 += kJcChunk) {
    __syncthreads();  // tmin initialised / previous chunk consumed
    const int np = min(kJcChunk, c - p0);
    for (int p = t; p < np; p += kJcThreads) {
      const int col = idx[(long long)i * cap + p0 + p];
      const __half v = val[(long long)i * cap + p0 + p];
      const int beg = inv_ofs[col];
      s_v[p] = v;
      s_beg[p] = beg;
      s_len[p] = __half2float(v) != 0.f ? inv_ofs[col + 1] - beg : 0;   // V[i, col] == 0 -> column skipped (:88)
    }
    __syncthreads();
    // columns in groups of kJcGroup: every thread first fetches its entries of each column of the
    // group (independent loads, one L2 latency for the whole group), then the columns are applied
    // in order
    for (int p8 = 0; p8 < np; p8 += kJcGroup) {
      Ent ent[kJcGroup][kAhead];
#pragma unroll
      for (int u = 0; u < kJcGroup; ++u) {
        const int p = p8 + u;
#pragma unroll
        for (int k = 0; k < kAhead; ++k)
          if (p < np && t + k * kJcThreads < s_len[p]) ent[u][k] = jc_load(inv_ent + s_beg[p] + t + k * kJcThreads);
      }
#pragma unroll
      for (int u = 0; u < kJcGroup; ++u) {  // ascending column order (:89-93)
        const int p = p8 + u;
        if (p >= np) break;
        const int len = s_len[p];
        if (len == 0) continue;  // block-uniform
        // native half min / add (IEEE round-to-nearest): identical to numpy's "compute in float32,
        // round once" for + of two halfs; 3 instructions instead of 8
        const __half vh = s_v[p];
#pragma unroll
        for (int k = 0; k < kAhead; ++k)
          if (t + k * kJcThreads < len) {
            const unsigned r = jc_row(ent[u][k]);
            tmin[r] = __hadd(tmin[r], __hmin(vh, jc_val(ent[u][k])));
          }
        if (len > kAhead * kJcThreads) {     // longer lists
          const Ent* lst = inv_ent + s_beg[p];
          for (int q = kAhead * kJcThreads + t; q < len; q += kJcThreads) {
            const Ent en = jc_load(lst + q);
            const unsigned r = jc_row(en);
            tmin[r] = __hadd(tmin[r], __hmin(vh, jc_val(en)));
          }
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  const __half w = __float2half_rn(one_minus_lambda_h);
  const float div = rowmax[li];
  // 8 independent row loads in flight per thread: with 64 KB of temp_min per block only three blocks
  // fit an SM, so the memory-level parallelism has to come from the thread
  const float* erow = E + (long long)li * lde + Q;
  float* orow = out + (long long)li * ldo;
  for (int g0 = t; g0 < G; g0 += 8 * kJcThreads) {
    float e[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * kJcThreads;
      e[u] = g < G ? __ldcs(erow + g) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int g = g0 + u * kJcThreads;
      if (g >= G) break;
      const __half tm = tmin[g];
      const unsigned tb = __half_as_ushort(tm);
      const float jt = tb < static_cast<unsigned>(kJacTab) ? __ldg(jac_tab + tb) : jaccard_term_slow(tm, w);
      const float od = e[u] / div;
      __stcs(orow + g, __fadd_rn(jt, __fmul_rn(od, lambda_f)));  // (:95), no FMA contraction
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int launch_rowmax(const float* E, long long lde, int N, float* rowmax, cudaStream_t stream) {
  rowmax_kernel<<<N, 256, 0, stream>>>(E, lde, N, rowmax);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

int launch_transpose_add(const float* X, long long ldx, float* E, long long lde, int N, bool accumulate,
                         cudaStream_t stream) {
  dim3 grid(ceil_div(N, 32), ceil_div(N, 32)), block(32, 8);
  transpose_add_kernel<<<grid, block, 0, stream>>>(X, ldx, E, lde, N, accumulate ? 1 : 0);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

int launch_topk_rows(const float* mat, long long ld, int rows, int cols, const float* row_div, int k, int* idx_out,
                     float* val_out, cudaStream_t stream) {
  DEMO_REQUIRE(k >= 1 && k <= 256 && k <= cols, "topk: k=%d out of range (cols=%d, max 256)", k, cols);
  if (rows <= 0) return DEMO_OK;
  // Long rows are not staged in shared memory: the second (filter) pass re-reads them from L2,
  // which keeps 8 blocks resident per SM instead of 4 and hides the load latency better.
  const size_t smem = static_cast<size_t>(cols) * 4;
  if (smem <= 16 * 1024) {
    static PerDeviceInt configured;
    DEMO_CHECK_CUDA(ensure_dynamic_smem(configured, topk_rows_kernel<true>, 200 * 1024));
    topk_rows_kernel<true><<<rows, kTopkThreads, smem, stream>>>(mat, ld, cols, row_div, k, idx_out, val_out, cols);
  } else {
    // prefix for the bound: the expected number of candidates of the filter pass is about
    // (cols / prefix) * 64 * ln(64 / (64 - k)) (k-th smallest of 64 group minima); keep it near
    // half the candidate buffer.  k > 63 uses 256 single-thread groups and the whole row.
    // Rows up to ~12k columns are re-read from L2 by the filter pass (resident blocks keep
    // < 64 MB in flight); beyond that the second pass would come from HBM again.
    int prefix = cols;
    if (k < 64 && cols > 12288) {
      const double want = static_cast<double>(cols) * 64.0 * log(64.0 / (64.0 - k)) / (kTopkCap / 2);
      prefix = static_cast<int>(want < 8192.0 ? 8192.0 : (want > cols ? cols : want));
      prefix &= ~1023;
    }
    topk_rows_kernel<false><<<rows, kTopkThreads, 0, stream>>>(mat, ld, cols, row_div, k, idx_out, val_out, prefix);
  }
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

int rerank_k(int k1, int k2) { return (k1 + 1 > k2 ? k1 + 1 : k2); }
int rerank_kh(int k1) {
  // int(np.around(k1 / 2)) + 1 with round-half-to-even (appendix A7)
  const int half2 = k1;  // k1/2 = half2/2
  int r = half2 / 2;
  if (half2 % 2 == 1 && (r % 2 == 1)) r += 1;  // x.5 -> even
  return r + 1;
}
int rerank_cap(int k1) { return (k1 + 1) * (rerank_kh(k1) + 1); }
int rerank_capq(int N, int k1, int k2) {
  const long long c = static_cast<long long>(k2 > 1 ? k2 : 1) * rerank_cap(k1);
  return static_cast<int>(c < N ? c : N);
}

constexpr size_t kJcMaxSmemTmin = 200 * 1024;
constexpr int kJcPackedRows = 65536;   // gallery rows a one-word inverted-list entry can address
// The two large-problem variants (two-word entries beyond 65 536 gallery rows, temp_min in global
// memory beyond ~100 000) can be forced at any size, so that tests reach them without a 17 GB matrix.
static bool jc_packed(size_t G) {
  static const bool force_wide = getenv("DEMO_JC_WIDE") != nullptr;
  return G <= static_cast<size_t>(kJcPackedRows) && !force_wide;
}
static bool jc_scratch(size_t G) {
  static const bool force_scratch = getenv("DEMO_JC_SCRATCH") != nullptr;
  return G * 2 > kJcMaxSmemTmin || force_scratch;
}

template <typename Ent, int kAhead, int kMinBlocks>
static int launch_jaccard_t(int nq, size_t smem, cudaStream_t stream, const float* E, long long lde, const float* rowmax,
                            int N, int Q, const int* f_idx, const __half* f_val, const int* f_cnt, int f_cap,
                            const RerankWs& w, float oml, float lambda_f, float* out, long long ldo, int row0) {
  auto kern = jaccard_kernel<Ent, kAhead, kMinBlocks>;
  DEMO_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kJcMaxSmemTmin)));
  jaccard_table_kernel<<<kJacTab / 256, 256, 0, stream>>>(oml, w.jac_tab);
  kern<<<nq, kJcThreads, smem, stream>>>(E, lde, rowmax, N, Q, f_idx, f_val, f_cnt, f_cap, w.inv_ofs,
                                         static_cast<const Ent*>(w.inv_ent), oml, lambda_f, w.tmin_scratch, out, ldo, row0,
                                         w.jac_tab);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

// inverted index of the gallery rows of the final V + Jaccard / blend rows of the queries
// [row0, row0 + nq): shared by the one-GPU and the row-sharded flow
static int launch_index_and_jaccard(const float* E, long long lde, const float* rowmax, int N, int Q, int k1, int k2,
                                    double lambda_value, int row0, int nq, const int* f_idx, const __half* f_val,
                                    const int* f_cnt, int f_cap, const RerankWs& w, float* out, long long ldo,
                                    cudaStream_t stream, bool counted = false) {
  const int G = N - Q;
  DEMO_REQUIRE(G >= 1, "re_ranking: empty gallery");
  const bool packed = jc_packed(G);
  DEMO_CHECK_CUDA(cudaMemsetAsync(w.cursor, 0, sizeof(int) * (N + 1), stream));
  if (!counted) {   // counted: the expansion kernel has already filled col_cnt (run_rerank_stages)
    DEMO_CHECK_CUDA(cudaMemsetAsync(w.col_cnt, 0, sizeof(int) * (N + 1), stream));
    inv_count_kernel<<<G, 128, 0, stream>>>(f_idx, f_val, f_cnt, f_cap, Q, w.col_cnt);
  }
  size_t tmp = w.cub_bytes;
  DEMO_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(w.cub_tmp, tmp, w.col_cnt, w.inv_ofs, N + 1, stream));
  if (packed)
    inv_fill_kernel<unsigned><<<G, 128, 0, stream>>>(f_idx, f_val, f_cnt, f_cap, Q, w.inv_ofs, w.cursor,
                                                     static_cast<unsigned*>(w.inv_ent));
  else
    inv_fill_kernel<JcWide><<<G, 128, 0, stream>>>(f_idx, f_val, f_cnt, f_cap, Q, w.inv_ofs, w.cursor,
                                                   static_cast<JcWide*>(w.inv_ent));
  DEMO_CHECK_CUDA(cudaGetLastError());
  if (nq <= 0) return DEMO_OK;
  // (1 - lambda) is a Python float: numpy multiplies the fp16 array by fp16(1 - lambda)
  const float oml = __half2float(__double2half(1.0 - lambda_value)), lam = static_cast<float>(lambda_value);
  const size_t smem = w.tmin_scratch ? 0 : static_cast<size_t>(G) * 2;
  // Entries per column list ~ non-zeros per row of the final V x (gallery share of the rows):
  // ~0.8 k1 k2 after the expansion (measured 96 at 20 / 6 and 583 at 50 / 15 on RGBNT100-like
  // data), ~k1 without it.  The estimate only picks the prefetch depth; any length is handled.
  const double est = (k2 > 1 ? 0.8 * k1 * k2 : 1.0 * k1) * G / N;
  static const int force = getenv("DEMO_JC_AHEAD") ? atoi(getenv("DEMO_JC_AHEAD")) : 0;   // A/B experiments
  // measured at 50 / 15 (lists of ~480 gallery rows): depth 2 0.77 ms, depth 3 0.80 ms, depth 1 (round 1) 1.07 ms
  const int ahead = force ? force : est <= 0.8 * kJcThreads ? 1 : est <= 3.0 * kJcThreads ? 2 : 3;
  // (128-thread blocks for the short lists of k1 = 20 / k2 = 6 -- 3 of 8 warps have entries there --
  // measured the same 0.098 ms: at that size half of the kernel's instructions are the two IEEE
  // divisions per output element of the blend epilogue, not the column walk.)
  // resident blocks per SM: 6 for depth 1 and 2 (40 registers, no spills; depth 1 at 8 blocks = 32 registers
  // spills: 0.092 instead of 0.082 ms at 20 / 6, depth 2 at 8 blocks: 1.22 instead of 0.72 ms at 50 / 15)
#define DEMO_JC(ENT, A, B) \
  return launch_jaccard_t<ENT, A, B>(nq, smem, stream, E, lde, rowmax, N, Q, f_idx, f_val, f_cnt, f_cap, w, oml, lam, out, ldo, row0)
  if (packed) {
    if (ahead == 1) DEMO_JC(unsigned, 1, 6);
    if (ahead == 2) DEMO_JC(unsigned, 2, 6);
    DEMO_JC(unsigned, 3, 4);
  }
  if (ahead == 1) DEMO_JC(JcWide, 1, 5);   // two-word entries: 48 / 64 registers without spills
  if (ahead == 2) DEMO_JC(JcWide, 2, 4);
  DEMO_JC(JcWide, 3, 3);
#undef DEMO_JC
}

size_t rerank_carve(Carver& c, int N, int Q, int k1, int k2, RerankWs* w) {
  RerankWs t;
  const size_t n = N > 0 ? N : 1;
  t.K = rerank_k(k1, k2);
  t.cap = rerank_cap(k1);
  t.capq = rerank_capq(N, k1, k2);
  t.rank = c.take<int>(n * t.K);
  t.v_idx = c.take<int>(n * t.cap);
  t.v_val = c.take<__half>(n * t.cap);
  t.v_cnt = c.take<int>(n);
  t.rh_idx = c.take<int>(n * rerank_kh(k1));
  t.rh_cnt = c.take<int>(n);
  t.q_idx = c.take<int>(n * t.capq);
  t.q_val = c.take<__half>(n * t.capq);
  t.q_cnt = c.take<int>(n);
  const size_t g = N > Q && Q > 0 ? N - Q : n;   // gallery rows: the only ones in the inverted index
  t.col_cnt = c.take<int>(n + 1);
  t.inv_ofs = c.take<int>(n + 1);
  t.cursor = c.take<int>(n + 1);
  t.inv_ent = c.take<unsigned>(g * t.capq * (jc_packed(g) ? 1 : 2));
  t.jac_tab = c.take<float>(kJacTab);
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, t.col_cnt, t.inv_ofs, N + 1);
  t.cub_bytes = tmp + 256;
  t.cub_tmp = c.take<char>(t.cub_bytes);
  t.tmin_scratch = nullptr;
  t.tmin_bytes = 0;
  if (jc_scratch(g)) {  // temp_min (gallery rows) does not fit shared memory
    const size_t q = Q > 0 ? Q : 1;
    t.tmin_bytes = q * g * 2;
    t.tmin_scratch = c.take<__half>(q * g);
  }
  if (w) *w = t;
  return c.off;
}

int run_rerank_stages(const float* E, long long lde, const float* rowmax, int N, int Q, int k1, int k2,
                      double lambda_value, const RerankWs& w, float* out, long long ldo, cudaStream_t stream) {
  DEMO_REQUIRE(k1 >= 1 && k1 + 1 <= kMaxK && k1 + 1 <= N, "re_ranking: need 1 <= k1 < min(N, %d) (k1=%d, N=%d)", kMaxK, k1, N);
  DEMO_REQUIRE(k2 >= 1 && k2 <= 64 && k2 <= N, "re_ranking: need 1 <= k2 <= min(N, 64) (k2=%d)", k2);
  DEMO_REQUIRE(Q >= 1 && Q < N, "re_ranking: need 1 <= Q < N");
  const int K = w.K, kh = rerank_kh(k1), words = ceil_div(N, 32);
  DEMO_TRY(launch_topk_rows(E, lde, N, N, rowmax, K, w.rank, nullptr, stream));
  {
    const size_t smem = static_cast<size_t>(2 * words) * 4 + static_cast<size_t>(w.cap) * 8;
    DEMO_REQUIRE(smem <= 200 * 1024, "re_ranking: N=%d too large for the shared-memory bitmaps", N);
    DEMO_CHECK_CUDA(cudaFuncSetAttribute(krecip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    recip_half_kernel<<<ceil_div(N * 32, 128), 128, 0, stream>>>(w.rank, K, N, kh, w.rh_idx, w.rh_cnt);
    krecip_kernel<<<N, kKrThreads, smem, stream>>>(E, lde, rowmax, w.rank, K, N, k1, kh, w.cap, w.v_idx, w.v_val, w.v_cnt, 0,
                                                   w.rh_idx, w.rh_cnt);
    DEMO_CHECK_CUDA(cudaGetLastError());
  }
  const int* f_idx = w.v_idx;
  const __half* f_val = w.v_val;
  const int* f_cnt = w.v_cnt;
  int f_cap = w.cap;
  const bool counted = k2 != 1;
  if (k2 != 1) {
    DEMO_CHECK_CUDA(cudaMemsetAsync(w.col_cnt, 0, sizeof(int) * (N + 1), stream));
    DEMO_TRY(launch_expand_rows(w.rank, N, k1, k2, 0, N, w.v_idx, w.v_val, w.v_cnt, w.q_idx, w.q_val, w.q_cnt, stream,
                                w.col_cnt, Q));
    f_idx = w.q_idx;
    f_val = w.q_val;
    f_cnt = w.q_cnt;
    f_cap = w.capq;
  }
  DEMO_TRY(launch_index_and_jaccard(E, lde, rowmax, N, Q, k1, k2, lambda_value, 0, Q, f_idx, f_val, f_cnt, f_cap, w, out, ldo,
                                    stream, counted));
  return DEMO_OK;
}


// ---------------------------------------------------------------------------------------
// Row-sharded stages (multi-GPU re-ranking): a rank owns the contiguous rows
// [row0, row0 + nrows) of the N x N problem.  E / rowmax are LOCAL ([nrows][lde] / [nrows]);
// the neighbour lists and the sparse V rows are full-size arrays that the host all-gathers
// between the stages, so every kernel indexes them by global row.
// ---------------------------------------------------------------------------------------
int launch_krecip_rows(const float* E, long long lde, const float* rowmax, const int* rank_all, int N, int k1,
                       int k2, int row0, int nrows, int* v_idx, __half* v_val, int* v_cnt, int* rh_idx,
                       int* rh_cnt, cudaStream_t stream) {
  DEMO_REQUIRE(k1 >= 1 && k1 + 1 <= kMaxK && k1 + 1 <= N, "re_ranking: need 1 <= k1 < min(N, %d) (k1=%d, N=%d)", kMaxK, k1, N);
  if (nrows <= 0) return DEMO_OK;
  const int K = rerank_k(k1, k2), kh = rerank_kh(k1), cap = rerank_cap(k1), words = ceil_div(N, 32);
  const size_t smem = static_cast<size_t>(2 * words) * 4 + static_cast<size_t>(cap) * 8;
  DEMO_REQUIRE(smem <= 200 * 1024, "re_ranking: N=%d too large for the shared-memory bitmaps", N);
  DEMO_CHECK_CUDA(cudaFuncSetAttribute(krecip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  // every rank needs R(c, k1/2) of all rows (any row can be a candidate of a local row)
  recip_half_kernel<<<ceil_div(N * 32, 128), 128, 0, stream>>>(rank_all, K, N, kh, rh_idx, rh_cnt);
  krecip_kernel<<<nrows, kKrThreads, smem, stream>>>(E, lde, rowmax, rank_all, K, N, k1, kh, cap, v_idx, v_val, v_cnt, row0,
                                                     rh_idx, rh_cnt);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

int launch_expand_rows(const int* rank_all, int N, int k1, int k2, int row0, int nrows, const int* v_idx,
                       const __half* v_val, const int* v_cnt, int* q_idx, __half* q_val, int* q_cnt,
                       cudaStream_t stream, int* col_cnt, int count_from) {
  DEMO_REQUIRE(k2 >= 2 && k2 <= 64 && k2 <= N, "re_ranking: expansion needs 2 <= k2 <= min(N, 64) (k2=%d)", k2);
  if (nrows <= 0) return DEMO_OK;
  const int K = rerank_k(k1, k2), cap = rerank_cap(k1), capq = rerank_capq(N, k1, k2), words = ceil_div(N, 32);
  // DEMO_QE_ACC: tests force the multi-pass accumulation (unions of more than kQeAcc = 2048 columns
  // do not occur in practice: ~1000 at k1 = 120 / k2 = 64 on adversarial low-dimensional data)
  static const int acc_env = getenv("DEMO_QE_ACC") ? atoi(getenv("DEMO_QE_ACC")) : 0;
  const int acc_slots = acc_env >= 32 && acc_env <= kQeAcc ? acc_env : kQeAcc;
  const size_t smem = static_cast<size_t>(2 * words) * 4 + static_cast<size_t>(acc_slots) * 4;
  DEMO_REQUIRE(smem <= 200 * 1024, "re_ranking: N=%d too large for the expansion kernel", N);
  DEMO_CHECK_CUDA(cudaFuncSetAttribute(expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  expand_kernel<<<nrows, kQeThreads, smem, stream>>>(rank_all, K, N, k2, cap, v_idx, v_val, v_cnt, capq, q_idx, q_val,
                                                     q_cnt, row0, acc_slots, col_cnt, count_from);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

// Inverted index over the gallery rows of the (gathered) final V, then the Jaccard / blend rows of
// the local queries [row0, row0 + nq_local), row0 + nq_local <= Q.
int launch_jaccard_rows(const float* E, long long lde, const float* rowmax, int N, int Q, int k1, int k2,
                        double lambda_value, int row0, int nq_local, const int* f_idx, const __half* f_val, const int* f_cnt, int f_cap,
                        const RerankWs& w, float* out, long long ldo, cudaStream_t stream) {
  return launch_index_and_jaccard(E, lde, rowmax, N, Q, k1, k2, lambda_value, row0, nq_local, f_idx, f_val, f_cnt, f_cap, w, out,
                                  ldo, stream);
}

}  // namespace demo
