// Epilogue policies for sqdist_gemm_kernel.  Each epilogue thread owns ONE accumulator row
// (one query) and walks the 256 columns of the tile in chunks of 32 TMEM columns; per-column
// metadata (|g|^2, 2^-e, labels) is staged once per tile in shared memory and read as
// warp-wide broadcasts.
//
//   dist = fma(-2, acc * ia[m] * ib[n], na[m] + nb[n])     (utils/metrics.py:398-400 order)
#pragma once

#include "gemm_sm100.cuh"

namespace demo {

__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}

// Distance flavours shared by the store / mining epilogues and the SIMT reference kernel.
enum : int {
  DIST_SQ = 0,       // |a|^2 + |b|^2 - 2ab              euclidean_distance  (metrics.py:395-401)
  DIST_SQRT = 1,     // sqrt(clamp(., 1e-12))            euclidean_dist      (triplet_loss.py:16-31)
  DIST_COS_SIM = 2,  // ab / (|a||b|)                    cosine_similarity   (north_star)
  DIST_COS_DIST = 3  // (1 - ab/(|a||b|)) / 2            cosine_dist         (triplet_loss.py:34-48)
};

__device__ __forceinline__ float finish_distance(int mode, float dot, float na, float nb) {
  if (mode == DIST_SQ) return fmaf(-2.f, dot, na + nb);
  if (mode == DIST_SQRT) return sqrtf(fmaxf(fmaf(-2.f, dot, na + nb), 1e-12f));
  const float c = dot / (sqrtf(na) * sqrtf(nb));
  return mode == DIST_COS_SIM ? c : (1.f - c) * 0.5f;
}

// Common base: stages b_norm / b_inv (and optionally an int label) for the tile's columns.
struct EpiColumns {
  float* s_bnorm;  // [2][kBN]
  float* s_binv;   // [2][kBN]
  int* s_blab;     // [2][kBN] (optional)
  static constexpr int kBytes = 2 * kBN * 4 * 3;

  __device__ EpiColumns(uint8_t* smem) {
    s_bnorm = reinterpret_cast<float*>(smem);
    s_binv = s_bnorm + 2 * kBN;
    s_blab = reinterpret_cast<int*>(s_binv + 2 * kBN);
  }
  __device__ __forceinline__ void stage(const TileInfo& t, int as, int epi_tid, const float* b_norm,
                                        const float* b_inv, const int* b_lab) {
    for (int c = epi_tid; c < kBN; c += kEpiThreads) {
      const bool ok = c < t.n_valid;
      s_bnorm[as * kBN + c] = ok ? __ldg(b_norm + t.n0 + c) : 0.f;
      s_binv[as * kBN + c] = ok ? __ldg(b_inv + t.n0 + c) : 0.f;
      if (b_lab) s_blab[as * kBN + c] = ok ? __ldg(b_lab + t.n0 + c) : -0x7fffffff;
    }
    epi_bar_sync();
  }
};

// ---------------------------------------------------------------------------------------
// STORE: materialise the matrix (euclidean_distance, cosine_*, re-ranking all-pairs)
// ---------------------------------------------------------------------------------------
struct EpiStore {
  static constexpr int kStages = 4;
  static constexpr int kSmemBytes = EpiColumns::kBytes;
  struct Params {
    const float* a_norm;
    const float* a_inv;
    const float* b_norm;
    const float* b_inv;
    float* out;
    long long ldo;
    int M;
    int mode;
    unsigned* rowmax_key;  // optional: per-row max of the stored values (ordered-uint keys)
  };
  const Params& p;
  EpiColumns cols;
  int epi_tid, row_in_tile;

  __device__ EpiStore(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_)
      : p(p_), cols(smem), epi_tid(epi_tid_), row_in_tile(row_in_tile_) {}

  __device__ void tile_begin(const TileInfo& t, int as) {
    cols.stage(t, as, epi_tid, p.b_norm, p.b_inv, nullptr);
  }
  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const int row = t.m0 + row_in_tile;
    const bool row_ok = row < p.M;
    const float na = row_ok ? __ldg(p.a_norm + row) : 0.f;
    const float ia = row_ok ? __ldg(p.a_inv + row) : 0.f;
    float* orow = p.out + static_cast<long long>(row) * p.ldo + t.n0;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(orow) & 15u) == 0);
    float vmax = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-row branches
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      if (!row_ok || c * 32 >= t.n_valid) continue;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cc = c * 32 + j;
        const float dot = __uint_as_float(r[j]) * ia * cols.s_binv[as * kBN + cc];
        v[j] = finish_distance(p.mode, dot, na, cols.s_bnorm[as * kBN + cc]);
      }
      const int nv = min(32, t.n_valid - c * 32);
      if (nv == 32 && vec_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(orow + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nv) orow[c * 32 + j] = v[j];
      }
      if (p.rowmax_key) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nv) vmax = fmaxf(vmax, v[j]);
      }
    }
    if (p.rowmax_key && row_ok && vmax > -INFINITY) atomicMax(p.rowmax_key + row, float_key(vmax));
  }
  __device__ void tile_end(const TileInfo&, int) {}
};

// ---------------------------------------------------------------------------------------
// COUNT: rank counts without materialising Q x G.
// For its row q the thread holds up to kWin thresholds (the distances of q's valid positives,
// sorted ascending by (distance, gallery index)) in a private shared-memory column and, for
// every gallery column g, finds   b = #{ j : (t_j, p_j) <=lex (d(q,g), g) }   by binary search
// and bumps hist[b].  At the end of the unit  #{g before threshold j} = sum_{b<=j} hist[b]
// is added to counts[].  No labels are read here: junk / positives are subtracted later from
// the per-query record list (they all belong to it).
// ---------------------------------------------------------------------------------------
constexpr int kWin = 63;

struct EpiCount {
  static constexpr int kStages = 3;
  static constexpr int kSmemBytes = EpiColumns::kBytes + kWin * kEpiThreads * 4 + (kWin + 1) * kEpiThreads * 4;
  struct Params {
    const float* a_norm;
    const float* a_inv;
    const float* b_norm;
    const float* b_inv;
    const int* b_gidx;        // global gallery index of every B row (tie-break)
    const int* thr_ofs;       // [M] offsets into thr_val / thr_gidx / counts
    const int* thr_cnt;       // [M] number of thresholds (valid positives) of the row
    const float* thr_val;     // thresholds, ascending per row
    const int* thr_gidx;      // their global gallery indices
    unsigned* counts;         // += #{gallery columns lexicographically before the threshold}
    int M;
    int window;               // thresholds [window*kWin, window*kWin + kWin) of every row
  };
  const Params& p;
  EpiColumns cols;
  float* s_thr;   // [kWin][kEpiThreads]   thread-private column = epi_tid
  unsigned* s_hist;  // [kWin+1][kEpiThreads]
  int epi_tid, row_in_tile;
  int nthr = 0, tbase = 0;

  __device__ EpiCount(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_)
      : p(p_), cols(smem), epi_tid(epi_tid_), row_in_tile(row_in_tile_) {
    s_thr = reinterpret_cast<float*>(smem + EpiColumns::kBytes);
    s_hist = reinterpret_cast<unsigned*>(s_thr + kWin * kEpiThreads);
  }

  __device__ void tile_begin(const TileInfo& t, int as) {
    if (t.first_in_unit) {
      const int row = t.m0 + row_in_tile;
      nthr = 0;
      if (row < p.M) {
        tbase = __ldg(p.thr_ofs + row) + p.window * kWin;
        nthr = max(0, min(kWin, __ldg(p.thr_cnt + row) - p.window * kWin));
      }
      for (int k = 0; k < kWin; ++k)
        s_thr[k * kEpiThreads + epi_tid] = k < nthr ? __ldg(p.thr_val + tbase + k) : INFINITY;
      for (int k = 0; k <= kWin; ++k) s_hist[k * kEpiThreads + epi_tid] = 0u;
    }
    cols.stage(t, as, epi_tid, p.b_norm, p.b_inv, nullptr);
  }

  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const int row = t.m0 + row_in_tile;
    const bool active = nthr > 0;
    const float na = active ? __ldg(p.a_norm + row) : 0.f;
    const float ia = active ? __ldg(p.a_inv + row) : 0.f;
    const float* thr = s_thr + epi_tid;
    unsigned* hist = s_hist + epi_tid;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-row branches
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      const int nv = min(32, t.n_valid - c * 32);
      if (!active || nv <= 0) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j >= nv) continue;
        const int cc = c * 32 + j;
        const float dot = __uint_as_float(r[j]) * ia * cols.s_binv[as * kBN + cc];
        const float d = fmaf(-2.f, dot, na + cols.s_bnorm[as * kBN + cc]);
        int pos = 0;
#pragma unroll
        for (int step = 32; step >= 1; step >>= 1)
          if (thr[(pos + step - 1) * kEpiThreads] <= d) pos += step;
        if (pos > 0 && thr[(pos - 1) * kEpiThreads] == d) {
          // exact tie with a threshold: thresholds with a larger gallery index come AFTER this column
          const int g = __ldg(p.b_gidx + t.n0 + cc);
          while (pos > 0 && thr[(pos - 1) * kEpiThreads] == d && __ldg(p.thr_gidx + tbase + pos - 1) > g) --pos;
        }
        hist[pos * kEpiThreads] += 1u;
      }
    }
  }

  __device__ void tile_end(const TileInfo& t, int) {
    if (t.last_in_unit && nthr > 0) {
      unsigned run = 0;
      for (int k = 0; k < nthr; ++k) {
        run += s_hist[k * kEpiThreads + epi_tid];
        if (run) atomicAdd(p.counts + tbase + k, run);
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
// EXTRACT: distances of the same-identity pairs (positives + junk) of every query.
// Gallery rows are sorted by pid, so the candidates of query q are the contiguous rows
// [g_lo[q], g_lo[q] + cnt[q]) and the record slot is rec_base[q] + (row - g_lo[q]).
// ---------------------------------------------------------------------------------------
struct EpiExtract {
  static constexpr int kStages = 4;
  static constexpr int kSmemBytes = EpiColumns::kBytes;
  struct Params {
    const float* a_norm;
    const float* a_inv;
    const float* b_norm;
    const float* b_inv;
    const int* a_pid;      // [M] pid of the (sorted) query rows
    const int* b_pid;      // [N] pid of the (sorted) gallery rows
    const int* g_lo;       // [M] first sorted-gallery row with the query's pid
    const int* rec_base;   // [M+1] record offsets
    float* rec_dist;       // [T]
    int M;
  };
  const Params& p;
  EpiColumns cols;
  int epi_tid, row_in_tile;

  __device__ EpiExtract(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_)
      : p(p_), cols(smem), epi_tid(epi_tid_), row_in_tile(row_in_tile_) {}

  __device__ void tile_begin(const TileInfo& t, int as) {
    cols.stage(t, as, epi_tid, p.b_norm, p.b_inv, p.b_pid);
  }
  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const int row = t.m0 + row_in_tile;
    const bool row_ok = row < p.M;
    const float na = row_ok ? __ldg(p.a_norm + row) : 0.f;
    const float ia = row_ok ? __ldg(p.a_inv + row) : 0.f;
    const int pid = row_ok ? __ldg(p.a_pid + row) : 0;
    const int lo = row_ok ? __ldg(p.g_lo + row) : 0;
    const int base = row_ok ? __ldg(p.rec_base + row) : 0;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-row branches
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cc = c * 32 + j;
        if (cc < t.n_valid && cols.s_blab[as * kBN + cc] == pid) {
          const float dot = __uint_as_float(r[j]) * ia * cols.s_binv[as * kBN + cc];
          p.rec_dist[base + (t.n0 + cc - lo)] = fmaf(-2.f, dot, na + cols.s_bnorm[as * kBN + cc]);
        }
      }
    }
  }
  __device__ void tile_end(const TileInfo&, int) {}
};

// ---------------------------------------------------------------------------------------
// MINE: batch-hard triplet mining fused into the distance epilogue
// (layers/triplet_loss.py:51-104 on euclidean_dist(x, x)).  Hardest positive = max distance
// among same-label columns (self included), hardest negative = min among the others; ties go
// to the lowest index.  Results are combined across column tiles with 64-bit atomics on
// (ordered distance key << 32 | index code).
// ---------------------------------------------------------------------------------------
struct EpiMine {
  static constexpr int kStages = 4;
  static constexpr int kSmemBytes = EpiColumns::kBytes;
  struct Params {
    const float* a_norm;
    const float* a_inv;
    const float* b_norm;
    const float* b_inv;
    const int* a_lab;
    const int* b_lab;
    unsigned long long* best_pos;  // init 0;            max of (key(d) << 32 | ~idx)
    unsigned long long* best_neg;  // init 0xFFFF...F;   min of (key(d) << 32 |  idx)
    int M;
  };
  const Params& p;
  EpiColumns cols;
  int epi_tid, row_in_tile;

  __device__ EpiMine(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_)
      : p(p_), cols(smem), epi_tid(epi_tid_), row_in_tile(row_in_tile_) {}

  __device__ void tile_begin(const TileInfo& t, int as) {
    cols.stage(t, as, epi_tid, p.b_norm, p.b_inv, p.b_lab);
  }
  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const int row = t.m0 + row_in_tile;
    const bool row_ok = row < p.M;
    const float na = row_ok ? __ldg(p.a_norm + row) : 0.f;
    const float ia = row_ok ? __ldg(p.a_inv + row) : 0.f;
    const int lab = row_ok ? __ldg(p.a_lab + row) : 0;
    unsigned long long bp = 0ull, bn = ~0ull;
#pragma unroll 1
    for (int c = 0; c < kBN / 32; ++c) {
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-row branches
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cc = c * 32 + j;
        if (cc >= t.n_valid) continue;
        const float dot = __uint_as_float(r[j]) * ia * cols.s_binv[as * kBN + cc];
        const float d = finish_distance(DIST_SQRT, dot, na, cols.s_bnorm[as * kBN + cc]);
        const unsigned key = float_key(d);
        const unsigned idx = static_cast<unsigned>(t.n0 + cc);
        if (cols.s_blab[as * kBN + cc] == lab) {
          const unsigned long long v = (static_cast<unsigned long long>(key) << 32) | (0xFFFFFFFFu - idx);
          bp = v > bp ? v : bp;
        } else {
          const unsigned long long v = (static_cast<unsigned long long>(key) << 32) | idx;
          bn = v < bn ? v : bn;
        }
      }
    }
    if (row_ok) {
      if (bp != 0ull) atomicMax(p.best_pos + row, bp);
      if (bn != ~0ull) atomicMin(p.best_neg + row, bn);
    }
  }
  __device__ void tile_end(const TileInfo&, int) {}
};

}  // namespace demo
