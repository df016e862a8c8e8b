// Epilogue policies for sqdist_gemm_kernel.  Each epilogue thread owns ONE accumulator row
// (one query) and one half (128 columns) of the 256-column tile, which it walks in steps of 16
// or 32 TMEM columns; per-column metadata (|g|^2, 2^-e, labels) is staged once per tile in
// shared memory and read as warp-wide broadcasts.
//
//   dist = fma(acc * 2^-ea[m], -2 * 2^-eb[n], |a_m|^2 + |b_n|^2)
//        = fma(-2, dot, |a|^2 + |b|^2)  exactly (power-of-two scalings)   utils/metrics.py:398-400
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

// Per-tile column metadata: {-2 * 2^-eb[n], |b_n|^2} (+ optional int label), double-buffered.
// Invalid (out-of-range) columns carry |b|^2 = +inf so their distance is +inf.
struct EpiColumns {
  float2* s_col;   // [2][kBN]
  int* s_lab;      // [2][kBN]
  static constexpr int kBytes = 2 * kBN * 8 + 2 * kBN * 4;

  __device__ EpiColumns(uint8_t* smem) {
    s_col = reinterpret_cast<float2*>(smem);
    s_lab = reinterpret_cast<int*>(smem + 2 * kBN * 8);
  }
  // one column per epilogue thread (kEpiThreads == kBN): global -> registers ...
  float2 r_col;
  int r_lab;
  __device__ __forceinline__ void load(const TileInfo& t, int epi_tid, const float* b_norm, const float* b_inv,
                                       const int* b_lab, float invalid_norm) {
    const bool ok = epi_tid < t.n_valid;
    r_col = make_float2(ok ? -2.f * __ldg(b_inv + t.n0 + epi_tid) : 0.f,
                        ok ? __ldg(b_norm + t.n0 + epi_tid) : invalid_norm);
    r_lab = (b_lab && ok) ? __ldg(b_lab + t.n0 + epi_tid) : -0x7fffffff;
  }
  // ... -> shared memory buffer `as` (made visible by the barrier at the next tile_begin)
  __device__ __forceinline__ void store(int as, int epi_tid) {
    s_col[as * kBN + epi_tid] = r_col;
    s_lab[as * kBN + epi_tid] = r_lab;
  }
};

// ---------------------------------------------------------------------------------------
// STORE: materialise the matrix (euclidean_distance, cosine_*, re-ranking all-pairs)
// ---------------------------------------------------------------------------------------
struct EpiStore {
  // One 32 x 32 fp32 staging block (row stride 33) per epilogue warp: a thread owns a ROW of the
  // tile, so storing straight from registers makes every store instruction touch 32 different
  // 128-byte lines (ncu on the 10 290^2 all-pairs GEMM: lg_throttle stalls, tensor pipe 48 % active
  // -- the accumulator buffers were not drained fast enough).  Through the staging block eight
  // lanes write one 128-byte line: 4 lines per instruction.
  static constexpr int kStageFloats = 32 * 33;
  static constexpr int kStages = 3;   // 1-CTA kernel: 3 x 48 KB ring + the staging blocks
  static constexpr int kSmemBytes = EpiColumns::kBytes + (kEpiThreads / 32) * kStageFloats * 4;
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
    const unsigned char* run_flag = nullptr;  // optional (device): the launch is a no-op when *run_flag == 0
    // Symmetric all-pairs matrices (A and B are row ranges of the SAME set, global indices
    // a_global0 + r / b_global0 + c).  The value of a pair is DEFINED as the GEMM result with the
    // lower index on the A side, so that it does not depend on how the matrix is tiled or sharded:
    //   store_normal  out[r][c] = v          (only where gr <= gc when sym_mask is set)
    //   sym_mirror    out_t[c][r] = v        where gr < gc (the transposed position; lanes are
    //                 consecutive rows r, so every store instruction writes one 128-byte line)
    // rowmax_key_t receives the maxima of the mirrored stores (per column c).
    int store_normal = 1, sym_mask = 0, sym_mirror = 0;
    int a_global0 = 0, b_global0 = 0;
    float* out_t = nullptr;
    long long ldo_t = 0;
    unsigned* rowmax_key_t = nullptr;
  };
  const Params& p;
  EpiColumns cols;
  int epi_tid, row_in_tile, col0;
  float* stage;   // this warp's staging block

  __device__ EpiStore(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_, int col0_)
      : p(p_), cols(smem), epi_tid(epi_tid_), row_in_tile(row_in_tile_), col0(col0_),
        stage(reinterpret_cast<float*>(smem + EpiColumns::kBytes) + (epi_tid_ >> 5) * kStageFloats) {}

  __device__ void stage_load(const TileInfo& t) { cols.load(t, epi_tid, p.b_norm, p.b_inv, nullptr, INFINITY); }
  __device__ void stage_store(int as) { cols.store(as, epi_tid); }
  __device__ void tile_begin(const TileInfo&, int) { epi_bar_sync(); }
  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const int row = t.m0 + row_in_tile;
    const bool row_ok = row < p.M;
    const float na = row_ok ? __ldg(p.a_norm + row) : 0.f;
    const float ia = row_ok ? __ldg(p.a_inv + row) : 0.f;
    float* orow = p.out + static_cast<long long>(row) * p.ldo + t.n0 + col0;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(orow) & 15u) == 0);
    const float2* col = cols.s_col + as * kBN + col0;
    const int n_here = t.n_valid - col0;  // valid columns in this thread's half
    const int gr = p.a_global0 + row;                 // global index of this thread's row
    const int gc0 = p.b_global0 + t.n0 + col0;        // global index of this thread's first column
    float vmax = -INFINITY;
    // coalesced path (warp-uniform conditions): the warp's 32 rows, 16-byte aligned row segments
    const int lane = epi_tid & 31;
    const int wrow0 = t.m0 + row_in_tile - lane;      // first row of this warp
    float* wbase = p.out + static_cast<long long>(wrow0) * p.ldo + t.n0 + col0;
    const bool warp_vec = (p.ldo & 3) == 0 && ((reinterpret_cast<uintptr_t>(wbase) & 15u) == 0);
#pragma unroll 1
    for (int c = 0; c < kEpiCols / 32; ++c) {
      if (c * 32 >= n_here) break;  // warp-uniform
      uint32_t r[32];
      __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the per-row branches
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float2 cm = col[c * 32 + j];
        if (p.mode == DIST_SQ) {
          v[j] = fmaf(__uint_as_float(r[j]) * ia, cm.x, na + cm.y);
        } else {
          const float dot = __uint_as_float(r[j]) * ia * cm.x * -0.5f;
          v[j] = finish_distance(p.mode, dot, na, cm.y);
        }
      }
      const int nv = min(32, n_here - c * 32);
      // whole 32 x 32 block valid (and on or above the diagonal): staged, line-coalesced stores
      const bool block_ok = p.store_normal && nv == 32 &&
                            (!p.sym_mask || p.a_global0 + wrow0 + 31 <= gc0 + c * 32);
      if (block_ok) {
#pragma unroll
        for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = v[j];
        __syncwarp();
        if (warp_vec) {                        // 8 lanes per 128-byte row segment, 4 rows per instruction
          const int sub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = 4 * i + sub;
            const float* sp = stage + rr * 33 + c4;
            const float4 o = make_float4(sp[0], sp[1], sp[2], sp[3]);
            if (wrow0 + rr < p.M)
              *reinterpret_cast<float4*>(wbase + static_cast<long long>(rr) * p.ldo + c * 32 + c4) = o;
          }
        } else {                               // rows not 16-byte aligned: one contiguous 128-byte row segment per instruction
          const int nrow = min(32, p.M - wrow0);
#pragma unroll 8
          for (int rr = 0; rr < nrow; ++rr)
            wbase[static_cast<long long>(rr) * p.ldo + c * 32 + lane] = stage[rr * 33 + lane];
        }
        __syncwarp();
        if (p.rowmax_key && row_ok) {
#pragma unroll
          for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, v[j]);
        }
      } else if (p.store_normal && row_ok) {
        const int gc = gc0 + c * 32;                  // column j of this chunk has global index gc + j
        const bool all_upper = !p.sym_mask || gr <= gc;
        if (nv == 32 && vec_ok && all_upper) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(orow + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          if (p.rowmax_key) {
#pragma unroll
            for (int j = 0; j < 32; ++j) vmax = fmaxf(vmax, v[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nv && (!p.sym_mask || gr <= gc + j)) {
              orow[c * 32 + j] = v[j];
              vmax = fmaxf(vmax, v[j]);
            }
        }
      }
      if (p.sym_mirror) {
        // transposed position: out_t[gc + j - b_global0 ...]; addressed by operand-local indices
        const int gc = gc0 + c * 32;
        float* tcol = p.out_t + static_cast<long long>(t.n0 + col0 + c * 32) * p.ldo_t + row;
        if (__any_sync(0xffffffffu, row_ok && gr < gc + nv - 1 + 1)) {   // some lane has an element above the diagonal
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const bool ok = row_ok && j < nv && gr < gc + j;
            if (ok) tcol[static_cast<long long>(j) * p.ldo_t] = v[j];
            if (p.rowmax_key_t) {
              const unsigned key = ok ? float_key(v[j]) : 0u;
              const unsigned m = __reduce_max_sync(0xffffffffu, key);
              if (m != 0u && (threadIdx.x & 31) == 0) atomicMax(p.rowmax_key_t + t.n0 + col0 + c * 32 + j, m);
            }
          }
        }
      }
    }
    if (p.rowmax_key && row_ok && vmax > -INFINITY) atomicMax(p.rowmax_key + row, float_key(vmax));
  }
  __device__ void tile_end(const TileInfo&, int) {}
  __device__ void finish() {}
  __device__ static bool skip(const Params& p) { return p.run_flag != nullptr && *p.run_flag == 0; }
};

// ---------------------------------------------------------------------------------------
// COUNT: rank counts without materialising Q x G.
// A row q keeps up to kWin thresholds (the distances of q's valid positives, ascending by
// (distance, gallery index)) in a shared-memory column; for every gallery column g the thread
// finds   b = #{ j : t_j <= d(q,g) }   by bisection (top three tree levels in registers) and
// bumps its private 16-bit histogram bucket b.  At the end of the unit
//   L_j = #{g : d(q,g) < t_j} = sum_{b<=j} hist[b]   is added to counts[].
// The lexicographic tie rule ((d, g) before (t_j, p_j) also when d == t_j and g < p_j) is applied
// OUTSIDE the hot loop: an element whose distance equals a threshold bit for bit (about one per
// tile: every positive ties with its own threshold, plus coincidences) is appended to a small
// shared-memory list that is flushed to a global tie list once per unit, and
// resolve_ties_kernel adds the missing +1s afterwards.  Nothing in the per-element path touches
// global memory: a global load issued from here queues behind ~7 TB/s of TMA operand traffic and
// takes microseconds, which used to stall the whole accumulator pipeline (measured: 206 ms ->
// 150 ms on 20k x 1M).  If the global list overflows (pathological inputs: masses of identical
// rows) a flag is raised, the resolver stands down and the same GEMM is re-run with
// EpiCountT<true>, which fixes every tie in place (slow, exact).  No labels are read: junk /
// positives are subtracted later from the per-query record list.
// ---------------------------------------------------------------------------------------
constexpr int kWin = 63;

template <int kOff>
__device__ __forceinline__ float lds_f32_off(uint32_t addr) {
  float v;
  asm("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(kOff));  // not volatile: free to schedule
  return v;
}
__device__ __forceinline__ void hist_inc_u16(uint32_t addr) {  // ordered read-modify-write
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  v += 1u;
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v));
}

// Global tie list header: [0] = entries appended, [1] = overflow flag.
// Entry = (sorted query row, B row, distance bits, 0).
struct TieList {
  int4* entries;
  unsigned* hdr;
  unsigned cap;
};

template <bool kTieFix>
struct EpiCountT {
  static constexpr int kStages = 3;
  static constexpr int kRowBytes = 512;  // one bucket row: 128 x f32 thresholds == 256 x u16 counters
  static constexpr int kTieCap = 112;    // per-unit shared tie list, two parities
  // columns x2 buffers | thresholds [kWin][128] f32 | histogram [kWin+1][256] u16 | tie lists | counters
  static constexpr int kSmemBytes =
      2 * kBN * 8 + kWin * kRowBytes + (kWin + 1) * kRowBytes + 2 * kTieCap * 16 + 16;
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
    TieList ties;
  };
  // The tie-fix pass only runs after an overflow of the tie list.
  __device__ static bool skip(const Params& p) { return kTieFix && p.ties.hdr[1] == 0u; }

  const Params& p;
  float2* s_col;              // [2][kBN]
  float* s_thr;               // [kWin][128]     column = row_in_tile (shared by both column halves)
  unsigned short* s_hist;     // [kWin+1][256]   column = epi_tid (private)
  int4* s_tie;                // [2][kTieCap]
  unsigned* s_tie_cnt;        // [2]
  int epi_tid, row_in_tile, col0;
  int hcol;                   // histogram column 2*row + half: the two column halves of a row share a 32-bit
                              // word (low / high u16), so counter b sits at a CONSTANT byte offset
                              // (kWin*512 + 2*half) from threshold b and a warp hits 32 different banks
  int nthr = 0, tbase = 0, par = 0;
  // top four levels of the search tree live in registers
  float t31 = 0, t15 = 0, t47 = 0, t7 = 0, t23 = 0, t39 = 0, t55 = 0;
  float t3 = 0, t11 = 0, t19 = 0, t27 = 0, t35 = 0, t43 = 0, t51 = 0, t59 = 0;  // level 4

  __device__ EpiCountT(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_, int col0_)
      : p(p_), epi_tid(epi_tid_), row_in_tile(row_in_tile_), col0(col0_) {
    s_col = reinterpret_cast<float2*>(smem);
    s_thr = reinterpret_cast<float*>(smem + 2 * kBN * 8);
    s_hist = reinterpret_cast<unsigned short*>(smem + 2 * kBN * 8 + kWin * kRowBytes);
    s_tie = reinterpret_cast<int4*>(smem + 2 * kBN * 8 + kWin * kRowBytes + (kWin + 1) * kRowBytes);
    s_tie_cnt = reinterpret_cast<unsigned*>(s_tie + 2 * kTieCap);
    hcol = (row_in_tile << 1) | (col0 ? 1 : 0);
    if (epi_tid < 2) s_tie_cnt[epi_tid] = 0;  // visible after the barrier of the first tile_begin
  }

  float2 r_col;
  float na = 0.f, ia = 0.f;   // per-row constants of the current unit
  __device__ void stage_load(const TileInfo& t) {
    const bool ok = epi_tid < t.n_valid;
    // invalid columns get a huge FINITE distance: they land in the (unused) bucket after the
    // last real threshold and can never tie with the +inf padding
    r_col = make_float2(ok ? -2.f * __ldg(p.b_inv + t.n0 + epi_tid) : 0.f,
                        ok ? __ldg(p.b_norm + t.n0 + epi_tid) : 3.0e38f);
  }
  __device__ void stage_store(int as) { s_col[as * kBN + epi_tid] = r_col; }

  __device__ __forceinline__ void push_global(const int4& e) const {
    const unsigned g = atomicAdd(p.ties.hdr, 1u);
    if (g < p.ties.cap) p.ties.entries[g] = e;
    else p.ties.hdr[1] = 1u;
  }
  // all epilogue threads; the list of parity `which` is complete (a barrier has been passed)
  __device__ void flush_ties(int which) const {
    const int n = min(static_cast<int>(s_tie_cnt[which]), kTieCap);
    for (int i = epi_tid; i < n; i += kEpiThreads) push_global(s_tie[which * kTieCap + i]);
  }

  __device__ void tile_begin(const TileInfo& t, int as) {
    // everybody has left the previous tile body (s_thr is shared by the two warps of a lane
    // quadrant) and the staged columns of this tile are visible
    epi_bar_sync();
    if (t.first_in_unit) {
      if (!kTieFix) {
        flush_ties(par);  // ties of the unit that just ended
        par ^= 1;
      }
      const int row = t.m0 + row_in_tile;
      nthr = 0;
      na = ia = 0.f;
      if (row < p.M) {
        tbase = __ldg(p.thr_ofs + row) + p.window * kWin;
        nthr = max(0, min(kWin, __ldg(p.thr_cnt + row) - p.window * kWin));
        na = __ldg(p.a_norm + row);
        ia = __ldg(p.a_inv + row);
      }
      // the two threads of a row (column halves) fill alternate threshold slots
      for (int k = (col0 ? 1 : 0); k < kWin; k += 2)
        s_thr[k * 128 + row_in_tile] = k < nthr ? __ldg(p.thr_val + tbase + k) : INFINITY;
      if (!kTieFix)
        for (int k = 0; k <= kWin; ++k) s_hist[k * kEpiThreads + hcol] = 0;
      epi_bar_sync();
      if (!kTieFix && epi_tid == 0) s_tie_cnt[par ^ 1] = 0;  // flushed above; unused until the unit after this one
      const float* thr = s_thr + row_in_tile;
      t31 = thr[31 * 128];
      t15 = thr[15 * 128];
      t47 = thr[47 * 128];
      t7 = thr[7 * 128];
      t23 = thr[23 * 128];
      t39 = thr[39 * 128];
      t55 = thr[55 * 128];
      t3 = thr[3 * 128];
      t11 = thr[11 * 128];
      t19 = thr[19 * 128];
      t27 = thr[27 * 128];
      t35 = thr[35 * 128];
      t43 = thr[43 * 128];
      t51 = thr[51 * 128];
      t59 = thr[59 * 128];
    }
  }

  // ONE copy of the loop body for both column halves (the histogram offset is a register): the
  // SM's L1.5 instruction cache holds 32 KB and the unrolled body is ~14 KB; two copies evicted
  // everything else and every rarely executed path paid an L2 instruction fetch (microseconds
  // behind the TMA traffic).
  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const uint32_t hoff = kWin * kRowBytes + (col0 ? 2 : 0);
    const int row = t.m0 + row_in_tile;
    const bool active = nthr > 0;
    if (p.window < 0) return;  // debug: mainloop-only timing (DEMO_DEBUG_NOEPI)
    if (!__any_sync(0xffffffffu, active)) return;  // whole warp beyond M / without positives
    const uint32_t thr0 = smem_u32(s_thr + row_in_tile);             // bucket b at thr0 + b*512
    const float2* col = s_col + as * kBN + col0;
    const int n_here = t.n_valid - col0;
    constexpr int kC = 16;
#pragma unroll 1
    for (int c = 0; c < kEpiCols / kC; ++c) {
      if (c * kC >= n_here) break;  // warp-uniform
      uint32_t r[kC];
      __syncwarp();
      tmem_ld_32x16(taddr + c * kC, r);
      tmem_ld_wait();
      uint32_t slot[kC];   // shared address of the threshold row == bucket, per element
      uint32_t ties = 0;
#pragma unroll
      for (int j = 0; j < kC; ++j) {
        const float2 cm = col[c * kC + j];
        const float d = fmaf(__uint_as_float(r[j]) * ia, cm.x, na + cm.y);
        // level 1..4 from registers (the shared-memory pipe is the scarce resource next to the
        // tensor cores' operand fetch); `last` tracks the largest threshold <= d (tie detection)
        const bool p1 = t31 <= d;
        uint32_t a = p1 ? thr0 + 32 * kRowBytes : thr0;
        float last = p1 ? t31 : -INFINITY;
        const float u2 = p1 ? t47 : t15;
        const bool p2 = u2 <= d;
        a += p2 ? 16 * kRowBytes : 0;
        last = p2 ? u2 : last;
        const float hi3 = p2 ? t55 : t39, lo3 = p2 ? t23 : t7;
        const float u3 = p1 ? hi3 : lo3;
        const bool p3 = u3 <= d;
        a += p3 ? 8 * kRowBytes : 0;
        last = p3 ? u3 : last;
        const float q00 = p3 ? t11 : t3, q01 = p3 ? t27 : t19, q10 = p3 ? t43 : t35, q11 = p3 ? t59 : t51;
        const float u4 = p1 ? (p2 ? q11 : q10) : (p2 ? q01 : q00);
        const bool p4 = u4 <= d;
        a += p4 ? 4 * kRowBytes : 0;
        last = p4 ? u4 : last;
        // level 5..6 from shared memory (immediate offsets, one predicated add per level)
        float u = lds_f32_off<1 * kRowBytes>(a);
        bool q = u <= d;
        a += q ? 2 * kRowBytes : 0;
        last = q ? u : last;
        u = lds_f32_off<0>(a);
        q = u <= d;
        a += q ? 1 * kRowBytes : 0;
        last = q ? u : last;
        ties |= (last == d) ? (1u << j) : 0u;
        slot[j] = a;
      }
      if (ties && active) {
        // the distance equals a threshold bit for bit (rare)
        do {
          const int jj = __ffs(ties) - 1;
          ties &= ties - 1u;
          uint32_t rj = 0;
#pragma unroll
          for (int j = 0; j < kC; ++j) rj = (j == jj) ? r[j] : rj;
          const float2 cm = col[c * kC + jj];
          const float d = fmaf(__uint_as_float(rj) * ia, cm.x, na + cm.y);
          const int bcol = t.n0 + col0 + c * kC + jj;
          if (!kTieFix) {
            const int4 e = make_int4(row, bcol, __float_as_int(d), 0);
            const unsigned i = atomicAdd(s_tie_cnt + par, 1u);
            if (i < static_cast<unsigned>(kTieCap)) s_tie[par * kTieCap + i] = e;
            else push_global(e);
          } else {
            const int g = __ldg(p.b_gidx + bcol);
#pragma unroll 1
            for (int k = 0; k < nthr; ++k)
              if (s_thr[k * 128 + row_in_tile] == d && g < __ldg(p.thr_gidx + tbase + k))
                atomicAdd(p.counts + tbase + k, 1u);
          }
        } while (ties);
      }
      if (!kTieFix) {
#pragma unroll
        for (int j = 0; j < kC; ++j) hist_inc_u16(slot[j] + hoff);
      }
    }
  }

  __device__ void tile_end(const TileInfo& t, int) {
    if (!kTieFix && t.last_in_unit && nthr > 0) {
      unsigned run = 0;
      for (int k = 0; k < nthr; ++k) {
        run += s_hist[k * kEpiThreads + hcol];
        if (run) atomicAdd(p.counts + tbase + k, run);
      }
    }
  }

  // after the last tile of the CTA (all epilogue threads)
  __device__ void finish() {
    if (kTieFix) return;
    epi_bar_sync();
    flush_ties(par);
  }
};
using EpiCount = EpiCountT<false>;
using EpiCountTieFix = EpiCountT<true>;

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
  int epi_tid, row_in_tile, col0;

  __device__ EpiExtract(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_, int col0_)
      : p(p_), cols(smem), epi_tid(epi_tid_), row_in_tile(row_in_tile_), col0(col0_) {}

  __device__ void stage_load(const TileInfo& t) { cols.load(t, epi_tid, p.b_norm, p.b_inv, p.b_pid, INFINITY); }
  __device__ void stage_store(int as) { cols.store(as, epi_tid); }
  __device__ void tile_begin(const TileInfo&, int) { epi_bar_sync(); }
  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const int row = t.m0 + row_in_tile;
    const bool row_ok = row < p.M;
    const float na = row_ok ? __ldg(p.a_norm + row) : 0.f;
    const float ia = row_ok ? __ldg(p.a_inv + row) : 0.f;
    const int pid = row_ok ? __ldg(p.a_pid + row) : 0;
    const int lo = row_ok ? __ldg(p.g_lo + row) : 0;
    const int base = row_ok ? __ldg(p.rec_base + row) : 0;
    const float2* col = cols.s_col + as * kBN + col0;
    const int* lab = cols.s_lab + as * kBN + col0;
    const int n_here = t.n_valid - col0;
#pragma unroll 1
    for (int c = 0; c < kEpiCols / 32; ++c) {
      if (c * 32 >= n_here) break;
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cc = c * 32 + j;
        if (cc < n_here && lab[cc] == pid) {
          const float2 cm = col[cc];
          p.rec_dist[base + (t.n0 + col0 + cc - lo)] = fmaf(__uint_as_float(r[j]) * ia, cm.x, na + cm.y);
        }
      }
    }
  }
  __device__ void tile_end(const TileInfo&, int) {}
  __device__ void finish() {}
  __device__ static bool skip(const Params&) { return false; }
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
  int epi_tid, row_in_tile, col0;

  __device__ EpiMine(const Params& p_, uint8_t* smem, int epi_tid_, int row_in_tile_, int col0_)
      : p(p_), cols(smem), epi_tid(epi_tid_), row_in_tile(row_in_tile_), col0(col0_) {}

  __device__ void stage_load(const TileInfo& t) { cols.load(t, epi_tid, p.b_norm, p.b_inv, p.b_lab, INFINITY); }
  __device__ void stage_store(int as) { cols.store(as, epi_tid); }
  __device__ void tile_begin(const TileInfo&, int) { epi_bar_sync(); }
  __device__ void tile_body(const TileInfo& t, int as, uint32_t taddr) {
    const int row = t.m0 + row_in_tile;
    const bool row_ok = row < p.M;
    const float na = row_ok ? __ldg(p.a_norm + row) : 0.f;
    const float ia = row_ok ? __ldg(p.a_inv + row) : 0.f;
    const int mylab = row_ok ? __ldg(p.a_lab + row) : 0;
    const float2* col = cols.s_col + as * kBN + col0;
    const int* lab = cols.s_lab + as * kBN + col0;
    const int n_here = t.n_valid - col0;
    unsigned long long bp = 0ull, bn = ~0ull;
#pragma unroll 1
    for (int c = 0; c < kEpiCols / 32; ++c) {
      if (c * 32 >= n_here) break;
      uint32_t r[32];
      __syncwarp();
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
      if (!row_ok) continue;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int cc = c * 32 + j;
        if (cc >= n_here) continue;
        const float2 cm = col[cc];
        const float d = sqrtf(fmaxf(fmaf(__uint_as_float(r[j]) * ia, cm.x, na + cm.y), 1e-12f));
        const unsigned key = float_key(d);
        const unsigned idx = static_cast<unsigned>(t.n0 + col0 + cc);
        if (lab[cc] == mylab) {
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
  __device__ void finish() {}
  __device__ static bool skip(const Params&) { return false; }
};

}  // namespace demo
