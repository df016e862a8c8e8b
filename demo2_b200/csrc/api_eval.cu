// extern "C" entry points for the rank-count evaluation (include/demo_b200.h).
#include <cstdlib>
#include <cstring>

#include "gemm_epilogues.cuh"
#include "gemm2_sm100.cuh"
#include "rank.cuh"

using namespace demo;

namespace {

// Caller-owned evaluation workspace (same carve for sizing and slicing).
struct EvalWs {
  PrepView a, b;          // prepared queries / gallery, pid-sorted row order
  int* b_gidx;            // [G] global gallery index of every sorted gallery row
  float* rec_dist;        // [T] records (only used by the one-call path)
  int* rec_gidx;
  int* rec_junk;
  int* thr_cnt;           // [Q]
  float* thr_val;         // [T]
  int* thr_gidx;
  int* thr_junk;
  unsigned* counts;       // [T]
  double* ap;             // [Q]
  int* first;             // [Q]
  double* scratch;        // [4096/2]
  float* cmc;             // [4096]
  double* map;            // [1]
  int* nvalid;            // [1]
  TieList ties;           // exact-tie list of the count GEMM (entries + {count, overflow})
  unsigned* pace;         // per-iteration arrival counters of the paced CTA-pair schedule
  size_t pace_cap;
  unsigned char* blk_flag;    // [ceil(Q/256)] query blocks with a row of more than kWin thresholds (slab path)
  unsigned char* blk_unflag;  // [ceil(Q/256)] the complement (schedule skip list of the slab GEMM)
  unsigned char* slab_any;    // [ceil(Q/slab_rows)] slabs with a flagged block
  float* slab;                // [slab_rows][slab_ld] distance slab of the flagged query blocks; nullptr = not carved
  int slab_ld, slab_rows;
};

constexpr int kSlabMaxRows = 1024;    // query rows per slab (4 CTA-pair blocks): enough blocks to keep the
                                      // streaming count kernels at HBM speed
constexpr int kSlabMaxCols = 1 << 20; // gallery columns per slab pass (4 GB of fp32 at most)

// Capacity of the tie list: every valid positive ties with its own threshold (<= T entries per
// window) plus coincidental bit-equal distances; beyond it the exact tie-fix pass takes over.
inline unsigned tie_capacity(int Q, long long T) {
  const long long c = 4 * (T > 0 ? T : 1) + 8ll * (Q > 0 ? Q : 1) + 65536;
  return static_cast<unsigned>(c < (1ll << 27) ? c : (1ll << 27));
}

// max_cnt > kWin adds the distance slab (LAST, so every other offset is independent of it).
size_t carve_eval(Carver& c, int Q, int G, int d, long long T, EvalWs* w, int max_cnt = 0) {
  EvalWs t;
  const size_t t1 = T > 0 ? static_cast<size_t>(T) : 1, q1 = Q > 0 ? Q : 1, g1 = G > 0 ? G : 1;
  prep_carve(c, Q, d, &t.a);
  prep_carve(c, G, d, &t.b);
  t.b_gidx = c.take<int>(g1);
  t.rec_dist = c.take<float>(t1);
  t.rec_gidx = c.take<int>(t1);
  t.rec_junk = c.take<int>(t1);
  t.thr_cnt = c.take<int>(q1);
  t.thr_val = c.take<float>(t1);
  t.thr_gidx = c.take<int>(t1);
  t.thr_junk = c.take<int>(t1);
  t.counts = c.take<unsigned>(t1);
  t.ap = c.take<double>(q1);
  t.first = c.take<int>(q1);
  t.scratch = c.take<double>(2048);
  t.cmc = c.take<float>(4096);
  t.map = c.take<double>(1);
  t.nvalid = c.take<int>(4);
  t.ties.cap = tie_capacity(Q, T);
  // header of the tie list and, right behind it (one memset), the pacing counters of the pair kernel
  t.pace_cap = static_cast<size_t>(ceil_div(static_cast<int>(q1), kBM)) * ceil_div(static_cast<int>(g1), kBN) + 64;
  t.ties.hdr = c.take<unsigned>(16 + t.pace_cap);
  t.pace = t.ties.hdr + 16;
  t.ties.entries = c.take<int4>(t.ties.cap);
  const int nb = ceil_div(static_cast<int>(q1), 256);
  t.blk_flag = c.take<unsigned char>(3 * (nb + 16));
  t.blk_unflag = t.blk_flag + nb + 16;
  t.slab_any = t.blk_unflag + nb + 16;
  t.slab = nullptr;
  t.slab_ld = 0;
  t.slab_rows = 256 * (nb < kSlabMaxRows / 256 ? nb : kSlabMaxRows / 256);
  if (max_cnt > kWin) {
    t.slab_ld = static_cast<int>(round_up(g1 < static_cast<size_t>(kSlabMaxCols) ? g1 : static_cast<size_t>(kSlabMaxCols), size_t(4)));
    t.slab = c.take<float>(static_cast<size_t>(t.slab_rows) * t.slab_ld);
  }
  if (w) *w = t;
  return c.off;
}

__global__ void gidx_kernel(const int* __restrict__ g_perm, int G, int base, const int* __restrict__ g_index,
                            int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < G) out[i] = g_index ? g_index[g_perm[i]] : base + g_perm[i];
}

// Adds the lexicographic tie corrections recorded by the count GEMM:
// counts[k] += 1 for every threshold k of the row with t_k == d and g < p_k.
__global__ void __launch_bounds__(256)
resolve_ties_kernel(TieList ties, const int* __restrict__ b_gidx, const int* __restrict__ thr_ofs,
                    const int* __restrict__ thr_cnt, const float* __restrict__ thr_val,
                    const int* __restrict__ thr_gidx, unsigned* __restrict__ counts, int window) {
  if (ties.hdr[1] != 0u) return;  // overflow: the tie-fix GEMM pass handles every tie
  const unsigned n = min(ties.hdr[0], ties.cap);
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int4 e = ties.entries[i];
    const int g = __ldg(b_gidx + e.y);
    const float d = __int_as_float(e.z);
    const int base = __ldg(thr_ofs + e.x) + window * kWin;
    const int cnt = max(0, min(kWin, __ldg(thr_cnt + e.x) - window * kWin));
    for (int k = 0; k < cnt; ++k)
      if (__ldg(thr_val + base + k) == d && g < __ldg(thr_gidx + base + k)) atomicAdd(counts + base + k, 1u);
  }
}

int get_plan(const void* plan, size_t plan_bytes, int Q, int G, PlanView* p) {
  Carver c(const_cast<void*>(plan), plan_bytes);
  plan_carve(c, Q, G, p);
  if (!plan || !c.ok()) {
    set_error("plan buffer missing or too small (%zu < %zu)", plan_bytes, c.off);
    return DEMO_ERR_WORKSPACE;
  }
  return DEMO_OK;
}

// max_cnt > kWin: the slab is used when the caller's workspace has room for it
// (demo_eval_workspace_bytes_ex); otherwise w->slab stays nullptr (window passes instead).
int get_ws(void* ws, size_t ws_bytes, int Q, int G, int d, long long T, EvalWs* w, int max_cnt = 0) {
  if (max_cnt > kWin) {
    Carver cs(ws, ws_bytes);
    carve_eval(cs, Q, G, d, T, w, max_cnt);
    if (ws && cs.ok()) return DEMO_OK;
  }
  Carver c(ws, ws_bytes);
  carve_eval(c, Q, G, d, T, w);
  if (!ws || !c.ok()) {
    set_error("eval workspace missing or too small (%zu < %zu)", ws_bytes, c.off);
    return DEMO_ERR_WORKSPACE;
  }
  return DEMO_OK;
}

int norm_mode_of(int flags) {
  if (flags & DEMO_FLAG_L2NORM) return PREP_NORM_F_NORMALIZE;
  if (flags & DEMO_FLAG_TRIPLET_NORM) return PREP_NORM_TRIPLET;
  return PREP_NORM_NONE;
}

PrepView sub_rows(const PrepView& v, int row0, int nrows) {
  PrepView s = v;
  s.hi = v.hi + static_cast<long long>(row0) * 2 * v.pitch;
  s.lo = s.hi + 32;
  s.norm = v.norm + row0;
  s.inv_scale = v.inv_scale + row0;
  s.rows = nrows;
  return s;
}

// Slab path for the query blocks flagged in w.blk_flag (a row with more than kWin thresholds):
// per slab of up to 1024 query rows the distances of the flagged 256-row blocks against the
// gallery range are stored once (EpiStore; unflagged blocks are skipped, a slab without flagged
// block is a no-op) and counted by the streaming kernels of rank.cu, whatever the number of
// thresholds -- one GEMM pass plus 4 B per pair of HBM traffic instead of one full GEMM per 63
// thresholds.
int count_slabs(const EvalWs& w, int Q, const PrepView& b, const int* b_gidx, const int* thr_ofs, const int* thr_cnt,
                const float* thr_val, const int* thr_gidx, unsigned* counts, int max_cnt, cudaStream_t stream) {
  for (int m0 = 0; m0 < Q; m0 += w.slab_rows) {
    const int rows = Q - m0 < w.slab_rows ? Q - m0 : w.slab_rows;
    const PrepView a = sub_rows(w.a, m0, rows);
    for (int c0 = 0; c0 < b.rows; c0 += w.slab_ld) {
      const int cols = b.rows - c0 < w.slab_ld ? b.rows - c0 : w.slab_ld;
      const PrepView bc = sub_rows(b, c0, cols);
      EpiStore::Params ep;
      ep.a_norm = a.norm;
      ep.a_inv = a.inv_scale;
      ep.b_norm = bc.norm;
      ep.b_inv = bc.inv_scale;
      ep.out = w.slab;
      ep.ldo = w.slab_ld;
      ep.M = rows;
      ep.mode = DIST_SQ;
      ep.rowmax_key = nullptr;
      ep.run_flag = w.slab_any + m0 / w.slab_rows;   // no flagged block in this slab: the launch returns at once
      GemmOperands ops;
      if (rows > kBM) {
        DEMO_TRY(make_gemm2_operands(a, bc, &ops));
        Schedule s = make_dense_schedule2(rows, cols);
        s.m_skip = w.blk_unflag + (m0 >> 8);          // unflagged blocks are counted by the GEMM epilogue
        DEMO_TRY(launch_sqdist_gemm2<EpiStore>(ops, s, s.num_units, ep, stream));
      } else {
        DEMO_TRY(make_gemm_operands(a, bc, &ops));
        Schedule s = make_dense_schedule(rows, cols);
        s.m_skip = w.blk_unflag + (m0 >> 8);
        DEMO_TRY(launch_sqdist_gemm<EpiStore>(ops, s, s.num_units, ep, stream));
      }
      CountRows cr;
      cr.row0 = m0;
      cr.col_gidx = b_gidx + c0;
      cr.blk_flag = w.blk_flag;
      DEMO_TRY(launch_count_matrix(w.slab, w.slab_ld, cols, 0, nullptr, thr_ofs, thr_cnt, thr_val, thr_gidx, counts,
                                   rows, max_cnt, stream, &cr));
    }
  }
  return DEMO_OK;
}

// counts[] += #{gallery rows [g0, g0 + gn) (sorted order) lexicographically before each threshold}
int count_features(const EvalWs& w, int Q, int g0, int gn, const int* thr_ofs, const int* thr_cnt,
                   const float* thr_val, const int* thr_gidx, unsigned* counts, int max_cnt, int chunk_tiles,
                   cudaStream_t stream, int reserve_sms = 0) {
  if (gn <= 0) return DEMO_OK;
  const PrepView b = sub_rows(w.b, g0, gn);
  const int* b_gidx = w.b_gidx + g0;
  const int G = gn;
  // Rows with more than one window of thresholds: their 256-row query blocks go to the slab path
  // (when the workspace holds a slab), everything else is counted in the GEMM epilogue.
  const bool use_slab = max_cnt > kWin && w.slab != nullptr;
  if (use_slab)
    DEMO_TRY(launch_block_flags(thr_cnt, Q, kWin, w.slab_rows / 256, w.blk_flag, w.blk_unflag, w.slab_any, stream));
  GemmOperands ops;
  DEMO_TRY(make_gemm_operands(w.a, b, &ops));
  EpiCount::Params ep;
  ep.a_norm = w.a.norm;
  ep.a_inv = w.a.inv_scale;
  ep.b_norm = b.norm;
  ep.b_inv = b.inv_scale;
  ep.b_gidx = b_gidx;
  ep.thr_ofs = thr_ofs;
  ep.thr_cnt = thr_cnt;
  ep.thr_val = thr_val;
  ep.thr_gidx = thr_gidx;
  ep.counts = counts;
  ep.M = Q;
  ep.ties = w.ties;
  // The fast pass runs on CTA pairs (cta_group::2, 256 x 256 tiles) unless there is a single
  // 128-row query block or DEMO_COUNT_1CTA is set (A/B timing experiments).
  static const bool force_1cta = getenv("DEMO_COUNT_1CTA") != nullptr;
  const bool pair = !force_1cta && Q > kBM;
  const int n_tiles = ceil_div(G, kBN), m_blocks = ceil_div(Q, pair ? 2 * kBM : kBM);
  // reserve_sms: SMs the CTA-pair grid leaves free (the streamed evaluation pulls the next gallery
  // slab over PCIe with a small kernel meanwhile; next to a full persistent grid it starves)
  if (reserve_sms < 0) reserve_sms = 0;
  const int max_pairs = pair && reserve_sms > 0 ? (num_sms() / 2 - (reserve_sms + 1) / 2 > 8 ? num_sms() / 2 - (reserve_sms + 1) / 2 : 8) : 0;
  const int workers = pair ? (max_pairs > 0 ? max_pairs : num_sms() / 2) : num_sms();
  if (const char* e = getenv("DEMO_CHUNK_TILES")) chunk_tiles = atoi(e);  // experiments
  const bool auto_chunk = chunk_tiles <= 0;
  if (auto_chunk) {
    // enough units to balance the persistent CTAs (>= ~8 units each) but long enough to
    // amortise the per-unit threshold load / histogram flush
    chunk_tiles = 8;   // L2 super-tile: see balanced_group_m (gemm_host.cu)
    while (chunk_tiles > 1 && static_cast<long long>(m_blocks) * ceil_div(n_tiles, chunk_tiles) < 8ll * workers)
      chunk_tiles >>= 1;
  }
  Schedule s = make_chunked_schedule(Q, G, chunk_tiles, w.a.pitch);
  if (use_slab) s.m_skip = w.blk_flag;
  GemmOperands ops2;
  Schedule s2 = s;
  if (pair) {
    DEMO_TRY(make_gemm2_operands(w.a, b, &ops2));
    s2 = make_chunked_schedule2(Q, G, chunk_tiles, w.a.pitch, workers);
    if (use_slab) s2.m_skip = w.blk_flag;
    // Paced schedule (gemm_sm100.cuh, Schedule::pace): a worker starts its i-th unit only when
    // every worker has issued the loads of its unit i - 2.  Free-running workers drift apart by
    // more than an L2 lifetime within milliseconds and then each re-fetches the gallery tiles its
    // group shares from DRAM (20k x 262k: 60 GB -> 17 GB of DRAM reads; at 20k x 1M the kernel
    // runs 137 ms instead of 148 ms because the saved HBM power raises the capped SM clock).
    // DEMO_PACE=<window> / DEMO_PACE=-1 (off) and DEMO_PACE_TILES=<tiles per step> are experiments.
    static const int pace_window = getenv("DEMO_PACE") ? atoi(getenv("DEMO_PACE")) : 1;
    static const int pace_tiles_env = getenv("DEMO_PACE_TILES") ? atoi(getenv("DEMO_PACE_TILES")) : 0;
    int pace_tiles = pace_tiles_env > 0 && pace_tiles_env < chunk_tiles ? pace_tiles_env : chunk_tiles;
    const int pace_steps = ceil_div(chunk_tiles, pace_tiles);
    if (pace_window >= 0 && static_cast<size_t>(s2.num_units) * pace_steps < w.pace_cap) {
      s2.pace = w.pace;
      s2.pace_window = pace_window;
      s2.pace_tiles = pace_tiles;
      s2.pace_steps = pace_steps;
    }
  }
  const int windows = use_slab ? 1 : ceil_div(max_cnt > 0 ? max_cnt : 1, kWin);
  static const bool no_epi = getenv("DEMO_DEBUG_NOEPI") != nullptr;  // timing experiments only
  for (int wdw = 0; wdw < windows; ++wdw) {
    ep.window = no_epi ? -1 : wdw;
    DEMO_CHECK_CUDA(cudaMemsetAsync(w.ties.hdr, 0, (16 + (s2.pace ? static_cast<size_t>(s2.num_units) * s2.pace_steps + 1 : 0)) * sizeof(unsigned), stream));
    if (pair) DEMO_TRY(launch_sqdist_gemm2<EpiCount>(ops2, s2, s2.num_units, ep, stream, max_pairs));
    else DEMO_TRY(launch_sqdist_gemm<EpiCount>(ops, s, s.num_units, ep, stream));
    resolve_ties_kernel<<<2 * num_sms(), 256, 0, stream>>>(w.ties, b_gidx, thr_ofs, thr_cnt, thr_val, thr_gidx,
                                                            counts, wdw);
    DEMO_CHECK_CUDA(cudaGetLastError());
    // exact in-place tie pass; every CTA returns at once unless the list overflowed
    EpiCountTieFix::Params fix;
    static_assert(sizeof(fix) == sizeof(ep), "count epilogue parameter blocks must match");
    memcpy(&fix, &ep, sizeof(fix));
    DEMO_TRY(launch_sqdist_gemm<EpiCountTieFix>(ops, s, s.num_units, fix, stream));
    if (getenv("DEMO_DEBUG_TIES")) {
      unsigned h[16];
      cudaMemcpyAsync(h, w.ties.hdr, 64, cudaMemcpyDeviceToHost, stream);
      cudaStreamSynchronize(stream);
      printf("[ties] window %d: %u entries (cap %u), overflow %u, units %d\n", wdw, h[0], w.ties.cap, h[1], s.num_units);
    }
  }
  if (use_slab)
    DEMO_TRY(count_slabs(w, Q, b, b_gidx, thr_ofs, thr_cnt, thr_val, thr_gidx, counts, max_cnt, stream));
  return DEMO_OK;
}

}  // namespace

extern "C" {

size_t demo_plan_bytes(int Q, int G) {
  Carver c(nullptr, ~size_t(0));
  return round_up(plan_carve(c, Q, G, nullptr), size_t(1024));
}

int demo_eval_plan(const int* q_pid, const int* g_pid, int Q, int G, void* plan, size_t plan_bytes,
                   int64_t* info_host, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(q_pid && g_pid, "eval_plan: null labels");
  PlanView p;
  DEMO_TRY(get_plan(plan, plan_bytes, Q, G, &p));
  DEMO_TRY(run_plan(q_pid, g_pid, p, stream));
  if (info_host) {
    int h[4];
    DEMO_CHECK_CUDA(cudaMemcpyAsync(h, p.info, sizeof(h), cudaMemcpyDeviceToHost, stream));
    DEMO_CHECK_CUDA(cudaStreamSynchronize(stream));
    for (int k = 0; k < 4; ++k) info_host[k] = h[k];
    if (h[2] > p.band_cap) {
      set_error("eval_plan: band list capacity exceeded (%d > %d)", h[2], p.band_cap);
      return DEMO_ERR_CAPACITY;
    }
  }
  return DEMO_OK;
}

int demo_plan_pointers(const void* plan, size_t plan_bytes, int Q, int G, const int** q_perm,
                       const int** g_perm, const int** rec_ofs, const int** g_lo) {
  PlanView p;
  DEMO_TRY(get_plan(plan, plan_bytes, Q, G, &p));
  if (q_perm) *q_perm = p.q_perm;
  if (g_perm) *g_perm = p.g_perm;
  if (rec_ofs) *rec_ofs = p.rec_ofs;
  if (g_lo) *g_lo = p.g_lo;
  return DEMO_OK;
}

// Device pointer of the plan's info[4] = {T, max same-pid count, band units, #queried gallery rows}
// for hosts that enqueue the plan without synchronising and read the numbers later.
int demo_plan_info(const void* plan, size_t plan_bytes, int Q, int G, const int** info) {
  PlanView p;
  DEMO_TRY(get_plan(plan, plan_bytes, Q, G, &p));
  if (info) *info = p.info;
  return DEMO_OK;
}

size_t demo_eval_workspace_bytes_ex(int Q, int G, int d, int64_t T, int max_cnt) {
  Carver c(nullptr, ~size_t(0));
  return round_up(carve_eval(c, Q, G, d, T, nullptr, max_cnt), size_t(1024));
}

size_t demo_eval_workspace_bytes(int Q, int G, int d, int64_t T) { return demo_eval_workspace_bytes_ex(Q, G, d, T, 0); }

size_t demo_eval_matrix_workspace_bytes(int Q, int G, int64_t T) { return demo_eval_workspace_bytes(Q, G, 8, T); }

// Prepares (normalise, scale, fp16 hi/lo split) the pid-sorted rows [row0, row0 + nrows) of the
// queries (which = 0) or the gallery (which = 1) into the evaluation workspace.  x [n][ld] only
// has to be DEVICE-ACCESSIBLE: with pinned (page-locked) host memory the rows are pulled over
// PCIe by the kernel itself, in sorted order, without an intermediate fp32 copy in HBM -- this is
// what lets a host-resident gallery be delivered "queried rows first" (demo_eval_plan) in slabs
// that the count GEMM consumes while the rest is still in flight.
int demo_eval_prepare(const float* x, int n, int d, int64_t ld, int flags, int which, int row0, int nrows,
                      const void* plan, size_t plan_bytes, int Q, int G, int64_t T, void* ws, size_t ws_bytes,
                      float* xn_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(x && d > 0 && ld >= d, "eval_prepare: bad input");
  DEMO_REQUIRE((which == 0 && n == Q) || (which == 1 && n == G), "eval_prepare: row count does not match the plan");
  DEMO_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= n, "eval_prepare: row range [%d, %d) outside [0, %d)",
               row0, row0 + nrows, n);
  PlanView p;
  EvalWs w;
  DEMO_TRY(get_plan(plan, plan_bytes, Q, G, &p));
  DEMO_TRY(get_ws(ws, ws_bytes, Q, G, d, T, &w));
  if (nrows == 0) return DEMO_OK;
  const PrepView v = sub_rows(which ? w.b : w.a, row0, nrows);
  return launch_prep_rows(x, nrows, d, ld, norm_mode_of(flags), (which ? p.g_perm : p.q_perm) + row0, v, xn_out, d,
                          stream, (flags & DEMO_FLAG_HOST_INPUT) != 0);
}

// Records of the same-identity pairs from PREPARED operands (queries + the queried gallery rows,
// which come first in sorted order): global index and junk flag of every pair, then the tcgen05
// extract GEMM over the pid bands.  g_index (optional, [G]): global gallery index of every local
// gallery row, used as the tie-break key instead of g_index_base + local row.
int demo_eval_extract(int Q, int G, int d, const int* q_cam, const int* g_cam, int g_index_base,
                      const int* g_index, const void* plan, size_t plan_bytes, int64_t T, void* ws,
                      size_t ws_bytes, float* rec_dist, int* rec_gidx, int* rec_junk, int q_row0, int q_nrows,
                      void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(q_cam && g_cam, "eval_extract: null pointer");
  DEMO_REQUIRE(q_row0 >= 0 && q_nrows >= 0 && q_row0 + q_nrows <= Q && q_row0 % kBM == 0,
               "eval_extract: query range [%d, %d) outside [0, %d) or not 128-aligned", q_row0, q_row0 + q_nrows, Q);
  DEMO_REQUIRE(Q > 0 && G > 0 && d > 0, "eval_extract: bad shape");
  PlanView p;
  EvalWs w;
  DEMO_TRY(get_plan(plan, plan_bytes, Q, G, &p));
  DEMO_TRY(get_ws(ws, ws_bytes, Q, G, d, T, &w));
  if (!rec_dist) rec_dist = w.rec_dist;
  if (!rec_gidx) rec_gidx = w.rec_gidx;
  if (!rec_junk) rec_junk = w.rec_junk;
  if (q_row0 == 0) {   // label-only parts, whole problem: once (the first group of a staged evaluation)
    gidx_kernel<<<ceil_div(G, 256), 256, 0, stream>>>(p.g_perm, G, g_index_base, g_index, w.b_gidx);
    DEMO_TRY(launch_fill_records(p, q_cam, g_cam, g_index_base, g_index, rec_gidx, rec_junk, stream));
  }
  if (T > 0 && q_nrows > 0) {
    GemmOperands ops;
    DEMO_TRY(make_gemm_operands(w.a, w.b, &ops));
    EpiExtract::Params ep;
    ep.a_norm = w.a.norm;
    ep.a_inv = w.a.inv_scale;
    ep.b_norm = w.b.norm;
    ep.b_inv = w.b.inv_scale;
    ep.a_pid = p.q_pid_sorted;
    ep.b_pid = p.g_pid_sorted;
    ep.g_lo = p.g_lo;
    ep.rec_base = p.rec_ofs;
    ep.rec_dist = rec_dist;
    ep.M = Q;
    Schedule s = make_list_schedule(Q, G, p.band_list, p.band_count);
    if (q_row0 > 0 || q_nrows < Q) {   // one query group: the bands of the other queries are skipped
      s.m_lo = q_row0;
      s.m_hi = q_row0 + q_nrows;
    }
    DEMO_TRY(launch_sqdist_gemm<EpiExtract>(ops, s, p.band_cap, ep, stream));
  }
  return DEMO_OK;
}

int demo_eval_records(const float* q, const float* g, int Q, int G, int d, int64_t ldq, int64_t ldg, int flags,
                      const int* q_cam, const int* g_cam, int g_index_base, const void* plan,
                      size_t plan_bytes, int64_t T, void* ws, size_t ws_bytes, float* rec_dist,
                      int* rec_gidx, int* rec_junk, float* qn_out, float* gn_out, void* stream_) {
  DEMO_REQUIRE(q && g && q_cam && g_cam, "eval_records: null pointer");
  DEMO_REQUIRE(Q > 0 && G > 0 && d > 0 && ldq >= d && ldg >= d, "eval_records: bad shape");
  DEMO_TRY(demo_eval_prepare(q, Q, d, ldq, flags, 0, 0, Q, plan, plan_bytes, Q, G, T, ws, ws_bytes, qn_out, stream_));
  DEMO_TRY(demo_eval_prepare(g, G, d, ldg, flags, 1, 0, G, plan, plan_bytes, Q, G, T, ws, ws_bytes, gn_out, stream_));
  return demo_eval_extract(Q, G, d, q_cam, g_cam, g_index_base, nullptr, plan, plan_bytes, T, ws, ws_bytes, rec_dist,
                           rec_gidx, rec_junk, 0, Q, stream_);
}

int demo_build_thresholds(const int* rec_ofs, const float* rec_dist, const int* rec_gidx, const int* rec_junk,
                          int Q, int* thr_cnt, float* thr_val, int* thr_gidx, int* thr_junk, void* stream_) {
  DEMO_REQUIRE(rec_ofs && thr_cnt, "build_thresholds: null pointer");
  return launch_build_thresholds(rec_ofs, rec_dist, rec_gidx, rec_junk, Q, thr_cnt, thr_val, thr_gidx, thr_junk,
                                 static_cast<cudaStream_t>(stream_));
}

// counts[] += #{gallery rows [g_row0, g_row0 + g_nrows) of the SORTED local gallery before each
// threshold}; the ranges of successive calls must partition [0, G) (any order, any streams that
// are ordered after the prepare of their rows).  Rows with more than 63 thresholds use the slab
// path when the workspace was sized with demo_eval_workspace_bytes_ex(.., max_cnt).
int demo_eval_count_range(int Q, int G, int d, int64_t T_local, void* ws, size_t ws_bytes, const int* thr_ofs,
                          const int* thr_cnt, const float* thr_val, const int* thr_gidx, unsigned* counts,
                          int max_cnt, int chunk_tiles, int g_row0, int g_nrows, int q_row0, int q_nrows,
                          int reserve_sms, void* stream_) {
  EvalWs w;
  DEMO_REQUIRE(g_row0 >= 0 && g_nrows >= 0 && g_row0 + g_nrows <= G, "eval_count: gallery range outside [0, %d)", G);
  DEMO_REQUIRE(q_row0 >= 0 && q_nrows >= 0 && q_row0 + q_nrows <= Q, "eval_count: query range outside [0, %d)", Q);
  DEMO_TRY(get_ws(ws, ws_bytes, Q, G, d, T_local, &w, max_cnt));
  if (q_nrows == 0) return DEMO_OK;
  if (q_row0 > 0 || q_nrows < Q) {
    // a block of pid-sorted queries: the same launch on a shifted view (thr_ofs holds absolute offsets)
    const int align = w.slab != nullptr && max_cnt > kWin ? w.slab_rows : 2 * kBM;
    DEMO_REQUIRE(q_row0 % align == 0, "eval_count: query range must start at a multiple of %d rows", align);
    w.a = sub_rows(w.a, q_row0, q_nrows);
    w.blk_flag += q_row0 >> 8;
    w.blk_unflag += q_row0 >> 8;
    w.slab_any += q_row0 / w.slab_rows;
    thr_ofs += q_row0;
    thr_cnt += q_row0;
  }
  return count_features(w, q_nrows, g_row0, g_nrows, thr_ofs, thr_cnt, thr_val, thr_gidx, counts, max_cnt, chunk_tiles,
                        static_cast<cudaStream_t>(stream_), reserve_sms);
}

int demo_eval_count(int Q, int G, int d, int64_t T_local, void* ws, size_t ws_bytes, const int* thr_ofs,
                    const int* thr_cnt, const float* thr_val, const int* thr_gidx, unsigned* counts,
                    int max_cnt, int chunk_tiles, void* stream_) {
  return demo_eval_count_range(Q, G, d, T_local, ws, ws_bytes, thr_ofs, thr_cnt, thr_val, thr_gidx, counts, max_cnt,
                               chunk_tiles, 0, G, 0, Q, 0, stream_);
}

int demo_cmc_map_finalize(const int* thr_ofs, const int* thr_cnt, const int* thr_junk, const unsigned* counts,
                          const int* q_perm, int Q, int max_rank, float* cmc_out, double* map_out,
                          int* num_valid_out, double* ap_out, int* first_out, void* scratch, void* stream_) {
  DEMO_REQUIRE(thr_ofs && thr_cnt && thr_junk && counts && q_perm && cmc_out && map_out && num_valid_out &&
                   ap_out && first_out && scratch,
               "finalize: null pointer");
  return launch_finalize(thr_ofs, thr_cnt, thr_junk, counts, q_perm, Q, max_rank, cmc_out, map_out, num_valid_out,
                         ap_out, first_out, static_cast<double*>(scratch), static_cast<cudaStream_t>(stream_));
}

// One call, one GPU: features -> (cmc, mAP, num_valid, per-query AP / first rank).
// Replaces euclidean_distance + eval_func of R1_mAP_eval.compute (utils/metrics.py:341-369)
// without materialising the Q x G matrix.  Outputs are DEVICE pointers (cmc_out[max_rank]).
int demo_eval_features(const float* q, const float* g, int Q, int G, int d, int64_t ldq, int64_t ldg, int flags,
                       const int* q_cam, const int* g_cam, const void* plan, size_t plan_bytes, int64_t T,
                       int max_cnt, int max_rank, void* ws, size_t ws_bytes, float* cmc_out, double* map_out,
                       int* num_valid_out, double* ap_out, int* first_out, float* qn_out, float* gn_out,
                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PlanView p;
  EvalWs w;
  DEMO_TRY(get_plan(plan, plan_bytes, Q, G, &p));
  DEMO_TRY(get_ws(ws, ws_bytes, Q, G, d, T, &w, max_cnt));
  DEMO_TRY(demo_eval_records(q, g, Q, G, d, ldq, ldg, flags, q_cam, g_cam, 0, plan, plan_bytes, T, ws, ws_bytes,
                             nullptr, nullptr, nullptr, qn_out, gn_out, stream_));
  DEMO_TRY(launch_build_thresholds(p.rec_ofs, w.rec_dist, w.rec_gidx, w.rec_junk, Q, w.thr_cnt, w.thr_val,
                                   w.thr_gidx, w.thr_junk, stream));
  DEMO_CHECK_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(unsigned) * (T > 0 ? T : 1), stream));
  if (T > 0)
    DEMO_TRY(count_features(w, Q, 0, G, p.rec_ofs, w.thr_cnt, w.thr_val, w.thr_gidx, w.counts, max_cnt, 0, stream));
  return launch_finalize(p.rec_ofs, w.thr_cnt, w.thr_junk, w.counts, p.q_perm, Q, max_rank,
                         cmc_out ? cmc_out : w.cmc, map_out ? map_out : w.map, num_valid_out ? num_valid_out : w.nvalid,
                         ap_out ? ap_out : w.ap, first_out ? first_out : w.first, w.scratch, stream);
}

// Materialised matrix: eval_func(distmat, ...) (utils/metrics.py:110-169).  distmat is a DEVICE
// matrix [Q][G]; the plan must have been built from the same labels.
int demo_eval_matrix(const float* distmat, int Q, int G, int64_t ld, const int* q_cam, const int* g_cam,
                     const void* plan, size_t plan_bytes, int64_t T, int max_cnt, int max_rank, void* ws,
                     size_t ws_bytes, float* cmc_out, double* map_out, int* num_valid_out, double* ap_out,
                     int* first_out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(distmat && ld >= G, "eval_matrix: bad matrix");
  PlanView p;
  EvalWs w;
  DEMO_TRY(get_plan(plan, plan_bytes, Q, G, &p));
  DEMO_TRY(get_ws(ws, ws_bytes, Q, G, 8, T, &w));
  DEMO_TRY(launch_fill_records(p, q_cam, g_cam, 0, nullptr, w.rec_gidx, w.rec_junk, stream));
  DEMO_TRY(launch_gather_records(p, distmat, ld, w.rec_dist, stream));
  DEMO_TRY(launch_build_thresholds(p.rec_ofs, w.rec_dist, w.rec_gidx, w.rec_junk, Q, w.thr_cnt, w.thr_val,
                                   w.thr_gidx, w.thr_junk, stream));
  DEMO_CHECK_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(unsigned) * (T > 0 ? T : 1), stream));
  if (T > 0)
    DEMO_TRY(launch_count_matrix(distmat, ld, G, 0, p.q_perm, p.rec_ofs, w.thr_cnt, w.thr_val, w.thr_gidx, w.counts,
                                 Q, max_cnt, stream));
  return launch_finalize(p.rec_ofs, w.thr_cnt, w.thr_junk, w.counts, p.q_perm, Q, max_rank,
                         cmc_out ? cmc_out : w.cmc, map_out ? map_out : w.map, num_valid_out ? num_valid_out : w.nvalid,
                         ap_out ? ap_out : w.ap, first_out ? first_out : w.first, w.scratch, stream);
}

// Device pointers of the result slots inside an eval workspace (for callers that passed NULL
// outputs to demo_eval_features / demo_eval_matrix) and of the intermediate arrays.
int demo_eval_ws_pointers(void* ws, size_t ws_bytes, int Q, int G, int d, int64_t T, float** cmc, double** map,
                          int** nvalid, double** ap, int** first, unsigned** counts, int** thr_cnt,
                          float** thr_val, int** thr_gidx, int** thr_junk, float** rec_dist, int** rec_gidx,
                          int** rec_junk) {
  EvalWs w;
  DEMO_TRY(get_ws(ws, ws_bytes, Q, G, d, T, &w));
  if (cmc) *cmc = w.cmc;
  if (map) *map = w.map;
  if (nvalid) *nvalid = w.nvalid;
  if (ap) *ap = w.ap;
  if (first) *first = w.first;
  if (counts) *counts = w.counts;
  if (thr_cnt) *thr_cnt = w.thr_cnt;
  if (thr_val) *thr_val = w.thr_val;
  if (thr_gidx) *thr_gidx = w.thr_gidx;
  if (thr_junk) *thr_junk = w.thr_junk;
  if (rec_dist) *rec_dist = w.rec_dist;
  if (rec_gidx) *rec_gidx = w.rec_gidx;
  if (rec_junk) *rec_junk = w.rec_junk;
  return DEMO_OK;
}

}  // extern "C"
