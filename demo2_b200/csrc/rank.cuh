// Rank-count evaluation: CMC/mAP from, for every valid positive p of
// every query q,  r_p = 1 + #{valid g before p},  c_p = 1 + #{positive g before p}  where
// "before" is lexicographic on (distance, gallery index).  Pipeline:
//
//   plan        sort queries / gallery by pid (stable); per sorted query the contiguous range of
//               same-pid gallery rows -> record CSR (positives + junk), band work list
//   records     distances of the same-pid pairs (GEMM extract epilogue, or gathered from a
//               materialised matrix) + global gallery index + junk flag
//   thresholds  per query: positives sorted by (d, gidx), #junk before each
//   count       for each threshold: #{ALL gallery items before it}  (GEMM count epilogue on
//               features, or one streaming pass over a materialised matrix); additive across
//               gallery shards
//   finalize    r = 1 + count - junk_before, c = position + 1 -> AP, CMC, mAP
#pragma once

#include "common.cuh"

namespace demo {

// Device-resident label plan (all arrays caller-owned, carved from one buffer).
struct PlanView {
  int Q = 0, G = 0;
  int* q_perm = nullptr;        // [Q] sorted position -> original query index
  int* q_pid_sorted = nullptr;  // [Q]
  int* g_perm = nullptr;        // [G] sorted position -> original (local) gallery index; the rows whose pid
                                // some query asks for come first (info[3] of them), each class ordered by pid
  int* g_pid_sorted = nullptr;  // [G]
  int* g_lo = nullptr;          // [Q] first sorted gallery row with the query's pid
  int* rec_ofs = nullptr;       // [Q+1] record CSR (sorted query order)
  int4* band_list = nullptr;    // [band_cap] extract work units (m_block, n0, n_rows, 0)
  int* band_count = nullptr;    // [1]
  int* info = nullptr;          // [4] T, max_cnt, band units, #queried gallery rows
  int band_cap = 0;
  void* cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  int* iota = nullptr;          // [max(Q,G)] scratch
  int* cnt = nullptr;           // [Q] scratch
  unsigned long long* gkey = nullptr;         // [G] gallery sort keys (queried-first, then pid)
  unsigned long long* gkey_sorted = nullptr;  // [G]
};

size_t plan_carve(Carver& c, int Q, int G, PlanView* v);
int plan_band_cap(int Q, int G);
int run_plan(const int* q_pid, const int* g_pid, const PlanView& p, cudaStream_t stream);

// g_index (optional, [G]): global gallery index of every LOCAL gallery row (tie-break key);
// nullptr: g_index_base + local row.
int launch_fill_records(const PlanView& p, const int* q_cam, const int* g_cam, int g_index_base,
                        const int* g_index, int* rec_gidx, int* rec_junk, cudaStream_t stream);
int launch_gather_records(const PlanView& p, const float* distmat, long long ld, float* rec_dist,
                          cudaStream_t stream);
int launch_build_thresholds(const int* rec_ofs, const float* rec_dist, const int* rec_gidx,
                            const int* rec_junk, int Q, int* thr_cnt, float* thr_val, int* thr_gidx,
                            int* thr_junk, cudaStream_t stream);
// Row / column mapping of the streaming count kernels.  Default: block i handles sorted query i and
// reads matrix row q_perm[i]; column g has global gallery index g_index_base + g.  For a slab of
// the fused evaluation (matrix rows = sorted queries [row0, row0 + Q) in order, q_perm == nullptr;
// columns = sorted gallery rows): row0 / col_gidx, and blk_flag marks the 256-row query blocks
// the slab path is responsible for (the others are counted by the GEMM epilogue).
struct CountRows {
  int row0 = 0;
  const int* col_gidx = nullptr;
  const unsigned char* blk_flag = nullptr;
};
int launch_count_matrix(const float* distmat, long long ld, int G, int g_index_base, const int* q_perm,
                        const int* thr_ofs, const int* thr_cnt, const float* thr_val, const int* thr_gidx,
                        unsigned* counts, int Q, int max_cnt, cudaStream_t stream, const CountRows* rows = nullptr);
int launch_block_flags(const int* thr_cnt, int Q, int win, int bps, unsigned char* flag, unsigned char* unflag,
                       unsigned char* slab_any, cudaStream_t stream);
int launch_finalize(const int* thr_ofs, const int* thr_cnt, const int* thr_junk, const unsigned* counts,
                    const int* q_perm, int Q, int max_rank, float* cmc_out, double* map_out,
                    int* num_valid_out, double* ap_out, int* first_out, double* scratch,
                    cudaStream_t stream);

}  // namespace demo
