// k-reciprocal re-ranking stages (see rerank.cu).
#pragma once

#include "common.cuh"

namespace demo {

struct RerankWs {
  int K = 0, cap = 0, capq = 0;
  int* rank = nullptr;       // [N][K]
  int* v_idx = nullptr;      // [N][cap]   V rows before expansion
  __half* v_val = nullptr;
  int* v_cnt = nullptr;      // [N]
  int* rh_idx = nullptr;     // [N][kh]  R(c, k1/2) of every row (reciprocal half-size sets)
  int* rh_cnt = nullptr;     // [N]
  int* q_idx = nullptr;      // [N][capq]  V rows after local query expansion
  __half* q_val = nullptr;
  int* q_cnt = nullptr;
  int* col_cnt = nullptr;    // [N+1]
  int* inv_ofs = nullptr;    // [N+1]
  int* cursor = nullptr;     // [N+1]
  void* inv_ent = nullptr;   // [(N-Q)*capq] inverted-list entries of the gallery rows: (row - Q) << 16 | fp16 weight
                             // in one word, or two words (row - Q, weight) for more than 65 536 gallery rows
  float* jac_tab = nullptr;  // [16384] Jaccard term of the blend per temp_min bit pattern (jaccard_table_kernel)
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  __half* tmin_scratch = nullptr;  // [Q][N-Q] only when (N-Q)*2 bytes exceed shared memory
  size_t tmin_bytes = 0;
};

int rerank_k(int k1, int k2);
int rerank_kh(int k1);
int rerank_cap(int k1);
int rerank_capq(int N, int k1, int k2);
size_t rerank_carve(Carver& c, int N, int Q, int k1, int k2, RerankWs* w);

int launch_rowmax(const float* E, long long lde, int N, float* rowmax, cudaStream_t stream);
int launch_transpose_add(const float* X, long long ldx, float* E, long long lde, int N, bool accumulate,
                         cudaStream_t stream);
int launch_topk_rows(const float* mat, long long ld, int rows, int cols, const float* row_div, int k, int* idx_out,
                     float* val_out, cudaStream_t stream);
int run_rerank_stages(const float* E, long long lde, const float* rowmax, int N, int Q, int k1, int k2,
                      double lambda_value, const RerankWs& w, float* out, long long ldo, cudaStream_t stream);

// row-sharded stages (see rerank.cu)
int launch_krecip_rows(const float* E, long long lde, const float* rowmax, const int* rank_all, int N, int k1,
                       int k2, int row0, int nrows, int* v_idx, __half* v_val, int* v_cnt, int* rh_idx,
                       int* rh_cnt, cudaStream_t stream);
int launch_expand_rows(const int* rank_all, int N, int k1, int k2, int row0, int nrows, const int* v_idx,
                       const __half* v_val, const int* v_cnt, int* q_idx, __half* q_val, int* q_cnt,
                       cudaStream_t stream, int* col_cnt = nullptr, int count_from = 0);
int launch_jaccard_rows(const float* E, long long lde, const float* rowmax, int N, int Q, int k1, int k2, double lambda_value,
                        int row0, int nq_local, const int* f_idx, const __half* f_val, const int* f_cnt, int f_cap,
                        const RerankWs& w, float* out, long long ldo, cudaStream_t stream);

}  // namespace demo
