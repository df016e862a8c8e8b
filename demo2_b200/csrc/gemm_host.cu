// Host side of the tcgen05 distance GEMM: TMA tensor maps, schedules, and the SIMT (FFMA)
// reference kernel used as an on-device self-check of the tensor-core path.
#include <cstdarg>
#include <cstdlib>

#include "gemm_epilogues.cuh"
#include "gemm2_sm100.cuh"

namespace demo {

// ---------------------------------------------------------------------------------------
// error string
// ---------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---------------------------------------------------------------------------------------
// TMA descriptors (driver entry point fetched through the runtime: no libcuda link needed)
// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_operand_tensor_map(CUtensorMap* map, const __half* base, int rows, int d, int pitch,
                            int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available (driver too old / no GPU)");
    return DEMO_ERR_CUDA;
  }
  DEMO_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15u) == 0 && pitch % 8 == 0,
               "operand not 16-byte aligned (base %p pitch %d)", (const void*)base, pitch);
  // interleaved hi|lo rows (prep.cuh): 2 * pitch halfs per row, one k-block = 64 halfs = 128 bytes
  (void)d;
  const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(2 * pitch), static_cast<cuuint64_t>(rows > 0 ? rows : 1)};
  const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(2 * pitch) * sizeof(__half)};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(2 * kBK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %d d %d pitch %d)", (int)r, rows, d, pitch);
    return DEMO_ERR_CUDA;
  }
  return DEMO_OK;
}

int make_gemm_operands(const PrepView& a, const PrepView& b, GemmOperands* ops) {
  DEMO_REQUIRE(a.d == b.d, "operand feature dims differ (%d vs %d)", a.d, b.d);
  DEMO_TRY(make_operand_tensor_map(&ops->a, a.hi, a.rows, a.d, a.pitch, kBM));
  DEMO_TRY(make_operand_tensor_map(&ops->b, b.hi, b.rows, b.d, b.pitch, kBN));
  ops->num_k_blocks = ceil_div(a.d, kBK);
  return DEMO_OK;
}

Schedule make_dense_schedule(int M, int N) {
  Schedule s;
  s.mode = 0;
  s.M = M;
  s.N = N;
  s.m_blocks = ceil_div(M, kBM);
  s.n_tiles = ceil_div(N, kBN);
  s.group_n = 8;
  s.num_units = s.m_blocks * s.n_tiles;
  return s;
}

// Dense raster for the CTA-pair kernel (256-row blocks).
Schedule make_dense_schedule2(int M, int N) {
  Schedule s = make_dense_schedule(M, N);
  s.m_block_rows = 2 * kBM;
  s.m_blocks = ceil_div(M, 2 * kBM);
  s.num_units = s.m_blocks * s.n_tiles;
  return s;
}

// Upper triangle of the square tile grid of an N x N all-pairs problem for the CTA-pair kernel
// (256 x 256 tiles), rows folded in pairs: only non-empty units, so the persistent workers get
// equal shares (the dense raster with empty lower-triangle units gave every worker a random
// 11.6 +- 2.4 of the 861 tiles at N = 10 290: the slowest pair set the time).
Schedule make_folded_schedule2(int N) {
  Schedule s = make_dense_schedule2(N, N);
  s.mode = 3;
  s.num_units = ceil_div(s.m_blocks, 2) * (s.n_tiles + 1);
  return s;
}

// The pair kernel pays off once there is enough work to fill the machine with 256 x 256 tiles.
bool prefer_pair_kernel(int M, int N) {
  static const bool force_1cta = getenv("DEMO_STORE_1CTA") != nullptr;  // A/B timing experiments
  return !force_1cta && M > kBM &&
         static_cast<long long>(ceil_div(M, 2 * kBM)) * ceil_div(N, kBN) >= num_sms() / 2;
}

// How many row blocks of A share a gallery chunk back to back.  The persistent workers (CTAs or
// CTA pairs) take consecutive units, so in one "round" they cover group_m A blocks x
// (workers / group_m) B chunks -- an L2-level super-tile.  group_m divides the worker count so
// that the rounds stay aligned with the chunks, and is the largest such divisor whose A blocks
// (re-streamed for every tile) stay below ~64 MB.  Measured on 20k x 1M (74 pairs): groups of
// 37 pairs x 8-tile chunks 162 ms / 298 GB of DRAM reads, groups of 8 x 32 tiles 167 ms / 386 GB.
static int balanced_group_m(int m_blocks, int block_rows, int d_pitch, int workers) {
  const double block_bytes = static_cast<double>(block_rows) * d_pitch * 4.0;   // hi + lo
  int best = 1;
  for (int g = 1; g <= workers; ++g)
    if (workers % g == 0 && g * block_bytes <= 64e6) best = g;
  return best < m_blocks ? best : (m_blocks > 0 ? m_blocks : 1);
}

Schedule make_chunked_schedule(int M, int N, int chunk_tiles, int d_pitch) {
  Schedule s;
  s.mode = 1;
  s.M = M;
  s.N = N;
  s.m_blocks = ceil_div(M, kBM);
  s.n_tiles = ceil_div(N, kBN);
  s.chunk_tiles = chunk_tiles < 1 ? 1 : chunk_tiles;
  s.n_chunks = ceil_div(s.n_tiles, s.chunk_tiles);
  s.group_m = balanced_group_m(s.m_blocks, kBM, d_pitch, num_sms());
  s.num_units = s.m_blocks * s.n_chunks;
  return s;
}

int make_gemm2_operands(const PrepView& a, const PrepView& b, GemmOperands* ops) {
  DEMO_REQUIRE(a.d == b.d, "operand feature dims differ (%d vs %d)", a.d, b.d);
  DEMO_TRY(make_operand_tensor_map(&ops->a, a.hi, a.rows, a.d, a.pitch, kBM));
  DEMO_TRY(make_operand_tensor_map(&ops->b, b.hi, b.rows, b.d, b.pitch, kBM));  // each CTA loads half a B tile
  ops->num_k_blocks = ceil_div(a.d, kBK);
  return DEMO_OK;
}

// Units of the CTA-pair kernel: 256 A rows x chunk_tiles B tiles.
Schedule make_chunked_schedule2(int M, int N, int chunk_tiles, int d_pitch, int workers) {
  if (workers <= 0) workers = num_sms() / 2;
  Schedule s;
  s.mode = 1;
  s.M = M;
  s.N = N;
  s.m_block_rows = 2 * kBM;
  s.m_blocks = ceil_div(M, 2 * kBM);
  s.n_tiles = ceil_div(N, kBN);
  s.chunk_tiles = chunk_tiles < 1 ? 1 : chunk_tiles;
  s.n_chunks = ceil_div(s.n_tiles, s.chunk_tiles);
  s.group_m = balanced_group_m(s.m_blocks, 2 * kBM, d_pitch, workers);
  if (const char* e = getenv("DEMO_GROUP_M")) s.group_m = atoi(e) > 0 ? atoi(e) : s.group_m;  // experiments
  s.num_units = s.m_blocks * s.n_chunks;
  // Neighbour order: the workers / group_m pairs that read the same query block are neighbouring
  // workers (w, w + 1) instead of w, w + group_m.  Measured DRAM reads per launch (run time equal
  // within 1 %): 20k x 262k 10.4 GB against 17.1 GB, but 20k x 1M 70.6 GB against 51.9 GB
  // (profiles/traffic_schedules_r2.txt) -- so it is used for galleries up to 512k rows only.
  // DEMO_ADJ=0 / 1 forces it off / on.
  {
    const char* e = getenv("DEMO_ADJ");
    const bool on = e ? atoi(e) != 0 : N <= (1 << 19);
    if (on && workers % s.group_m == 0) s.adj = workers / s.group_m;
  }
  return s;
}

int max_active_pairs(const void* kernel, int smem) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * num_sms());
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = 2;
  attr.val.clusterDim.y = 1;
  attr.val.clusterDim.z = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return num_sms() / 2;
  }
  return n;
}

Schedule make_list_schedule(int M, int N, const int4* list, const int* list_count) {
  Schedule s;
  s.mode = 2;
  s.M = M;
  s.N = N;
  s.m_blocks = ceil_div(M, kBM);
  s.n_tiles = ceil_div(N, kBN);
  s.list = list;
  s.list_count = list_count;
  return s;
}

}  // namespace demo
