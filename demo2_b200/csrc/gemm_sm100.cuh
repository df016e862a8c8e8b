// Persistent, warp-specialised tcgen05 distance GEMM for sm_100a.
//
//   acc[m][n] = sum_k  a_hi*b_hi + a_hi*b_lo + a_lo*b_hi          (fp16 operands, fp32 TMEM accumulator)
//   dot[m][n] = acc * 2^-ea[m] * 2^-eb[n]                          (exact power-of-two rescale)
//
// Operands are the prepared fp16 hi/lo rows of prep.cuh (K-major, 64-byte swizzled tiles loaded
// by TMA).  Warp roles per CTA (256 threads, one CTA per SM):
//   warp 0  TMA producer       (4 tile loads per k-block: A_hi, A_lo, B_hi, B_lo)
//   warp 1  MMA issuer         (3 x BK/16 tcgen05.mma per k-block, one thread)
//   warp 2  TMEM allocator
//   warps 4-11 epilogue        (thread t of warp w owns accumulator row 32*(w%4)+t and the
//                               column half (w-4)/4 of the tile: two warps per TMEM lane quadrant,
//                               i.e. two epilogue warps per SM sub-partition to hide ALU latency)
// The 128 x 256 fp32 accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue
// of tile i overlaps the MMAs of tile i+1.  What the epilogue does with the distances is a
// policy class (store / rank-count / extract-positives / hard-mining), see gemm_epilogues.cuh.
#pragma once

#include "common.cuh"
#include "prep.cuh"

namespace demo {

constexpr int kBM = 128;            // rows of A per tile (TMEM lanes)
constexpr int kBN = 256;            // rows of B per tile (TMEM columns)
constexpr int kBK = 32;             // k elements per k-block; hi[32] | lo[32] = 128 B = one SWIZZLE_128B span
constexpr int kUmmaK = 16;          // K per tcgen05.mma (kind::f16)
constexpr int kTileABytes = kBM * kBK * 2;
constexpr int kTileBBytes = kBN * kBK * 2;
constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;  // 48 KB
constexpr int kGemmThreads = 384;
constexpr int kEpiThreads = 256;
constexpr int kEpiCols = kBN / 2;   // columns per epilogue thread
constexpr int kTmemCols = 2 * kBN;  // 512: double-buffered accumulator

// One unit of work: a fixed block of 128 A rows against `n_rows` consecutive B rows
// starting at n0 (processed as ceil(n_rows / kBN) tiles by the same CTA).
struct WorkUnit {
  int m0, n0, n_rows, index;
};

// How units are enumerated.  All three roles of a CTA walk the same sequence.
struct Schedule {
  int mode = 0;          // 0: dense tiles, n-grouped raster; 1: chunked rows of tiles; 2: device list;
                         // 3: upper triangle of a square tile grid, rows folded in pairs (make_folded_schedule2)
  int M = 0, N = 0;      // valid rows of A / B
  int m_blocks = 0, n_tiles = 0;
  int group_n = 8;       // mode 0: n-tiles per raster group
  int chunk_tiles = 1;   // mode 1: n-tiles per unit
  int n_chunks = 0;      // mode 1
  int group_m = 16;      // mode 1: m-blocks that share a chunk consecutively
  int adj = 0;           // mode 1: > 1 = that many chunks are in flight per group and the workers that
                         // share an m-block are neighbours (w, w + 1, ...) instead of w, w + group_m, ...
  int m_block_rows = kBM;  // A rows per unit (2 * kBM for the CTA-pair kernel)
  int num_units = 0;     // modes 0/1
  const int4* list = nullptr;      // mode 2: (m_block, n0, n_rows, _)
  // Pacing of the persistent workers (CTA-pair kernel).  A unit is cut into steps of pace_tiles
  // tiles (pace_steps per unit); pace[i] counts the workers whose producer has issued every load
  // of its i-th step, and a producer starts step i only when all workers are done with step
  // i - 1 - pace_window.  Keeps the workers that share a gallery chunk within an L2 lifetime of
  // each other.  nullptr: free-running.  Zeroed by the caller (num_units * pace_steps entries).
  unsigned* pace = nullptr;
  int pace_window = 0, pace_tiles = 1, pace_steps = 1;
  const int* list_count = nullptr; // mode 2: device-side unit count
  // Optional per-256-row-block flags (device): units of a flagged A block are empty for this launch
  // (the fused rank count leaves query blocks with more thresholds than one window to the slab path).
  const unsigned char* m_skip = nullptr;
  // Units whose A block starts outside [m_lo, m_hi) are empty (extract pass of one query group).
  int m_lo = 0, m_hi = 0x7fffffff;
  // Symmetric all-pairs problems (A and B are row ranges of one set with global offsets tri_a0 /
  // tri_b0): tiles that hold no element with (global row) <= (global column) are empty.
  int tri = 0, tri_a0 = 0, tri_b0 = 0;
};

__device__ __forceinline__ int schedule_num_units(const Schedule& s) {
  return s.mode == 2 ? __ldg(s.list_count) : s.num_units;
}

__device__ __forceinline__ WorkUnit schedule_get(const Schedule& s, int u) {
  WorkUnit w;
  w.index = u;
  if (s.mode == 0) {
    const int per_group = s.group_n * s.m_blocks;
    const int g = u / per_group, r = u - g * per_group;
    const int n_in = min(s.group_n, s.n_tiles - g * s.group_n);
    const int m = r / n_in, n = g * s.group_n + (r - m * n_in);
    w.m0 = m * s.m_block_rows;
    w.n0 = n * kBN;
    w.n_rows = min(kBN, s.N - w.n0);
  } else if (s.mode == 1) {
    const int per_group = s.group_m * s.n_chunks;
    const int g = u / per_group, r = u - g * per_group;
    const int m_in = min(s.group_m, s.m_blocks - g * s.group_m);
    int c = r / m_in, m = g * s.group_m + (r - c * m_in);
    if (s.adj > 1 && m_in == s.group_m) {
      const int blk_units = s.group_m * s.adj, blk = r / blk_units;
      if ((blk + 1) * s.adj <= s.n_chunks) {     // full block of adj chunks x group_m m-blocks
        const int idx = r - blk * blk_units;
        m = g * s.group_m + idx / s.adj;
        c = blk * s.adj + idx % s.adj;
      }
    }
    w.m0 = m * s.m_block_rows;
    w.n0 = c * s.chunk_tiles * kBN;
    w.n_rows = min(s.chunk_tiles * kBN, s.N - w.n0);
  } else if (s.mode == 3) {
    // square grid (m_blocks == n_tiles, square tiles): row i holds the tiles n >= i; rows i and
    // m_blocks - 1 - i together hold n_tiles + 1 tiles, so "super-row" sr = u / (n_tiles + 1) is
    // a closed form and every unit is non-empty (except the second half of the middle row of an
    // odd grid)
    const int per = s.n_tiles + 1;
    const int sr = u / per, t = u - sr * per;
    int m = sr, n = sr + t;
    if (t >= s.n_tiles - sr) {
      m = s.m_blocks - 1 - sr;
      n = m + (t - (s.n_tiles - sr));
    }
    w.m0 = m * s.m_block_rows;
    w.n0 = n * kBN;
    w.n_rows = (t >= s.n_tiles - sr && m == sr) ? 0 : min(kBN, s.N - w.n0);
  } else {
    const int4 e = __ldg(s.list + u);
    w.m0 = e.x * kBM;
    w.n0 = e.y;
    w.n_rows = e.z;
  }
  if (s.m_skip != nullptr && __ldg(s.m_skip + (w.m0 >> 8)) != 0) w.n_rows = 0;
  if (w.m0 < s.m_lo || w.m0 >= s.m_hi) w.n_rows = 0;
  if (s.tri && s.tri_a0 + w.m0 > s.tri_b0 + w.n0 + w.n_rows - 1) w.n_rows = 0;
  return w;
}

struct TileInfo {
  int m0, n0, n_valid;   // n_valid: valid columns in this tile (<= kBN)
  int unit_index;
  bool first_in_unit, last_in_unit;
};

// Walks the tiles of a CTA in schedule order (same order as the producer / MMA loops).
struct TileCursor {
  const Schedule& s;
  int num_units, u, stride, n_off;
  WorkUnit w;
  __device__ TileCursor(const Schedule& s_, int num_units_, int first, int stride_)
      : s(s_), num_units(num_units_), u(first), stride(stride_), n_off(0) {
    skip_empty();
  }
  __device__ TileCursor& operator=(const TileCursor& o) {
    num_units = o.num_units;
    u = o.u;
    stride = o.stride;
    n_off = o.n_off;
    w = o.w;
    return *this;
  }
  __device__ TileCursor(const TileCursor& o) = default;
  __device__ void skip_empty() {
    while (u < num_units) {
      w = schedule_get(s, u);
      if (w.n_rows > 0) break;
      u += stride;
    }
  }
  __device__ bool valid() const { return u < num_units; }
  __device__ void advance() {
    n_off += kBN;
    if (n_off >= w.n_rows) {
      n_off = 0;
      u += stride;
      skip_empty();
    }
  }
  __device__ TileInfo info() const {
    TileInfo t;
    t.m0 = w.m0;
    t.n0 = w.n0 + n_off;
    t.n_valid = min(kBN, w.n_rows - n_off);
    t.unit_index = u;
    t.first_in_unit = n_off == 0;
    t.last_in_unit = n_off + kBN >= w.n_rows;
    return t;
  }
};

template <class Epi>
struct GemmSmem {
  static constexpr int kStages = Epi::kStages;
  static constexpr int kBarrierBytes = 8 * (2 * kStages + 4) + 16;
  static constexpr int kEpiOffset = kStages * kStageBytes + round_up(kBarrierBytes, 128);
  static constexpr int kTotal = kEpiOffset + Epi::kSmemBytes + 1024;  // + alignment slack
};

template <class Epi>
__global__ void __launch_bounds__(kGemmThreads, 1)
sqdist_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const Schedule sched, const int num_k_blocks, const typename Epi::Params ep) {
  constexpr int kStages = Epi::kStages;
  if (Epi::skip(ep)) return;  // conditional passes (uniform over the grid, before any setup)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // keep the pointer in the shared address space (plain pointer arithmetic, no integer round trip),
  // otherwise every epilogue access degrades to a generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* bar_empty = bar_full + kStages;
  uint64_t* bar_tfull = bar_empty + kStages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  uint8_t* epi_smem = smem + GemmSmem<Epi>::kEpiOffset;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bar_tfull[s], 1);
      mbar_init(&bar_tempty[s], kEpiThreads / 32);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_units = schedule_num_units(sched);

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const WorkUnit w = schedule_get(sched, u);
        for (int n_off = 0; n_off < w.n_rows; n_off += kBN) {
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            mbar_wait(&bar_empty[stage], phase ^ 1u);
            uint8_t* st = smem + stage * kStageBytes;
            mbar_expect_tx(&bar_full[stage], kStageBytes);
            // one 128-byte box row = hi[32] | lo[32] of the k-block (interleaved operands, prep.cuh)
            tma_load_2d(st, &tm_a, &bar_full[stage], kb * 2 * kBK, w.m0);
            tma_load_2d(st + 2 * kTileABytes, &tm_b, &bar_full[stage], kb * 2 * kBK, w.n0 + n_off);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const WorkUnit w = schedule_get(sched, u);
        for (int n_off = 0; n_off < w.n_rows; n_off += kBN) {
          mbar_wait(&bar_tempty[as], aphase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(as * kBN);
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            mbar_wait(&bar_full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kStageBytes);
            // 128-byte swizzled rows: hi at byte 0, lo at byte 64 of every row
            const uint64_t a_hi = make_kmajor_desc<128>(sa);
            const uint64_t a_lo = a_hi + 4;
            const uint64_t b_hi = make_kmajor_desc<128>(sa + 2 * kTileABytes);
            const uint64_t b_lo = b_hi + 4;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);  // +32 B per step
              // small cross terms first, then the main term
              umma_f16(tmem_acc, a_hi + adv, b_lo + adv, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_f16(tmem_acc, a_lo + adv, b_hi + adv, idesc, 1u);
              umma_f16(tmem_acc, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            umma_commit(&bar_empty[stage]);  // frees the smem stage once these MMAs retire
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&bar_tfull[as]);  // accumulator complete -> epilogue
          if (++as == 2) {
            as = 0;
            aphase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue ------------------------------
    const int q = warp & 3;             // TMEM lane quadrant accessible to this warp
    const int row_in_tile = q * 32 + lane;
    const int epi_tid = threadIdx.x - (kGemmThreads - kEpiThreads);
    const int col0 = ((warp - 4) >> 2) * kEpiCols;  // this thread's column half of the tile
    Epi epi(ep, epi_smem, epi_tid, row_in_tile, col0);
    int as = 0;
    uint32_t aphase = 0;
    // The per-column metadata of tile i+1 is fetched from global memory while tile i is being
    // processed (stage_load -> registers, stage_store -> the other shared-memory buffer), so its
    // L2 latency never sits between two tiles.
    TileCursor cur(sched, num_units, blockIdx.x, gridDim.x);
    if (cur.valid()) {
      epi.stage_load(cur.info());
      epi.stage_store(0);
    }
    while (cur.valid()) {
      const TileInfo t = cur.info();
      TileCursor nxt = cur;
      nxt.advance();
      epi.tile_begin(t, as);  // starts with an epilogue-wide barrier: staged columns become visible
      if (nxt.valid()) epi.stage_load(nxt.info());
      mbar_wait(&bar_tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * kBN + col0);
      epi.tile_body(t, as, taddr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_tempty[as]);
      epi.tile_end(t, as);
      if (nxt.valid()) epi.stage_store(as ^ 1);
      cur = nxt;
      if (++as == 2) {
        as = 0;
        aphase ^= 1u;
      }
    }
    epi.finish();
  }

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
int make_operand_tensor_map(CUtensorMap* map, const __half* base, int rows, int d, int pitch,
                            int box_rows);

struct GemmOperands {
  CUtensorMap a, b;
  int num_k_blocks;
};
int make_gemm_operands(const PrepView& a, const PrepView& b, GemmOperands* ops);

Schedule make_dense_schedule(int M, int N);
Schedule make_chunked_schedule(int M, int N, int chunk_tiles, int d_pitch);
Schedule make_list_schedule(int M, int N, const int4* list, const int* list_count);

template <class Epi>
int launch_sqdist_gemm(const GemmOperands& ops, const Schedule& sched, int max_units,
                       const typename Epi::Params& ep, cudaStream_t stream) {
  if (max_units <= 0) return DEMO_OK;
  auto kernel = sqdist_gemm_kernel<Epi>;
  constexpr int smem = GemmSmem<Epi>::kTotal;
  static_assert(smem <= 232448, "shared memory budget exceeded");
  static PerDeviceInt configured;
  DEMO_CHECK_CUDA(ensure_dynamic_smem(configured, kernel, smem));
  const int grid = max_units < num_sms() ? max_units : num_sms();
  kernel<<<grid, kGemmThreads, smem, stream>>>(ops.a, ops.b, sched, ops.num_k_blocks, ep);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // namespace demo
