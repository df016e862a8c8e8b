// CTA-pair (cta_group::2) variant of the persistent tcgen05 distance GEMM.
//
// Two CTAs of a cluster (the two SMs of a TPC) work on ONE 256 x 256 tile: each CTA owns 128 of
// the 256 A rows (its own 128 x 256 fp32 accumulator in its own TMEM, double-buffered) and loads
// HALF of the B tile (128 of its 256 rows); the leader CTA issues tcgen05.mma.cta_group::2 with
// M = 256, which reads A from both CTAs and the two B halves from both CTAs.  Per CTA and k-block
// that is 16 KB of A (hi+lo) + 16 KB of B (hi+lo) instead of 16 + 32 KB:
//   * L2 -> SM operand traffic and shared-memory fill bandwidth drop by a third,
//   * the tensor pipe reads 8 KB instead of 12 KB of operands per MMA from shared memory, which
//     leaves room for the epilogue's own shared-memory traffic (the 1-CTA count kernel lost a
//     quarter of its MMA rate to it),
//   * a stage shrinks to 32 KB, so the same shared memory holds a deeper TMA pipeline.
// Everything else (warp roles, schedules, epilogue policies) is shared with gemm_sm100.cuh.
//
// Barriers (per CTA unless noted):
//   full[s]    leader only; its producer arrives with expect_tx(2 x 32 KB); the TMA loads of BOTH
//              CTAs complete_tx on the leader's barrier (.cta_group::2 loads)
//   empty[s]   both; signalled by the leader's tcgen05.commit (multicast to both CTAs)
//   tfull[a]   both; leader's commit (multicast) -> both epilogues
//   tempty[a]  leader only; 2 x 8 epilogue warps arrive (the peer's warps remotely)
#pragma once

#include <cstdlib>

#include "gemm_sm100.cuh"

namespace demo {

constexpr int kStageBytes2 = 2 * kTileABytes + 2 * kTileABytes;  // A hi/lo + half-B hi/lo = 32 KB

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2D tile load whose completion bytes are signalled on a barrier of the LEADER CTA
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const void* map, uint32_t leader_bar,
                                                int32_t c_inner, int32_t c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(leader_bar), "r"(c_inner), "r"(c_outer)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                             uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at the same offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}

template <class Epi>
struct Gemm2Smem {
  static constexpr int kBudget = 232448 - 1024;
  static constexpr int kBarrierBytes = 8 * (2 * 8 + 4) + 16;  // room for up to 8 stages
  static constexpr int kFixed = round_up(kBarrierBytes, 128) + Epi::kSmemBytes;
  static constexpr int kStages = (kBudget - kFixed) / kStageBytes2 > 8 ? 8 : (kBudget - kFixed) / kStageBytes2;
  static constexpr int kEpiOffset = kStages * kStageBytes2 + round_up(kBarrierBytes, 128);
  static constexpr int kTotal = kEpiOffset + Epi::kSmemBytes + 1024;
  static_assert(kStages >= 3, "not enough shared memory for the TMA pipeline");
};

template <class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
sqdist_gemm2_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                    const Schedule sched, const int num_k_blocks, const typename Epi::Params ep) {
  constexpr int kStages = Gemm2Smem<Epi>::kStages;
  if (Epi::skip(ep)) return;  // uniform over the grid, before any setup
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes2);
  uint64_t* bar_empty = bar_full + kStages;
  uint64_t* bar_tfull = bar_empty + kStages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);
  uint8_t* epi_smem = smem + Gemm2Smem<Epi>::kEpiOffset;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

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
      mbar_init(&bar_tempty[s], 2 * kEpiThreads / 32);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc_cg2(tmem_slot, kTmemCols);
  tc_fence_before();
  cluster_sync_all();  // barrier inits and TMEM allocations of both CTAs are visible
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_units = schedule_num_units(sched);
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int row_ofs = static_cast<int>(rank) * kBM;  // this CTA's half of the 256 A rows / B rows

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      const uint32_t leader_full0 = mapa_u32(smem_u32(&bar_full[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      bool paced = sched.pace != nullptr;
      for (int u = cid; u < num_units; u += ncl, ++it) {
        const WorkUnit w = schedule_get(sched, u);
        int t = 0, step = it * sched.pace_steps;
        for (int n_off = 0; n_off < w.n_rows; n_off += kBN, ++t) {
          if (paced && t % sched.pace_tiles == 0) {
            const int j = step - 1 - sched.pace_window;
            if (j >= 0) {
              const unsigned expect = static_cast<unsigned>(min(ncl, num_units - (j / sched.pace_steps) * ncl));
              const volatile unsigned* flag = sched.pace + j;
              if (*flag < expect) {
                // Pacing is a performance device, never a correctness dependency: if the others do not
                // show up within 3 ms = ten units of the largest problem (part of the grid not resident
                // because another kernel or another tenant holds SMs), this worker stops waiting for
                // the rest of the launch.
                const unsigned long long t0 = globaltimer_ns();
                while (*flag < expect) {
                  __nanosleep(500);
                  if (globaltimer_ns() - t0 > 3000000ull) {
                    paced = false;
                    break;
                  }
                }
              }
            }
          }
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            mbar_wait(&bar_empty[stage], phase ^ 1u);
            uint8_t* st = smem + stage * kStageBytes2;
            const uint32_t lbar = leader_full0 + 8u * stage;
            if (leader) mbar_expect_tx(&bar_full[stage], 2 * kStageBytes2);
            tma_load_2d_cg2(st, &tm_a, lbar, kb * 2 * kBK, w.m0 + row_ofs);
            tma_load_2d_cg2(st + 2 * kTileABytes, &tm_b, lbar, kb * 2 * kBK, w.n0 + n_off + row_ofs);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          if (sched.pace != nullptr && (t + 1) % sched.pace_tiles == 0) {
            if (leader) atomicAdd(sched.pace + step, 1u);
            ++step;
          }
        }
        // short unit (end of a gallery row of chunks): the steps it does not have count as done
        if (sched.pace != nullptr && leader)
          for (; step < (it + 1) * sched.pace_steps; ++step) atomicAdd(sched.pace + step, 1u);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader CTA only) ------------------------------
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(2 * kBM, kBN);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int u = cid; u < num_units; u += ncl) {
        const WorkUnit w = schedule_get(sched, u);
        for (int n_off = 0; n_off < w.n_rows; n_off += kBN) {
          mbar_wait(&bar_tempty[as], aphase ^ 1u);
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(as * kBN);
          for (int kb = 0; kb < num_k_blocks; ++kb) {
            mbar_wait(&bar_full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kStageBytes2);
            const uint64_t a_hi = make_kmajor_desc<128>(sa);
            const uint64_t a_lo = a_hi + 4;
            const uint64_t b_hi = make_kmajor_desc<128>(sa + 2 * kTileABytes);
            const uint64_t b_lo = b_hi + 4;
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);  // +32 B per step
              // same accumulation order as the 1-CTA kernel: cross terms, then the main term
              umma_f16_cg2(tmem_acc, a_hi + adv, b_lo + adv, idesc, (kb | k) != 0 ? 1u : 0u);
              umma_f16_cg2(tmem_acc, a_lo + adv, b_hi + adv, idesc, 1u);
              umma_f16_cg2(tmem_acc, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            umma_commit_cg2(&bar_empty[stage]);  // frees the stage in both CTAs
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit_cg2(&bar_tfull[as]);  // both accumulators complete -> both epilogues
          if (++as == 2) {
            as = 0;
            aphase ^= 1u;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (both CTAs, own 128 rows) ------------------------------
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const int epi_tid = threadIdx.x - (kGemmThreads - kEpiThreads);
    const int col0 = ((warp - 4) >> 2) * kEpiCols;
    Epi epi(ep, epi_smem, epi_tid, row_in_tile, col0);
    const uint32_t leader_tempty0 = mapa_u32(smem_u32(&bar_tempty[0]), 0);
    int as = 0;
    uint32_t aphase = 0;
    TileCursor cur(sched, num_units, cid, ncl);
    auto info_of = [&](const TileCursor& c) {
      TileInfo t = c.info();
      t.m0 += row_ofs;
      return t;
    };
    if (cur.valid()) {
      epi.stage_load(info_of(cur));
      epi.stage_store(0);
    }
    while (cur.valid()) {
      const TileInfo t = info_of(cur);
      TileCursor nxt = cur;
      nxt.advance();
      epi.tile_begin(t, as);
      if (nxt.valid()) epi.stage_load(info_of(nxt));
      mbar_wait(&bar_tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(as * kBN + col0);
      epi.tile_body(t, as, taddr);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty0 + 8u * as);
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

  // The peer's shared memory and barriers are used by the leader's MMAs / commits until the very
  // end: nobody leaves (or frees TMEM) before both CTAs are done.
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  if (warp == 2) tmem_dealloc_cg2(tmem_base, kTmemCols);
}

// Operands for the CTA-pair kernel: all four tensor maps use 128-row boxes.
int make_gemm2_operands(const PrepView& a, const PrepView& b, GemmOperands* ops);
Schedule make_chunked_schedule2(int M, int N, int chunk_tiles, int d_pitch, int workers = 0);
Schedule make_dense_schedule2(int M, int N);
Schedule make_folded_schedule2(int N);
bool prefer_pair_kernel(int M, int N);
int max_active_pairs(const void* kernel, int smem);

template <class Epi>
int launch_sqdist_gemm2(const GemmOperands& ops, const Schedule& sched, int max_units,
                        const typename Epi::Params& ep, cudaStream_t stream, int max_pairs = 0) {
  if (max_units <= 0) return DEMO_OK;
  auto kernel = sqdist_gemm2_kernel<Epi>;
  constexpr int smem = Gemm2Smem<Epi>::kTotal;
  static_assert(smem <= 232448, "shared memory budget exceeded");
  static PerDeviceInt configured, pairs_of;   // per device: shared-memory opt-in, resident CTA pairs
  const int dev = current_device();
  DEMO_CHECK_CUDA(ensure_dynamic_smem(configured, kernel, smem));
  int pairs = pairs_of.get(dev);
  if (!pairs) {
    pairs = max_active_pairs(reinterpret_cast<const void*>(kernel), smem);
    if (pairs <= 0) {
      set_error("CTA-pair kernel cannot be scheduled on this device");
      return DEMO_ERR_CUDA;
    }
    if (const char* e = getenv("DEMO_PAIRS")) pairs = atoi(e) > 0 && atoi(e) < pairs ? atoi(e) : pairs;  // experiments
    pairs_of.set(dev, pairs);
  }
  if (max_pairs > 0 && max_pairs < pairs) pairs = max_pairs;   // the caller keeps SMs free for a concurrent kernel
  const int grid = 2 * (max_units < pairs ? max_units : pairs);
  kernel<<<grid, kGemmThreads, smem, stream>>>(ops.a, ops.b, sched, ops.num_k_blocks, ep);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // namespace demo
