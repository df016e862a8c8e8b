// Batch-hard triplet loss for training-size batches (N <= 256 anchors, any d), ONE launch for
// all modalities: fp32 Gram matrix, sqrt / clamp, hardest positive / negative with indices, the
// ranking loss and its mean -- layers/triplet_loss.py:16-31 (euclidean_dist), :51-104
// (hard_example_mining), :121-135 (TripletLoss.__call__), executed 3x per training step on
// [128, C] (one call per modality).  At this size the tensor-core path is pure
// launch latency (prep + init + GEMM + decode + ~8 small torch kernels for the loss); here the
// whole forward is one kernel and the whole backward another.
//
//   grid  = (ceil(N / 8) row blocks, B modalities), 256 threads
//   CTA   = 8 anchor rows x all N columns; thread (ty = tid / 128, tx = tid % 128) owns rows
//           r0 + 4 ty .. + 3 and columns tx, tx + 128; X is staged through shared memory in
//           32-wide k-chunks (row stride 36 floats: conflict-free 128-bit reads), next chunk
//           prefetched into registers while the current one is multiplied.
//   The last CTA of a modality to finish (atomic ticket) adds the per-block loss sums in block
//   order (deterministic), writes loss[b] and the status word, and resets the ticket.
#include "common.cuh"

using namespace demo;

namespace {

constexpr int kTsThreads = 256;
constexpr int kTsRows = 8;      // anchor rows per CTA
constexpr int kTsKC = 32;       // k-chunk
constexpr int kTsStride = 36;   // padded row stride (floats)
constexpr int kTsMaxN = 256;
constexpr int kTsMaxB = 8;
constexpr int kTsMaxBlocks = kTsMaxN / kTsRows;

struct TsPointers {
  const float* x[kTsMaxB];
};

struct TsWorkspace {      // caller-owned, zero-initialised ONCE (the kernel leaves the tickets at 0)
  unsigned ticket[kTsMaxB];
  float partial[kTsMaxB][kTsMaxBlocks];
  unsigned status[kTsMaxB][kTsMaxBlocks];
};

__device__ __forceinline__ int label_at(const void* labels, int is64, int i) {
  return is64 ? static_cast<int>(static_cast<const long long*>(labels)[i]) : static_cast<const int*>(labels)[i];
}

// z = dist_an' - dist_ap' (after the hard_factor scaling); margin < 0 (or NaN) = SoftMarginLoss
__device__ __forceinline__ float loss_term(float z, float margin, bool soft) {
  return soft ? log1pf(expf(-z)) : fmaxf(0.f, margin - z);
}
__device__ __forceinline__ float loss_dz(float z, float margin, bool soft) {
  return soft ? -1.f / (1.f + expf(z)) : (margin - z > 0.f ? -1.f : 0.f);
}

__global__ void __launch_bounds__(kTsThreads)
triplet_small_fwd_kernel(TsPointers xs, int N, int d, long long ld, const void* __restrict__ labels, int lab64,
                         float margin, int soft, float hard_factor, float* __restrict__ loss,
                         float* __restrict__ dist_ap, float* __restrict__ dist_an, long long* __restrict__ p_idx,
                         long long* __restrict__ n_idx, int* __restrict__ status, TsWorkspace* __restrict__ ws) {
  __shared__ __align__(16) float s_x[kTsMaxN * kTsStride];
  __shared__ int s_lab[kTsMaxN];
  __shared__ unsigned long long s_bp[kTsRows][4], s_bn[kTsRows][4];
  __shared__ int s_cnt[kTsRows][4];
  __shared__ float s_term[kTsRows];
  __shared__ unsigned s_stat[kTsRows];
  __shared__ unsigned s_last;

  const int b = blockIdx.y, tid = threadIdx.x;
  const int tx = tid & 127, ty = tid >> 7;
  const int r0 = blockIdx.x * kTsRows;
  const float* __restrict__ x = xs.x[b];

  for (int i = tid; i < N; i += kTsThreads) s_lab[i] = label_at(labels, lab64, i);

  // staging map: thread t loads float4 #(t & 7) of rows (t >> 3) + 32 j, j = 0 .. 7
  const int lrow = tid >> 3, lvec = tid & 7;
  const bool vec_ok = (d % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15u) == 0);
  float4 pre[kTsMaxN / 32];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < kTsMaxN / 32; ++j) {
      const int row = lrow + 32 * j, k = k0 + 4 * lvec;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < N) {
        const float* p = x + static_cast<long long>(row) * ld + k;
        if (vec_ok && k + 3 < d) {
          v = __ldg(reinterpret_cast<const float4*>(p));
        } else {
          if (k < d) v.x = __ldg(p);
          if (k + 1 < d) v.y = __ldg(p + 1);
          if (k + 2 < d) v.z = __ldg(p + 2);
          if (k + 3 < d) v.w = __ldg(p + 3);
        }
      }
      pre[j] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < kTsMaxN / 32; ++j) {
      const int row = lrow + 32 * j;
      if (row < kTsMaxN) *reinterpret_cast<float4*>(&s_x[row * kTsStride + 4 * lvec]) = pre[j];
    }
  };

  float acc[4][2] = {}, na[4] = {}, nb[2] = {};
  const int two = N > 128 ? 2 : 1;
  fetch(0);
  for (int k0 = 0; k0 < d; k0 += kTsKC) {
    __syncthreads();      // everybody is done with the previous chunk
    stash();
    __syncthreads();
    if (k0 + kTsKC < d) fetch(k0 + kTsKC);
#pragma unroll
    for (int k = 0; k < kTsKC; k += 4) {
      float4 a[4], c[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(&s_x[(r0 + 4 * ty + i) % kTsMaxN * kTsStride + k]);
      c[0] = *reinterpret_cast<const float4*>(&s_x[tx * kTsStride + k]);
      c[1] = two == 2 ? *reinterpret_cast<const float4*>(&s_x[(tx + 128) * kTsStride + k]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        nb[j] = fmaf(c[j].x, c[j].x, nb[j]);
        nb[j] = fmaf(c[j].y, c[j].y, nb[j]);
        nb[j] = fmaf(c[j].z, c[j].z, nb[j]);
        nb[j] = fmaf(c[j].w, c[j].w, nb[j]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        na[i] = fmaf(a[i].x, a[i].x, na[i]);
        na[i] = fmaf(a[i].y, a[i].y, na[i]);
        na[i] = fmaf(a[i].z, a[i].z, na[i]);
        na[i] = fmaf(a[i].w, a[i].w, na[i]);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          acc[i][j] = fmaf(a[i].x, c[j].x, acc[i][j]);
          acc[i][j] = fmaf(a[i].y, c[j].y, acc[i][j]);
          acc[i][j] = fmaf(a[i].z, c[j].z, acc[i][j]);
          acc[i][j] = fmaf(a[i].w, c[j].w, acc[i][j]);
        }
      }
    }
  }

  // hardest positive (max, anchor included) / negative (min); ties -> lowest index
  const int warp_in_ty = (tid >> 5) & 3, lane = tid & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = r0 + 4 * ty + i;
    unsigned long long bp = 0ull, bn = ~0ull;
    int cnt = 0;
    if (row < N) {
      const int mylab = s_lab[row];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = tx + 128 * j;
        if (col >= N) continue;
        const float dist = sqrtf(fmaxf(fmaf(-2.f, acc[i][j], na[i] + nb[j]), 1e-12f));   // triplet_loss.py:25-30
        const unsigned key = float_key(dist);
        if (s_lab[col] == mylab) {
          const unsigned long long v = (static_cast<unsigned long long>(key) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(col));
          bp = v > bp ? v : bp;
          ++cnt;
        } else {
          const unsigned long long v = (static_cast<unsigned long long>(key) << 32) | static_cast<unsigned>(col);
          bn = v < bn ? v : bn;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long op = __shfl_xor_sync(0xffffffffu, bp, o), on = __shfl_xor_sync(0xffffffffu, bn, o);
      bp = op > bp ? op : bp;
      bn = on < bn ? on : bn;
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) {
      s_bp[4 * ty + i][warp_in_ty] = bp;
      s_bn[4 * ty + i][warp_in_ty] = bn;
      s_cnt[4 * ty + i][warp_in_ty] = cnt;
    }
  }
  __syncthreads();
  if (tid < kTsRows) {
    const int row = r0 + tid;
    float term = 0.f;
    unsigned st = 0u;
    if (row < N) {
      unsigned long long bp = 0ull, bn = ~0ull;
      int cnt = 0;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        bp = s_bp[tid][w] > bp ? s_bp[tid][w] : bp;
        bn = s_bn[tid][w] < bn ? s_bn[tid][w] : bn;
        cnt += s_cnt[tid][w];
      }
      const float ap = key_float(static_cast<unsigned>(bp >> 32));
      const float an = bn == ~0ull ? INFINITY : key_float(static_cast<unsigned>(bn >> 32));
      const long long o = static_cast<long long>(b) * N + row;
      const float aps = ap * (1.0f + hard_factor), ans = an * (1.0f - hard_factor);   // triplet_loss.py:127-128
      dist_ap[o] = aps;
      dist_an[o] = ans;
      p_idx[o] = static_cast<long long>(0xFFFFFFFFu - static_cast<unsigned>(bp & 0xFFFFFFFFu));
      n_idx[o] = bn == ~0ull ? -1 : static_cast<long long>(bn & 0xFFFFFFFFu);
      term = loss_term(ans - aps, margin, soft != 0);
      // status: bit 0 = number of positives differs from anchor 0's (reference: view(N, -1) fails),
      //         bit 1 = an anchor without a negative
      int c0 = 0;
      const int lab0 = s_lab[0];
      for (int j = 0; j < N; ++j) c0 += s_lab[j] == lab0;
      st = (cnt != c0 ? 1u : 0u) | (bn == ~0ull ? 2u : 0u);
    }
    s_term[tid] = term;
    s_stat[tid] = st;
  }
  __syncthreads();
  if (tid == 0) {
    float sum = 0.f;
    unsigned st = 0u;
    for (int i = 0; i < kTsRows; ++i) {
      sum += s_term[i];
      st |= s_stat[i];
    }
    ws->partial[b][blockIdx.x] = sum;
    ws->status[b][blockIdx.x] = st;
    __threadfence();
    s_last = atomicAdd(&ws->ticket[b], 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && tid == 0) {
    __threadfence();
    float total = 0.f;
    unsigned st = 0u;
    for (unsigned i = 0; i < gridDim.x; ++i) {
      total += *(volatile float*)&ws->partial[b][i];
      st |= *(volatile unsigned*)&ws->status[b][i];
    }
    loss[b] = total / static_cast<float>(N);     // reduction='mean' of SoftMarginLoss / MarginRankingLoss
    if (status) status[b] = static_cast<int>(st);
    ws->ticket[b] = 0u;                          // ready for the next launch
  }
}

// grad_x[b][r] = sum over the selected pairs that touch row r of c * (x_r - x_other):
//   anchor a, positive p[a]:  c = dL/d dist_ap[a] / D_ap[a]   (rows a and p[a], opposite signs)
//   anchor a, negative n[a]:  c = dL/d dist_an[a] / D_an[a]
// with dL/d dist from the ranking loss (g_loss[b] / N * dloss/dz * (-(1+h) | +(1-h))) plus the
// optional upstream gradients of the returned (scaled) distances; zero where the clamp was
// active.  Deterministic gather form, one CTA per (row, modality).
__global__ void __launch_bounds__(256)
triplet_small_bwd_kernel(TsPointers xs, int N, int d, long long ld, float margin, int soft, float hard_factor,
                         const float* __restrict__ dist_ap, const float* __restrict__ dist_an,
                         const long long* __restrict__ p_idx, const long long* __restrict__ n_idx,
                         const float* __restrict__ g_loss, const float* __restrict__ g_ap_ext,
                         const float* __restrict__ g_an_ext, float* __restrict__ grad, long long ldg) {
  __shared__ float s_cp[kTsMaxN], s_cn[kTsMaxN];
  __shared__ int s_p[kTsMaxN], s_n[kTsMaxN];
  const int r = blockIdx.x, b = blockIdx.y;
  const float* __restrict__ x = xs.x[b];
  const float gl = g_loss ? g_loss[b] / static_cast<float>(N) : 0.f;
  for (int a = threadIdx.x; a < N; a += blockDim.x) {
    const long long o = static_cast<long long>(b) * N + a;
    const float aps = dist_ap[o], ans = dist_an[o];           // scaled by (1 +- hard_factor)
    const long long ni = n_idx[o];
    const float dz = (ni >= 0) ? gl * loss_dz(ans - aps, margin, soft != 0) : 0.f;
    float gap = -dz + (g_ap_ext ? g_ap_ext[o] : 0.f);           // d / d (scaled dist_ap)
    float gan = dz + (g_an_ext ? g_an_ext[o] : 0.f);
    gap *= 1.0f + hard_factor;                                  // -> d / d raw distance
    gan *= 1.0f - hard_factor;
    const float dp = aps / (1.0f + hard_factor), dn = (1.0f - hard_factor) != 0.f ? ans / (1.0f - hard_factor) : 0.f;
    s_cp[a] = (dp * dp > 1e-12f) ? gap / dp : 0.f;
    s_cn[a] = (ni >= 0 && dn * dn > 1e-12f) ? gan / dn : 0.f;
    s_p[a] = static_cast<int>(p_idx[o]);
    s_n[a] = static_cast<int>(ni);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += blockDim.x) {
    const float xr = x[static_cast<long long>(r) * ld + k];
    float acc = 0.f;
    {
      const float cp = s_cp[r], cn = s_cn[r];
      if (cp != 0.f) acc += cp * (xr - x[static_cast<long long>(s_p[r]) * ld + k]);
      if (cn != 0.f) acc += cn * (xr - x[static_cast<long long>(s_n[r]) * ld + k]);
    }
    for (int a = 0; a < N; ++a) {
      const float cp = s_p[a] == r ? s_cp[a] : 0.f, cn = s_n[a] == r ? s_cn[a] : 0.f;
      if (cp != 0.f || cn != 0.f) acc += (cp + cn) * (xr - x[static_cast<long long>(a) * ld + k]);
    }
    grad[(static_cast<long long>(b) * N + r) * ldg + k] = acc;
  }
}

}  // namespace

extern "C" {

size_t demo_triplet_loss_workspace_bytes(void) { return sizeof(TsWorkspace); }
int demo_triplet_loss_max_batch(void) { return kTsMaxN; }

// Fused TripletLoss forward for B same-shaped feature matrices (modalities) that share the labels
// (layers/triplet_loss.py:121-135 called once per modality, layers/make_loss.py:47-52).
//   xs        HOST array of B device pointers, each [N][ld] fp32, N <= 256, B <= 8
//   labels    device, int32 (label_is_i64 = 0) or int64 (= 1, the reference's LongTensor)
//   margin    < 0 or NaN: SoftMarginLoss (margin=None in the reference), else MarginRankingLoss(margin)
//   loss [B]; dist_ap / dist_an [B][N] (already scaled by 1 +- hard_factor, as the reference
//   returns them); p_idx / n_idx [B][N] int64 (n_idx = -1: no negative in the batch)
//   status [B] (optional): bit 0 = anchors with different numbers of positives (the reference's
//   view(N, -1) at :79 raises), bit 1 = an anchor without a negative
//   ws        demo_triplet_loss_workspace_bytes() bytes, zero-initialised once by the caller
int demo_triplet_loss_fwd(const float* const* xs, int B, int N, int d, int64_t ld, const void* labels,
                          int label_is_i64, float margin, float hard_factor, float* loss, float* dist_ap,
                          float* dist_an, int64_t* p_idx, int64_t* n_idx, int* status, void* ws, size_t ws_bytes,
                          void* stream_) {
  DEMO_REQUIRE(xs && labels && loss && dist_ap && dist_an && p_idx && n_idx && ws, "triplet_loss_fwd: null pointer");
  DEMO_REQUIRE(B >= 1 && B <= kTsMaxB && N >= 1 && N <= kTsMaxN && d >= 1 && ld >= d,
               "triplet_loss_fwd: unsupported shape (B=%d N=%d d=%d; N <= %d, B <= %d)", B, N, d, kTsMaxN, kTsMaxB);
  DEMO_REQUIRE(ws_bytes >= sizeof(TsWorkspace), "triplet_loss_fwd: workspace too small");
  TsPointers p;
  for (int b = 0; b < kTsMaxB; ++b) p.x[b] = b < B ? xs[b] : nullptr;
  for (int b = 0; b < B; ++b) DEMO_REQUIRE(p.x[b], "triplet_loss_fwd: null feature pointer");
  const bool soft = !(margin >= 0.f);
  const dim3 grid(ceil_div(N, kTsRows), B);
  triplet_small_fwd_kernel<<<grid, kTsThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      p, N, d, ld, labels, label_is_i64, margin, soft ? 1 : 0, hard_factor, loss, dist_ap, dist_an,
      reinterpret_cast<long long*>(p_idx), reinterpret_cast<long long*>(n_idx), status, static_cast<TsWorkspace*>(ws));
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

// d loss / d x for demo_triplet_loss_fwd: g_loss [B] upstream gradient of the losses; g_ap_ext /
// g_an_ext (optional, [B][N]) upstream gradients of the returned distances; grad [B][N][ldg].
int demo_triplet_loss_bwd(const float* const* xs, int B, int N, int d, int64_t ld, float margin, float hard_factor,
                          const float* dist_ap, const float* dist_an, const int64_t* p_idx, const int64_t* n_idx,
                          const float* g_loss, const float* g_ap_ext, const float* g_an_ext, float* grad,
                          int64_t ldg, void* stream_) {
  DEMO_REQUIRE(xs && dist_ap && dist_an && p_idx && n_idx && grad, "triplet_loss_bwd: null pointer");
  DEMO_REQUIRE(B >= 1 && B <= kTsMaxB && N >= 1 && N <= kTsMaxN && d >= 1 && ld >= d && ldg >= d,
               "triplet_loss_bwd: unsupported shape");
  TsPointers p;
  for (int b = 0; b < kTsMaxB; ++b) p.x[b] = b < B ? xs[b] : nullptr;
  const bool soft = !(margin >= 0.f);
  triplet_small_bwd_kernel<<<dim3(N, B), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      p, N, d, ld, margin, soft ? 1 : 0, hard_factor, dist_ap, dist_an, reinterpret_cast<const long long*>(p_idx),
      reinterpret_cast<const long long*>(n_idx), g_loss, g_ap_ext, g_an_ext, grad, ldg);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // extern "C"
