// prep_rows: fp32 feature rows -> (optional L2 normalisation) -> power-of-two row scaling
// -> fp16 hi/lo split + squared norms.  One warp per row; HBM-bound:
// reads 4*d B and writes 4*d B (hi+lo) per row.
#include "prep.cuh"

#include <cstdlib>

namespace demo {

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Calibrated compensation of the tensor cores' round-toward-zero accumulation: every
// tcgen05.mma adds its 16 products to the fp32 accumulator with truncation, which biases a dot
// product with mostly same-signed terms low by ~(2.8 .. 6.2)e-9 * d (relative; measured on B200
// with tools/gemm_check `dup` cases for d = 64 .. 4096 and cosines 0.05 .. 1).  The mid value is
// folded into the row scales (half on each operand) at zero cost in the epilogues; it leaves a
// residual of <= 2e-9 * d * |a.b| on the distance (3e-6 for unit rows at d = 1536, was 1.9e-5).
constexpr float kAccumBiasPerDim = 5.2e-9f;

__device__ __forceinline__ float row_inv_scale(int e, int d) {
  return ldexpf(1.f + 0.5f * kAccumBiasPerDim * static_cast<float>(d), -e);
}

__device__ __forceinline__ int row_exponent(float ymax) {
  int e = 0;
  if (ymax > 0.f && ymax < 3.0e38f) {
    e = 14 - ilogbf(ymax);
    e = max(-100, min(100, e));
  }
  return e;
}

__device__ __forceinline__ uint2 pack_half4(__half a, __half b, __half c, __half d) {
  uint2 r;
  r.x = static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
  r.y = static_cast<uint32_t>(__half_as_ushort(c)) | (static_cast<uint32_t>(__half_as_ushort(d)) << 16);
  return r;
}

// Fast path: d % 4 == 0, 16-byte aligned rows, d <= 128 * kNV.  The row lives in registers
// (kNV float4 per lane): ONE pass over the input, float4 loads, 8-byte hi / lo stores.
template <int kNV>
__device__ __forceinline__ void prep_row_vec(const float* __restrict__ x, int r, int d, long long ldx, int norm_mode,
                                             const int* __restrict__ perm, __half* __restrict__ hi,
                                             __half* __restrict__ lo, int pitch, float* __restrict__ norm,
                                             float* __restrict__ inv_scale, float* __restrict__ xn_out,
                                             long long ldxn, int lane) {
  const int src = perm ? __ldg(perm + r) : r;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(src) * ldx);
  const int nvec = d >> 2, pvec = pitch >> 2;
  float4 v[kNV];
  float ss = 0.f, amax = 0.f;
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const int k = lane + 32 * i;
    v[i] = k < nvec ? __ldcs(xr + k) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // same summation order per lane as the generic kernel is NOT required: |x|^2 of the raw row is
  // only used for the normalisation denominator / un-normalised norms, in fp32 either way
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    ss = fmaf(v[i].x, v[i].x, ss);
    ss = fmaf(v[i].y, v[i].y, ss);
    ss = fmaf(v[i].z, v[i].z, ss);
    ss = fmaf(v[i].w, v[i].w, ss);
    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
  }
  ss = warp_sum(ss);
  amax = warp_max(amax);
  float denom = 1.f;
  if (norm_mode == PREP_NORM_F_NORMALIZE) denom = fmaxf(sqrtf(ss), 1e-12f);
  if (norm_mode == PREP_NORM_TRIPLET) denom = sqrtf(ss) + 1e-12f;
  const bool do_norm = norm_mode != PREP_NORM_NONE;
  const int e = row_exponent(do_norm ? amax / denom : amax);
  float ss2 = 0.f;
  // interleaved layout: element k -> half index 64 * (k / 32) + k % 32 (hi), + 32 (lo); in units of
  // 4 halfs (uint2): vector kv = k / 4 -> 16 * (kv / 8) + kv % 8 (hi), + 8 (lo)
  uint2* hr = reinterpret_cast<uint2*>(hi + static_cast<long long>(r) * 2 * pitch);
  uint2* lr = reinterpret_cast<uint2*>(lo + static_cast<long long>(r) * 2 * pitch);
  float4* xo = xn_out ? reinterpret_cast<float4*>(xn_out + static_cast<long long>(src) * ldxn) : nullptr;
#pragma unroll
  for (int i = 0; i < kNV; ++i) {
    const int k = lane + 32 * i;
    if (k >= pvec) continue;
    float4 y = v[i];
    if (do_norm) {
      y.x = y.x / denom;
      y.y = y.y / denom;
      y.z = y.z / denom;
      y.w = y.w / denom;
    }
    if (xo && k < nvec) xo[k] = y;
    ss2 = fmaf(y.x, y.x, ss2);
    ss2 = fmaf(y.y, y.y, ss2);
    ss2 = fmaf(y.z, y.z, ss2);
    ss2 = fmaf(y.w, y.w, ss2);
    const float s0 = ldexpf(y.x, e), s1 = ldexpf(y.y, e), s2 = ldexpf(y.z, e), s3 = ldexpf(y.w, e);
    const __half h0 = __float2half_rn(s0), h1 = __float2half_rn(s1), h2 = __float2half_rn(s2), h3 = __float2half_rn(s3);
    const int kq = 16 * (k >> 3) + (k & 7);
    hr[kq] = pack_half4(h0, h1, h2, h3);
    lr[kq] = pack_half4(__float2half_rn(s0 - __half2float(h0)), __float2half_rn(s1 - __half2float(h1)),
                       __float2half_rn(s2 - __half2float(h2)), __float2half_rn(s3 - __half2float(h3)));
  }
  ss2 = warp_sum(ss2);
  if (lane == 0) {
    norm[r] = do_norm ? ss2 : ss;
    inv_scale[r] = row_inv_scale(e, d);
  }
}

template <int kNV>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
prep_rows_vec_kernel(const float* __restrict__ x, int rows, int d, long long ldx, int norm_mode,
                     const int* __restrict__ perm, __half* __restrict__ hi, __half* __restrict__ lo,
                     int pitch, float* __restrict__ norm, float* __restrict__ inv_scale,
                     float* __restrict__ xn_out, long long ldxn) {
  const int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= rows) return;
  prep_row_vec<kNV>(x, r, d, ldx, norm_mode, perm, hi, lo, pitch, norm, inv_scale, xn_out, ldxn, threadIdx.x & 31);
}

// Streaming variant for rows that live in pinned HOST memory (zero-copy: the loads go over PCIe).
// A SMALL grid (kStreamBlocks blocks of 64 threads) walks the rows with a grid stride, so that a
// slab of the gallery can be pulled in and prepared while the tensor cores rank the previous slab.
// PCIe needs ~0.2 MB in flight; 64 blocks x 2 warps x 6 KB keep 0.8 MB, and they reach the same
// 51 GB/s as a grid that fills the machine (32 blocks: 43 GB/s, 16: 32, 8: 16).  What matters for
// the overlap (tools/probe_stream.py --overlap-only, half of the 20k x 1M count GEMM against the
// zero-copy prepare of the other half): a machine-filling prepare grid (2 blocks per SM) ran
// almost back to back with the persistent GEMM grid (118 ms for 71 + 60); 64 blocks finish
// together with it (76 ms with the GEMM on 70 CTA pairs, and in the streamed evaluation of
// 20k x 1M the count stage takes 141 ms = the GEMM alone, whether the GEMM leaves SMs free
// (demo_eval_count_range reserve_sms) or not).  The kernel asks for the SAME
// shared-memory carve-out as the GEMM (maximum shared memory): an SM cannot change its L1 /
// shared split while blocks are resident.
constexpr int kStreamWarps = 2;
constexpr int kStreamBlocks = 64;
template <int kNV>
__global__ void __launch_bounds__(kStreamWarps * 32, 14)
prep_rows_stream_kernel(const float* __restrict__ x, int rows, int d, long long ldx, int norm_mode,
                        const int* __restrict__ perm, __half* __restrict__ hi, __half* __restrict__ lo,
                        int pitch, float* __restrict__ norm, float* __restrict__ inv_scale,
                        float* __restrict__ xn_out, long long ldxn) {
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * kStreamWarps;
  for (int r = blockIdx.x * kStreamWarps + (threadIdx.x >> 5); r < rows; r += stride)
    prep_row_vec<kNV>(x, r, d, ldx, norm_mode, perm, hi, lo, pitch, norm, inv_scale, xn_out, ldxn, lane);
}

// Generic path (any d, any alignment): two passes over the row, scalar accesses.
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
prep_rows_kernel(const float* __restrict__ x, int rows, int d, long long ldx, int norm_mode,
                 const int* __restrict__ perm, __half* __restrict__ hi, __half* __restrict__ lo,
                 int pitch, float* __restrict__ norm, float* __restrict__ inv_scale,
                 float* __restrict__ xn_out, long long ldxn) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int src = perm ? perm[r] : r;
  const float* xr = x + static_cast<long long>(src) * ldx;

  // pass 1: sum of squares and max |x| of the raw row
  float ss = 0.f, amax = 0.f;
  for (int k = lane; k < d; k += 32) {
    float v = __ldg(xr + k);
    ss = fmaf(v, v, ss);
    amax = fmaxf(amax, fabsf(v));
  }
  ss = warp_sum(ss);
  amax = warp_max(amax);

  float denom = 1.f;
  if (norm_mode == PREP_NORM_F_NORMALIZE) denom = fmaxf(sqrtf(ss), 1e-12f);
  if (norm_mode == PREP_NORM_TRIPLET) denom = sqrtf(ss) + 1e-12f;
  const bool do_norm = norm_mode != PREP_NORM_NONE;
  const int e = row_exponent(do_norm ? amax / denom : amax);  // division is monotone: max|y| exactly

  // pass 2: normalise, split, accumulate |y|^2
  float ss2 = 0.f;
  __half* hr = hi + static_cast<long long>(r) * 2 * pitch;   // interleaved: see PrepView
  __half* lr = lo + static_cast<long long>(r) * 2 * pitch;
  for (int k = lane; k < pitch; k += 32) {
    float y = 0.f;
    if (k < d) {
      y = __ldg(xr + k);
      if (do_norm) y = y / denom;
      if (xn_out) xn_out[static_cast<long long>(src) * ldxn + k] = y;
    }
    ss2 = fmaf(y, y, ss2);
    const float ys = ldexpf(y, e);
    const __half h = __float2half_rn(ys);
    const __half l = __float2half_rn(ys - __half2float(h));
    const int kq = 64 * (k >> 5) + (k & 31);
    hr[kq] = h;
    lr[kq] = l;
  }
  ss2 = warp_sum(ss2);
  if (lane == 0) {
    norm[r] = do_norm ? ss2 : ss;
    inv_scale[r] = row_inv_scale(e, d);
  }
}

}  // namespace

int launch_prep_rows(const float* x, int rows, int d, long long ldx, int norm_mode,
                     const int* perm, const PrepView& out, float* xn_out, long long ldxn,
                     cudaStream_t stream, bool host_input) {
  if (rows <= 0) return DEMO_OK;
  DEMO_REQUIRE(x && out.hi && out.lo && out.norm && out.inv_scale, "prep_rows: null pointer");
  DEMO_REQUIRE(d > 0 && out.pitch >= d && out.pitch % 8 == 0, "prep_rows: bad d/pitch (%d, %d)", d,
               out.pitch);
  const int blocks = ceil_div(rows, kWarpsPerBlock);
  const bool vec = d % 4 == 0 && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 && d <= 2048 &&
                   (!xn_out || (ldxn % 4 == 0 && (reinterpret_cast<uintptr_t>(xn_out) & 15u) == 0));
#define DEMO_PREP_VEC(NV)                                                                              \
  prep_rows_vec_kernel<NV><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(                                \
      x, rows, d, ldx, norm_mode, perm, out.hi, out.lo, out.pitch, out.norm, out.inv_scale, xn_out, ldxn)
  if (host_input && vec) {
    // zero-copy rows: small persistent grid (see prep_rows_stream_kernel)
    static const int total_env = getenv("DEMO_STREAM_TOTAL") ? atoi(getenv("DEMO_STREAM_TOTAL")) : 0;  // experiments
    const int sblocks = min(ceil_div(rows, kStreamWarps), total_env > 0 ? total_env : kStreamBlocks);
#define DEMO_PREP_STREAM(NV)                                                                           \
  {                                                                                                    \
    static PerDeviceInt carved;                                                                        \
    const int dev = current_device();                                                                  \
    if (!carved.get(dev)) {                                                                            \
      DEMO_CHECK_CUDA(cudaFuncSetAttribute(prep_rows_stream_kernel<NV>,                                \
                                           cudaFuncAttributePreferredSharedMemoryCarveout,             \
                                           cudaSharedmemCarveoutMaxShared));                           \
      carved.set(dev, 1);                                                                              \
    }                                                                                                  \
    prep_rows_stream_kernel<NV><<<sblocks, kStreamWarps * 32, 0, stream>>>(                            \
        x, rows, d, ldx, norm_mode, perm, out.hi, out.lo, out.pitch, out.norm, out.inv_scale, xn_out, ldxn); \
  }
    if (out.pitch <= 512) DEMO_PREP_STREAM(4)
    else if (out.pitch <= 1024) DEMO_PREP_STREAM(8)
    else if (out.pitch <= 1536) DEMO_PREP_STREAM(12)
    else DEMO_PREP_STREAM(16)
#undef DEMO_PREP_STREAM
    DEMO_CHECK_CUDA(cudaGetLastError());
    return DEMO_OK;
  }
  if (vec && out.pitch <= 512) DEMO_PREP_VEC(4);
  else if (vec && out.pitch <= 1024) DEMO_PREP_VEC(8);
  else if (vec && out.pitch <= 1536) DEMO_PREP_VEC(12);
  else if (vec && out.pitch <= 2048) DEMO_PREP_VEC(16);
  else
    prep_rows_kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
        x, rows, d, ldx, norm_mode, perm, out.hi, out.lo, out.pitch, out.norm, out.inv_scale, xn_out,
        ldxn);
#undef DEMO_PREP_VEC
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // namespace demo
