// prep_rows: fp32 feature rows -> (optional L2 normalisation) -> power-of-two row scaling
// -> fp16 hi/lo split + squared norms.  One warp per row; HBM-bound:
// reads 4*d B and writes 4*d B (hi+lo) per row.
#include "prep.cuh"

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
  const float ymax = do_norm ? amax / denom : amax;  // division is monotone: max|y| exactly

  int e = 0;
  if (ymax > 0.f && ymax < 3.0e38f) {
    e = 14 - ilogbf(ymax);
    e = max(-100, min(100, e));
  }

  // pass 2: normalise, split, accumulate |y|^2
  float ss2 = 0.f;
  __half* hr = hi + static_cast<long long>(r) * pitch;
  __half* lr = lo + static_cast<long long>(r) * pitch;
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
    hr[k] = h;
    lr[k] = l;
  }
  ss2 = warp_sum(ss2);
  if (lane == 0) {
    norm[r] = do_norm ? ss2 : ss;
    inv_scale[r] = ldexpf(1.f, -e);
  }
}

}  // namespace

int launch_prep_rows(const float* x, int rows, int d, long long ldx, int norm_mode,
                     const int* perm, const PrepView& out, float* xn_out, long long ldxn,
                     cudaStream_t stream) {
  if (rows <= 0) return DEMO_OK;
  DEMO_REQUIRE(x && out.hi && out.lo && out.norm && out.inv_scale, "prep_rows: null pointer");
  DEMO_REQUIRE(d > 0 && out.pitch >= d && out.pitch % 8 == 0, "prep_rows: bad d/pitch (%d, %d)", d,
               out.pitch);
  const int blocks = ceil_div(rows, kWarpsPerBlock);
  prep_rows_kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(
      x, rows, d, ldx, norm_mode, perm, out.hi, out.lo, out.pitch, out.norm, out.inv_scale, xn_out,
      ldxn);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // namespace demo
