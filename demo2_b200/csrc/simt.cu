// SIMT (FFMA) distance kernel on the raw fp32 rows: the on-device cross-check of the
// tensor-core path (flag DEMO_DIST_SIMT) and the path for operands that cannot be fed to TMA.
// Plain 64x64x16 shared-memory tiling, 4x4 outputs per thread, fp32 FMA accumulation in k order.
#include "gemm_epilogues.cuh"
#include "simt.cuh"

namespace demo {

namespace {

constexpr int TS = 64, TK = 16;

__global__ void __launch_bounds__(256)
simt_dist_kernel(const float* __restrict__ a, const float* __restrict__ b, int M, int N, int d,
                 long long lda, long long ldb, const float* __restrict__ a_norm,
                 const float* __restrict__ b_norm, float* __restrict__ out, long long ldo, int mode) {
  __shared__ float sa[TK][TS + 1];
  __shared__ float sb[TK][TS + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < d; k0 += TK) {
    for (int i = threadIdx.x; i < TS * TK; i += 256) {
      const int r = i / TK, k = i % TK;
      sa[k][r] = (m0 + r < M && k0 + k < d) ? __ldg(a + (long long)(m0 + r) * lda + k0 + k) : 0.f;
      sb[k][r] = (n0 + r < N && k0 + k < d) ? __ldg(b + (long long)(n0 + r) * ldb + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = sa[k][ty * 4 + i];
        bv[i] = sb[k][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) out[(long long)m * ldo + n] = finish_distance(mode, acc[i][j], a_norm[m], b_norm[n]);
    }
  }
}

}  // namespace

int launch_simt_dist(const float* a, const float* b, int M, int N, int d, long long lda, long long ldb,
                     const float* a_norm, const float* b_norm, float* out, long long ldo, int mode,
                     cudaStream_t stream) {
  if (M <= 0 || N <= 0) return DEMO_OK;
  dim3 grid(ceil_div(N, TS), ceil_div(M, TS));
  simt_dist_kernel<<<grid, 256, 0, stream>>>(a, b, M, N, d, lda, ldb, a_norm, b_norm, out, ldo, mode);
  DEMO_CHECK_CUDA(cudaGetLastError());
  return DEMO_OK;
}

}  // namespace demo
