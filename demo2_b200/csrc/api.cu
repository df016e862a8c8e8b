// extern "C" entry points (include/demo_b200.h).
#include "../../include/demo_b200.h"

#include "gemm_epilogues.cuh"
#include "gemm2_sm100.cuh"
#include "simt.cuh"

using namespace demo;

namespace {

__global__ void keys_to_float_kernel(const unsigned* keys, float* out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = key_float(keys[i]);
}

int norm_mode_of(int flags) {
  if (flags & DEMO_FLAG_L2NORM) return PREP_NORM_F_NORMALIZE;
  if (flags & DEMO_FLAG_TRIPLET_NORM) return PREP_NORM_TRIPLET;
  return PREP_NORM_NONE;
}

// rows longer than the validated range of the tensor-core path go to the fp32 FMA kernel
int sqdist_flags(int flags, int d) { return d > kMaxCompensatedDim ? (flags | DEMO_FLAG_SIMT) : flags; }

// row maxima of a stored matrix (SIMT path only; the tensor-core epilogue fuses them)
__global__ void __launch_bounds__(256) rowmax_kernel(const float* __restrict__ m, long long ld, int cols,
                                                     float* __restrict__ out) {
  const float* row = m + static_cast<long long>(blockIdx.x) * ld;
  float v = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += 256) v = fmaxf(v, row[c]);
  __shared__ float s[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) v = fmaxf(v, s[w]);
    out[blockIdx.x] = v;
  }
}

struct SqdistWs {
  PrepView a, b;
  unsigned* rowmax_keys;
  float* a_n;  // normalised copies for the SIMT path
  float* b_n;
};

size_t carve_sqdist(Carver& c, int Q, int G, int d, int flags, SqdistWs* w) {
  SqdistWs t;
  prep_carve(c, Q, d, &t.a);
  prep_carve(c, G, d, &t.b);
  t.rowmax_keys = c.take<unsigned>(Q > 0 ? Q : 1);
  const bool simt_norm = (flags & DEMO_FLAG_SIMT) && norm_mode_of(flags) != PREP_NORM_NONE;
  t.a_n = simt_norm ? c.take<float>(static_cast<size_t>(Q) * d) : nullptr;
  t.b_n = simt_norm ? c.take<float>(static_cast<size_t>(G) * d) : nullptr;
  if (w) *w = t;
  return c.off;
}

}  // namespace

extern "C" {

const char* demo_last_error(void) { return last_error(); }
int demo_version(void) { return 100; }

int demo_device_ok(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

size_t demo_sqdist_workspace_bytes(int Q, int G, int d, int flags) {
  Carver c(nullptr, ~size_t(0));
  return round_up(carve_sqdist(c, Q, G, d, sqdist_flags(flags, d), nullptr), size_t(1024));
}

int demo_sqdist_f32(const float* q, const float* g, int Q, int G, int d, int64_t ldq, int64_t ldg,
                    float* out, int64_t ldo, int flags, float* rowmax, float* qn_out, float* gn_out,
                    void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DEMO_REQUIRE(Q >= 0 && G >= 0 && d > 0, "sqdist: bad shape Q=%d G=%d d=%d", Q, G, d);
  if (Q == 0 || G == 0) return DEMO_OK;
  DEMO_REQUIRE(q && g && out && workspace, "sqdist: null pointer");
  DEMO_REQUIRE(ldq >= d && ldg >= d && ldo >= G, "sqdist: leading dimension too small");
  flags = sqdist_flags(flags, d);
  Carver c(workspace, workspace_bytes);
  SqdistWs w;
  carve_sqdist(c, Q, G, d, flags, &w);
  if (!c.ok()) {
    set_error("sqdist: workspace too small (%zu < %zu)", workspace_bytes, c.off);
    return DEMO_ERR_WORKSPACE;
  }
  const int mode = flags & 3;
  const int nm = norm_mode_of(flags);
  const bool simt = (flags & DEMO_FLAG_SIMT) != 0;
  float* qn = qn_out ? qn_out : w.a_n;
  float* gn = gn_out ? gn_out : w.b_n;
  DEMO_TRY(launch_prep_rows(q, Q, d, ldq, nm, nullptr, w.a, qn, d, stream));
  DEMO_TRY(launch_prep_rows(g, G, d, ldg, nm, nullptr, w.b, gn, d, stream));
  if (simt) {
    const bool use_n = nm != PREP_NORM_NONE;
    DEMO_TRY(launch_simt_dist(use_n ? qn : q, use_n ? gn : g, Q, G, d, use_n ? d : ldq, use_n ? d : ldg,
                              w.a.norm, w.b.norm, out, ldo, mode, stream));
    if (rowmax) {
      rowmax_kernel<<<Q, 256, 0, stream>>>(out, ldo, G, rowmax);
      DEMO_CHECK_CUDA(cudaGetLastError());
    }
    return DEMO_OK;
  }
  if (rowmax) DEMO_CHECK_CUDA(cudaMemsetAsync(w.rowmax_keys, 0, sizeof(unsigned) * Q, stream));
  GemmOperands ops;
  DEMO_TRY(make_gemm_operands(w.a, w.b, &ops));
  EpiStore::Params ep;
  ep.a_norm = w.a.norm;
  ep.a_inv = w.a.inv_scale;
  ep.b_norm = w.b.norm;
  ep.b_inv = w.b.inv_scale;
  ep.out = out;
  ep.ldo = ldo;
  ep.M = Q;
  ep.mode = mode;
  ep.rowmax_key = rowmax ? w.rowmax_keys : nullptr;
  if (prefer_pair_kernel(Q, G)) {
    DEMO_TRY(make_gemm2_operands(w.a, w.b, &ops));
    const Schedule s = make_dense_schedule2(Q, G);
    DEMO_TRY(launch_sqdist_gemm2<EpiStore>(ops, s, s.num_units, ep, stream));
  } else {
    const Schedule s = make_dense_schedule(Q, G);
    DEMO_TRY(launch_sqdist_gemm<EpiStore>(ops, s, s.num_units, ep, stream));
  }
  if (rowmax) {
    keys_to_float_kernel<<<ceil_div(Q, 256), 256, 0, stream>>>(w.rowmax_keys, rowmax, Q);
    DEMO_CHECK_CUDA(cudaGetLastError());
  }
  return DEMO_OK;
}

}  // extern "C"
