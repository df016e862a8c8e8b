// Standalone device check of libdemo_b200's distance GEMM (no torch): accuracy of the
// tcgen05 split-fp16 path and of the SIMT kernel against an fp64 host reference, and a quick
// throughput number.   Build: make -C tools   Run (GPU box): tools/gemm_check
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../include/demo_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)
#define DK(x)                                                            \
  do {                                                                   \
    int r = (x);                                                         \
    if (r != 0) {                                                        \
      printf("demo error %d: %s (%s:%d)\n", r, demo_last_error(), __FILE__, __LINE__); \
      exit(3);                                                           \
    }                                                                    \
  } while (0)

struct Stats {
  double max_abs = 0, max_rel = 0, mean_signed = 0, rms = 0;
  long long n = 0, mismatch_vs_other = 0;
};

static void make_rows(std::vector<float>& x, int rows, int d, unsigned seed, bool normalise, float corr) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> centre(d);
  for (auto& c : centre) c = nd(rng);
  x.resize((size_t)rows * d);
  for (int r = 0; r < rows; ++r) {
    double ss = 0;
    for (int k = 0; k < d; ++k) {
      float v = corr * centre[k] + nd(rng);
      x[(size_t)r * d + k] = v;
      ss += (double)v * v;
    }
    if (normalise) {
      float inv = 1.f / (float)std::sqrt(ss);
      for (int k = 0; k < d; ++k) x[(size_t)r * d + k] *= inv;
    }
  }
}

static bool run_case(int Q, int G, int d, bool normalise, float corr, int mode, int sample) {
  std::vector<float> hq, hg;
  make_rows(hq, Q, d, 1234 + Q, normalise, corr);
  make_rows(hg, G, d, 99 + G, normalise, corr);
  float *dq, *dg, *dout, *dsimt, *drowmax;
  CK(cudaMalloc(&dq, hq.size() * 4));
  CK(cudaMalloc(&dg, hg.size() * 4));
  CK(cudaMalloc(&dout, (size_t)Q * G * 4));
  CK(cudaMalloc(&dsimt, (size_t)Q * G * 4));
  CK(cudaMalloc(&drowmax, (size_t)Q * 4));
  CK(cudaMemcpy(dq, hq.data(), hq.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dg, hg.data(), hg.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, (size_t)Q * G * 4));
  size_t wsb = demo_sqdist_workspace_bytes(Q, G, d, mode | DEMO_FLAG_SIMT | DEMO_FLAG_L2NORM);
  void* ws;
  CK(cudaMalloc(&ws, wsb));
  DK(demo_sqdist_f32(dq, dg, Q, G, d, d, d, dout, G, mode, drowmax, nullptr, nullptr, ws, wsb, nullptr));
  CK(cudaDeviceSynchronize());
  DK(demo_sqdist_f32(dq, dg, Q, G, d, d, d, dsimt, G, mode | DEMO_FLAG_SIMT, nullptr, nullptr, nullptr, ws, wsb, nullptr));
  CK(cudaDeviceSynchronize());
  std::vector<float> out((size_t)Q * G), simt((size_t)Q * G), rowmax(Q);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(simt.data(), dsimt, simt.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(rowmax.data(), drowmax, Q * 4, cudaMemcpyDeviceToHost));

  Stats st, ss;
  std::mt19937 rng(7);
  const long long total = (long long)Q * G;
  const long long nsamp = sample > 0 && sample < total ? sample : total;
  std::vector<double> qq(Q), gg(G);
  for (int i = 0; i < Q; ++i) {
    double s = 0;
    for (int k = 0; k < d; ++k) s += (double)hq[(size_t)i * d + k] * hq[(size_t)i * d + k];
    qq[i] = s;
  }
  for (int i = 0; i < G; ++i) {
    double s = 0;
    for (int k = 0; k < d; ++k) s += (double)hg[(size_t)i * d + k] * hg[(size_t)i * d + k];
    gg[i] = s;
  }
  for (long long it = 0; it < nsamp; ++it) {
    long long idx = nsamp == total ? it : (long long)(rng() % total);
    int i = (int)(idx / G), j = (int)(idx % G);
    double dot = 0;
    for (int k = 0; k < d; ++k) dot += (double)hq[(size_t)i * d + k] * hg[(size_t)j * d + k];
    double ref;
    if (mode == DEMO_DIST_SQ) ref = qq[i] + gg[j] - 2 * dot;
    else if (mode == DEMO_DIST_SQRT) ref = std::sqrt(std::max(qq[i] + gg[j] - 2 * dot, 1e-12));
    else if (mode == DEMO_DIST_COS_SIM) ref = dot / std::sqrt(qq[i] * gg[j]);
    else ref = (1 - dot / std::sqrt(qq[i] * gg[j])) / 2;
    const double scale = std::max(std::fabs(ref), mode == DEMO_DIST_SQ ? (qq[i] + gg[j]) * 1e-1 : 1e-1);
    auto upd = [&](Stats& s, float v) {
      double e = (double)v - ref;
      s.max_abs = std::max(s.max_abs, std::fabs(e));
      s.max_rel = std::max(s.max_rel, std::fabs(e) / scale);
      s.mean_signed += e;
      s.rms += e * e;
      s.n++;
    };
    upd(st, out[idx]);
    upd(ss, simt[idx]);
  }
  // rowmax check (exact: max of the stored row)
  int rowmax_bad = 0;
  for (int i = 0; i < Q; ++i) {
    float m = -INFINITY;
    for (int j = 0; j < G; ++j) m = std::fmax(m, out[(size_t)i * G + j]);
    if (m != rowmax[i]) rowmax_bad++;
  }
  printf("case Q=%d G=%d d=%d norm=%d corr=%.1f mode=%d: tc  max_abs %.3e max_rel %.3e bias %.3e rms %.3e | simt max_abs %.3e bias %.3e rms %.3e | rowmax_bad %d\n",
         Q, G, d, (int)normalise, corr, mode, st.max_abs, st.max_rel, st.mean_signed / st.n, std::sqrt(st.rms / st.n),
         ss.max_abs, ss.mean_signed / ss.n, std::sqrt(ss.rms / ss.n), rowmax_bad);
  bool ok = st.max_rel < 1e-5 && rowmax_bad == 0 && std::isfinite(st.max_abs);
  if (!ok) printf("  ^^^ FAIL\n");
  cudaFree(dq); cudaFree(dg); cudaFree(dout); cudaFree(dsimt); cudaFree(drowmax); cudaFree(ws);
  return ok;
}

static void bench(int Q, int G, int d, int iters) {
  float *dq, *dg, *dout;
  CK(cudaMalloc(&dq, (size_t)Q * d * 4));
  CK(cudaMalloc(&dg, (size_t)G * d * 4));
  CK(cudaMalloc(&dout, (size_t)Q * G * 4));
  std::vector<float> h;
  make_rows(h, Q, d, 1, true, 0.3f);
  CK(cudaMemcpy(dq, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  make_rows(h, G, d, 2, true, 0.3f);
  CK(cudaMemcpy(dg, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  size_t wsb = demo_sqdist_workspace_bytes(Q, G, d, 0);
  void* ws;
  CK(cudaMalloc(&ws, wsb));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) DK(demo_sqdist_f32(dq, dg, Q, G, d, d, d, dout, G, 0, nullptr, nullptr, nullptr, ws, wsb, nullptr));
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) DK(demo_sqdist_f32(dq, dg, Q, G, d, d, d, dout, G, 0, nullptr, nullptr, nullptr, ws, wsb, nullptr));
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= iters;
  double flop = 2.0 * Q * G * d;
  printf("bench Q=%d G=%d d=%d: %.3f ms  -> %.1f TFLOP/s algorithmic (x3 executed = %.1f)  store %.1f GB/s\n", Q, G, d, ms,
         flop / ms * 1e-9, 3 * flop / ms * 1e-9, (double)Q * G * 4 / ms * 1e-6);
  cudaFree(dq); cudaFree(dg); cudaFree(dout); cudaFree(ws);
}

// Accuracy on near-duplicate pairs (high cosine): A == B, report the self-distance error and the
// dot-product relative error as a function of d.
static void self_case(int n, int d, bool normalise) {
  std::vector<float> h;
  make_rows(h, n, d, 4242, normalise, 0.0f);
  float *dx, *dout;
  CK(cudaMalloc(&dx, h.size() * 4));
  CK(cudaMalloc(&dout, (size_t)n * n * 4));
  CK(cudaMemcpy(dx, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  size_t wsb = demo_sqdist_workspace_bytes(n, n, d, DEMO_FLAG_SIMT);
  void* ws;
  CK(cudaMalloc(&ws, wsb));
  for (int simt = 0; simt < 2; ++simt) {
    DK(demo_sqdist_f32(dx, dx, n, n, d, d, d, dout, n, simt ? DEMO_FLAG_SIMT : 0, nullptr, nullptr, nullptr, ws, wsb, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)n * n);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double worst = 0, mean = 0, norm2 = 0;
    for (int i = 0; i < n; ++i) {
      double s = 0;
      for (int k = 0; k < d; ++k) s += (double)h[(size_t)i * d + k] * h[(size_t)i * d + k];
      norm2 += s / n;
      worst = std::max(worst, std::fabs((double)out[(size_t)i * n + i]));
      mean += out[(size_t)i * n + i] / (double)n;
    }
    printf("self-distance n=%d d=%d norm=%d %s: |x|^2=%.3g  max|d_ii| %.3e  mean d_ii %.3e  (relative to 2|x|^2: %.2e)\n", n, d,
           (int)normalise, simt ? "simt" : "tc  ", norm2, worst, mean, worst / (2 * norm2));
  }
  cudaFree(dx); cudaFree(dout); cudaFree(ws);
}

// Calibration probe for the accumulation bias of the tensor-core path: B row i is a noisy copy of A
// row i with a prescribed cosine; reports the mean / rms RELATIVE error of the recovered dot
// product (from d_ii = |a|^2 + |b|^2 - 2ab) against fp64, as a function of d and the cosine.
static void dup_case(int n, int d, double cosv) {
  std::vector<float> a, noise, b((size_t)n * d);
  make_rows(a, n, d, 777, true, 0.0f);
  make_rows(noise, n, d, 778, true, 0.0f);
  const double s = std::sqrt(std::max(0.0, 1 - cosv * cosv));
  for (int i = 0; i < n; ++i) {
    double ss = 0;
    for (int k = 0; k < d; ++k) {
      double v = cosv * a[(size_t)i * d + k] + s * noise[(size_t)i * d + k];
      b[(size_t)i * d + k] = (float)v;
      ss += v * v;
    }
    float inv = (float)(1 / std::sqrt(ss));
    for (int k = 0; k < d; ++k) b[(size_t)i * d + k] *= inv;
  }
  float *da, *db, *dout;
  CK(cudaMalloc(&da, a.size() * 4));
  CK(cudaMalloc(&db, b.size() * 4));
  CK(cudaMalloc(&dout, (size_t)n * n * 4));
  CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
  size_t wsb = demo_sqdist_workspace_bytes(n, n, d, DEMO_FLAG_SIMT);
  void* ws;
  CK(cudaMalloc(&ws, wsb));
  for (int simt = 0; simt < 2; ++simt) {
    DK(demo_sqdist_f32(da, db, n, n, d, d, d, dout, n, simt ? DEMO_FLAG_SIMT : 0, nullptr, nullptr, nullptr, ws, wsb, nullptr));
    CK(cudaDeviceSynchronize());
    std::vector<float> out((size_t)n * n);
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    double mean_rel = 0, rms_rel = 0, mean_abs = 0, max_abs = 0, mean_dot = 0;
    for (int i = 0; i < n; ++i) {
      double aa = 0, bb = 0, ab = 0;
      for (int k = 0; k < d; ++k) {
        double x = a[(size_t)i * d + k], y = b[(size_t)i * d + k];
        aa += x * x; bb += y * y; ab += x * y;
      }
      const double ref = aa + bb - 2 * ab;
      const double e = (double)out[(size_t)i * n + i] - ref;   // = -2 * (dot_tc - dot)
      const double rel = -0.5 * e / ab;
      mean_rel += rel / n; rms_rel += rel * rel / n; mean_abs += e / n; mean_dot += ab / n;
      max_abs = std::max(max_abs, std::fabs(e));
    }
    printf("dup d=%d cos=%.2f %s: dot %.3f  rel dot err mean %+.3e rms %.3e  | dist err mean %+.3e max %.3e | per-d %.3e\n", d, cosv,
           simt ? "simt" : "tc  ", mean_dot, mean_rel, std::sqrt(rms_rel), mean_abs, max_abs, mean_rel / d);
  }
  cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(ws);
}

int main(int argc, char** argv) {
  if (!demo_device_ok()) {
    printf("no sm_100 device\n");
    return 1;
  }
  bool ok = true;
  ok &= run_case(128, 256, 32, true, 0.5f, DEMO_DIST_SQ, 0);
  ok &= run_case(128, 256, 64, true, 0.5f, DEMO_DIST_SQ, 0);
  ok &= run_case(200, 300, 1536, true, 0.5f, DEMO_DIST_SQ, 0);
  ok &= run_case(836, 836, 1536, true, 0.3f, DEMO_DIST_SQ, 200000);
  ok &= run_case(836, 836, 1536, true, 2.0f, DEMO_DIST_SQ, 200000);   // high cosine: stresses accumulation bias
  ok &= run_case(333, 1001, 100, true, 0.5f, DEMO_DIST_SQ, 0);         // ragged everything
  ok &= run_case(128, 128, 768, false, 0.0f, DEMO_DIST_SQRT, 0);       // triplet-like, un-normalised
  ok &= run_case(130, 515, 520, false, 1.0f, DEMO_DIST_COS_SIM, 0);
  ok &= run_case(130, 515, 520, false, 1.0f, DEMO_DIST_COS_DIST, 0);
  ok &= run_case(1715, 8575, 1536, true, 0.3f, DEMO_DIST_SQ, 100000);
  printf(ok ? "ALL CASES PASS\n" : "SOME CASES FAILED\n");
  self_case(256, 64, true);
  self_case(256, 512, true);
  self_case(256, 1536, true);
  self_case(256, 4096, true);
  self_case(128, 768, false);
  for (int d : {64, 256, 768, 1536, 2048, 4096})
    for (double c : {1.0, 0.9, 0.5, 0.2, 0.05}) dup_case(256, d, c);
  if (argc > 1) {
    bench(1672, 1672, 1536, 20);
    bench(10290, 10290, 1536, 10);
    bench(8192, 32768, 1536, 5);
  }
  return ok ? 0 : 1;
}
