// Operand preparation for the split-fp16 distance GEMM.
//
// Every fp32 feature row x is (optionally L2-normalised, then) scaled by an exact power of
// two s = 2^e so that max|x*s| lies in [2^14, 2^15), and split into two fp16 numbers
//   hi = fp16(x*s),  lo = fp16(x*s - hi)          (x*s = hi + lo up to 2^-22 relative)
// The tensor cores then accumulate hi*hi + hi*lo + lo*hi in fp32 (3 fp16 MMA passes, the
// fp16 analogue of 3xTF32 at twice the MMA rate) and the epilogue undoes the scaling with
// the exact factors 2^-e_a * 2^-e_b.
#pragma once

#include "common.cuh"

namespace demo {

// View of a prepared operand (device pointers into a caller-owned buffer).
struct PrepView {
  // hi and lo are INTERLEAVED per 32-element k-block in one array [rows][2 * pitch] fp16:
  // row r, k-block kb holds hi[32] | lo[32] = one 128-byte line, fetched by ONE TMA box row.
  __half* hi = nullptr;       // base of the interleaved array (hi of k-block 0)
  __half* lo = nullptr;       // hi + 32 (lo of k-block 0); same array
  float* norm = nullptr;      // [rows]  sum x^2 (after the optional normalisation), fp32
  float* inv_scale = nullptr; // [rows]  2^-e
  int rows = 0, d = 0, pitch = 0;
};

inline int prep_pitch(int d) { return round_up(d, 64); }

// Largest feature dimension for which the accumulation-bias compensation of prep.cu is validated
// (tests: near-duplicate rows within 1e-5 of the fp64 distance for d = 64 .. 2048; the bias grows
// with the number of tcgen05.mma accumulation steps, 3 * d / 16).  demo_sqdist_f32 routes longer
// rows to the fp32 FMA kernel (the north_star's "FFMA fallback").
constexpr int kMaxCompensatedDim = 2048;

// Bytes of a prepared operand and its carve-up (same function sizes and slices).
inline size_t prep_carve(Carver& c, int rows, int d, PrepView* v) {
  PrepView t;
  t.rows = rows;
  t.d = d;
  t.pitch = prep_pitch(d);
  size_t r = static_cast<size_t>(rows > 0 ? rows : 1);
  t.hi = c.take<__half>(r * 2 * t.pitch);
  t.lo = t.hi ? t.hi + 32 : nullptr;
  t.norm = c.take<float>(r);
  t.inv_scale = c.take<float>(r);
  if (v) *v = t;
  return c.off;
}
inline size_t prep_bytes(int rows, int d) {
  Carver c(nullptr, ~size_t(0));
  return round_up(prep_carve(c, rows, d, nullptr), size_t(1024));
}

enum : int {
  PREP_NORM_NONE = 0,
  PREP_NORM_F_NORMALIZE = 1,  // x / max(|x|, 1e-12)   (utils/metrics.py:345)
  PREP_NORM_TRIPLET = 2,      // x / (|x| + 1e-12)     (layers/triplet_loss.py:12)
};

// perm (optional): output row r is input row perm[r].  xn_out (optional): normalised fp32 rows
// written in INPUT order (row perm[r]).
// host_input: x is pinned host memory read in place over PCIe (small persistent grid that can
// share the SMs with a resident GEMM CTA).
int launch_prep_rows(const float* x, int rows, int d, long long ldx, int norm_mode,
                     const int* perm, const PrepView& out, float* xn_out, long long ldxn,
                     cudaStream_t stream, bool host_input = false);

}  // namespace demo
