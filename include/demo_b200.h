/*
 * libdemo_b200 -- C ABI of the B200-native DeMo retrieval hot path.
 *
 * The reference (maxingan2412/DeMo2) has no FFI: its boundary for this path is the Python
 * function surface imported by engine/processor.py:7 and layers/make_loss.py:9.  The Python
 * shim `demo2_b200` re-exports that surface and calls the entry points below through ctypes;
 * every entry point cites the reference interface it replaces.  See INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (demo_last_error() has the message);
 *   - all pointers are DEVICE pointers unless the name ends in _host; matrices are row-major
 *     fp32 with an explicit leading dimension in elements; labels are int32;
 *   - the caller owns every buffer including the workspace (query the size first); the
 *     library never allocates or frees device memory and keeps no pointer after returning;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous
 *     unless documented otherwise;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef DEMO_B200_H_
#define DEMO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DEMO_API __attribute__((visibility("default")))

enum {
  DEMO_OK = 0,
  DEMO_ERR_INVALID = -1,
  DEMO_ERR_CUDA = -2,
  DEMO_ERR_WORKSPACE = -3,
  DEMO_ERR_UNSUPPORTED = -4,
  DEMO_ERR_CAPACITY = -5
};

/* distance flavour (low 2 bits of `flags`) */
enum {
  DEMO_DIST_SQ = 0,       /* |q|^2+|g|^2-2qg          utils/metrics.py:395-401 euclidean_distance   */
  DEMO_DIST_SQRT = 1,     /* sqrt(clamp(.,1e-12))     layers/triplet_loss.py:16-31 euclidean_dist    */
  DEMO_DIST_COS_SIM = 2,  /* qg/(|q||g|)              cosine_similarity (north_star; absent upstream) */
  DEMO_DIST_COS_DIST = 3, /* (1-qg/(|q||g|))/2        layers/triplet_loss.py:34-48 cosine_dist       */
  DEMO_FLAG_L2NORM = 0x10,       /* F.normalize rows first   utils/metrics.py:345 */
  DEMO_FLAG_TRIPLET_NORM = 0x20, /* x/(|x|+1e-12) first      layers/triplet_loss.py:5-13 */
  DEMO_FLAG_SIMT = 0x40          /* FFMA cross-check kernel instead of the tcgen05 GEMM */
};

DEMO_API const char* demo_last_error(void);
DEMO_API int demo_version(void);
/* 1 when a CUDA device with compute capability 10.x is current, else 0 (host-only query). */
DEMO_API int demo_device_ok(void);

/* ---- distance matrix ------------------------------------------------------------------
 * Replaces euclidean_distance(qf, gf) (utils/metrics.py:395-401), euclidean_dist / cosine_dist
 * (layers/triplet_loss.py:16-48) and the all-pairs matrix of re_ranking (utils/reranking.py:36-41).
 * out[Q][G] (ld = ldo).  rowmax (optional, [Q]) receives max_g out[q][g].
 * qn_out / gn_out (optional, [Q][d] / [G][d] contiguous) receive the normalised rows. */
DEMO_API size_t demo_sqdist_workspace_bytes(int Q, int G, int d, int flags);
DEMO_API int demo_sqdist_f32(const float* q, const float* g, int Q, int G, int d, int64_t ldq,
                             int64_t ldg, float* out, int64_t ldo, int flags, float* rowmax,
                             float* qn_out, float* gn_out, void* workspace, size_t workspace_bytes,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEMO_B200_H_ */
