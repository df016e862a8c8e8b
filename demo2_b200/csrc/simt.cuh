#pragma once
#include "common.cuh"
namespace demo {
int launch_simt_dist(const float* a, const float* b, int M, int N, int d, long long lda, long long ldb,
                     const float* a_norm, const float* b_norm, float* out, long long ldo, int mode,
                     cudaStream_t stream);
}
