#pragma once
#include "common.cuh"
namespace mfr {
int knn2_l2(const float* q, int nq, const float* t, int nt, int dim, int* idx1, float* dist1, int* idx2,
            float* dist2, cudaStream_t st);
}
