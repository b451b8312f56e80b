// Block-cooperative small dense linear algebra in shared memory (p <= 64): Cholesky and triangular solves.
// Shared by the NormalNormal draw (dense_draw.cu) and the manifold-MALA proposal (mh.cu).
//   ref: gmrf.py:465-486 (cholesky), :437-462 (cho_solve), :414-434 (solve on L.T), :29-61 (sample_normal)
#pragma once
#include "omc_common.cuh"

// In-place right-looking Cholesky of the symmetric p x p matrix Q (row stride ld) -> lower factor L in the lower
// triangle.  All threads of the CTA must call it.  Returns false (uniformly) if a pivot is <= 0 or NaN.
__device__ __forceinline__ bool omc_chol_block(double* Q, int p, int ld) {
  const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  for (int j = 0; j < p; ++j) {
    const double djj = Q[j * ld + j];
    if (!(djj > 0.0)) return false;  // same smem value for every thread -> uniform
    const double d = sqrt(djj);
    __syncthreads();  // everyone has read Q[j][j] before it is overwritten
    for (int i = j + tid; i < p; i += nthr) {
      if (i == j) Q[j * ld + j] = d;
      else Q[i * ld + j] = Q[i * ld + j] / d;
    }
    __syncthreads();
    for (int i = j + 1 + warp; i < p; i += nwarp) {
      const double lij = Q[i * ld + j];
      for (int c = j + 1 + lane; c <= i; c += 32) Q[i * ld + c] -= lij * Q[c * ld + j];
    }
    __syncthreads();
  }
  return true;
}

// The same factorisation by ONE warp (no CTA barriers): lane c owns column c (and c + 32) of the trailing update.  Every
// element sees the same operations in the same order as in omc_chol_block, so the factors are bit-identical.
__device__ __forceinline__ bool omc_chol_warp(double* Q, int p, int ld) {
  const int lane = threadIdx.x & 31;
  for (int j = 0; j < p; ++j) {
    const double djj = Q[j * ld + j];
    if (!(djj > 0.0)) return false;
    const double d = sqrt(djj);
    __syncwarp();
    for (int i = j + lane; i < p; i += 32) {
      if (i == j) Q[j * ld + j] = d;
      else Q[i * ld + j] = Q[i * ld + j] / d;
    }
    __syncwarp();
    for (int c = j + 1 + lane; c < p; c += 32) {
      const double lcj = Q[c * ld + j];
      for (int i = c; i < p; ++i) Q[i * ld + c] -= Q[i * ld + j] * lcj;
    }
    __syncwarp();
  }
  return true;
}

// One warp solves L w = b (forward) in registers: lane owns rows lane and lane+32; x0/x1 in/out.
__device__ __forceinline__ void omc_warp_solve_lower(const double* L, int p, int ld, double& x0, double& x1) {
  const int lane = threadIdx.x & 31;
  for (int j = 0; j < p; ++j) {
    const double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) / L[j * ld + j];
    if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
    if (lane > j && lane < p) x0 -= L[lane * ld + j] * xj;
    if (lane + 32 > j && lane + 32 < p) x1 -= L[(lane + 32) * ld + j] * xj;
  }
}
// One warp solves L' x = b (backward), same register layout.
__device__ __forceinline__ void omc_warp_solve_lower_T(const double* L, int p, int ld, double& x0, double& x1) {
  const int lane = threadIdx.x & 31;
  for (int j = p - 1; j >= 0; --j) {
    const double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) / L[j * ld + j];
    if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
    if (lane < j) x0 -= L[j * ld + lane] * xj;
    if (lane + 32 < j) x1 -= L[j * ld + lane + 32] * xj;
  }
}
// One warp computes w = L' r (r in registers, same layout); result in w0/w1.
__device__ __forceinline__ void omc_warp_mul_lower_T(const double* L, int p, int ld, double r0, double r1, double& w0,
                                                     double& w1) {
  const int lane = threadIdx.x & 31;
  w0 = 0.0;
  w1 = 0.0;
  for (int i = 0; i < p; ++i) {  // (L' r)_c = sum_{i >= c} L[i][c] r_i
    const double ri = __shfl_sync(0xffffffffu, (i < 32) ? r0 : r1, i & 31);
    if (lane <= i && lane < p) w0 += L[i * ld + lane] * ri;
    if (lane + 32 <= i && lane + 32 < p) w1 += L[i * ld + lane + 32] * ri;
  }
}
