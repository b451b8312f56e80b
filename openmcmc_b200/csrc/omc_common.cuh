// Shared device/host helpers for libomc (sm_100a only).
//   * error plumbing behind the C-ABI (include/omc.h)
//   * Philox4x32-10 counter RNG -> uniforms / normals / gammas
//   * warp / block reductions
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>
#include <string>

// ---------------------------------------------------------------- errors
// Thread-local message returned by omc_last_error(); codes follow include/omc.h.
void omc_set_error(const char* fmt, ...);
#define OMC_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      omc_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return (int)_e;                                                                    \
    }                                                                                    \
  } while (0)
#define OMC_REQUIRE(cond, ...)                     \
  do {                                             \
    if (!(cond)) {                                 \
      omc_set_error(__VA_ARGS__);                  \
      return -1;                                   \
    }                                              \
  } while (0)
#define OMC_LAUNCH_CHECK() OMC_CHECK_CUDA(cudaGetLastError())
// the Philox counter gives a site 12 bits (omc_rng_block below)
#define OMC_REQUIRE_SITE(rng, who) OMC_REQUIRE((rng).site < 4096u, "%s: rng.site = %u, sites are 12 bits", who, (rng).site)

// ---------------------------------------------------------------- RNG
// One RNG "site" per sampler in the sweep plan.  Counter layout (128 bit):
//   c0,c1 = sweep counter (read from device memory so a captured CUDA graph replays with fresh draws)
//   c2    = global chain id (chain_offset + local chain)  -> results independent of the GPU count
//   c3    = (site << 20) | block  (block = position of the 128-bit block inside this site's draw)
// Key = 64-bit seed.
struct OmcRng {
  unsigned long long seed;
  const unsigned long long* sweep;  // device pointer, may be null (=> sweep 0)
  unsigned int chain_offset;
  unsigned int site;
};

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned int hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    unsigned int hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ uint4 omc_rng_block(const OmcRng& r, unsigned int chain, unsigned int block) {
  unsigned long long sw = r.sweep ? *r.sweep : 0ull;
  // word 3 holds the site (12 bits: the host wrapper refuses sites >= 4096) and the low 20 bits of the block index;
  // the bits above spill into word 1 (zero for every block < 2^20, so those streams are what they always were)
  uint4 ctr = make_uint4((unsigned int)sw, (unsigned int)(sw >> 32) ^ (block >> 20) * 0x9E3779B9u, r.chain_offset + chain,
                         (r.site << 20) | (block & 0xFFFFFu));
  uint2 key = make_uint2((unsigned int)r.seed, (unsigned int)(r.seed >> 32));
  return philox4x32_10(ctr, key);
}

// 64 random bits -> double in the open interval (0,1), 53-bit resolution.
__device__ __forceinline__ double omc_u01(unsigned int lo, unsigned int hi) {
  unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return ((double)(v >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// Two independent standard normals from one Philox block (Box-Muller in fp64).
__device__ __forceinline__ void omc_normal2(const OmcRng& r, unsigned int chain, unsigned int block, double& z0,
                                            double& z1) {
  uint4 b = omc_rng_block(r, chain, block);
  double u1 = omc_u01(b.x, b.y), u2 = omc_u01(b.z, b.w);
  double rad = sqrt(-2.0 * log(u1));
  double s, c;
  sincospi(2.0 * u2, &s, &c);
  z0 = rad * c;
  z1 = rad * s;
}

// Standard gamma variate Gamma(a, 1): Marsaglia & Tsang (2000) squeeze/rejection, with the a<1 boost
// Gamma(a) = Gamma(a+1) * U^(1/a).  `block0` is the first Philox block this draw may use.
__device__ inline double omc_std_gamma(const OmcRng& r, unsigned int chain, unsigned int block0, double a) {
  if (!(a > 0.0)) return (a == 0.0) ? 0.0 : nan("");
  double boost = 1.0;
  unsigned int blk = block0;
  if (a < 1.0) {
    uint4 b = omc_rng_block(r, chain, blk++);
    boost = exp(log(omc_u01(b.x, b.y)) / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 1 << 16; ++it) {
    double x, unused;
    omc_normal2(r, chain, blk++, x, unused);
    double t = 1.0 + c * x;
    if (t <= 0.0) continue;
    double v = t * t * t;
    uint4 b = omc_rng_block(r, chain, blk++);
    double u = omc_u01(b.x, b.y);
    double x2 = x * x;
    if (u < 1.0 - 0.0331 * x2 * x2) return boost * d * v;
    if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return boost * d * v;
  }
  return boost * d;  // unreachable in practice
}

// ---------------------------------------------------------------- reductions
__device__ __forceinline__ double omc_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ double omc_warp_sum(double v) {
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) v += omc_shfl_xor(v, m);
  return v;
}
// Block-wide sum; `scratch` must hold >= 32 doubles.  Result valid in every thread.
__device__ __forceinline__ double omc_block_sum(double v, double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = omc_warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = (lane < nw) ? scratch[lane] : 0.0;
  t = omc_warp_sum(t);
  return t;
}

// x*log(y) with the scipy.special.xlogy convention 0*log(0) = 0.
__device__ __forceinline__ double omc_xlogy(double x, double y) { return (x == 0.0 && !isnan(y)) ? 0.0 : x * log(y); }

// status bits reported per chain (include/omc.h)
#define OMC_STATUS_NOT_PD 1
#define OMC_STATUS_NAN 2
#define OMC_STATUS_OUT_OF_SUPPORT 4
