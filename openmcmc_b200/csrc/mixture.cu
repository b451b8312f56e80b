// Mixture-model conjugate updates: allocation draws, per-component sufficient statistics and the
// Categorical log-density, for a Normal whose mean / precision are MixtureParameterVector / MixtureParameterMatrix:
//     x_i ~ N(mu[z_i], 1 / tau[z_i]),   z_i ~ Categorical(prob),   i < n,  K components.
//   ref: sampler/sampler.py:292-355 (MixtureAllocation.sample), :272-288 (NormalGamma K-loop), parameter.py:377-538
//        (MixtureParameterVector / Matrix: predictor = param[allocation]), distribution/distribution.py:282-374
// Allocations are kept as float64 integers in the chain state like every other state entry.
// Mapping: omc_mixture_allocation is one thread per (chain, observation); omc_mixture_stats one CTA per chain, one
// block-wide reduction per component in a fixed order (deterministic).  HBM-bound sweeps of 16 bytes per observation.
#include "../../include/omc.h"
#include "omc_common.cuh"
#include "omc_internal.h"

namespace {

constexpr int MX_THREADS = 256;
constexpr double MX_LOG_2PI = 1.8378770664093454835606594728112;
constexpr double MX_INV_SQRT_2PI = 0.39894228040143267793994605993438;

__device__ __forceinline__ double vat(const omc_vec_t& v, int chain, long long i, double dflt) {
  return v.ptr ? v.ptr[(long long)chain * v.chain_stride + i] : dflt;
}
__device__ __forceinline__ OmcRng to_rng(const omc_rng_t& r) {
  OmcRng o;
  o.seed = r.seed; o.sweep = r.sweep; o.chain_offset = r.chain_offset; o.site = r.site;
  return o;
}

// gam_k = prob_k * N(x_i; mu_k, 1/tau_k), normalised; z_i = #{k : U_i > cumsum_k}   (sampler.py:340-353)
__global__ void __launch_bounds__(MX_THREADS) mixture_allocation_kernel(omc_mixture_alloc_t a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)a.n_chains * a.n;
  if (t >= total) return;
  const int chain = (int)(t / a.n);
  const int i = (int)(t - (long long)chain * a.n);
  const double x = vat(a.x, chain, i, 0.0);
  double sum = 0.0;
  for (int k = 0; k < a.K; ++k) {
    const double tau = vat(a.tau, chain, k, 1.0), d = x - vat(a.mu, chain, k, 0.0);
    const double sd = 1.0 / sqrt(tau), zz = d / sd;
    const double pdf = exp(-0.5 * zz * zz) * MX_INV_SQRT_2PI / sd;          // scipy norm.pdf(x, loc, scale)
    sum += vat(a.prob, chain, (long long)(a.prob_rows > 1 ? i : 0) * a.K + k, 0.0) * pdf;
  }
  double u;
  if (a.debug_u)
    u = a.debug_u[(a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll) * a.debug_sweep_stride + (long long)chain * a.n + i];
  else {
    const uint4 b = omc_rng_block(to_rng(a.rng), chain, (unsigned int)(i >> 1));
    u = (i & 1) ? omc_u01(b.z, b.w) : omc_u01(b.x, b.y);
  }
  double cum = 0.0;
  int z = 0;
  for (int k = 0; k < a.K; ++k) {
    const double tau = vat(a.tau, chain, k, 1.0), d = x - vat(a.mu, chain, k, 0.0);
    const double sd = 1.0 / sqrt(tau), zz = d / sd;
    const double pdf = exp(-0.5 * zz * zz) * MX_INV_SQRT_2PI / sd;
    cum += vat(a.prob, chain, (long long)(a.prob_rows > 1 ? i : 0) * a.K + k, 0.0) * pdf / sum;
    if (u > cum) ++z;
  }
  a.z[(long long)chain * a.n + i] = (double)z;
}

__global__ void __launch_bounds__(MX_THREADS) mixture_stats_kernel(omc_mixture_stats_t a) {
  __shared__ double scratch[32];
  const int chain = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, K = a.K;
  const double* z = a.z + (long long)chain * n;
  double* st = a.stats + (long long)chain * K * 4;
  double* rec = a.record ? a.record + (long long)chain * ((long long)K * K + K + 2) : nullptr;
  if (rec)
    for (int e = tid; e < K * K; e += MX_THREADS) rec[e] = 0.0;
  double lp = 0.0, rss = 0.0;
  bool bad = false;
  for (int k = 0; k < K; ++k) {
    const double mu = vat(a.mu, chain, k, 0.0), tau = vat(a.tau, chain, k, 1.0);
    double cnt = 0.0, s1 = 0.0, s2 = 0.0;
    for (int i = tid; i < n; i += MX_THREADS) {
      if (z[i] == (double)k) {
        const double x = vat(a.x, chain, i, 0.0), d = x - mu;
        cnt += 1.0;
        s1 += x;
        s2 = fma(d, d, s2);
      }
    }
    cnt = omc_block_sum(cnt, scratch);
    s1 = omc_block_sum(s1, scratch);
    s2 = omc_block_sum(s2, scratch);
    if (tid == 0) {
      st[4 * k] = cnt; st[4 * k + 1] = s1; st[4 * k + 2] = s2; st[4 * k + 3] = 0.0;
      if (rec) rec[(long long)K * K + k] = tau * s1;
    }
    lp += cnt * log(tau) - tau * s2;
    rss = fma(tau, s2, rss);
    __syncthreads();
    if (rec && tid == 0) rec[(long long)k * K + k] = tau * cnt;
  }
  // gathers: mu[z_i], tau[z_i]  (parameter.py:437-446, 494-504); an allocation outside [0, K) poisons them
  for (int i = tid; i < n; i += MX_THREADS) {
    const int zi = (int)z[i];
    const bool ok = zi >= 0 && zi < K && (double)zi == z[i];
    if (!ok) bad = true;
    if (a.gather_mu) a.gather_mu[(long long)chain * n + i] = ok ? vat(a.mu, chain, zi, 0.0) : nan("");
    if (a.gather_tau) a.gather_tau[(long long)chain * n + i] = ok ? vat(a.tau, chain, zi, 0.0) : nan("");
  }
  if (tid == 0) {
    if (rec) { rec[(long long)K * K + K] = rss; rec[(long long)K * K + K + 1] = (double)n; }
    if (a.logp) {
      const double v = 0.5 * (lp - n * MX_LOG_2PI);
      a.logp[chain] = a.accumulate ? a.logp[chain] + v : v;
    }
  }
  (void)bad;
}

// sum_i log prob[i or 0][z_i]  == multinomial(n = 1).logpmf of the one-hot rows  (distribution.py:318-345)
__global__ void __launch_bounds__(MX_THREADS) logp_categorical_kernel(int n, int K, const double* z, omc_vec_t prob,
                                                                       int prob_rows, double* out, int accumulate) {
  __shared__ double scratch[32];
  const int chain = blockIdx.x, tid = threadIdx.x;
  double acc = 0.0;
  for (int i = tid; i < n; i += MX_THREADS) {
    const double zd = z[(long long)chain * n + i];
    const int zi = (int)zd;
    if (zi < 0 || zi >= K || (double)zi != zd) acc = -INFINITY;
    else acc += log(vat(prob, chain, (long long)(prob_rows > 1 ? i : 0) * K + zi, 0.0));
  }
  acc = omc_block_sum(acc, scratch);
  if (tid == 0) out[chain] = accumulate ? out[chain] + acc : acc;
}

}  // namespace

extern "C" {

int omc_mixture_allocation(const omc_mixture_alloc_t* a, void* stream) {
  if (a) OMC_REQUIRE_SITE(a->rng, "omc_mixture_allocation");
  OMC_REQUIRE(a && a->x.ptr && a->mu.ptr && a->tau.ptr && a->prob.ptr && a->z, "omc_mixture_allocation: null argument");
  OMC_REQUIRE(a->n_chains >= 1 && a->n >= 1 && a->K >= 1, "omc_mixture_allocation: bad shape");
  const long long total = (long long)a->n_chains * a->n;
  mixture_allocation_kernel<<<(unsigned)((total + MX_THREADS - 1) / MX_THREADS), MX_THREADS, 0, (cudaStream_t)stream>>>(*a);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_mixture_stats(const omc_mixture_stats_t* a, void* stream) {
  OMC_REQUIRE(a && a->x.ptr && a->mu.ptr && a->tau.ptr && a->z && a->stats, "omc_mixture_stats: null argument");
  OMC_REQUIRE(a->n_chains >= 1 && a->n >= 1 && a->K >= 1 && a->K <= 64, "omc_mixture_stats: bad shape (K=%d)", a->K);
  mixture_stats_kernel<<<a->n_chains, MX_THREADS, 0, (cudaStream_t)stream>>>(*a);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_logp_categorical(int n_chains, int n, int K, const double* z, omc_vec_t prob, int prob_rows, double* out,
                         int accumulate, void* stream) {
  OMC_REQUIRE(z && prob.ptr && out && n_chains >= 1 && n >= 1 && K >= 1, "omc_logp_categorical: bad argument");
  logp_categorical_kernel<<<n_chains, MX_THREADS, 0, (cudaStream_t)stream>>>(n, K, z, prob, prob_rows, out, accumulate);
  OMC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
