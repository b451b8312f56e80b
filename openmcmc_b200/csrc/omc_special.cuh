// Device special functions behind the truncated-normal proposals.
// The reference reaches these through scipy.stats.truncnorm (gmrf.py:269-318): inverse-CDF sampling on ONE uniform with
// log-space mass computations (_log_gauss_mass, _truncnorm ppf; scipy/stats/_continuous_distns.py:10219-10404) on top of
// scipy.special.log_ndtr / ndtr / ndtri_exp.  The functions below restate that algorithm with CUDA's erfc / erfcx /
// normcdfinv; agreement with scipy is checked on the GPU against committed golden vectors (tests/golden/truncnorm.npz).
#pragma once
#include "omc_common.cuh"

#define OMC_SQRT1_2 0.70710678118654752440
#define OMC_LOG_SQRT_2PI 0.91893853320467274178

__device__ __forceinline__ double omc_ndtr(double x) { return 0.5 * erfc(-x * OMC_SQRT1_2); }

// log Phi(x), accurate in both tails (erfcx keeps the left tail out of underflow).
__device__ __forceinline__ double omc_log_ndtr(double x) {
  if (x > 0.0) return log1p(-0.5 * erfc(x * OMC_SQRT1_2));
  const double t = -x * OMC_SQRT1_2;
  return log(0.5 * erfcx(t)) - t * t;
}

// inverse of log_ndtr (scipy.special.ndtri_exp): same three branches as scipy's _ndtri_exp.pxd
__device__ inline double omc_ndtri_exp(double y) {
  if (!(y <= 0.0)) return (y == 0.0) ? INFINITY : nan("");
  if (y < -2.0) {
    // far left tail: asymptotic start, then Newton on f(x) = log_ndtr(x) - y  (f' = phi/Phi)
    const double s = sqrt(-2.0 * y);
    double x = -(s - log(s) / s);  // x ~ -sqrt(-2y) corrected
    for (int it = 0; it < 4; ++it) {
      const double l = omc_log_ndtr(x);
      const double log_ratio = l - (-0.5 * x * x - OMC_LOG_SQRT_2PI);  // log(Phi/phi)
      x -= (l - y) * exp(log_ratio);
    }
    return x;
  }
  if (y > log1p(-exp(-2.0))) return -normcdfinv(-expm1(y));
  return normcdfinv(exp(y));
}

__device__ __forceinline__ double omc_log_diff(double log_p, double log_q) {  // log(exp(log_p) - exp(log_q))
  return log_p + log1p(-exp(log_q - log_p));
}
__device__ __forceinline__ double omc_logaddexp(double a, double b) {
  if (a == b) return a + 0.69314718055994530942;
  const double m = fmax(a, b), d = -fabs(a - b);
  return isinf(m) && m < 0 ? m : m + log1p(exp(d));
}

// log(Phi(b) - Phi(a)), evaluated in the tail that keeps precision (scipy _log_gauss_mass)
__device__ inline double omc_log_gauss_mass(double a, double b) {
  if (b <= 0.0) return omc_log_diff(omc_log_ndtr(b), omc_log_ndtr(a));
  if (a > 0.0) return omc_log_diff(omc_log_ndtr(-a), omc_log_ndtr(-b));
  return log1p(-omc_ndtr(a) - omc_ndtr(-b));
}

// standard normal truncated to [a, b]: inverse CDF at q (scipy truncnorm_gen._ppf)
__device__ inline double omc_truncnorm_ppf(double q, double a, double b) {
  const double lm = omc_log_gauss_mass(a, b);
  if (a < 0.0) return omc_ndtri_exp(omc_logaddexp(omc_log_ndtr(a), log(q) + lm));
  return -omc_ndtri_exp(omc_logaddexp(omc_log_ndtr(-b), log1p(-q) + lm));
}

// ref: gmrf.py:269-292  truncated_normal_rv(mean, scale, lower, upper) with the uniform u behind truncnorm.rvs
__device__ inline double omc_truncated_normal_rv(double mean, double scale, double lower, double upper, double u) {
  const double a = (lower - mean) / scale, b = (upper - mean) / scale;
  return omc_truncnorm_ppf(u, a, b) * scale + mean;
}
// ref: gmrf.py:295-318  truncated_normal_log_pdf
__device__ inline double omc_truncated_normal_log_pdf(double x, double mean, double scale, double lower, double upper) {
  const double a = (lower - mean) / scale, b = (upper - mean) / scale;
  const double z = (x - mean) / scale;
  if (z < a || z > b) return -INFINITY;
  return -0.5 * z * z - OMC_LOG_SQRT_2PI - omc_log_gauss_mass(a, b) - log(scale);
}
