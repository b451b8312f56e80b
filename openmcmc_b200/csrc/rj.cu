// ReversibleJump birth / death step for the Gaussian-kernel basis ("location-scale mixture source") model
// (BASELINE configs[4]): one CTA per chain on a padded, fixed-capacity state.
//
//   ref: sampler/reversible_jump.py — proposal :76-94, birth_proposal :96-146, death_proposal :148-193,
//        matched_birth_transition :195-263, matched_death_transition :265-311, get_move_type :313-333,
//        get_move_probabilities :335-373; accept/reject metropolis_hastings.py:127-173; truncated normal gmrf.py:269-318;
//        the basis is the reference tests' make_basis (tests/test_reversible_jump.py:23-40), which replaces the Python
//        state_birth_function / state_death_function callbacks with a declarative built-in.
//
// State per chain (capacity n_max, the first n entries are live): n, theta[n_max] (knots), omega[n_max] (widths),
// beta[n_max] (coefficients), B[n_data][n_max] (basis, column j = N(X; theta_j, omega_j)).  Model terms in the accept
// ratio: response Normal(y | B beta, (tau_y I)^-1) or the Null response, beta ~ iid N(mu_beta, 1/tau_beta),
// n ~ Poisson(rho), theta ~ U(lo, hi), omega ~ Gamma(a, b).
//
// Matched transitions: the reference solves (S + eps I) G = S[:, cols] with S the Gram matrix of the larger basis
// (eps = 1e-10).  Here G = (I - eps (S + eps I)^-1)[:, cols]: ONE m x m shared-memory buffer holds S, then its in-place
// Gauss-Jordan inverse, then F (whose determinant / solve go through an in-place LU with partial pivoting).  Quirks of
// the reference are kept: the proposal density of the associated parameters is taken at the LAST component of the
// CURRENT state (F8), log(det F) has no abs (negative determinant -> NaN -> reject, Q10), strict accept test.
#include "../../include/omc.h"
#include "omc_common.cuh"
#include "omc_internal.h"
#include "omc_special.cuh"

namespace {

constexpr int RJ_NT = 128;
constexpr int RJ_ROWS = 16;       // data rows per shared-memory chunk
constexpr double RJ_EPS = 1e-10;
constexpr double RJ_LOG_2PI = 1.83787706640934548356;

__device__ __forceinline__ double vat(const omc_vec_t& v, int chain, double dflt) {
  return v.ptr ? v.ptr[(long long)chain * v.chain_stride] : dflt;
}
__device__ __forceinline__ double normpdf(double x, double loc, double scale) {
  const double z = (x - loc) / scale;
  return exp(-0.5 * z * z) / (2.50662827463100050242 * scale);
}
__device__ __forceinline__ double gamma_logpdf(double x, double shape, double rate) {
  const double y = x * rate;
  if (!(y >= 0.0)) return isnan(y) ? y : -INFINITY;
  return omc_xlogy(shape - 1.0, y) - y - lgamma(shape) + log(rate);
}

// CTA-wide loop over the elements (i, c) of a rows x cols block, element e = i * cols + c handled by thread e % RJ_NT in
// the order e, e + RJ_NT, ...; (i, c) advance incrementally, so the loop costs ONE integer division per thread instead
// of one per element (the run-time divisor made e / cols the most expensive instruction of these loops).
template <class F>
__device__ __forceinline__ void rj_for2d(int rows, int cols, F f) {
  const int q = RJ_NT / cols, rr = RJ_NT - q * cols;
  int i = threadIdx.x / cols, c = threadIdx.x - i * cols;
  while (i < rows) {
    f(i, c);
    c += rr;
    i += q;
    if (c >= cols) { c -= cols; ++i; }
  }
}

// In-place inverse of the SPD matrix A (m x m, row stride ld) by Gauss-Jordan without pivoting.  colv: m doubles of
// scratch.  Returns false (uniformly) on a non-positive pivot.
__device__ bool gj_inverse_spd(double* A, int m, int ld, double* colv) {
  const int tid = threadIdx.x;
  for (int p = 0; p < m; ++p) {
    const double piv = A[p * ld + p];
    if (!(piv > 0.0)) return false;
    __syncthreads();
    for (int i = tid; i < m; i += RJ_NT) colv[i] = A[i * ld + p];
    __syncthreads();
    const double ip = 1.0 / piv;
    for (int c = tid; c < m; c += RJ_NT) A[p * ld + c] = (c == p ? 1.0 : A[p * ld + c]) * ip;
    __syncthreads();
    rj_for2d(m, m, [&](int i, int c) {
      if (i != p) A[i * ld + c] = (c == p ? 0.0 : A[i * ld + c]) - colv[i] * A[p * ld + c];
    });
    __syncthreads();
  }
  return true;
}

// In-place LU with partial pivoting of F (m x m, row stride ld); optionally solves F x = rhs (rhs overwritten by x).
// Returns log(det F), NaN when det F <= 0 (np.log(np.linalg.det(F))).  pivs: scratch ints (>= 2).
__device__ double lu_logdet_solve(double* F, int m, int ld, double* rhs, int* pivs) {
  const int tid = threadIdx.x, lane = tid & 31;
  int neg = 0;
  double logabs = 0.0;
  for (int j = 0; j < m; ++j) {
    if (tid < 32) {   // pivot: largest |F[i][j]|, i >= j, lowest index on ties
      double best = -1.0;
      int bi = j;
      for (int i = j + lane; i < m; i += 32) {
        const double v = fabs(F[i * ld + j]);
        if (v > best) { best = v; bi = i; }
      }
      for (int s = 16; s > 0; s >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      if (lane == 0) pivs[0] = bi;
    }
    __syncthreads();
    const int pr = pivs[0];
    if (pr != j) {
      for (int c = tid; c < m; c += RJ_NT) {
        const double t = F[j * ld + c];
        F[j * ld + c] = F[pr * ld + c];
        F[pr * ld + c] = t;
      }
      if (rhs && tid == 0) { const double t = rhs[j]; rhs[j] = rhs[pr]; rhs[pr] = t; }
      neg ^= 1;
    }
    __syncthreads();
    const double d = F[j * ld + j];
    if (d < 0.0) neg ^= 1;
    logabs += log(fabs(d));
    for (int i = j + 1 + tid; i < m; i += RJ_NT) F[i * ld + j] /= d;
    __syncthreads();
    const int w = m - 1 - j;
    if (w > 0)
      rj_for2d(w, w, [&](int ii, int cc) {
        const int i = j + 1 + ii, c = j + 1 + cc;
        F[i * ld + c] -= F[i * ld + j] * F[j * ld + c];
      });
    if (rhs)
      for (int i = j + 1 + tid; i < m; i += RJ_NT) rhs[i] -= F[i * ld + j] * rhs[j];
    __syncthreads();
  }
  if (rhs) {   // back substitution, one warp
    if (tid < 32) {
      for (int j = m - 1; j >= 0; --j) {
        double s = 0.0;
        for (int c = j + 1 + lane; c < m; c += 32) s += F[j * ld + c] * rhs[c];
        s = omc_warp_sum(s);
        if (lane == 0) rhs[j] = (rhs[j] - s) / F[j * ld + j];
        __syncwarp();
      }
    }
    __syncthreads();
  }
  return (neg || isnan(logabs)) ? nan("") : logabs;
}

struct Shared {
  int birth, d, accept, ok;
  double theta_new, omega_new, beta_new, u_trunc, u_accept, lq_f, lq_r;
};

// ld: leading dimension of the shared-memory matrices (>= largest basis handled + 1).  With a size_class scratch array
// the step is launched once per size class (classes fixed by rj_class_kernel from the counts BEFORE the step) and a CTA
// serves its chain only in its class, so that chains with few live components run from a small shared-memory
// footprint (several CTAs per SM) and only the rare large ones pay for the full capacity.
__global__ void __launch_bounds__(RJ_NT) rj_kernel(omc_rj_t a, int ld, int cls) {
  extern __shared__ __align__(16) double sm[];
  __shared__ Shared sh;
  __shared__ double s_red[32];
  __shared__ int s_piv[2];
  const int tid = threadIdx.x;
  const int chain = blockIdx.x;
  const int nd = a.n_data, cap = a.n_max;
  double* A = sm;                         // ld * ld
  double* chunk = A + ld * ld;            // RJ_ROWS * ld
  double* bnew = chunk + RJ_ROWS * ld;    // nd
  const int lv = cap + 1;                 // vectors are kept at full capacity whatever the matrix size class
  double* th = bnew + nd;                 // lv each from here on
  double* om = th + lv;
  double* be = om + lv;                   // current coefficients
  double* bp = be + lv;                   // proposed coefficients (laid out on the LARGER basis)
  double* colv = bp + lv;
  double* snew = colv + lv;               // birth: inner products of the new column with the larger basis (row k of S)
  unsigned short* ptab = reinterpret_cast<unsigned short*>(snew + lv);   // (i, j) of every lower-triangle pair
  const int k = (int)a.n_basis[chain];
  if (cls >= 0 && a.size_class[chain] != cls) return;   // another launch of this step owns the chain
  double* thg = a.theta + (long long)chain * cap;
  double* omg = a.omega + (long long)chain * cap;
  double* beg = a.beta + (long long)chain * cap;
  double* Bg = a.B + (long long)chain * nd * cap;
  const double* yp = a.y.ptr ? a.y.ptr + (long long)chain * a.y.chain_stride : nullptr;
  if (k < 1 || k > cap) {   // the reference raises ValueError for n == 0 (reversible_jump.py:330-331)
    if (tid == 0 && a.status) atomicOr(&a.status[chain], OMC_STATUS_NAN);
    return;
  }
  for (int j = tid; j < lv; j += RJ_NT) {
    th[j] = j < k ? thg[j] : 0.0;
    om[j] = j < k ? omg[j] : 1.0;
    be[j] = j < k ? beg[j] : 0.0;
    bp[j] = 0.0;
  }
  const double shape_w = vat(a.omega_shape, chain, 1.0), rate_w = vat(a.omega_rate, chain, 1.0);
  if (a.logp_only) {   // model log-density of the current state (the per-iteration log_post of mcmc.py:108)
    __syncthreads();
    double rss = 0.0;
    if (yp)
      for (int r_ = tid; r_ < nd; r_ += RJ_NT) {
        double f = 0.0;
        const double* row = Bg + (long long)r_ * cap;
        for (int j = 0; j < k; ++j) f = fma(row[j], be[j], f);
        const double q = yp[r_] - f;
        rss = fma(q, q, rss);
      }
    rss = omc_block_sum(rss, s_red);
    if (tid == 0) {
      const double tau_y = vat(a.tau_y, chain, 1.0), tau_b = vat(a.tau_beta, chain, 1.0), mu_b = vat(a.mu_beta, chain, 0.0);
      const double rho = vat(a.rho, chain, 1.0);
      double ss = 0.0, lw = 0.0;
      for (int j = 0; j < k; ++j) {
        ss += (be[j] - mu_b) * (be[j] - mu_b);
        if (a.sample_omega) lw += gamma_logpdf(om[j], shape_w, rate_w);
      }
      double lp = lw;
      if (yp) lp += 0.5 * (nd * log(tau_y) - nd * RJ_LOG_2PI - tau_y * rss);
      lp += 0.5 * (k * log(tau_b) - k * RJ_LOG_2PI - tau_b * ss);
      lp += omc_xlogy((double)k, rho) - lgamma(k + 1.0) - rho;
      lp += -k * log(a.theta_hi - a.theta_lo);
      a.logp_out[chain] = lp;
    }
    return;
  }
  // ---- move type and the variates of this step (thread 0).  Draw order of the reference: move uniform
  //      (skipped at n = 1 / n_max), then birth: knot uniform, width gamma, coefficient | death: index; accept uniform.
  if (tid == 0) {
    const double* dbg = nullptr;
    if (a.debug) {
      const long long sw = a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll;
      dbg = a.debug + sw * a.debug_sweep_stride + (long long)chain * 6;
    }
    OmcRng r;
    r.seed = a.rng.seed; r.sweep = a.rng.sweep; r.chain_offset = a.rng.chain_offset; r.site = a.rng.site;
    const uint4 b0 = omc_rng_block(r, chain, 0), b1 = omc_rng_block(r, chain, 1);
    const double u_move = dbg ? dbg[0] : omc_u01(b0.x, b0.y);
    int birth;
    if (k == cap) birth = 0;
    else if (k == 1) birth = 1;
    else birth = (u_move <= a.birth_probability) ? 1 : 0;
    sh.birth = birth;
    sh.d = -1;
    sh.theta_new = sh.omega_new = sh.beta_new = 0.0;
    sh.u_trunc = -1.0;
    if (birth) {
      sh.theta_new = dbg ? dbg[1] : a.theta_lo + (a.theta_hi - a.theta_lo) * omc_u01(b0.z, b0.w);
      if (a.sample_omega) sh.omega_new = dbg ? dbg[2] : omc_std_gamma(r, chain, 8, shape_w) / rate_w;
      else sh.omega_new = omg[k - 1];
      if (dbg) sh.beta_new = dbg[3];          // final value of the reference's truncnorm / normal draw (NaN: the mean)
      else sh.u_trunc = omc_u01(b1.x, b1.y);   // uniform behind the truncated normal; z below for the plain normal
    } else {
      int d = dbg ? (int)dbg[4] : (int)(omc_u01(b0.z, b0.w) * k);
      sh.d = min(max(d, 0), k - 1);
    }
    sh.u_accept = dbg ? dbg[5] : omc_u01(b1.z, b1.w);
    sh.ok = 1;
  }
  __syncthreads();
  const int birth = sh.birth, d = sh.d;
  const int m = birth ? k + 1 : k;          // size of the LARGER basis
  if (birth) {
    for (int r_ = tid; r_ < nd; r_ += RJ_NT) bnew[r_] = normpdf(a.X[r_], sh.theta_new, sh.omega_new);
    if (tid == 0) { th[k] = sh.theta_new; om[k] = sh.omega_new; }
  }
  for (int e = tid; e < m * ld; e += RJ_NT) A[e] = 0.0;
  // pair index -> (i, j) of the lower triangle, decoded once per step (it was a double-precision square root per pair
  // and per 16-row chunk)
  const int npair = m * (m + 1) / 2;
  for (int pi = tid; pi < npair; pi += RJ_NT) {
    int i = (int)((sqrt(8.0 * pi + 1.0) - 1.0) * 0.5);
    while ((i + 1) * (i + 2) / 2 <= pi) ++i;
    while (i * (i + 1) / 2 > pi) --i;
    ptab[pi] = (unsigned short)((i << 8) | (pi - i * (i + 1) / 2));
  }
  __syncthreads();
  // ---- pass 1 over the data rows: Gram matrix of the larger basis (lower triangle) and the current residual sum.
  //      With the live Gram matrix in the chain state (a.gram, valid) only the NEW column's inner products are formed:
  //      a warp per data row, lanes over the columns (coalesced), the row's fitted value by a shuffle sum.
  double rss_c = 0.0;
  double* Sg = a.gram ? a.gram + (long long)chain * cap * cap : nullptr;
  const bool gram_on = Sg && a.gram_valid && cap <= 128;      // (the new-row pass covers 4 x 32 columns)
  const bool have_gram = gram_on && a.gram_valid[chain] != 0;
  if (have_gram) {
    for (int pi = tid; pi < k * (k + 1) / 2; pi += RJ_NT) {
      const int i = ptab[pi] >> 8, j = ptab[pi] & 255;
      A[i * ld + j] = Sg[(long long)i * cap + j];
    }
    const int lane = tid & 31, warp = tid >> 5;
    constexpr int NWARP = RJ_NT / 32;
    double part[4] = {0.0, 0.0, 0.0, 0.0};           // this lane's columns lane, lane + 32, ... of the new row of S
    double nn = 0.0;
    // RB rows per warp and round: their loads go out together (one row at a time left 4 KB in flight per SM and the
    // pass bound by HBM latency: 3.6 ms per sweep at the C5 shape, ncu long-scoreboard 5.6 per issue)
    constexpr int RB = 8;
    for (int r0 = warp * RB; r0 < nd; r0 += NWARP * RB) {
      double v[RB][4], yv[RB];
#pragma unroll
      for (int t = 0; t < RB; ++t) {
        const int r_ = r0 + t;
        const double* row = Bg + (long long)min(r_, nd - 1) * cap;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = lane + 32 * q;
          v[t][q] = (j < k && r_ < nd) ? row[j] : 0.0;
        }
        yv[t] = (yp && r_ < nd) ? yp[r_] : 0.0;
      }
#pragma unroll
      for (int t = 0; t < RB; ++t) {
        const int r_ = r0 + t;
        if (r_ >= nd) break;
        const double bn = birth ? bnew[r_] : 0.0;
        double f = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = lane + 32 * q;
          if (j < k) {
            f = fma(v[t][q], be[j], f);
            part[q] = fma(bn, v[t][q], part[q]);
          }
        }
        if (yp) {
          f = omc_warp_sum(f);
          const double q_ = yv[t] - f;
          if (lane == 0) rss_c = fma(q_, q_, rss_c);
        }
        if (lane == 0) nn = fma(bn, bn, nn);
      }
    }
    if (birth) {                                     // combine the warps in a fixed order (chunk is free in this path)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (lane + 32 * q < k) chunk[warp * lv + lane + 32 * q] = part[q];
      if (lane == 0) chunk[warp * lv + k] = nn;
      __syncthreads();
      for (int j = tid; j <= k; j += RJ_NT) {
        double sacc = 0.0;
        for (int w_ = 0; w_ < NWARP; ++w_) sacc += chunk[w_ * lv + j];
        snew[j] = sacc;
        A[k * ld + j] = sacc;
      }
    }
    __syncthreads();
  } else {
    for (int r0 = 0; r0 < nd; r0 += RJ_ROWS) {
      const int rows = min(RJ_ROWS, nd - r0);
      rj_for2d(rows, m, [&](int r_, int j) {
        chunk[r_ * ld + j] = (j < k) ? Bg[(long long)(r0 + r_) * cap + j] : bnew[r0 + r_];
      });
      __syncthreads();
      for (int pi = tid; pi < npair; pi += RJ_NT) {
        const int i = ptab[pi] >> 8, j = ptab[pi] & 255;
        double s = 0.0;
        for (int r_ = 0; r_ < rows; ++r_) s = fma(chunk[r_ * ld + i], chunk[r_ * ld + j], s);
        A[i * ld + j] += s;
      }
      if (yp && tid < rows) {
        double f = 0.0;
        for (int j = 0; j < k; ++j) f = fma(chunk[tid * ld + j], be[j], f);
        const double q = yp[r0 + tid] - f;
        rss_c = fma(q, q, rss_c);
      }
      __syncthreads();
    }
    if (gram_on) {                                   // first step after the basis was rewritten: keep S from here on
      for (int pi = tid; pi < k * (k + 1) / 2; pi += RJ_NT) {
        const int i = ptab[pi] >> 8, j = ptab[pi] & 255;
        Sg[(long long)i * cap + j] = A[i * ld + j];
      }
      if (birth)
        for (int j = tid; j <= k; j += RJ_NT) snew[j] = A[k * ld + j];
      if (tid == 0) a.gram_valid[chain] = 1;
      __syncthreads();
    }
  }
  for (int pi = tid; pi < npair; pi += RJ_NT) {   // symmetrise, add the ridge
    const int i = ptab[pi] >> 8, j = ptab[pi] & 255;
    if (i == j) A[i * ld + i] += RJ_EPS;
    else A[j * ld + i] = A[i * ld + j];
  }
  __syncthreads();
  // ---- Z = (S + eps I)^-1 in place;  G = I - eps Z
  const bool pd = gj_inverse_spd(A, m, ld, colv);
  if (!pd) {
    if (tid == 0 && a.status) atomicOr(&a.status[chain], OMC_STATUS_NOT_PD);
    if (tid == 0 && a.counters) a.counters[(long long)chain * 2 + 1] += 1;
    return;
  }
  rj_for2d(m, m, [&](int i, int c) { A[i * ld + c] = (i == c ? 1.0 : 0.0) - RJ_EPS * A[i * ld + c]; });
  __syncthreads();
  double logdetF;
  if (birth) {
    // mu* = G[:, :k] beta  (matched_birth_transition :243-244)
    for (int i = tid; i < m; i += RJ_NT) {
      double s = 0.0;
      for (int j = 0; j < k; ++j) s = fma(A[i * ld + j], be[j], s);
      bp[i] = s;
    }
    __syncthreads();
    if (tid == 0) {
      const double mu_new = bp[k];
      double x;
      if (sh.u_trunc >= 0.0) {
        if (a.match_truncated) x = omc_truncated_normal_rv(mu_new, a.match_scale, a.match_lo, a.match_hi, sh.u_trunc);
        else x = mu_new + a.match_scale * normcdfinv(sh.u_trunc);
      } else {
        x = isnan(sh.beta_new) ? mu_new : sh.beta_new;
      }
      bp[k] = x;
      if (a.match_truncated) sh.lq_f = omc_truncated_normal_log_pdf(x, mu_new, a.match_scale, a.match_lo, a.match_hi);
      else {
        const double z = (x - mu_new) / a.match_scale;
        sh.lq_f = -0.5 * z * z - log(a.match_scale) - 0.5 * RJ_LOG_2PI;
      }
    }
    // det F = det G[:k, :k]  (F = [G | e_last], :259)
    logdetF = lu_logdet_solve(A, k, ld, nullptr, s_piv);
  } else {
    // F = G with column d replaced by e_d (np.insert, :291);  F mu_aug = beta (:294)
    for (int i = tid; i < m; i += RJ_NT) {
      A[i * ld + d] = (i == d) ? 1.0 : 0.0;
      bp[i] = be[i];
    }
    __syncthreads();
    logdetF = lu_logdet_solve(A, m, ld, bp, s_piv);
    if (tid == 0) {
      const double param_del = bp[d];
      if (a.match_truncated) sh.lq_r = omc_truncated_normal_log_pdf(param_del, 0.0, a.match_scale, a.match_lo, a.match_hi);
      else {
        const double z = param_del / a.match_scale;
        sh.lq_r = -0.5 * z * z - log(a.match_scale) - 0.5 * RJ_LOG_2PI;
      }
      bp[d] = 0.0;   // the deleted component carries no weight in the proposed fit
    }
  }
  __syncthreads();
  // ---- pass 2 over the data rows: residual sum of the proposed state (coefficients bp on the larger basis)
  double rss_p = 0.0;
  if (yp) {
    for (int r_ = tid; r_ < nd; r_ += RJ_NT) {
      double f = birth ? bnew[r_] * bp[k] : 0.0;
      const double* row = Bg + (long long)r_ * cap;
      for (int j = 0; j < k; ++j) f = fma(row[j], bp[j], f);
      const double q = yp[r_] - f;
      rss_p = fma(q, q, rss_p);
    }
  }
  rss_c = omc_block_sum(rss_c, s_red);
  __syncthreads();
  rss_p = omc_block_sum(rss_p, s_red);
  // ---- prior terms of both states, one component per thread (they were a serial loop of thread 0 with two lgamma /
  //      log evaluations per component while the other 127 threads waited at the barrier below: 13 % of the stall
  //      samples of the step in profiles/r02_ncu_rj_v1.txt; r02_ncu_rj_lines.txt is the kernel as shipped)
  const double mu_b = vat(a.mu_beta, chain, 0.0);
  double ss_c = 0.0, ss_p = 0.0, lw_c = 0.0, lw_p = 0.0;
  for (int j = tid; j < m; j += RJ_NT) {
    const double lw = a.sample_omega ? gamma_logpdf(om[j], shape_w, rate_w) : 0.0;
    if (j < k) {
      ss_c += (be[j] - mu_b) * (be[j] - mu_b);
      lw_c += lw;
    }
    if (birth || j != d) {
      ss_p += (bp[j] - mu_b) * (bp[j] - mu_b);
      lw_p += lw;
    }
  }
  __syncthreads();
  ss_c = omc_block_sum(ss_c, s_red);
  __syncthreads();
  ss_p = omc_block_sum(ss_p, s_red);
  __syncthreads();
  lw_c = omc_block_sum(lw_c, s_red);
  __syncthreads();
  lw_p = omc_block_sum(lw_p, s_red);
  // ---- log-densities of the whole model at both states, transition densities, accept / reject (thread 0)
  if (tid == 0) {
    const double tau_y = vat(a.tau_y, chain, 1.0), tau_b = vat(a.tau_beta, chain, 1.0);
    const double rho = vat(a.rho, chain, 1.0);
    const int kp = birth ? k + 1 : k - 1;
    auto logp = [&](int n, double rss, double ss, double lw) {
      double lp = 0.0;
      if (yp) lp += 0.5 * (nd * log(tau_y) - nd * RJ_LOG_2PI - tau_y * rss);
      lp += 0.5 * (n * log(tau_b) - n * RJ_LOG_2PI - tau_b * ss);
      lp += omc_xlogy((double)n, rho) - lgamma(n + 1.0) - rho;
      lp += -n * log(a.theta_hi - a.theta_lo);
      return lp + lw;
    };
    const double lp_c = logp(k, rss_c, ss_c, lw_c), lp_p = logp(kp, rss_p, ss_p, lw_p);
    // proposal density of the associated parameters: LAST component of the CURRENT state (F8)
    double lpd_last = -log(a.theta_hi - a.theta_lo);
    if (a.sample_omega) lpd_last += gamma_logpdf(om[k - 1], shape_w, rate_w);
    double p_birth = a.birth_probability, p_death = 1.0 - a.birth_probability;   // :361-373
    if (k == cap) p_death = 1.0;
    if (k == cap - 1 && birth) p_death = 1.0;
    if (k == 1) p_birth = 1.0;
    if (k == 2 && !birth) p_birth = 1.0;
    double lq_f, lq_r;
    if (birth) {
      lq_f = sh.lq_f + log(p_birth) + lpd_last;
      lq_r = logdetF + log(p_death);
    } else {
      lq_f = logdetF + log(p_death);
      lq_r = sh.lq_r + log(p_birth) + lpd_last;
    }
    const double log_accept = lp_p + lq_r - (lp_c + lq_f);
    const int acc = (log(sh.u_accept) < log_accept) ? 1 : 0;   // strict; NaN rejects
    sh.accept = acc;
    if (a.counters) {
      a.counters[(long long)chain * 2 + 1] += 1;
      a.counters[(long long)chain * 2] += acc;
    }
    if (a.probe) {
      double* o = a.probe + (long long)chain * 8;
      o[0] = birth; o[1] = d; o[2] = lp_c; o[3] = lp_p; o[4] = lq_f; o[5] = lq_r; o[6] = log_accept; o[7] = acc;
    }
  }
  __syncthreads();
  if (!sh.accept) return;
  // ---- accepted: write the new state (padded layout)
  const bool keep_gram = gram_on;
  if (birth) {
    if (keep_gram)
      for (int j = tid; j <= k; j += RJ_NT) Sg[(long long)k * cap + j] = snew[j];     // new row of S
    for (int j = tid; j <= k; j += RJ_NT) beg[j] = bp[j];
    for (int r_ = tid; r_ < nd; r_ += RJ_NT) Bg[(long long)r_ * cap + k] = bnew[r_];
    if (tid == 0) {
      thg[k] = th[k];
      omg[k] = om[k];
      a.n_basis[chain] = (double)(k + 1);
    }
  } else {
    for (int j = tid; j < k - 1; j += RJ_NT) {
      const int src = j < d ? j : j + 1;
      beg[j] = bp[src];
      thg[j] = th[src];
      omg[j] = om[src];
    }
    for (int r_ = tid; r_ < nd; r_ += RJ_NT) {
      double* row = Bg + (long long)r_ * cap;
      for (int j = d; j < k - 1; ++j) row[j] = row[j + 1];
    }
    if (keep_gram) {     // delete row / column d of S: rows >= d through shared memory (A is free now), then back
      const int npk = k * (k + 1) / 2;
      for (int pi = tid; pi < npk; pi += RJ_NT) {
        const int i = ptab[pi] >> 8, j = ptab[pi] & 255;
        if (i >= d) A[i * ld + j] = Sg[(long long)i * cap + j];
      }
      __syncthreads();
      for (int pi = tid; pi < npk; pi += RJ_NT) {
        const int i = ptab[pi] >> 8, j = ptab[pi] & 255;       // (i, j) of the NEW matrix, i < k - 1
        if (i >= d && i < k - 1) {
          const int si = i + 1, sj = j < d ? j : j + 1;
          Sg[(long long)i * cap + j] = A[si * ld + sj];
        }
      }
    }
    if (tid == 0) a.n_basis[chain] = (double)(k - 1);
  }
}

__global__ void rj_class_kernel(const double* n_basis, int n_chains, int* size_class) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  const int k = (int)n_basis[c];
  size_class[c] = (k <= 31) ? 0 : (k <= 63) ? 1 : 2;   // invalid counts fall into class 0, which reports them
}

// B[c][r][j] = N(X_r; theta_cj, omega_cj) for the live columns  (make_basis; also the state_update_function of the
// RandomWalkLoop samplers on theta / omega in the reference's test model)
__global__ void rj_basis_kernel(omc_rj_t a) {
  const int chain = blockIdx.y;
  const int k = (int)a.n_basis[chain];
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long long)a.n_data * a.n_max) return;
  const int r_ = (int)(e / a.n_max), j = (int)(e % a.n_max);
  double v = 0.0;
  if (j < k) v = normpdf(a.X[r_], a.theta[(long long)chain * a.n_max + j], a.omega[(long long)chain * a.n_max + j]);
  a.B[((long long)chain * a.n_data + r_) * a.n_max + j] = v;
  if (e == 0 && a.gram_valid) a.gram_valid[chain] = 0;          // the basis is rewritten: the live Gram matrix is stale
}

int rj_check(const omc_rj_t* a, const char* who) {
  OMC_REQUIRE(a && a->n_basis && a->theta && a->omega && a->beta && a->B && a->X, "%s: null argument", who);
  OMC_REQUIRE(a->n_chains >= 1 && a->n_data >= 1 && a->n_max >= 2, "%s: bad shape", who);
  OMC_REQUIRE(a->theta_hi > a->theta_lo, "%s: empty knot domain", who);
  return 0;
}

}  // namespace

extern "C" {

static int rj_smem_for_ld(int n_data, int n_max, int ld) {
  const int pair_table_doubles = (ld * (ld + 1) / 2 * 2 + 7) / 8;   // 16-bit (i, j) per lower-triangle pair
  return (ld * ld + RJ_ROWS * ld + n_data + 6 * (n_max + 1) + pair_table_doubles) * 8;
}

int omc_rj_smem_bytes(int n_data, int n_max) { return rj_smem_for_ld(n_data, n_max, n_max + 1); }

int omc_reversible_jump(const omc_rj_t* a, void* stream) {
  if (a) OMC_REQUIRE_SITE(a->rng, "omc_reversible_jump");
  if (int rc = rj_check(a, "omc_reversible_jump")) return rc;
  OMC_REQUIRE(a->birth_probability >= 0.0 && a->birth_probability <= 1.0, "omc_reversible_jump: birth_probability");
  OMC_REQUIRE(a->logp_only || a->match_scale > 0.0, "omc_reversible_jump: match_scale");
  OMC_REQUIRE(!a->logp_only || a->logp_out, "omc_reversible_jump: logp_out missing");
  const int smem = omc_rj_smem_bytes(a->n_data, a->n_max);
  OMC_REQUIRE(smem <= 220 * 1024, "omc_reversible_jump: n_max=%d, n_data=%d need %d bytes of shared memory", a->n_max,
              a->n_data, smem);
  OMC_CHECK_CUDA(cudaFuncSetAttribute(rj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaStream_t st = (cudaStream_t)stream;
  if (a->logp_only) {   // no matrices needed: vectors only
    rj_kernel<<<a->n_chains, RJ_NT, rj_smem_for_ld(a->n_data, a->n_max, 1), st>>>(*a, 1, -1);
    OMC_LAUNCH_CHECK();
    return 0;
  }
  if (!a->size_class || a->n_max <= 32) {
    rj_kernel<<<a->n_chains, RJ_NT, smem, st>>>(*a, a->n_max + 1, -1);
    OMC_LAUNCH_CHECK();
    return 0;
  }
  // size classes by live components BEFORE the step: n <= 31 (ld 33), n <= 63 (ld 65), the rest at full capacity
  rj_class_kernel<<<(a->n_chains + 255) / 256, 256, 0, st>>>(a->n_basis, a->n_chains, a->size_class);
  OMC_LAUNCH_CHECK();
  const int hi[3] = {31, 63, a->n_max};
  for (int q = 0; q < 3; ++q) {
    const int top = hi[q] < a->n_max ? hi[q] : a->n_max;
    const int ld = (top + 1 < a->n_max ? top + 1 : a->n_max) + 1;   // a birth from n = top needs top + 1 columns
    rj_kernel<<<a->n_chains, RJ_NT, rj_smem_for_ld(a->n_data, a->n_max, ld), st>>>(*a, ld, q);
    OMC_LAUNCH_CHECK();
    if (top >= a->n_max) break;
  }
  return 0;
}

int omc_rj_basis(const omc_rj_t* a, void* stream) {
  if (int rc = rj_check(a, "omc_rj_basis")) return rc;
  const long long per = (long long)a->n_data * a->n_max;
  dim3 grid((unsigned)((per + 255) / 256), a->n_chains);
  rj_basis_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*a);
  OMC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
