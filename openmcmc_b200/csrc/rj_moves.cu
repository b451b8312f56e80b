// Companion moves of the ReversibleJump model on its padded, fixed-capacity state (the other three samplers of the
// reference's RJ model, tests/test_reversible_jump.py:213-252, and of BASELINE configs[4]'s source model):
//
//   omc_rj_knot_walk  : RandomWalkLoop over the knots theta (which = 0) or the widths omega (which = 1), one truncated-
//                       Gaussian MH step per live component with the basis column N(X; theta_j, omega_j) rebuilt for the
//                       proposal -- the reference reaches this through a Python state_update_function (its tests'
//                       move_function = make_basis); here the basis is declared (GaussianKernelBasis).
//                       ref: metropolis_hastings.py:212-289 (proposal with param_index, loop), :127-173 (accept on the
//                       FULL model, which RandomWalk keeps when a state_update_function is given, :201-210),
//                       gmrf.py:269-318 (truncated normal).
//   omc_rj_coef_mmala : ManifoldMALA on the live coefficients beta[0:n] with the conditional model
//                       y ~ N(B beta, (tau_y I)^-1) (or the Null response) and beta ~ iid N(mu_beta, 1/tau_beta):
//                       g = tau_y B'(y - B beta) - tau_beta (beta - mu_beta), H = tau_y B'B + tau_beta I.
//                       ref: metropolis_hastings.py:292-373, location_scale.py:222-250.
//
// One CTA per chain.  The accept ratios use differences of the changed terms (the residual sum of squares, the prior of
// the moved component); the terms that do not change cancel in the reference's full-model sums.
#include "../../include/omc.h"
#include "omc_common.cuh"
#include "omc_internal.h"
#include "omc_smallmat.cuh"
#include "omc_special.cuh"

namespace {

constexpr int RM_NT = 128;
constexpr int RM_ROWS = 16;

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
__device__ __forceinline__ OmcRng to_rng(const omc_rng_t& r) {
  OmcRng o;
  o.seed = r.seed; o.sweep = r.sweep; o.chain_offset = r.chain_offset; o.site = r.site;
  return o;
}

struct WalkShared {
  int accept;
};

__global__ void __launch_bounds__(RM_NT) rj_knot_walk_kernel(omc_rj_walk_t w) {
  extern __shared__ __align__(16) double sm[];
  __shared__ WalkShared sh;
  __shared__ double s_red[32];
  const omc_rj_t& a = w.model;
  const int tid = threadIdx.x, chain = blockIdx.x;
  const int nd = a.n_data, cap = a.n_max;
  double* r = sm;              // nd : current residual y - B beta
  double* cnew = r + nd;       // nd : proposed basis column
  double* th = cnew + nd;      // cap
  double* om = th + cap;       // cap
  double* be = om + cap;       // cap
  double* pz = be + cap;       // cap : proposals of all components (a component's proposal depends on its own value only)
  double* plq = pz + cap;      // cap : lq_rev - lq_fwd + prior difference of the moved component
  double* pu = plq + cap;      // cap : accept uniforms
  const int k = (int)a.n_basis[chain];
  if (k < 1 || k > cap) {
    if (tid == 0 && a.status) atomicOr(&a.status[chain], OMC_STATUS_NAN);
    return;
  }
  double* thg = a.theta + (long long)chain * cap;
  double* omg = a.omega + (long long)chain * cap;
  const double* beg = a.beta + (long long)chain * cap;
  double* Bg = a.B + (long long)chain * nd * cap;
  const double* yp = a.y.ptr ? a.y.ptr + (long long)chain * a.y.chain_stride : nullptr;
  for (int j = tid; j < cap; j += RM_NT) {
    th[j] = j < k ? thg[j] : 0.0;
    om[j] = j < k ? omg[j] : 1.0;
    be[j] = j < k ? beg[j] : 0.0;
  }
  __syncthreads();
  double rss = 0.0;
  if (yp) {
    double part = 0.0;
    for (int i = tid; i < nd; i += RM_NT) {
      const double* row = Bg + (long long)i * cap;
      double f = 0.0;
      for (int j = 0; j < k; ++j) f = fma(row[j], be[j], f);
      const double q = yp[i] - f;
      r[i] = q;
      part = fma(q, q, part);
    }
    rss = omc_block_sum(part, s_red);
  }
  const double tau_y = vat(a.tau_y, chain, 1.0);
  const double shape_w = vat(a.omega_shape, chain, 1.0), rate_w = vat(a.omega_rate, chain, 1.0);
  const long long sw = a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll;
  const double* dtn = w.debug_tn_u ? w.debug_tn_u + sw * w.debug_sweep_stride + (long long)chain * cap : nullptr;
  const double* du = w.debug_u ? w.debug_u + sw * w.debug_sweep_stride + (long long)chain * cap : nullptr;
  const OmcRng rng = to_rng(a.rng);
  long long n_acc = 0;
  // ---- all proposals up front, one component per thread: the truncated-normal inverse CDF and the two proposal
  //      densities are the expensive scalar part of a step and do not depend on what the other components did
  for (int j = tid; j < k; j += RM_NT) {
    const double cur = w.which ? om[j] : th[j];
    const uint4 b = omc_rng_block(rng, chain, (unsigned int)j);
    const double var = dtn ? dtn[j] : omc_u01(b.x, b.y);
    const double z = omc_truncated_normal_rv(cur, w.step, w.lim_lo, w.lim_hi, var);
    const double lq_f = omc_truncated_normal_log_pdf(z, cur, w.step, w.lim_lo, w.lim_hi);
    const double lq_r = omc_truncated_normal_log_pdf(cur, z, w.step, w.lim_lo, w.lim_hi);
    // prior of the moved component: Uniform knots are constant, widths carry their Gamma prior when it is in the model
    const double dprior = (w.which && a.sample_omega) ? gamma_logpdf(z, shape_w, rate_w) - gamma_logpdf(cur, shape_w, rate_w) : 0.0;
    pz[j] = z;
    plq[j] = dprior + lq_r - lq_f;
    pu[j] = du ? du[j] : omc_u01(b.z, b.w);
  }
  __syncthreads();
  for (int j = 0; j < k; ++j) {
    const double zj = pz[j];
    const double tj = w.which ? th[j] : zj, oj = w.which ? zj : om[j];
    double part = 0.0;
    for (int i = tid; i < nd; i += RM_NT) {
      const double c = normpdf(a.X[i], tj, oj);
      cnew[i] = c;
      if (yp) {
        const double q = r[i] - be[j] * (c - Bg[(long long)i * cap + j]);
        part = fma(q, q, part);
      }
    }
    double rss_new = rss;
    if (yp) rss_new = omc_block_sum(part, s_red);
    if (tid == 0) {
      double log_accept = plq[j];
      if (yp) log_accept += -0.5 * tau_y * (rss_new - rss);
      sh.accept = (log(pu[j]) < log_accept) ? 1 : 0;
    }
    __syncthreads();
    if (sh.accept) {
      for (int i = tid; i < nd; i += RM_NT) {
        double* bij = Bg + (long long)i * cap + j;
        if (yp) r[i] -= be[j] * (cnew[i] - *bij);
        *bij = cnew[i];
      }
      if (tid == 0) {
        if (w.which) om[j] = zj; else th[j] = zj;
      }
      rss = rss_new;
      ++n_acc;
    }
    __syncthreads();
  }
  for (int j = tid; j < k; j += RM_NT) {
    if (w.which) omg[j] = om[j]; else thg[j] = th[j];
  }
  if (tid == 0 && n_acc > 0 && w.model.gram_valid) w.model.gram_valid[chain] = 0;   // basis columns moved: S = B'B is stale
  if (tid == 0 && w.counters) {
    w.counters[2 * (long long)chain] += n_acc;
    w.counters[2 * (long long)chain + 1] += k;
  }
}

// ---- small block-cooperative helpers for a k x k lower factor in shared memory (any k)
// x <- L^-1 x (forward) by warp 0; all threads must call.
__device__ void tri_solve_lower(const double* L, int k, int ld, double* x) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 32) {
    for (int j = 0; j < k; ++j) {
      double s = 0.0;
      for (int c = lane; c < j; c += 32) s = fma(L[j * ld + c], x[c], s);
      s = omc_warp_sum(s);
      if (lane == 0) x[j] = (x[j] - s) / L[j * ld + j];
      __syncwarp();
    }
  }
  __syncthreads();
}
// x <- L^-T x (backward) by warp 0.
__device__ void tri_solve_lower_T(const double* L, int k, int ld, double* x) {
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < 32) {
    for (int j = k - 1; j >= 0; --j) {
      double s = 0.0;
      for (int c = j + 1 + lane; c < k; c += 32) s = fma(L[c * ld + j], x[c], s);
      s = omc_warp_sum(s);
      if (lane == 0) x[j] = (x[j] - s) / L[j * ld + j];
      __syncwarp();
    }
  }
  __syncthreads();
}

struct MalaShared {
  double lq_f, lq_r;
  int ok;
};

// CTA-wide loop over the elements (i, c) of a rows x cols block in the order e = i * cols + c, e += RM_NT, with (i, c)
// advanced incrementally: one integer division per thread instead of one per element (same helper as in rj.cu).
template <class F>
__device__ __forceinline__ void rm_for2d(int rows, int cols, F f) {
  const int q = RM_NT / cols, rr = RM_NT - q * cols;
  int i = threadIdx.x / cols, c = threadIdx.x - i * cols;
  while (i < rows) {
    f(i, c);
    c += rr;
    i += q;
    if (c >= cols) { c -= cols; ++i; }
  }
}

__global__ void __launch_bounds__(RM_NT) rj_coef_mmala_kernel(omc_rj_mmala_t m, int ld, int cls) {
  extern __shared__ __align__(16) double sm[];
  __shared__ MalaShared sh;
  __shared__ double s_red[32];
  const omc_rj_t& a = m.model;
  const int tid = threadIdx.x, chain = blockIdx.x;
  const int nd = a.n_data, cap = a.n_max;
  if (cls >= 0 && a.size_class[chain] != cls) return;
  const int k = (int)a.n_basis[chain];
  if (k < 1 || k > cap) {
    if (tid == 0 && a.status) atomicOr(&a.status[chain], OMC_STATUS_NAN);
    return;
  }
  double* S = sm;                        // ld x ld : Gram matrix, then Hs = H / step^2, then its Cholesky factor
  double* chunk = S + ld * ld;           // RM_ROWS x ld
  double* cv = chunk + RM_ROWS * ld;     // B'y
  double* be = cv + ld;                  // current coefficients
  double* pr = be + ld;                  // proposal
  double* g = pr + ld;                   // gradient / work vector
  double* mu = g + ld;                   // proposal mean
  double* zv = mu + ld;
  unsigned short* ptab = reinterpret_cast<unsigned short*>(zv + ld);   // (i, j) of every lower-triangle pair
  double* beg = a.beta + (long long)chain * cap;
  const double* Bg = a.B + (long long)chain * nd * cap;
  const double* yp = a.y.ptr ? a.y.ptr + (long long)chain * a.y.chain_stride : nullptr;
  const double tau_y = vat(a.tau_y, chain, 1.0), tau_b = vat(a.tau_beta, chain, 1.0), mu_b = vat(a.mu_beta, chain, 0.0);
  for (int j = tid; j < k; j += RM_NT) { be[j] = beg[j]; cv[j] = 0.0; }
  for (int e = tid; e < k * ld; e += RM_NT) S[e] = 0.0;
  __syncthreads();
  // ---- Gram matrix S = B'B (lower, then symmetrised) and c = B'y; current residual sum of squares
  double rss_c = 0.0;
  if (yp) {
    const int npair = k * (k + 1) / 2;
    for (int pi = tid; pi < npair; pi += RM_NT) {   // pair index -> (i, j), decoded once per step
      int i = (int)((sqrt(8.0 * pi + 1.0) - 1.0) * 0.5);
      while ((i + 1) * (i + 2) / 2 <= pi) ++i;
      while (i * (i + 1) / 2 > pi) --i;
      ptab[pi] = (unsigned short)((i << 8) | (pi - i * (i + 1) / 2));
    }
    __syncthreads();
    for (int r0 = 0; r0 < nd; r0 += RM_ROWS) {
      const int rows = min(RM_ROWS, nd - r0);
      rm_for2d(rows, k, [&](int r_, int j) { chunk[r_ * ld + j] = Bg[(long long)(r0 + r_) * cap + j]; });
      __syncthreads();
      for (int pi = tid; pi < npair; pi += RM_NT) {
        const int i = ptab[pi] >> 8, j = ptab[pi] & 255;
        double s = 0.0;
        for (int r_ = 0; r_ < rows; ++r_) s = fma(chunk[r_ * ld + i], chunk[r_ * ld + j], s);
        S[i * ld + j] += s;
      }
      for (int j = tid; j < k; j += RM_NT) {
        double s = 0.0;
        for (int r_ = 0; r_ < rows; ++r_) s = fma(chunk[r_ * ld + j], yp[r0 + r_], s);
        cv[j] += s;
      }
      if (tid < rows) {
        double f = 0.0;
        for (int j = 0; j < k; ++j) f = fma(chunk[tid * ld + j], be[j], f);
        const double q = yp[r0 + tid] - f;
        rss_c = fma(q, q, rss_c);
      }
      __syncthreads();
    }
    rm_for2d(k, k, [&](int i, int j) {
      if (j > i) S[i * ld + j] = S[j * ld + i];
    });
    rss_c = omc_block_sum(rss_c, s_red);
  }
  __syncthreads();
  // ---- gradient at the current point, g = tau_y (c - S beta) - tau_b (beta - mu_b), while S still is the Gram matrix
  for (int i = tid; i < k; i += RM_NT) {
    double s_ = 0.0;
    if (yp) {
      for (int j = 0; j < k; ++j) s_ = fma(S[i * ld + j], be[j], s_);
      s_ = tau_y * (cv[i] - s_);
    }
    g[i] = s_ - tau_b * (be[i] - mu_b);
  }
  __syncthreads();
  // ---- Hs = (tau_y S + tau_b I) / step^2, L = chol(Hs)   (metropolis_hastings.py:325-348)
  const double inv_s2 = 1.0 / (m.step * m.step);
  rm_for2d(k, k, [&](int i, int j) {
    S[i * ld + j] = ((yp ? tau_y * S[i * ld + j] : 0.0) + (i == j ? tau_b : 0.0)) * inv_s2;
  });
  __syncthreads();
  const bool pd = omc_chol_block(S, k, ld);
  __syncthreads();
  if (!pd) {
    if (tid == 0) {
      if (a.status) atomicOr(&a.status[chain], OMC_STATUS_NOT_PD);
      if (m.counters) m.counters[2 * (long long)chain + 1] += 1;
    }
    return;
  }
  // log N(x | mean, (L L')^-1) up to the constant: sum log L_ii - |L'(x - mean)|^2 / 2   (metropolis_hastings.py:350-373)
  auto log_density = [&](const double* x, const double* mean) {
    double part = 0.0;
    for (int c = tid; c < k; c += RM_NT) {
      double s_ = 0.0;
      for (int i = c; i < k; ++i) s_ = fma(S[i * ld + c], x[i] - mean[i], s_);
      part += log(S[c * ld + c]) - 0.5 * s_ * s_;
    }
    return omc_block_sum(part, s_red);
  };
  // w = Hs^-1 g (kept in g); forward proposal mean mu = beta + w / 2
  tri_solve_lower(S, k, ld, g);
  tri_solve_lower_T(S, k, ld, g);
  for (int i = tid; i < k; i += RM_NT) mu[i] = be[i] + 0.5 * g[i];
  // z, proposal = mu + L^-T z
  const long long sw = a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll;
  if (m.debug_z) {
    const double* dz = m.debug_z + sw * m.debug_sweep_stride_z + (long long)chain * cap;
    for (int i = tid; i < k; i += RM_NT) zv[i] = dz[i];
  } else {
    const OmcRng rng = to_rng(a.rng);
    for (int t = tid; 2 * t < k; t += RM_NT) {
      double z0, z1;
      omc_normal2(rng, chain, t, z0, z1);
      zv[2 * t] = z0;
      if (2 * t + 1 < k) zv[2 * t + 1] = z1;
    }
  }
  __syncthreads();
  tri_solve_lower_T(S, k, ld, zv);
  for (int i = tid; i < k; i += RM_NT) pr[i] = mu[i] + zv[i];
  __syncthreads();
  const double lq_f = log_density(pr, mu);
  __syncthreads();
  // reverse proposal mean at the proposed point.  The conditional of beta is Gaussian, so the Hessian (and L) is the
  // same there and the gradient moves linearly, g' = g - H (beta' - beta):  Hs^-1 g' = w - step^2 (beta' - beta).
  for (int i = tid; i < k; i += RM_NT) mu[i] = pr[i] + 0.5 * (g[i] - m.step * m.step * (pr[i] - be[i]));
  __syncthreads();
  const double lq_r = log_density(be, mu);
  __syncthreads();
  // model log-densities: response through the residuals, prior through the sums of squares
  double rss_p = 0.0, ss_c = 0.0, ss_p = 0.0;
  if (yp) {
    double part = 0.0;
    for (int i = tid; i < nd; i += RM_NT) {
      const double* row = Bg + (long long)i * cap;
      double f = 0.0;
      for (int j = 0; j < k; ++j) f = fma(row[j], pr[j], f);
      const double q = yp[i] - f;
      part = fma(q, q, part);
    }
    rss_p = omc_block_sum(part, s_red);
    __syncthreads();
  }
  {
    double pc = 0.0, pp = 0.0;
    for (int j = tid; j < k; j += RM_NT) {
      pc = fma(be[j] - mu_b, be[j] - mu_b, pc);
      pp = fma(pr[j] - mu_b, pr[j] - mu_b, pp);
    }
    ss_c = omc_block_sum(pc, s_red);
    __syncthreads();
    ss_p = omc_block_sum(pp, s_red);
  }
  if (tid == 0) {
    double u;
    if (m.debug_u) u = m.debug_u[sw * m.debug_sweep_stride_u + chain];
    else {
      const uint4 b = omc_rng_block(to_rng(a.rng), chain, 0xFFFFu);
      u = omc_u01(b.x, b.y);
    }
    const double dlp = (yp ? -0.5 * tau_y * (rss_p - rss_c) : 0.0) - 0.5 * tau_b * (ss_p - ss_c);
    const double log_accept = dlp + lq_r - lq_f;
    sh.ok = (log(u) < log_accept) ? 1 : 0;
    if (m.counters) {
      m.counters[2 * (long long)chain] += sh.ok;
      m.counters[2 * (long long)chain + 1] += 1;
    }
    if (m.probe) {
      double* o = m.probe + (long long)chain * 6;
      o[0] = rss_c; o[1] = rss_p; o[2] = lq_f; o[3] = lq_r; o[4] = log_accept; o[5] = sh.ok;
    }
  }
  __syncthreads();
  if (sh.ok)
    for (int j = tid; j < k; j += RM_NT) beg[j] = pr[j];
}

__global__ void rm_class_kernel(const double* n_basis, int n_chains, int* size_class) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  const int k = (int)n_basis[c];
  size_class[c] = (k <= 32) ? 0 : (k <= 64) ? 1 : 2;
}

int rm_check(const omc_rj_t* a, const char* who) {
  OMC_REQUIRE(a && a->n_basis && a->theta && a->omega && a->beta && a->B && a->X, "%s: null argument", who);
  OMC_REQUIRE(a->n_chains >= 1 && a->n_data >= 1 && a->n_max >= 2, "%s: bad shape", who);
  return 0;
}

}  // namespace

extern "C" {

int omc_rj_knot_walk(const omc_rj_walk_t* w, void* stream) {
  OMC_REQUIRE(w, "omc_rj_knot_walk: null argument");
  OMC_REQUIRE_SITE(w->model.rng, "omc_rj_knot_walk");
  if (int rc = rm_check(&w->model, "omc_rj_knot_walk")) return rc;
  OMC_REQUIRE(w->which == 0 || w->which == 1, "omc_rj_knot_walk: which=%d", w->which);
  OMC_REQUIRE(w->step > 0.0 && w->lim_hi > w->lim_lo, "omc_rj_knot_walk: step / limits");
  const int smem = (2 * w->model.n_data + 6 * w->model.n_max) * 8;
  OMC_REQUIRE(smem <= 200 * 1024, "omc_rj_knot_walk: n_data=%d too large", w->model.n_data);
  OMC_CHECK_CUDA(cudaFuncSetAttribute(rj_knot_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  rj_knot_walk_kernel<<<w->model.n_chains, RM_NT, smem, (cudaStream_t)stream>>>(*w);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_rj_coef_mmala(const omc_rj_mmala_t* m, void* stream) {
  OMC_REQUIRE(m, "omc_rj_coef_mmala: null argument");
  OMC_REQUIRE_SITE(m->model.rng, "omc_rj_coef_mmala");
  const omc_rj_t* a = &m->model;
  if (int rc = rm_check(a, "omc_rj_coef_mmala")) return rc;
  OMC_REQUIRE(m->step > 0.0, "omc_rj_coef_mmala: step=%g", m->step);
  cudaStream_t st = (cudaStream_t)stream;
  auto smem_for = [&](int ld) { return (ld * ld + RM_ROWS * ld + 6 * ld + (ld * (ld + 1) / 2 * 2 + 7) / 8) * 8; };
  const int full = smem_for(a->n_max + 1);
  OMC_REQUIRE(smem_for(a->n_max <= 64 ? a->n_max + 1 : 65) <= 220 * 1024 && (a->n_max <= 64 || full <= 227 * 1024 || a->size_class),
              "omc_rj_coef_mmala: n_max=%d needs too much shared memory", a->n_max);
  OMC_CHECK_CUDA(cudaFuncSetAttribute(rj_coef_mmala_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      full <= 227 * 1024 ? full : 227 * 1024));
  if (!a->size_class || a->n_max <= 32) {
    OMC_REQUIRE(full <= 227 * 1024, "omc_rj_coef_mmala: n_max=%d needs the size-class scratch", a->n_max);
    rj_coef_mmala_kernel<<<a->n_chains, RM_NT, full, st>>>(*m, a->n_max + 1, -1);
    OMC_LAUNCH_CHECK();
    return 0;
  }
  rm_class_kernel<<<(a->n_chains + 255) / 256, 256, 0, st>>>(a->n_basis, a->n_chains, a->size_class);
  OMC_LAUNCH_CHECK();
  const int hi[3] = {32, 64, a->n_max};
  for (int q = 0; q < 3; ++q) {
    const int top = hi[q] < a->n_max ? hi[q] : a->n_max;
    const int smem = smem_for(top + 1);
    OMC_REQUIRE(smem <= 227 * 1024, "omc_rj_coef_mmala: %d live coefficients need %d bytes of shared memory", top, smem);
    rj_coef_mmala_kernel<<<a->n_chains, RM_NT, smem, st>>>(*m, top + 1, q);
    OMC_LAUNCH_CHECK();
    if (top >= a->n_max) break;
  }
  return 0;
}

}  // extern "C"
