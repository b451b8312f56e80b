// Small per-chain kernels of the conjugate Gibbs updates:
//   omc_nn_dense_draw : Q = lambda*P0 + tau*G, Cholesky, posterior mean, MVN draw  (Rue & Held Alg. 2.5)
//   omc_quadform      : (x-mu)' P (x-mu) and #(diag P > 0)
//   omc_ng_draw       : Gamma(a0 + cnt/2, b0 + ss/2) draw
// One CTA per chain for the p x p work (matrix resident in shared memory, p <= 64); one thread per chain for the
// scalar Gamma draw.  These are latency-bound and account for a few % of a C2 sweep (profiles/), the pass in
// reg_pass.cu is the roofline kernel.
#include "../../include/omc.h"
#include "omc_common.cuh"
#include "omc_internal.h"
#include "omc_special.cuh"

namespace {

constexpr int DD_THREADS = 256;
constexpr int PMAX = 64;

__device__ __forceinline__ double vec_at(const omc_vec_t& v, int chain, int i, double dflt) {
  return v.ptr ? v.ptr[(long long)chain * v.chain_stride + i] : dflt;
}
__device__ __forceinline__ OmcRng to_rng(const omc_rng_t& r) {
  OmcRng o;
  o.seed = r.seed; o.sweep = r.sweep; o.chain_offset = r.chain_offset; o.site = r.site;
  return o;
}

// P0[i][j] for the three storage kinds
__device__ __forceinline__ double mat_at(int kind, const omc_vec_t& P, int chain, int p, int i, int j) {
  if (kind == OMC_MAT_DENSE) return P.ptr[(long long)chain * P.chain_stride + (long long)i * p + j];
  if (i != j) return 0.0;
  if (kind == OMC_MAT_DIAG) return P.ptr[(long long)chain * P.chain_stride + i];
  return P.ptr ? P.ptr[(long long)chain * P.chain_stride] : 1.0;
}

__global__ void __launch_bounds__(DD_THREADS) nn_dense_draw_kernel(omc_nn_dense_t a) {
  extern __shared__ double sm[];
  const int p = a.p, ld = p + 1;
  double* Q = sm;                 // p x ld, becomes L (lower triangle)
  double* b = Q + p * ld;         // p : rhs -> w -> mu
  double* z = b + p;              // p : z -> v
  double* mu0 = z + p;            // p
  double* invd = mu0 + p;         // p : 1 / L_jj
  __shared__ int bad;
  const int chain = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = DD_THREADS / 32;
  if (tid == 0) bad = 0;

  const double* rec = a.stats.ptr + (long long)chain * a.stats.chain_stride;
  const double tau = vec_at(a.tau, chain, 0, 1.0);
  const double lam = vec_at(a.lambda, chain, 0, 1.0);

  for (int i = tid; i < p; i += DD_THREADS) mu0[i] = vec_at(a.mu0, chain, i, 0.0);
  // Q = lam*P0 + tau*G  (full symmetric; reference adds the scaled prior then the likelihood Hessian, sampler.py:180-186)
  for (int e = tid; e < p * p; e += DD_THREADS) {
    int i = e / p, j = e - i * p;
    double q = lam * mat_at(a.prior_kind, a.prior_P, chain, p, i, j) + tau * rec[e];
    Q[i * ld + j] = q;
    if (a.probe_Q) a.probe_Q[(long long)chain * p * p + e] = q;
  }
  __syncthreads();
  // b = (lam*P0) mu0 + tau*g
  for (int i = tid; i < p; i += DD_THREADS) {
    double s = 0.0;
    if (a.prior_kind == OMC_MAT_DENSE) {
      for (int j = 0; j < p; ++j) s += lam * mat_at(OMC_MAT_DENSE, a.prior_P, chain, p, i, j) * mu0[j];
    } else {
      s = lam * mat_at(a.prior_kind, a.prior_P, chain, p, i, i) * mu0[i];
    }
    s += tau * rec[p * p + i];
    b[i] = s;
    if (a.probe_b) a.probe_b[(long long)chain * p + i] = s;
  }
  if (a.truncated) {
    // ---- truncated prior: one coordinate-wise Gibbs scan from the current beta (gmrf.py:201-266), warp 0
    __syncthreads();
    if (warp == 0) {
      double* bt = a.beta + (long long)chain * p;
      double x0 = (lane < p) ? bt[lane] : 0.0, x1 = (lane + 32 < p) ? bt[lane + 32] : 0.0;
      const double* du = a.debug_u
                             ? a.debug_u + (a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll) * a.debug_sweep_stride +
                                   (long long)chain * p
                             : nullptr;
      OmcRng rng = to_rng(a.rng);
      for (int i = 0; i < p; ++i) {
        double part = 0.0;
        if (lane < p) part = Q[i * ld + lane] * x0;
        if (lane + 32 < p) part = fma(Q[i * ld + lane + 32], x1, part);
        const double dot = omc_warp_sum(part);
        const double xi = __shfl_sync(0xffffffffu, (i < 32) ? x0 : x1, i & 31);
        const double qii = Q[i * ld + i];
        const double v = 1.0 / qii;
        const double mean = (p == 1) ? b[0] / qii : v * (b[i] - dot + qii * xi);
        const double lo = a.trunc_lo.ptr ? vec_at(a.trunc_lo, chain, a.trunc_lo_len > 1 ? i : 0, 0.0) : -INFINITY;
        const double hi = a.trunc_hi.ptr ? vec_at(a.trunc_hi, chain, a.trunc_hi_len > 1 ? i : 0, 0.0) : INFINITY;
        double u;
        if (du) u = du[i];
        else {
          const uint4 blk = omc_rng_block(rng, chain, (unsigned int)i);
          u = omc_u01(blk.x, blk.y);
        }
        const double xn = omc_truncated_normal_rv(mean, sqrt(v), lo, hi, u);
        if (!(qii > 0.0) && lane == 0) bad = 1;
        if (lane == (i & 31)) { if (i < 32) x0 = xn; else x1 = xn; }
      }
      if (lane < p) bt[lane] = x0;
      if (lane + 32 < p) bt[lane + 32] = x1;
      if (lane == 0 && bad && a.status) atomicOr(&a.status[chain], OMC_STATUS_NOT_PD);
    }
    return;
  }
  // z: injected or Philox/Box-Muller
  if (a.debug_z) {
    const double* dz = a.debug_z + (a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll) * a.debug_sweep_stride;
    for (int i = tid; i < p; i += DD_THREADS) z[i] = dz[(long long)chain * p + i];
  } else {
    OmcRng rng = to_rng(a.rng);
    for (int t = tid; 2 * t < p; t += DD_THREADS) {
      double z0, z1;
      omc_normal2(rng, chain, t, z0, z1);
      z[2 * t] = z0;
      if (2 * t + 1 < p) z[2 * t + 1] = z1;
    }
  }
  __syncthreads();

  // ---- right-looking Cholesky in shared memory (lower), 2 barriers per column
  for (int j = 0; j < p; ++j) {
    const double djj = Q[j * ld + j];
    if (!(djj > 0.0)) {  // uniform across the CTA (same smem value)
      if (tid == 0) bad = 1;
      break;
    }
    const double d = sqrt(djj), rd = 1.0 / d;
    __syncthreads();  // everyone has read Q[j][j] before it is overwritten
    for (int i = j + tid; i < p; i += DD_THREADS) {
      if (i == j) { Q[j * ld + j] = d; invd[j] = rd; }
      else Q[i * ld + j] = Q[i * ld + j] / d;
    }
    __syncthreads();
    for (int i = j + 1 + warp; i < p; i += nwarp) {
      const double lij = Q[i * ld + j];
      for (int c = j + 1 + lane; c <= i; c += 32) Q[i * ld + c] -= lij * Q[c * ld + j];
    }
    __syncthreads();
  }
  __syncthreads();
  if (bad) {
    if (tid == 0 && a.status) atomicOr(&a.status[chain], OMC_STATUS_NOT_PD);
    for (int i = tid; i < p; i += DD_THREADS) a.beta[(long long)chain * p + i] = nan("");
    return;
  }
  if (a.probe_L) {
    for (int e = tid; e < p * p; e += DD_THREADS) {
      int i = e / p, j = e - i * p;
      a.probe_L[(long long)chain * p * p + e] = (j <= i) ? Q[i * ld + j] : 0.0;
    }
  }

  // ---- triangular solves, one warp each: warp 0: L w = b, L' mu = w ; warp 1: L' v = z
  // each lane owns rows lane and lane+32
  if (warp == 0) {
    double x0 = (lane < p) ? b[lane] : 0.0, x1 = (lane + 32 < p) ? b[lane + 32] : 0.0;
    for (int j = 0; j < p; ++j) {  // forward, column oriented
      double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) / Q[j * ld + j];
      if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
      if (lane > j && lane < p) x0 -= Q[lane * ld + j] * xj;
      if (lane + 32 > j && lane + 32 < p) x1 -= Q[(lane + 32) * ld + j] * xj;
    }
    for (int j = p - 1; j >= 0; --j) {  // backward with L' : uses row j of L
      double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) / Q[j * ld + j];
      if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
      if (lane < j) x0 -= Q[j * ld + lane] * xj;
      if (lane + 32 < j) x1 -= Q[j * ld + lane + 32] * xj;
    }
    if (lane < p) b[lane] = x0;
    if (lane + 32 < p) b[lane + 32] = x1;
  } else if (warp == 1) {
    double x0 = (lane < p) ? z[lane] : 0.0, x1 = (lane + 32 < p) ? z[lane + 32] : 0.0;
    for (int j = p - 1; j >= 0; --j) {
      double xj = __shfl_sync(0xffffffffu, (j < 32) ? x0 : x1, j & 31) / Q[j * ld + j];
      if (lane == (j & 31)) { if (j < 32) x0 = xj; else x1 = xj; }
      if (lane < j) x0 -= Q[j * ld + lane] * xj;
      if (lane + 32 < j) x1 -= Q[j * ld + lane + 32] * xj;
    }
    if (lane < p) z[lane] = x0;
    if (lane + 32 < p) z[lane + 32] = x1;
  }
  __syncthreads();
  for (int i = tid; i < p; i += DD_THREADS) {
    const double m = b[i];
    if (a.probe_mu) a.probe_mu[(long long)chain * p + i] = m;
    a.beta[(long long)chain * p + i] = m + z[i];
  }
}

__global__ void quadform_kernel(omc_quadform_t a) {
  // one warp per chain
  const int chain = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (chain >= a.n_chains) return;
  const int p = a.p;
  double ss = 0.0, cnt = 0.0;
  if (a.kind == OMC_MAT_DENSE) {
    for (int i = lane; i < p; i += 32) {
      const double ri = vec_at(a.x, chain, i, 0.0) - vec_at(a.mu, chain, i, 0.0);
      double s = 0.0;
      for (int j = 0; j < p; ++j)
        s += mat_at(OMC_MAT_DENSE, a.P, chain, p, i, j) * (vec_at(a.x, chain, j, 0.0) - vec_at(a.mu, chain, j, 0.0));
      ss += ri * s;
      cnt += (mat_at(OMC_MAT_DENSE, a.P, chain, p, i, i) > 0.0) ? 1.0 : 0.0;
    }
  } else {
    for (int i = lane; i < p; i += 32) {
      const double ri = vec_at(a.x, chain, i, 0.0) - vec_at(a.mu, chain, i, 0.0);
      const double d = mat_at(a.kind, a.P, chain, p, i, i);
      ss += d * ri * ri;
      cnt += (d > 0.0) ? 1.0 : 0.0;
    }
  }
  ss = omc_warp_sum(ss);
  cnt = omc_warp_sum(cnt);
  if (lane == 0) {
    a.ss[chain] = ss;
    a.cnt[chain] = cnt;
  }
}

__global__ void ng_draw_kernel(omc_ng_draw_t a) {
  const int ne = a.n_elem > 0 ? a.n_elem : 1;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (long long)a.n_chains * ne) return;
  const int chain = (int)(t / ne), k = (int)(t - (long long)chain * ne);
  const double shape = vec_at(a.a0, chain, a.a0_len > 1 ? k : 0, 0.0) + 0.5 * vec_at(a.cnt, chain, k * a.cnt_stride, 0.0);
  const double rate = vec_at(a.b0, chain, a.b0_len > 1 ? k : 0, 0.0) + 0.5 * vec_at(a.ss, chain, k * a.ss_stride, 0.0);
  if (a.probe_a) a.probe_a[t] = shape;
  if (a.probe_b) a.probe_b[t] = rate;
  double gvar;
  if (a.debug_g) gvar = a.debug_g[(a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll) * a.debug_sweep_stride + t];
  else gvar = omc_std_gamma(to_rng(a.rng), chain, (unsigned int)k << 12, shape);   // 4096 Philox blocks per element
  // reference: scale = inf when rate == 0 (sampler.py:285-286)
  a.out[t] = (rate == 0.0) ? INFINITY : gvar * (1.0 / rate);
}

}  // namespace

extern "C" int omc_nn_dense_draw(const omc_nn_dense_t* args, void* stream) {
  OMC_REQUIRE(args && args->stats.ptr && args->beta, "omc_nn_dense_draw: null argument");
  OMC_REQUIRE(args->p >= 1 && args->p <= PMAX, "omc_nn_dense_draw: p=%d outside [1,%d]", args->p, PMAX);
  OMC_REQUIRE(args->n_chains >= 1, "omc_nn_dense_draw: n_chains=%d", args->n_chains);
  OMC_REQUIRE(args->prior_kind >= 0 && args->prior_kind <= 2, "omc_nn_dense_draw: prior_kind=%d", args->prior_kind);
  OMC_REQUIRE(args->prior_kind == OMC_MAT_EYE || args->prior_P.ptr, "omc_nn_dense_draw: prior_P missing");
  const int p = args->p;
  const size_t smem = (size_t)(p * (p + 1) + 4 * p) * sizeof(double);
  nn_dense_draw_kernel<<<args->n_chains, DD_THREADS, smem, (cudaStream_t)stream>>>(*args);
  OMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int omc_quadform(const omc_quadform_t* args, void* stream) {
  OMC_REQUIRE(args && args->x.ptr && args->ss && args->cnt, "omc_quadform: null argument");
  OMC_REQUIRE(args->kind >= 0 && args->kind <= 2, "omc_quadform: kind=%d", args->kind);
  OMC_REQUIRE(args->kind == OMC_MAT_EYE || args->P.ptr, "omc_quadform: P missing");
  const int threads = 128;
  const long long total = (long long)args->n_chains * 32;
  quadform_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(*args);
  OMC_LAUNCH_CHECK();
  return 0;
}

extern "C" int omc_ng_draw(const omc_ng_draw_t* args, void* stream) {
  OMC_REQUIRE(args && args->out, "omc_ng_draw: null argument");
  const int threads = 128;
  const long long total = (long long)args->n_chains * (args->n_elem > 0 ? args->n_elem : 1);
  ng_draw_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(*args);
  OMC_LAUNCH_CHECK();
  return 0;
}
