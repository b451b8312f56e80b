// Chain diagnostics on the device-resident sample store: per-chain mean / variance / autocorrelation
// ESS, and the cross-chain split-R-hat + total ESS from the all-gathered per-chain summaries.
//
// The reference has no diagnostics at all (no ess|rhat|autocorr in src/); BASELINE.json's metric asks for
// ESS/s and the north star for an R-hat/ESS all-gather, so this is new functionality with PARITY UNPINNED.  The
// estimators are restated in numpy in oracle/diagnostics.py, which is what the tests compare against.
//
//   ESS_c,j  = N / (1 + 2 sum_t rho_t), rho from the biased autocovariance (divide by N), truncated by Geyer's initial
//              monotone positive sequence over lag pairs, lags <= max_lag
//   split-R-hat_j: every chain split in two halves of nh = N/2 draws; W = mean of the half variances, B/nh = variance
//              of the half means; R-hat = sqrt(((nh-1)/nh W + B/nh) / W)      (Gelman et al., BDA3 §11.4)
//
// Layout: samples [n_iter][n_chains][size] (the store written by omc_store_copy); a "series" is one selected element
// (index j*elem_stride) of one chain.  Consecutive series are consecutive addresses, so a warp reads 32 series of one
// iteration with one coalesced request; a CTA = 32 series x 8 lag groups works from a shared-memory time chunk.
#include "../../include/omc.h"
#include "omc_common.cuh"
#include "omc_internal.h"

namespace {

constexpr int DG_SER = 32;        // series per CTA (lane)
constexpr int DG_GROUPS = 8;      // lag groups (warp)
constexpr int DG_TC = 192;        // time steps per shared-memory chunk
constexpr int DG_MAXLAG = 127;    // lags 0..127 -> 16 accumulators per thread
constexpr int DG_Q = (DG_MAXLAG + 1) / DG_GROUPS;

__global__ void __launch_bounds__(DG_SER* DG_GROUPS) chain_stats_kernel(omc_chain_stats_t a) {
  extern __shared__ double tile[];                  // [(DG_TC + DG_MAXLAG)][DG_SER]
  __shared__ double s_acc[DG_MAXLAG + 1][DG_SER];   // autocovariances
  __shared__ double s_part[DG_GROUPS][3][DG_SER];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const long long n_series = (long long)a.n_chains * a.n_sel;
  const long long ser = (long long)blockIdx.x * DG_SER + lane;
  const bool live = ser < n_series;
  const long long chain = live ? ser / a.n_sel : 0;
  const long long jsel = live ? ser % a.n_sel : 0;
  const long long N = a.n_iter;
  const long long step = (long long)a.n_chains * a.size;   // elements between consecutive iterations
  const double* x = a.samples + chain * a.size + jsel * a.elem_stride;
  const int L = (int)min((long long)min(a.max_lag, DG_MAXLAG), N - 1);
  const long long nh = N / 2;
  // ---- pass 0: sums for the mean and for the two half means (an odd-length chain leaves its middle draw out of both)
  double s1 = 0.0, s2 = 0.0, sm = 0.0;
  if (live)
    for (long long t = grp; t < N; t += DG_GROUPS) {
      const double v = x[t * step];
      if (t < nh) s1 += v;
      else if (t >= N - nh) s2 += v;
      else sm += v;
    }
  s_part[grp][0][lane] = s1;
  s_part[grp][1][lane] = s2;
  s_part[grp][2][lane] = sm;
  __syncthreads();
  double m1 = 0.0, m2 = 0.0, mid = 0.0;
  for (int g = 0; g < DG_GROUPS; ++g) { m1 += s_part[g][0][lane]; m2 += s_part[g][1][lane]; mid += s_part[g][2][lane]; }
  const double mean = (m1 + m2 + mid) / (double)N;
  if (nh > 0) { m1 /= (double)nh; m2 /= (double)nh; }
  // ---- pass 1: autocovariance sums, lag l = grp + 8 q, over shared-memory chunks of centred values;
  //              the half variances ride along (group 0)
  double acc[DG_Q];
#pragma unroll
  for (int q = 0; q < DG_Q; ++q) acc[q] = 0.0;
  double v1 = 0.0, v2 = 0.0;
  for (long long t0 = 0; t0 < N; t0 += DG_TC) {
    const int rows = (int)min((long long)(DG_TC + L), N - t0);
    for (int r = grp; r < rows; r += DG_GROUPS) tile[r * DG_SER + lane] = live ? x[(t0 + r) * step] - mean : 0.0;
    __syncthreads();
    const int tc = (int)min((long long)DG_TC, N - t0);
#pragma unroll
    for (int q = 0; q < DG_Q; ++q) {
      const int l = grp + DG_GROUPS * q;
      if (l <= L) {
        double s = 0.0;
        const int tmax = min(tc, rows - l);
        for (int t = 0; t < tmax; ++t) s = fma(tile[t * DG_SER + lane], tile[(t + l) * DG_SER + lane], s);
        acc[q] += s;
      }
    }
    if (grp == 0)
      for (int t = 0; t < tc; ++t) {
        const double v = tile[t * DG_SER + lane] + mean;
        if (t0 + t < nh) v1 = fma(v - m1, v - m1, v1);
        else if (t0 + t >= N - nh) v2 = fma(v - m2, v - m2, v2);
      }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < DG_Q; ++q) {
    const int l = grp + DG_GROUPS * q;
    if (l <= DG_MAXLAG) s_acc[l][lane] = acc[q];
  }
  __syncthreads();
  // ---- Geyer truncation and outputs (one thread per series)
  if (grp == 0 && live) {
    const double c0 = s_acc[0][lane];
    double ess = (double)N;
    if (c0 > 0.0 && N > 3) {
      double sum_rho = 0.0, prev = INFINITY;   // sum over lag pairs (rho_{2k} + rho_{2k+1}), k >= 0, includes rho_0 = 1
      for (int k = 0; 2 * k + 1 <= L; ++k) {
        double pair = (s_acc[2 * k][lane] + s_acc[2 * k + 1][lane]) / c0;
        if (!(pair > 0.0)) break;
        pair = fmin(pair, prev);                // initial monotone sequence
        sum_rho += pair;
        prev = pair;
      }
      const double tau_int = fmax(2.0 * sum_rho - 1.0, 1.0 / log10((double)N + 9.0));   // floor as in Stan: ESS <= N log10 N
      ess = (double)N / tau_int;
    }
    double* o = a.out + ser * 8;
    o[0] = (double)N;
    o[1] = mean;
    o[2] = (N > 1) ? c0 / (double)(N - 1) : 0.0;
    o[3] = ess;
    o[4] = m1;
    o[5] = (nh > 1) ? v1 / (double)(nh - 1) : 0.0;
    o[6] = m2;
    o[7] = (nh > 1) ? v2 / (double)(nh - 1) : 0.0;
  }
}

// stats: [n_chains_total][n_sel][8] (all ranks' records after the all-gather); out: [n_sel][4] = rhat, ess_total,
// grand mean, pooled variance.  One thread per selected element, chains in index order (deterministic).
__global__ void rhat_combine_kernel(const double* stats, int n_chains, int n_sel, double* out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_sel) return;
  double sum_m = 0.0, sum_v = 0.0, ess = 0.0, gm = 0.0;
  const double nh = floor(stats[(long long)j * 8] / 2.0);
  for (int c = 0; c < n_chains; ++c) {
    const double* s = stats + ((long long)c * n_sel + j) * 8;
    sum_m += s[4] + s[6];
    sum_v += s[5] + s[7];
    ess += s[3];
    gm += s[1];
  }
  const double H = 2.0 * n_chains;
  const double mbar = sum_m / H, W = sum_v / H;
  double bsum = 0.0;
  for (int c = 0; c < n_chains; ++c) {
    const double* s = stats + ((long long)c * n_sel + j) * 8;
    bsum += (s[4] - mbar) * (s[4] - mbar) + (s[6] - mbar) * (s[6] - mbar);
  }
  const double b_over_n = bsum / (H - 1.0);   // variance of the half-chain means = B / nh
  const double var_plus = (nh - 1.0) / nh * W + b_over_n;
  double* o = out + (long long)j * 4;
  o[0] = (W > 0.0) ? sqrt(var_plus / W) : nan("");
  o[1] = ess;
  o[2] = gm / n_chains;
  o[3] = var_plus;
}

// ---------------------------------------------------------------------------------------------- rank normalisation
// Vehtari, Gelman, Simpson, Carpenter, Buerkner (2021): replace every draw by the normal score of its rank,
//   z = Phi^-1((r - 3/8) / (S + 1/4)),  r = average rank (ties share the mean of their ranks) among the S draws ranked together,
// and run split-R-hat / ESS on z ("bulk" diagnostics: insensitive to heavy tails, invariant under monotone transforms).
// Two kernels: every series (one selected element of one chain, n_iter draws) is sorted in shared memory (bitonic, one CTA
// per series) and written out; then every draw finds its rank by binary search in the sorted series -- of its own chain
// (pooled = 0: S = n_iter) or of ALL chains of the launch (pooled = 1: S = n_chains n_iter, the form split-R-hat wants
// when the chains share a target).  Output z [n_iter][n_chains][n_sel].
constexpr int RN_NT = 256;

__global__ void __launch_bounds__(RN_NT) rank_sort_kernel(const double* __restrict__ samples, long long n_iter, int n_chains,
                                                          long long size, long long n_sel, long long elem_stride, int npad,
                                                          double* __restrict__ sorted) {
  extern __shared__ double v[];                      // npad values (power of two), padded with +inf
  const long long ser = blockIdx.x;
  const long long chain = ser / n_sel, jsel = ser % n_sel;
  const long long step = (long long)n_chains * size;
  const double* x = samples + chain * size + jsel * elem_stride;
  for (int t = threadIdx.x; t < npad; t += RN_NT) v[t] = t < n_iter ? x[(long long)t * step] : INFINITY;
  __syncthreads();
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < npad; t += RN_NT) {
        const int o = t ^ j;
        if (o > t) {
          const double a = v[t], b = v[o];
          const bool up = (t & k) == 0;
          if ((a > b) == up) { v[t] = b; v[o] = a; }
        }
      }
      __syncthreads();
    }
  double* out = sorted + ser * n_iter;
  for (int t = threadIdx.x; t < n_iter; t += RN_NT) out[t] = v[t];
}

__device__ __forceinline__ void rn_counts(const double* __restrict__ s, long long n, double x, long long& less, long long& leq) {
  long long lo = 0, hi = n;                          // first index with s[i] >= x
  while (lo < hi) { const long long m = (lo + hi) >> 1; if (s[m] < x) lo = m + 1; else hi = m; }
  less += lo;
  hi = n;                                            // first index with s[i] > x
  while (lo < hi) { const long long m = (lo + hi) >> 1; if (s[m] <= x) lo = m + 1; else hi = m; }
  leq += lo;
}

__global__ void __launch_bounds__(RN_NT) rank_score_kernel(const double* __restrict__ samples, long long n_iter, int n_chains,
                                                           long long size, long long n_sel, long long elem_stride, int pooled,
                                                           const double* __restrict__ sorted, double* __restrict__ z) {
  const long long total = n_iter * n_chains * n_sel;
  for (long long e = (long long)blockIdx.x * RN_NT + threadIdx.x; e < total; e += (long long)gridDim.x * RN_NT) {
    const long long jsel = e % n_sel, chain = (e / n_sel) % n_chains, t = e / (n_sel * n_chains);
    const double x = samples[(t * n_chains + chain) * size + jsel * elem_stride];
    long long less = 0, leq = 0;
    if (pooled) {
      for (int c = 0; c < n_chains; ++c) rn_counts(sorted + ((long long)c * n_sel + jsel) * n_iter, n_iter, x, less, leq);
    } else {
      rn_counts(sorted + (chain * n_sel + jsel) * n_iter, n_iter, x, less, leq);
    }
    const double S = (double)(pooled ? n_iter * n_chains : n_iter);
    const double r = 0.5 * (double)(less + leq + 1);                 // ranks less+1 .. leq share their mean
    z[e] = (x == x) ? normcdfinv((r - 0.375) / (S + 0.25)) : x;     // NaN draws stay NaN
  }
}

}  // namespace

extern "C" {

int omc_chain_stats(const omc_chain_stats_t* a, void* stream) {
  OMC_REQUIRE(a && a->samples && a->out, "omc_chain_stats: null argument");
  OMC_REQUIRE(a->n_iter >= 1 && a->n_chains >= 1 && a->size >= 1 && a->n_sel >= 1 && a->elem_stride >= 1,
              "omc_chain_stats: bad shape");
  OMC_REQUIRE((a->n_sel - 1) * a->elem_stride < a->size, "omc_chain_stats: selection runs past the parameter (n_sel=%lld, "
              "elem_stride=%lld, size=%lld)", a->n_sel, a->elem_stride, a->size);
  OMC_REQUIRE(a->max_lag >= 1, "omc_chain_stats: max_lag=%d", a->max_lag);
  const long long n_series = (long long)a->n_chains * a->n_sel;
  const unsigned grid = (unsigned)((n_series + DG_SER - 1) / DG_SER);
  const int smem = (DG_TC + DG_MAXLAG) * DG_SER * 8;
  OMC_CHECK_CUDA(cudaFuncSetAttribute(chain_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  chain_stats_kernel<<<grid, DG_SER * DG_GROUPS, smem, (cudaStream_t)stream>>>(*a);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_rhat_combine(const double* stats, int n_chains_total, int n_sel, double* out, void* stream) {
  OMC_REQUIRE(stats && out && n_chains_total >= 1 && n_sel >= 1, "omc_rhat_combine: bad argument");
  rhat_combine_kernel<<<(n_sel + 127) / 128, 128, 0, (cudaStream_t)stream>>>(stats, n_chains_total, n_sel, out);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_rank_normalize(const double* samples, long long n_iter, int n_chains, long long size, long long n_sel,
                       long long elem_stride, int pooled, double* sorted, double* z, void* stream) {
  OMC_REQUIRE(samples && sorted && z, "omc_rank_normalize: null argument");
  OMC_REQUIRE(n_iter >= 1 && n_chains >= 1 && size >= 1 && n_sel >= 1 && elem_stride >= 1 && (n_sel - 1) * elem_stride < size,
              "omc_rank_normalize: bad shape");
  OMC_REQUIRE(n_iter <= 16384, "omc_rank_normalize: n_iter=%lld > 16384 draws per series (shared-memory sort)", n_iter);
  int npad = 2;
  while (npad < n_iter) npad <<= 1;
  const int smem = npad * 8;
  OMC_CHECK_CUDA(cudaFuncSetAttribute(rank_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long n_series = (long long)n_chains * n_sel;
  OMC_REQUIRE(n_series < 0x7fffffffll, "omc_rank_normalize: too many series");
  rank_sort_kernel<<<(unsigned)n_series, RN_NT, smem, (cudaStream_t)stream>>>(samples, n_iter, n_chains, size, n_sel,
                                                                            elem_stride, npad, sorted);
  OMC_LAUNCH_CHECK();
  const long long total = n_iter * n_series;
  const long long blocks = (total + RN_NT - 1) / RN_NT;
  rank_score_kernel<<<(unsigned)(blocks < 65535 ? blocks : 65535), RN_NT, 0, (cudaStream_t)stream>>>(
      samples, n_iter, n_chains, size, n_sel, elem_stride, pooled, sorted, z);
  OMC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
