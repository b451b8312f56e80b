// Temporal-GMRF NormalNormal draw with a TRIDIAGONAL posterior precision (BASELINE configs[2]).
//
//   Q = lambda*P + tau*W  (P tridiagonal RW1 precision shared by all chains, W diagonal or identity)
//   b = lambda*P*mu0 + tau*W*y ;  L = chol(Q) in natural order ;  x = L^-T (L^-1 b + z)
//   ref: sampler.py:154-207 (NormalNormal.sample), gmrf.py:167-198 (sample_normal_canonical), gmrf.py:489-520
//        (sparse_cholesky: SuperLU, natural ordering, no pivoting == the Thomas-order recurrence), gmrf.py:437-462, 29-61
//
// The reference factorises sequentially (1e6 dependent steps per chain).  Here every recurrence is an exact parallel
// scan over tiles of TILE = 256 threads x 8 elements, single pass with decoupled look-back between tiles:
//   pivots   u_i = d_i - e_{i-1}^2 / u_{i-1}          Moebius maps compose as 2x2 matrix products (normalised)
//   forward  f_i = b_i - m_{i-1} f_{i-1}, m = e/u     affine maps (A, B)
//   backward x_i = g_i - m_i x_{i+1}, g = f/u + z/sqrt(u)   affine maps, tile aggregates published by the forward
//            kernel so that the backward kernel's tiles are independent
// (LDL' form: L = L~ D^1/2, so x = L~^-T (D^-1 L~^-1 b + D^-1/2 z) is the reference's mu + L^-T z exactly.)
// After the scan fixes the value entering a thread's 8 elements, the thread re-runs the *sequential* recurrence on
// them, so every stored number comes from the same operations as the sequential algorithm; the scan only supplies the
// boundary values (relative error ~1e-15, damped by the contraction of the pivot recurrence).
//
// HBM traffic per chain-iteration (n doubles each): forward reads y, writes g and m; backward reads g, m, y, writes x;
// P (pd, pe) is shared by all chains and read tile-major, so it stays in L2.  DESIGN.md gives the byte accounting.
#include "../../include/omc.h"
#include "omc_common.cuh"
#include "omc_internal.h"

namespace {

constexpr int TG_THREADS = 256;
constexpr int TG_E = 8;
constexpr int TG_TILE = TG_THREADS * TG_E;

struct __align__(16) TileRec {
  unsigned long long flag0, flag1;  // epoch*4 + {1: aggregate ready, 2: inclusive prefix ready}
  double m[4];                      // Moebius aggregate of the tile
  double u_end;                     // pivot of the last element of the tile
  double fa, fb, f_end;             // forward affine aggregate, f of the last element
  double ba, bb;                    // backward affine aggregate: x_first = ba * x_in + bb
  double x_in;                      // x of the first element of the NEXT tile (boundary value for the backward kernel)
  double part[3];                   // per-tile partial sums: log-det, prior quadratic form, likelihood quadratic form
};

struct Workspace {
  unsigned long long epoch;
  unsigned int ticket, done;
  unsigned int pad[12];
  // followed by: unsigned int chain_done[n_chains] (padded), TileRec rec[n_tiles][n_chains], scratch g/m
};

__host__ __device__ inline long long align_up(long long x, long long a) { return (x + a - 1) / a * a; }

struct Layout {
  long long off_chain_done, off_rec, off_g, off_m, total;
  long long n_tiles;
};
__host__ __device__ inline Layout make_layout(int n_chains, long long n) {
  Layout L;
  L.n_tiles = (n + TG_TILE - 1) / TG_TILE;
  L.off_chain_done = sizeof(Workspace);
  L.off_rec = align_up(L.off_chain_done + 2ll * n_chains * sizeof(unsigned int), 128);
  L.off_g = align_up(L.off_rec + L.n_tiles * n_chains * (long long)sizeof(TileRec), 128);
  L.off_m = L.off_g + align_up((long long)n_chains * n * 8, 128);
  L.total = L.off_m + align_up((long long)n_chains * n * 8, 128);
  return L;
}

// ---------------------------------------------------------------------------------------------- small operators
struct Mob { double a, b, c, d; };   // u -> (a u + b) / (c u + d)
struct Aff { double a, b; };         // v -> a v + b

__device__ __forceinline__ Mob mob_identity() { return Mob{1.0, 0.0, 0.0, 1.0}; }
__device__ __forceinline__ Mob mob_mul(const Mob& x, const Mob& y) {  // x after y
  return Mob{x.a * y.a + x.b * y.c, x.a * y.b + x.b * y.d, x.c * y.a + x.d * y.c, x.c * y.b + x.d * y.d};
}
__device__ __forceinline__ void mob_normalize(Mob& m) {
  const double mx = fmax(fmax(fabs(m.a), fabs(m.b)), fmax(fabs(m.c), fabs(m.d)));
  int e = ((__double2hiint(mx) >> 20) & 0x7ff) - 1023;
  e = max(-1000, min(1000, e));
  const double s = __hiloint2double((1023 - e) << 20, 0);
  m.a *= s; m.b *= s; m.c *= s; m.d *= s;
}
__device__ __forceinline__ double mob_apply(const Mob& m, double u) { return (m.a * u + m.b) / (m.c * u + m.d); }
__device__ __forceinline__ Aff aff_mul(const Aff& x, const Aff& y) { return Aff{x.a * y.a, x.a * y.b + x.b}; }  // x after y

__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_dn_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ double ld_cg(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ double vget(const omc_vec_t& v, int chain, long long i, double dflt) {
  return v.ptr ? v.ptr[(long long)chain * v.chain_stride + i] : dflt;
}

// 8 consecutive doubles of a per-chain / shared vector starting at element i0 (16-byte vector loads when aligned and
// fully inside the array, scalar loads at the ragged end); out-of-range entries get `fill`.
__device__ __forceinline__ void load8(const double* base, long long i0, long long n, double fill, double (&v)[TG_E]) {
  if (base == nullptr) {
#pragma unroll
    for (int k = 0; k < TG_E; ++k) v[k] = fill;
    return;
  }
  if (i0 + TG_E <= n && ((reinterpret_cast<uintptr_t>(base + i0) & 15) == 0)) {
    const double2* p = reinterpret_cast<const double2*>(base + i0);
#pragma unroll
    for (int k = 0; k < TG_E / 2; ++k) {
      const double2 t = __ldg(p + k);
      v[2 * k] = t.x;
      v[2 * k + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < TG_E; ++k) v[k] = (i0 + k < n) ? __ldg(base + i0 + k) : fill;
  }
}
__device__ __forceinline__ void store8(double* base, long long i0, long long n, const double (&v)[TG_E]) {
  if (i0 + TG_E <= n && ((reinterpret_cast<uintptr_t>(base + i0) & 15) == 0)) {
    double2* p = reinterpret_cast<double2*>(base + i0);
#pragma unroll
    for (int k = 0; k < TG_E / 2; ++k) p[k] = make_double2(v[2 * k], v[2 * k + 1]);
  } else {
#pragma unroll
    for (int k = 0; k < TG_E; ++k)
      if (i0 + k < n) base[i0 + k] = v[k];
  }
}

// CTA-wide inclusive scan of Moebius maps in thread order; returns the thread's EXCLUSIVE prefix (maps of all lower
// threads composed) and writes the tile aggregate to `total` (valid in all threads).  smem: 8 warps x 4 doubles + 4.
__device__ __forceinline__ Mob block_scan_mob(Mob mine, double* sm, Mob& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Mob inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Mob o{shfl_up_d(inc.a, d), shfl_up_d(inc.b, d), shfl_up_d(inc.c, d), shfl_up_d(inc.d, d)};
    if (lane >= d) {
      inc = mob_mul(inc, o);
      mob_normalize(inc);
    }
  }
  __syncthreads();
  if (lane == 31) { sm[warp * 4 + 0] = inc.a; sm[warp * 4 + 1] = inc.b; sm[warp * 4 + 2] = inc.c; sm[warp * 4 + 3] = inc.d; }
  __syncthreads();
  Mob wex = mob_identity();
  Mob tot = mob_identity();
  for (int w = 0; w < TG_THREADS / 32; ++w) {
    Mob ww{sm[w * 4 + 0], sm[w * 4 + 1], sm[w * 4 + 2], sm[w * 4 + 3]};
    if (w == warp) wex = tot;
    tot = mob_mul(ww, tot);
    mob_normalize(tot);
  }
  total = tot;
  Mob lex{shfl_up_d(inc.a, 1), shfl_up_d(inc.b, 1), shfl_up_d(inc.c, 1), shfl_up_d(inc.d, 1)};
  if (lane == 0) lex = mob_identity();
  Mob ex = mob_mul(lex, wex);
  mob_normalize(ex);
  return ex;
}

// Same for affine maps.  forward = true: thread order ascending (exclusive prefix = all LOWER threads);
// forward = false: descending (exclusive prefix = all HIGHER threads, composed from the top down).
template <bool FORWARD>
__device__ __forceinline__ Aff block_scan_aff(Aff mine, double* sm, Aff& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = TG_THREADS / 32;
  Aff inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    Aff o;
    if (FORWARD) { o.a = shfl_up_d(inc.a, d); o.b = shfl_up_d(inc.b, d); }
    else { o.a = shfl_dn_d(inc.a, d); o.b = shfl_dn_d(inc.b, d); }
    const bool ok = FORWARD ? (lane >= d) : (lane + d < 32);
    if (ok) inc = aff_mul(inc, o);
  }
  __syncthreads();
  if (lane == (FORWARD ? 31 : 0)) { sm[warp * 2] = inc.a; sm[warp * 2 + 1] = inc.b; }
  __syncthreads();
  Aff wex{1.0, 0.0}, tot{1.0, 0.0};
  for (int k = 0; k < NW; ++k) {
    const int w = FORWARD ? k : NW - 1 - k;
    Aff ww{sm[w * 2], sm[w * 2 + 1]};
    if (w == warp) wex = tot;
    tot = aff_mul(ww, tot);
  }
  total = tot;
  Aff lex;
  if (FORWARD) { lex.a = shfl_up_d(inc.a, 1); lex.b = shfl_up_d(inc.b, 1); if (lane == 0) lex = Aff{1.0, 0.0}; }
  else { lex.a = shfl_dn_d(inc.a, 1); lex.b = shfl_dn_d(inc.b, 1); if (lane == 31) lex = Aff{1.0, 0.0}; }
  return aff_mul(lex, wex);
}

__device__ __forceinline__ OmcRng to_rng(const omc_rng_t& r) {
  OmcRng o;
  o.seed = r.seed; o.sweep = r.sweep; o.chain_offset = r.chain_offset; o.site = r.site;
  return o;
}

// standard normals for elements [i0, i0+8) of a chain: Philox block index = element pair index (position based, so
// the draw does not depend on the tiling); blocks beyond 2^20 spill into the second counter word.
__device__ __forceinline__ void normals8(const OmcRng& r, unsigned int chain, long long i0, double (&z)[TG_E]) {
  const unsigned long long sw = r.sweep ? *r.sweep : 0ull;
  const uint2 key = make_uint2((unsigned int)r.seed, (unsigned int)(r.seed >> 32));
#pragma unroll
  for (int k = 0; k < TG_E / 2; ++k) {
    const unsigned long long blk = (unsigned long long)(i0 >> 1) + k;
    uint4 ctr = make_uint4((unsigned int)sw, (unsigned int)(sw >> 32) ^ (unsigned int)(blk >> 20) * 0x9E3779B9u,
                           r.chain_offset + chain, (r.site << 20) | (unsigned int)(blk & 0xFFFFFu));
    const uint4 b = philox4x32_10(ctr, key);
    const double u1 = omc_u01(b.x, b.y), u2 = omc_u01(b.z, b.w);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z[2 * k] = rad * c;
    z[2 * k + 1] = rad * s;
  }
}

// ---------------------------------------------------------------------------------------------- forward kernel
__global__ void __launch_bounds__(TG_THREADS) tridiag_forward_kernel(omc_tridiag_nn_t a, Workspace* ws, Layout L) {
  __shared__ double sm[40];
  __shared__ unsigned int s_ticket;
  __shared__ double s_bcast[2];
  __shared__ int s_bad;
  const int tid = threadIdx.x;
  const unsigned long long epoch = *reinterpret_cast<volatile unsigned long long*>(&ws->epoch);
  if (tid == 0) {
    s_ticket = atomicAdd(&ws->ticket, 1u);
    s_bad = 0;
  }
  __syncthreads();
  const long long work = s_ticket;
  const int C = a.n_chains;
  const long long tile = work / C;     // tile-major order: the 64 chains read the same tile of the shared P together
  const int chain = (int)(work % C);
  const long long n = a.n;
  char* wsb = reinterpret_cast<char*>(ws);
  TileRec* recs = reinterpret_cast<TileRec*>(wsb + L.off_rec);
  TileRec* rec = recs + tile * C + chain;
  unsigned int* chain_done = reinterpret_cast<unsigned int*>(wsb + L.off_chain_done);
  const unsigned long long FLAG_A = epoch * 4 + 1, FLAG_P = epoch * 4 + 2;

  const long long i0 = tile * TG_TILE + (long long)tid * TG_E;
  const double lam = vget(a.lambda, chain, 0, 1.0), tau = vget(a.tau, chain, 0, 1.0);
  double d[TG_E], e[TG_E], ep;  // d_i, e_i (couples i, i+1), ep = e_{i0-1}
  {
    double pd[TG_E], wv[TG_E];
    load8(a.pd, i0, n, 1.0, pd);
    load8(a.pe, i0, n - 1, 0.0, e);
    load8(a.w.ptr ? a.w.ptr + (long long)chain * a.w.chain_stride : nullptr, i0, n, 1.0, wv);
#pragma unroll
    for (int k = 0; k < TG_E; ++k) {
      d[k] = lam * pd[k] + tau * wv[k];
      e[k] *= lam;
    }
    ep = (a.pe && i0 > 0 && i0 - 1 < n - 1) ? lam * __ldg(a.pe + i0 - 1) : 0.0;
  }
  // ---- stage 0: pivots.  thread aggregate of the Moebius maps of its elements
  Mob agg = mob_identity();
  {
    double eprev = ep;
#pragma unroll
    for (int k = 0; k < TG_E; ++k) {
      if (i0 + k < n) {
        const double e2 = eprev * eprev;
        const Mob nm{d[k] * agg.a - e2 * agg.c, d[k] * agg.b - e2 * agg.d, agg.a, agg.b};
        agg = nm;
      }
      eprev = e[k];
    }
    mob_normalize(agg);
  }
  Mob tile_tot;
  const Mob ex = block_scan_mob(agg, sm, tile_tot);
  if (tid == 0) {
    double u_in = 1.0;
    if (tile > 0) {
      rec->m[0] = tile_tot.a; rec->m[1] = tile_tot.b; rec->m[2] = tile_tot.c; rec->m[3] = tile_tot.d;
      st_release(&rec->flag0, FLAG_A);
      Mob R = mob_identity();
      long long j = tile - 1;
      while (true) {
        TileRec* pr = recs + j * C + chain;
        unsigned long long f;
        do { f = ld_acquire(&pr->flag0); } while (f != FLAG_A && f != FLAG_P);
        if (f == FLAG_P) { u_in = mob_apply(R, ld_cg(&pr->u_end)); break; }
        Mob mj{ld_cg(&pr->m[0]), ld_cg(&pr->m[1]), ld_cg(&pr->m[2]), ld_cg(&pr->m[3])};
        R = mob_mul(R, mj);
        mob_normalize(R);
        --j;   // tile 0 always publishes FLAG_P, so the walk terminates
      }
    }
    s_bcast[0] = u_in;
  }
  __syncthreads();
  const double u_tile_in = s_bcast[0];
  double u_prev = mob_apply(ex, u_tile_in);   // pivot of element i0-1 (dummy 1.0 for the very first element)
  if (i0 == 0) u_prev = 1.0;
  double iu[TG_E];                            // 1 / u_i
  bool bad = false;
  double logdet = 0.0;
  const double iu_prev = 1.0 / u_prev;
  {
    double eprev = ep, iup = iu_prev;
#pragma unroll
    for (int k = 0; k < TG_E; ++k) {
      if (i0 + k < n) {
        const double u = d[k] - (eprev * eprev) * iup;
        if (!(u > 0.0)) bad = true;
        iup = 1.0 / u;
        iu[k] = iup;
        if (a.logdet) logdet += log(u);
        if (i0 + k == n - 1 || (tid == TG_THREADS - 1 && k == TG_E - 1)) {   // last element of the tile
          rec->u_end = u;
          st_release(&rec->flag0, FLAG_P);
        }
      } else {
        iu[k] = 1.0;
      }
      eprev = e[k];
    }
  }
  if (bad) s_bad = 1;
  // ---- stage 1: forward substitution  f_i = b_i - m_{i-1} f_{i-1},  b_i = tau w_i y_i + lambda h_i
  double f[TG_E];
  {
    double yv[TG_E], wv[TG_E], hv[TG_E];
    load8(a.y.ptr + (long long)chain * a.y.chain_stride, i0, n, 0.0, yv);
    load8(a.w.ptr ? a.w.ptr + (long long)chain * a.w.chain_stride : nullptr, i0, n, 1.0, wv);
    load8(a.h.ptr ? a.h.ptr + (long long)chain * a.h.chain_stride : nullptr, i0, n, 0.0, hv);
#pragma unroll
    for (int k = 0; k < TG_E; ++k) f[k] = tau * wv[k] * yv[k] + lam * hv[k];   // holds b_i for now
  }
  double mprev[TG_E];  // m_{i-1} = e_{i-1} / u_{i-1}
  Aff fagg{1.0, 0.0};
  {
    double eprev = ep, iup = iu_prev;
#pragma unroll
    for (int k = 0; k < TG_E; ++k) {
      mprev[k] = eprev * iup;
      if (i0 + k < n) {
        fagg.a = -mprev[k] * fagg.a;
        fagg.b = f[k] - mprev[k] * fagg.b;
      }
      eprev = e[k];
      iup = iu[k];
    }
  }
  Aff ftot;
  const Aff fex = block_scan_aff<true>(fagg, sm, ftot);
  if (tid == 0) {
    double f_in = 0.0;
    if (tile > 0) {
      rec->fa = ftot.a; rec->fb = ftot.b;
      st_release(&rec->flag1, FLAG_A);
      Aff R{1.0, 0.0};
      long long j = tile - 1;
      while (true) {
        TileRec* pr = recs + j * C + chain;
        unsigned long long fl;
        do { fl = ld_acquire(&pr->flag1); } while (fl != FLAG_A && fl != FLAG_P);
        if (fl == FLAG_P) { f_in = R.a * ld_cg(&pr->f_end) + R.b; break; }
        R = aff_mul(R, Aff{ld_cg(&pr->fa), ld_cg(&pr->fb)});
        --j;
      }
    }
    s_bcast[1] = f_in;
  }
  __syncthreads();
  {
    double fp = fex.a * s_bcast[1] + fex.b;   // f of element i0-1
#pragma unroll
    for (int k = 0; k < TG_E; ++k) {
      if (i0 + k < n) {
        fp = f[k] - mprev[k] * fp;
        f[k] = fp;
        if (i0 + k == n - 1 || (tid == TG_THREADS - 1 && k == TG_E - 1)) {
          rec->f_end = fp;
          st_release(&rec->flag1, FLAG_P);
        }
      }
    }
  }
  // ---- g_i = f_i / u_i + z_i / sqrt(u_i),  m_i = e_i / u_i ; backward tile aggregate
  double z[TG_E];
  if (a.debug_z) {
    const long long sw = a.rng.sweep ? (long long)(*a.rng.sweep) : 0ll;
    load8(a.debug_z + sw * a.debug_sweep_stride + (long long)chain * n, i0, n, 0.0, z);
  } else if (i0 < n) {
    normals8(to_rng(a.rng), chain, i0, z);
  }
  double g[TG_E], m[TG_E];
  Aff bagg{1.0, 0.0};  // x_{i0} = bagg.a * x_{i0+8} + bagg.b, built from the top element down
#pragma unroll
  for (int k = TG_E - 1; k >= 0; --k) {
    if (i0 + k < n) {
      const double su = sqrt(iu[k]);
      g[k] = f[k] * iu[k] + z[k] * su;
      m[k] = e[k] * iu[k];
      bagg.a = -m[k] * bagg.a;
      bagg.b = g[k] - m[k] * bagg.b;
      if (a.probe_l) a.probe_l[(long long)chain * n + i0 + k] = 1.0 / su;
      if (a.probe_c && i0 + k < n - 1) a.probe_c[(long long)chain * (n - 1) + i0 + k] = e[k] * su;
    } else {
      g[k] = 0.0;
      m[k] = 0.0;
    }
  }
  double* gs = reinterpret_cast<double*>(wsb + L.off_g) + (long long)chain * n;
  double* ms = reinterpret_cast<double*>(wsb + L.off_m) + (long long)chain * n;
  if (i0 < n) {
    store8(gs, i0, n, g);
    store8(ms, i0, n, m);
  }
  Aff btot;
  (void)block_scan_aff<false>(bagg, sm, btot);
  const double ld_tile = a.logdet ? omc_block_sum(logdet, sm) : 0.0;
  __syncthreads();
  if (tid == 0) {
    rec->ba = btot.a;
    rec->bb = btot.b;
    rec->part[0] = ld_tile;
    if (s_bad && a.status) atomicOr(&a.status[chain], OMC_STATUS_NOT_PD);
    __threadfence();
    const unsigned int dn = atomicAdd(&chain_done[chain], 1u);
    s_ticket = (dn == (unsigned int)(L.n_tiles - 1)) ? 1u : 0u;
  }
  __syncthreads();
  // ---- chain finisher: boundary values x_in for every tile of this chain (backward scan over the tile aggregates)
  if (s_ticket && tid < 32) {
    __threadfence();
    const int lane = tid;
    const long long T = L.n_tiles;
    const long long per = (T + 31) / 32;
    // lane l owns tiles [hi - per, hi) counted from the top: tile index t = T-1 - (l*per + q)
    Aff la{1.0, 0.0};
    for (long long q = 0; q < per; ++q) {
      const long long t = T - 1 - (lane * per + q);
      if (t >= 0) {
        TileRec* r = recs + t * C + chain;
        la = aff_mul(Aff{ld_cg(&r->ba), ld_cg(&r->bb)}, la);
      }
    }
    Aff inc = la;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      Aff o{shfl_up_d(inc.a, dd), shfl_up_d(inc.b, dd)};
      if (lane >= dd) inc = aff_mul(inc, o);
    }
    Aff exl{shfl_up_d(inc.a, 1), shfl_up_d(inc.b, 1)};
    if (lane == 0) exl = Aff{1.0, 0.0};
    double xin = exl.b;  // x entering the lane's top tile (x beyond the last element is 0 and its multiplier m is 0)
    double ldsum = 0.0;
    for (long long q = 0; q < per; ++q) {
      const long long t = T - 1 - (lane * per + q);
      if (t >= 0) {
        TileRec* r = recs + t * C + chain;
        r->x_in = xin;
        xin = ld_cg(&r->ba) * xin + ld_cg(&r->bb);
        ldsum += ld_cg(&r->part[0]);
      }
    }
    if (a.logdet) {
      // fixed summation order (lane partials, then a shuffle tree): deterministic
      ldsum = omc_warp_sum(ldsum);
      if (lane == 0) a.logdet[chain] = ldsum;
    }
    if (lane == 0) chain_done[chain] = 0;
  }
  // ---- last CTA of the launch re-arms the ticket counter and advances the epoch
  if (tid == 0) {
    __threadfence();
    const unsigned int dn = atomicAdd(&ws->done, 1u);
    if (dn == (unsigned int)(L.n_tiles * C - 1)) {
      ws->ticket = 0;
      ws->done = 0;
      ws->epoch = epoch + 1;
      __threadfence();
    }
  }
}

// ---------------------------------------------------------------------------------------------- backward kernel
// SOLVE = true : x_i = g_i - m_i x_{i+1} from the scratch written by the forward kernel, then the quadratic forms.
// SOLVE = false: x is given (quadratic forms of the current state only).
template <bool SOLVE>
__global__ void __launch_bounds__(TG_THREADS) tridiag_backward_kernel(omc_tridiag_nn_t a, Workspace* ws, Layout L) {
  __shared__ double sm[40];
  __shared__ unsigned int s_last;
  const int tid = threadIdx.x;
  const int C = a.n_chains;
  const long long T = L.n_tiles;
  const long long work = blockIdx.x;
  const long long tile = T - 1 - work / C;   // top tiles first: they were written last by the forward kernel (L2)
  const int chain = (int)(work % C);
  const long long n = a.n;
  char* wsb = reinterpret_cast<char*>(ws);
  TileRec* recs = reinterpret_cast<TileRec*>(wsb + L.off_rec);
  TileRec* rec = recs + tile * C + chain;
  unsigned int* chain_done = reinterpret_cast<unsigned int*>(wsb + L.off_chain_done) + C;
  const long long i0 = tile * TG_TILE + (long long)tid * TG_E;
  double* xg = a.x + (long long)chain * n;
  double x[TG_E];
  double x_next;   // x_{i0+8}
  if (SOLVE) {
    const double* gs = reinterpret_cast<const double*>(wsb + L.off_g) + (long long)chain * n;
    const double* ms = reinterpret_cast<const double*>(wsb + L.off_m) + (long long)chain * n;
    double g[TG_E], m[TG_E];
    load8(gs, i0, n, 0.0, g);
    load8(ms, i0, n, 0.0, m);
    Aff bagg{1.0, 0.0};
#pragma unroll
    for (int k = TG_E - 1; k >= 0; --k) {
      bagg.a = -m[k] * bagg.a;
      bagg.b = g[k] - m[k] * bagg.b;
    }
    Aff btot;
    const Aff bex = block_scan_aff<false>(bagg, sm, btot);
    x_next = bex.a * rec->x_in + bex.b;
    double xn = x_next;
#pragma unroll
    for (int k = TG_E - 1; k >= 0; --k) {
      xn = g[k] - m[k] * xn;
      x[k] = xn;
    }
    if (i0 < n) store8(xg, i0, n, x);
  } else {
    load8(xg, i0, n, 0.0, x);
    x_next = (i0 + TG_E < n) ? xg[i0 + TG_E] : 0.0;
  }
  // ---- quadratic forms: (x-mu0)' P (x-mu0) and (y-x)' W (y-x)
  double ssp = 0.0, ssl = 0.0;
  {
    double pd[TG_E], pe[TG_E], mu[TG_E], yv[TG_E], wv[TG_E];
    load8(a.pd, i0, n, 0.0, pd);
    load8(a.pe, i0, n - 1, 0.0, pe);
    const double* mup = a.mu0.ptr ? a.mu0.ptr + (long long)chain * a.mu0.chain_stride : nullptr;
    load8(mup, i0, n, 0.0, mu);
    load8(a.y.ptr ? a.y.ptr + (long long)chain * a.y.chain_stride : nullptr, i0, n, 0.0, yv);
    load8(a.w.ptr ? a.w.ptr + (long long)chain * a.w.chain_stride : nullptr, i0, n, 1.0, wv);
    const double mu_next = (mup && i0 + TG_E < n) ? __ldg(mup + i0 + TG_E) : 0.0;
    double r_next = x_next - mu_next;
#pragma unroll
    for (int k = TG_E - 1; k >= 0; --k) {
      if (i0 + k < n) {
        const double r = x[k] - mu[k];
        ssp += pd[k] * r * r + 2.0 * pe[k] * r * r_next;
        const double q = yv[k] - x[k];
        ssl += wv[k] * q * q;
        r_next = r;
      }
    }
  }
  ssp = omc_block_sum(ssp, sm);
  ssl = omc_block_sum(ssl, sm);
  if (tid == 0) {
    rec->part[1] = ssp;
    rec->part[2] = ssl;
    __threadfence();
    const unsigned int dn = atomicAdd(&chain_done[chain], 1u);
    s_last = (dn == (unsigned int)(T - 1)) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && tid < 32) {   // deterministic per-chain reduction of the tile partials
    __threadfence();
    double sp = 0.0, sl = 0.0;
    for (long long t = tid; t < T; t += 32) {
      TileRec* r = recs + t * C + chain;
      sp += ld_cg(&r->part[1]);
      sl += ld_cg(&r->part[2]);
    }
    sp = omc_warp_sum(sp);
    sl = omc_warp_sum(sl);
    if (tid == 0) {
      if (a.ss_prior) a.ss_prior[chain] = sp;
      if (a.ss_lik) a.ss_lik[chain] = sl;
      chain_done[chain] = 0;
    }
  }
}

__global__ void tridiag_ws_init_kernel(Workspace* ws, Layout L, int n_chains) {
  // zero the header, the per-chain counters and every tile flag
  const long long words = L.off_g / 8;
  unsigned long long* p = reinterpret_cast<unsigned long long*>(ws);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (long long)gridDim.x * blockDim.x)
    p[i] = 0ull;
}

__global__ void tridiag_matvec_kernel(const double* pd, const double* pe, omc_vec_t v, int n_chains, long long n,
                                      double* out) {
  const long long total = (long long)n_chains * n;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int chain = (int)(t / n);
    const long long i = t % n;
    const double* vp = v.ptr + (long long)chain * v.chain_stride;
    double s = pd[i] * vp[i];
    if (i > 0) s += pe[i - 1] * vp[i - 1];
    if (i + 1 < n) s += pe[i] * vp[i + 1];
    out[t] = s;
  }
}

int check_args(const omc_tridiag_nn_t* a, const char* who) {
  OMC_REQUIRE(a && a->pd && a->workspace, "%s: null argument", who);
  OMC_REQUIRE(a->n_chains >= 1 && a->n >= 1, "%s: n_chains=%d n=%lld", who, a->n_chains, a->n);
  OMC_REQUIRE((long long)a->n_chains * ((a->n + TG_TILE - 1) / TG_TILE) < 0x7fffffffll, "%s: too many tiles", who);
  return 0;
}

}  // namespace

extern "C" {

int omc_tridiag_workspace(int n_chains, long long n, long long* bytes) {
  OMC_REQUIRE(n_chains >= 1 && n >= 1 && bytes, "omc_tridiag_workspace: bad argument");
  *bytes = make_layout(n_chains, n).total;
  return 0;
}

int omc_tridiag_workspace_init(void* workspace, int n_chains, long long n, void* stream) {
  OMC_REQUIRE(workspace, "omc_tridiag_workspace_init: null workspace");
  const Layout L = make_layout(n_chains, n);
  tridiag_ws_init_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<Workspace*>(workspace), L, n_chains);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_tridiag_nn_draw(const omc_tridiag_nn_t* a, void* stream) {
  if (int rc = check_args(a, "omc_tridiag_nn_draw")) return rc;
  OMC_REQUIRE(a->y.ptr, "omc_tridiag_nn_draw: y missing");
  const Layout L = make_layout(a->n_chains, a->n);
  const unsigned int grid = (unsigned int)(L.n_tiles * a->n_chains);
  Workspace* ws = reinterpret_cast<Workspace*>(a->workspace);
  tridiag_forward_kernel<<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(*a, ws, L);
  OMC_LAUNCH_CHECK();
  if (a->x) {
    tridiag_backward_kernel<true><<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(*a, ws, L);
    OMC_LAUNCH_CHECK();
  }
  return 0;
}

int omc_tridiag_quadforms(const omc_tridiag_nn_t* a, void* stream) {
  if (int rc = check_args(a, "omc_tridiag_quadforms")) return rc;
  OMC_REQUIRE(a->x, "omc_tridiag_quadforms: x missing");
  const Layout L = make_layout(a->n_chains, a->n);
  const unsigned int grid = (unsigned int)(L.n_tiles * a->n_chains);
  tridiag_backward_kernel<false><<<grid, TG_THREADS, 0, (cudaStream_t)stream>>>(*a, reinterpret_cast<Workspace*>(a->workspace), L);
  OMC_LAUNCH_CHECK();
  return 0;
}

int omc_tridiag_matvec(const double* pd, const double* pe, omc_vec_t v, int n_chains, long long n, double* out,
                       void* stream) {
  OMC_REQUIRE(pd && v.ptr && out && n_chains >= 1 && n >= 1 && (n == 1 || pe), "omc_tridiag_matvec: bad argument");
  const long long total = (long long)n_chains * n;
  const unsigned int blocks = (unsigned int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
  tridiag_matvec_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pd, pe, v, n_chains, n, out);
  OMC_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
