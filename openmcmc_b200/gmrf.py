"""Host-side mirror of `openmcmc.gmrf` for the functions on the per-sweep path.  ref: gmrf.py

Same names and argument meaning as the reference; the arithmetic runs in the CUDA kernels behind the C-ABI
(`include/omc.h`).  Inside an `MCMC` run the samplers launch these kernels directly from the sweep plan; this module is the
one-call form for code that used the reference's `gmrf` functions on their own:

    precision_irregular / precision_temporal   RW1 precision of irregular locations (one-off host setup, gmrf.py:351-411)
    sample_normal_canonical / sample_normal     Rue & Held Alg. 2.5 draws, dense (p <= 64) or tridiagonal precision
    multivariate_normal_pdf                     log-density through the Cholesky factor (gmrf.py:321-348)
    truncated_normal_rv / truncated_normal_log_pdf   scipy's log-space truncnorm on the device (gmrf.py:269-318)

Keyword extras (not in the reference): `z` / `u` inject the standard-normal / uniform variates (the reference draws them
from numpy's global generator), `seed` keys the Philox generator otherwise.  There is no CPU path: every function raises
without a CUDA device.
"""

import numpy as np
import torch
from scipy import sparse

from openmcmc_b200 import kernels as K
from openmcmc_b200.engine import classify_matrix

_calls = 0   # Philox sweep counter of the one-call draws: successive calls give fresh variates


# ---------------------------------------------------------------------------------------------- precision builders
def precision_irregular(s, is_sparse: bool = True):
    """First-order random-walk precision for irregular locations `s` (Rue & Held 2005, pp. 97-99): with gaps
    d_i = s_{i+1} - s_i the diagonal is 1/d_{i-1} + 1/d_i (one term at either end) and the off-diagonals are -1/d_i.
    Sparse CSC by default, dense otherwise; a single location gives [[1]].  ref: gmrf.py:375-411"""
    s = np.asarray(s, dtype=np.float64).reshape(-1)
    if s.size <= 1:
        return np.array(1, ndmin=2)
    inv_gap = 1.0 / (s[1:] - s[:-1])
    main = np.zeros(s.size)
    main[:-1] += inv_gap
    main[1:] += inv_gap
    if is_sparse:
        return sparse.diags([-inv_gap, main, -inv_gap], offsets=[-1, 0, 1], format="csc")
    return np.diag(main) - np.diag(inv_gap, k=-1) - np.diag(inv_gap, k=1)


def precision_temporal(time, unit_length: float = 1.0, is_sparse: bool = True):
    """`precision_irregular` on a vector of times (pandas DatetimeIndex / DatetimeArray), measured in units of
    `unit_length` seconds from the earliest one.  ref: gmrf.py:351-372"""
    seconds = np.asarray((time - time.min()).total_seconds(), dtype=np.float64) / unit_length
    return precision_irregular(seconds, is_sparse=is_sparse)


# ---------------------------------------------------------------------------------------------- helpers
def _dev():
    return torch.device("cuda", K.init_device(None))


def _t(a, dev):
    return torch.as_tensor(np.array(a, dtype=np.float64, order="C", copy=True)).to(dev)


def _factor_and_solve(Q, b, z, seed, want_logdet=False):
    """One chain: L = chol(Q) (natural order), mu = Q^-1 b, x = mu + L^-T z on the device.  Returns (x, mu, logdet)."""
    global _calls
    dev = _dev()
    kind, main, off = classify_matrix(Q)
    n = Q.shape[0]
    b = np.zeros(n) if b is None else np.asarray(b, dtype=np.float64).reshape(-1)
    if b.size != n:
        raise ValueError(f"b has {b.size} entries for a {n} x {n} precision")
    dz = None if z is None else _t(np.asarray(z, dtype=np.float64).reshape(1, n), dev)
    sweep = torch.full((1,), _calls, dtype=torch.int64, device=dev)
    _calls += 1
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    if kind == "tridiag" or (kind in ("eye", "diag") and n > 64):
        pd = _t(np.ones(n) if kind == "eye" else main, dev)
        pe = _t(off if kind == "tridiag" else np.zeros(max(n - 1, 0)), dev)
        zeros, h = torch.zeros(1, n, dtype=torch.float64, device=dev), _t(b.reshape(1, n), dev)
        tau0 = torch.zeros(1, dtype=torch.float64, device=dev)
        x = torch.empty(1, n, dtype=torch.float64, device=dev)
        ld = torch.zeros(1, dtype=torch.float64, device=dev)
        ws = torch.zeros(K.tridiag_workspace(1, n), dtype=torch.uint8, device=dev)
        lib_args = dict(tau=K.vec(tau0, 1), y=K.vec(zeros, n), h=K.vec(h, n), x=x, status=status,
                        rng_=K.rng(seed=seed, sweep=sweep, site=1))
        # the posterior mean is the draw with z = 0 (Q = 1*P + 0*W, b = 1*h + 0)
        mu = torch.empty_like(x)
        K.tridiag_nn_draw(K.tridiag_args(1, n, pd, pe, ws, **{**lib_args, "x": mu}, debug_z=torch.zeros_like(x),
                                         logdet=ld if want_logdet else None))
        if z is not None or not want_logdet:
            K.tridiag_nn_draw(K.tridiag_args(1, n, pd, pe, ws, **lib_args, debug_z=dz))
        torch.cuda.synchronize()
        if int(status.item()) & 1:
            raise np.linalg.LinAlgError("Matrix is not positive definite")
        return x.cpu().numpy().reshape(n, 1), mu.cpu().numpy().reshape(n, 1), float(ld.item())
    if n > 64:
        raise NotImplementedError(f"{n} x {n} precision with bandwidth > 1 is not supported by the device path")
    Qd = np.eye(n) if kind == "eye" else np.diag(main) if kind == "diag" else np.asarray(main, dtype=np.float64)
    rec = torch.zeros(1, n * n + n + 2, dtype=torch.float64, device=dev)
    rec[0, : n * n] = _t(Qd.reshape(-1), dev)
    rec[0, n * n : n * n + n] = _t(b, dev)
    one, zero = torch.ones(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.float64, device=dev)
    x = torch.empty(1, n, dtype=torch.float64, device=dev)
    mu = torch.empty(1, n, dtype=torch.float64, device=dev)
    L = torch.empty(1, n, n, dtype=torch.float64, device=dev)
    K.nn_dense_draw(1, n, rec, K.vec(one, 1), K.MAT_EYE, K.vec(None), K.vec(zero, 1), K.vec(None), x,
                    K.rng(seed=seed, sweep=sweep, site=1), debug_z=dz, probe_mu=mu, probe_L=L, status=status)
    torch.cuda.synchronize()
    if int(status.item()) & 1:
        raise np.linalg.LinAlgError("Matrix is not positive definite")
    logdet = float(2.0 * torch.log(torch.diagonal(L[0])).sum().item()) if want_logdet else 0.0
    return x.cpu().numpy().reshape(n, 1), mu.cpu().numpy().reshape(n, 1), logdet


# ---------------------------------------------------------------------------------------------- draws
def sample_normal_canonical(b, Q=None, L=None, z=None, seed: int = 0):
    """x ~ N(Q^-1 b, Q^-1) by Rue & Held Alg. 2.5: L = chol(Q), L w = b, L' mu = w, L' v = z, x = mu + v.
    `Q` dense (p <= 64) or sparse diagonal / tridiagonal of any size; a precomputed factor `L` is accepted for
    signature compatibility (Q = L L' is re-formed).  Returns a p x 1 array.  ref: gmrf.py:167-198"""
    if Q is None:
        if L is None:
            raise ValueError("sample_normal_canonical needs Q or L")
        Ld = L.toarray() if sparse.issparse(L) else np.asarray(L, dtype=np.float64)
        Q = Ld @ Ld.T
    x, _, _ = _factor_and_solve(Q, b, z, seed)
    return x.reshape(np.shape(b)) if np.ndim(b) == 2 else x


def sample_normal(mu, Q=None, L=None, n: int = 1, z=None, seed: int = 0):
    """x ~ N(mu, Q^-1): mu + L^-T z, one column per draw.  Returns p x n.  ref: gmrf.py:29-61"""
    mu = np.asarray(mu, dtype=np.float64).reshape(-1, 1)
    if Q is None:
        if L is None:
            raise ValueError("sample_normal needs Q or L")
        Ld = L.toarray() if sparse.issparse(L) else np.asarray(L, dtype=np.float64)
        Q = Ld @ Ld.T
    cols = []
    for j in range(n):
        zj = None if z is None else np.asarray(z, dtype=np.float64).reshape(mu.size, -1)[:, j]
        x, _, _ = _factor_and_solve(Q, None, zj, seed)
        cols.append(x + mu)
    return np.concatenate(cols, axis=1)


def multivariate_normal_pdf(x, mu, Q, by_observation: bool = False):
    """Log-density of N(mu, Q^-1) at the columns of x (dim x n): 1/2 (log|Q| - dim log 2 pi - |L'(x - mu)|^2); summed
    over the observations unless `by_observation`.  log|Q| comes from the device factorisation, the quadratic forms
    (x - mu)' Q (x - mu) = |L'(x - mu)|^2 from `omc_quadform` / `omc_tridiag_matvec`.  ref: gmrf.py:321-348"""
    x = np.asarray(x, dtype=np.float64)
    x = x.reshape(-1, 1) if x.ndim == 1 else x
    mu = np.asarray(mu, dtype=np.float64).reshape(-1, 1)
    dim, n_obs = x.shape
    _, _, logdet = _factor_and_solve(Q, None, np.zeros(dim), 0, want_logdet=True)
    # quadratic forms (x_r - mu)' Q (x_r - mu) on the device, one observation per "chain" of the kernel
    dev = _dev()
    kind, main, off = classify_matrix(Q)
    xd, mud = _t(x.T, dev), _t(mu.reshape(-1), dev)                 # [n_obs, dim], [dim]
    quad = torch.empty(n_obs, dtype=torch.float64, device=dev)
    if kind == "tridiag":
        r = (xd - mud).contiguous()
        Qr = torch.empty_like(r)
        K.tridiag_matvec(_t(main, dev), _t(off, dev), K.vec(r, dim), n_obs, dim, Qr)
        quad = (r * Qr).sum(dim=1)
    else:
        cnt = torch.empty(n_obs, dtype=torch.float64, device=dev)
        if kind == "eye":
            K.quadform(n_obs, dim, K.vec(xd, dim), K.vec(mud, 0), K.MAT_EYE, K.vec(None), quad, cnt)
        elif kind == "diag":
            K.quadform(n_obs, dim, K.vec(xd, dim), K.vec(mud, 0), K.MAT_DIAG, K.vec(_t(main, dev), 0), quad, cnt)
        else:
            K.quadform(n_obs, dim, K.vec(xd, dim), K.vec(mud, 0), K.MAT_DENSE, K.vec(_t(np.asarray(main), dev), 0), quad, cnt)
    torch.cuda.synchronize()
    log_p = 0.5 * (logdet - dim * np.log(2.0 * np.pi) - quad.cpu().numpy().reshape(-1))
    return log_p if by_observation else np.sum(log_p)


# ---------------------------------------------------------------------------------------------- truncated normal
def _bcast(v, shape, default):
    return np.broadcast_to(np.asarray(default if v is None else v, dtype=np.float64), shape).reshape(-1)


def truncated_normal_rv(mean, scale, lower, upper, size=1, u=None, seed: int = 0):
    """Truncated-normal draws on [lower, upper] by the log-space inverse CDF scipy's `truncnorm.rvs` uses, from
    uniforms `u` (Philox when not given).  None bounds are infinite.  ref: gmrf.py:269-292"""
    if np.ndim(size) > 0:
        shape = tuple(int(v) for v in size)
    elif int(size) != 1:
        shape = (int(size),)
    else:
        shape = np.broadcast(np.asarray(mean, dtype=np.float64), np.asarray(scale, dtype=np.float64)).shape or (1,)
    n = int(np.prod(shape))
    dev = _dev()
    if u is None:
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(seed) + 7919 * (_calls_bump()))
        ud = torch.rand(n, dtype=torch.float64, device=dev, generator=gen)
    else:
        ud = _t(_bcast(u, shape, 0.5), dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    K.truncnorm_rv(_t(_bcast(mean, shape, 0.0), dev), _t(_bcast(scale, shape, 1.0), dev),
                   _t(_bcast(lower, shape, -np.inf), dev), _t(_bcast(upper, shape, np.inf), dev), ud, out)
    return out.cpu().numpy().reshape(shape)


def truncated_normal_log_pdf(x, mean, scale, lower, upper):
    """log-pdf of the truncated normal (scipy's `truncnorm.logpdf`: -inf outside the bounds).  ref: gmrf.py:295-318"""
    x = np.asarray(x, dtype=np.float64)
    shape = np.broadcast(x, np.asarray(mean, dtype=np.float64), np.asarray(scale, dtype=np.float64)).shape
    dev = _dev()
    out = torch.empty(int(np.prod(shape)), dtype=torch.float64, device=dev)
    K.truncnorm_logpdf(_t(_bcast(x, shape, 0.0), dev), _t(_bcast(mean, shape, 0.0), dev), _t(_bcast(scale, shape, 1.0), dev),
                       _t(_bcast(lower, shape, -np.inf), dev), _t(_bcast(upper, shape, np.inf), dev), out)
    return out.cpu().numpy().reshape(shape)


def _calls_bump():
    global _calls
    _calls += 1
    return _calls
