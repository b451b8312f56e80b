"""Oracle (test infrastructure): numpy restatement of the reference's GMRF linear algebra.

Every function cites the reference lines it follows (relative to /root/reference/src/openmcmc/).  Dense work
uses numpy.linalg exactly like the reference; the banded (tridiagonal) path restates what the reference gets from
SuperLU with natural ordering and no pivoting (gmrf.py:514-516: L * diag(U)^(1/2)), which for a tridiagonal SPD
matrix is the plain Thomas-order Cholesky recurrence (SURVEY.md B.6 measured agreement 1e-15 on regular grids).
"""

import numpy as np
from scipy import special


# ----------------------------------------------------------------------------- dense
def cholesky(Q):
    """Lower Cholesky factor.  ref: gmrf.py:465-486 (dense branch -> np.linalg.cholesky)."""
    return np.linalg.cholesky(np.asarray(Q, dtype=np.float64))


def cho_solve(L, b):
    """Solve (L L') x = b.  ref: gmrf.py:437-462 (dense branch -> scipy.linalg.cho_solve = dpotrs)."""
    w = _solve_lower(L, b)
    return _solve_upper(L.T, w)


def _solve_lower(L, b):
    b = np.array(b, dtype=np.float64, copy=True)
    p = L.shape[0]
    for j in range(p):
        b[j] = b[j] / L[j, j]
        b[j + 1:] -= L[j + 1:, [j]] * b[j]
    return b


def _solve_upper(U, b):
    b = np.array(b, dtype=np.float64, copy=True)
    p = U.shape[0]
    for j in range(p - 1, -1, -1):
        b[j] = b[j] / U[j, j]
        b[:j] -= U[:j, [j]] * b[j]
    return b


def solve_upper(U, b):
    """ref: gmrf.py:414-434 — np.linalg.solve on the triangular L.T (LU with partial pivoting never swaps an
    upper-triangular matrix, so it is a back-substitution)."""
    return _solve_upper(U, b)


def sample_normal(mu, L, z):
    """x = mu + L^-T z.  ref: gmrf.py:29-61 (z = norm.rvs(size=[p, n]) is injected)."""
    z = np.asarray(z, dtype=np.float64).reshape(mu.shape[0], -1)
    return solve_upper(L.T, z) + mu


def sample_normal_canonical(b, Q, z):
    """Rue & Held Alg 2.5.  ref: gmrf.py:167-198.  Returns (x, mu, L)."""
    L = cholesky(Q)
    mu = cho_solve(L, b).reshape(b.shape)
    return sample_normal(mu, L, z), mu, L


def multivariate_normal_logpdf(x, mu, Q):
    """Sum over replicate columns of the MVN log-pdf in precision form.  ref: gmrf.py:321-348."""
    L = cholesky(Q)
    dim = L.shape[0]
    log_det = 2.0 * np.sum(np.log(np.diag(L)))
    qres = L.T @ (x - mu)
    return float(np.sum(0.5 * (log_det - dim * np.log(2 * np.pi) - np.sum(qres ** 2, axis=0))))


# ----------------------------------------------------------------------------- RW1 precision / tridiagonal
def precision_irregular_diagonals(s):
    """Main and off diagonal of the first-order random-walk precision.  ref: gmrf.py:375-411.

    Returns (d, e): d[i] = 1/del_{i-1} + 1/del_i, e[i] = -1/del_i (the reference stores them in a CSC matrix).
    """
    s = np.asarray(s, dtype=np.float64).squeeze()
    if s.size <= 1:
        return np.ones(1), np.zeros(0)
    dr = 1.0 / np.diff(s)
    d = np.append(np.append(dr[0], dr[:-1] + dr[1:]), dr[-1])
    return d, -dr


def tridiag_cholesky(d, e):
    """Natural-order Cholesky of the SPD tridiagonal matrix (main d, off e): returns (l, c) with L = diag(l) + sub(c).

    ref: gmrf.py:489-520 (splu, diag_pivot_thresh=0, no row/col permutation, then L * sqrt(diag U)).
    Returns None when a pivot is <= 0 (the reference then falls back to dense Cholesky, gmrf.py:515-518).
    """
    n = d.size
    l = np.empty(n)
    c = np.empty(max(n - 1, 0))
    piv = d[0]
    for i in range(n):
        if not piv > 0:
            return None
        l[i] = np.sqrt(piv)
        if i + 1 < n:
            c[i] = e[i] / l[i]
            piv = d[i + 1] - c[i] * c[i]
    return l, c


def tridiag_forward(l, c, b):
    """Solve L w = b for bidiagonal L (diag l, sub-diagonal c)."""
    w = np.array(b, dtype=np.float64, copy=True)
    n = w.shape[0]
    w[0] = w[0] / l[0]
    for i in range(1, n):
        w[i] = (w[i] - c[i - 1] * w[i - 1]) / l[i]
    return w


def tridiag_backward(l, c, b):
    """Solve L' x = b for bidiagonal L."""
    x = np.array(b, dtype=np.float64, copy=True)
    n = x.shape[0]
    x[n - 1] = x[n - 1] / l[n - 1]
    for i in range(n - 2, -1, -1):
        x[i] = (x[i] - c[i] * x[i + 1]) / l[i]
    return x


def tridiag_sample_canonical(d, e, b, z):
    """sample_normal_canonical for a tridiagonal precision: returns (x, mu, l, c).  ref: gmrf.py:167-198 sparse branch."""
    fac = tridiag_cholesky(d, e)
    if fac is None:
        raise np.linalg.LinAlgError("tridiagonal precision is not positive definite")
    l, c = fac
    mu = tridiag_backward(l, c, tridiag_forward(l, c, b))
    v = tridiag_backward(l, c, z)
    return mu + v, mu, l, c


def tridiag_quadform(d, e, r):
    """r' P r for tridiagonal P (main d, off e)."""
    r = np.asarray(r, dtype=np.float64).ravel()
    return float(np.sum(d * r * r) + 2.0 * np.sum(e * r[:-1] * r[1:]))


def tridiag_logdet(d, e):
    fac = tridiag_cholesky(d, e)
    if fac is None:
        return np.nan
    return float(2.0 * np.sum(np.log(fac[0])))


# ----------------------------------------------------------------------------- the reference's sparse call stack
def _tridiag_csc(d, e):
    from scipy import sparse

    return sparse.diags([e, d, e], offsets=[-1, 0, 1], format="csc")


def sparse_cholesky_scipy(Q):
    """ref: gmrf.py:489-520 — splu(diag_pivot_thresh=0, natural ordering, no row permutation) -> L * sqrt(diag U)."""
    from scipy import sparse
    from scipy.sparse import linalg as sla

    lu = sla.splu(Q, diag_pivot_thresh=0, options={"RowPerm": False, "ColPerm": False})
    return lu.L.dot(sparse.diags(lu.U.diagonal() ** 0.5)).tocsc()


def sparse_sample_normal_canonical(d, e, b, z):
    """Rue & Held Alg 2.5 on the sparse branch exactly as the reference runs it: one splu + two spsolve for the mean
    (gmrf.py:459-460) + one spsolve for the draw (gmrf.py:432).  Used by the CPU timing leg (bench.py); the tridiagonal
    restatement above is what the parity tests use.  Returns (x, L)."""
    from scipy.sparse import linalg as sla

    Q = _tridiag_csc(d, e)
    L = sparse_cholesky_scipy(Q)
    w = sla.spsolve(L, b)
    mu = sla.spsolve(L.T.tocsc(), w)
    return mu + sla.spsolve(L.T.tocsc(), z), L


def sparse_logdet(d, e):
    """log|Q| the way Normal.log_p gets it: a fresh sparse Cholesky per call (gmrf.py:339-342)."""
    L = sparse_cholesky_scipy(_tridiag_csc(d, e))
    return float(2.0 * np.sum(np.log(L.diagonal())))


# ----------------------------------------------------------------------------- truncated normal (scipy.stats.truncnorm)
# ref: gmrf.py:269-318 standardise (a, b) = ((lower-mean)/scale, (upper-mean)/scale) and call
# scipy.stats.truncnorm.rvs / logpdf.  scipy (_continuous_distns.py, truncnorm_gen) draws by inverse CDF of ONE
# uniform per variate using log-space mass computations; the restatement below follows that published algorithm.
def _log_diff(log_p, log_q):
    return log_p + np.log1p(-np.exp(log_q - log_p))


def _log_gauss_mass(a, b):
    """log(Phi(b) - Phi(a)) evaluated in the tail that keeps precision (scipy _log_gauss_mass)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a, b = np.broadcast_arrays(a, b)
    out = np.empty(a.shape)
    left = b <= 0
    right = a > 0
    central = ~(left | right)
    with np.errstate(all="ignore"):
        out[left] = _log_diff(special.log_ndtr(b[left]), special.log_ndtr(a[left]))
        out[right] = _log_diff(special.log_ndtr(-a[right]), special.log_ndtr(-b[right]))
        out[central] = np.log1p(-special.ndtr(a[central]) - special.ndtr(-b[central]))
    return out


def truncnorm_ppf(q, a, b):
    """Inverse CDF of the standard normal truncated to [a, b] (scipy truncnorm_gen._ppf)."""
    q, a, b = np.broadcast_arrays(np.asarray(q, float), np.asarray(a, float), np.asarray(b, float))
    out = np.empty(q.shape)
    case_left = a < 0
    with np.errstate(all="ignore"):
        lm = _log_gauss_mass(a, b)
        # left: log Phi(x) = logaddexp(log Phi(a), log q + log mass)
        lp = np.logaddexp(special.log_ndtr(a), np.log(q) + lm)
        out_left = special.ndtri_exp(lp)
        # right: log Phi(-x) = logaddexp(log Phi(-b), log(1-q) + log mass)
        lp2 = np.logaddexp(special.log_ndtr(-b), np.log1p(-q) + lm)
        out_right = -special.ndtri_exp(lp2)
    out[case_left] = out_left[case_left]
    out[~case_left] = out_right[~case_left]
    return out


def truncated_normal_rv(mean, scale, lower, upper, u):
    """ref: gmrf.py:269-292 with the uniform `u` behind truncnorm.rvs injected."""
    lower = -np.inf if lower is None else lower
    upper = np.inf if upper is None else upper
    a, b = (lower - mean) / scale, (upper - mean) / scale
    return truncnorm_ppf(u, a, b) * scale + mean


def truncated_normal_log_pdf(x, mean, scale, lower, upper):
    """ref: gmrf.py:295-318 -> truncnorm.logpdf = norm_logpdf(z) - log_gauss_mass(a,b) - log(scale), -inf outside."""
    lower = -np.inf if lower is None else lower
    upper = np.inf if upper is None else upper
    a, b = (lower - mean) / scale, (upper - mean) / scale
    z = (x - mean) / scale
    with np.errstate(all="ignore"):
        out = -0.5 * z * z - 0.5 * np.log(2 * np.pi) - _log_gauss_mass(a, b) - np.log(scale)
    out = np.where((z < a) | (z > b), -np.inf, out)
    return out


def gibbs_canonical_truncated_normal(b, Q, x, lower, upper, u):
    """One coordinate-wise Gibbs scan of x ~ N_c(Q^-1 b, Q^-1) truncated to [lower, upper] (Rue & Held lemma 2.1).

    ref: gmrf.py:201-266.  `u` [p] are the uniforms behind the p truncnorm.rvs calls (one per coordinate, in order);
    bounds are scalars or [p] (None => -inf / +inf, gmrf.py:236-243).  p == 1 draws from N(b/Q, 1/Q) truncated (:245-248).
    """
    Q = np.asarray(Q, dtype=np.float64)
    p = Q.shape[0]
    b = np.asarray(b, dtype=np.float64).reshape(p)
    x = np.array(x, dtype=np.float64).reshape(p)
    u = np.asarray(u, dtype=np.float64).reshape(p)
    lower = np.broadcast_to(-np.inf if lower is None else np.asarray(lower, dtype=np.float64).reshape(-1), (p,))
    upper = np.broadcast_to(np.inf if upper is None else np.asarray(upper, dtype=np.float64).reshape(-1), (p,))
    if p == 1:
        x[0] = truncated_normal_rv(b[0] / Q[0, 0], 1.0 / np.sqrt(Q[0, 0]), lower[0], upper[0], u[0])
        return x.reshape(p, 1)
    for i in range(p):
        q_ii = Q[i, i]
        v_i = 1.0 / q_ii
        cond_mean = v_i * (b[i] - Q[i, :] @ x + q_ii * x[i])
        x[i] = truncated_normal_rv(cond_mean, np.sqrt(v_i), lower[i], upper[i], u[i])
    return x.reshape(p, 1)

