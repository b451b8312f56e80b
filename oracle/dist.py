"""Oracle (test infrastructure): numpy restatement of the reference's log-densities and their derivatives.

ref files are relative to /root/reference/src/openmcmc/.  scipy.special supplies gammaln / xlogy exactly as the
reference reaches them through scipy.stats.
"""

import numpy as np
from scipy import special

from oracle import gmrf


def normal_log_p(x, mu, Q):
    """Normal.log_p without truncation.  ref: distribution/location_scale.py:145-167 -> gmrf.py:321-348."""
    return gmrf.multivariate_normal_logpdf(np.asarray(x, float), np.asarray(mu, float), np.asarray(Q, float))


def normal_log_p_from_ss(dim, scalar, logdet_P, ss):
    """Same value from the quadratic form with the un-scaled matrix: 0.5(dim log s + log|P| - dim log 2pi - s*ss)."""
    return 0.5 * (dim * np.log(scalar) + logdet_P - dim * np.log(2 * np.pi) - scalar * ss)


def gamma_log_p(x, shape, rate):
    """Gamma.log_p.  ref: distribution/distribution.py:241-261 (stats.gamma.logpdf(x, a, scale=1/rate) summed).

    scipy: logpdf = xlogy(a-1, y) - y - gammaln(a) - log(scale), y = x/scale; -inf for y < 0.
    """
    x, shape, rate = np.broadcast_arrays(np.asarray(x, float), np.asarray(shape, float), np.asarray(rate, float))
    scale = 1.0 / rate
    y = x / scale
    with np.errstate(all="ignore"):
        lp = special.xlogy(shape - 1.0, y) - y - special.gammaln(shape) - np.log(scale)
    lp = np.where(y < 0, -np.inf, lp)
    return float(np.sum(lp))


def poisson_log_p(k, rate):
    """Poisson.log_p.  ref: distribution/distribution.py:490-508 (xlogy(k, mu) - gammaln(k+1) - mu; -inf off-support)."""
    k, rate = np.broadcast_arrays(np.asarray(k, float), np.asarray(rate, float))
    with np.errstate(all="ignore"):
        lp = special.xlogy(k, rate) - special.gammaln(k + 1.0) - rate
    lp = np.where((k < 0) | (np.floor(k) != k), -np.inf, lp)
    return float(np.sum(lp))


def uniform_log_p(lower, upper, d, n=1):
    """Uniform.log_p.  ref: distribution/distribution.py:406-442."""
    rng = np.broadcast_to(np.asarray(upper, float) - np.asarray(lower, float), (d, 1))
    return -float(np.sum(np.log(rng))) * n


# ----------------------------------------------------------------------------- finite differences (reference default)
def grad_fd(log_p, x, step=1e-4):
    """Central differences of log_p w.r.t. every element of x.  ref: distribution/distribution.py:124-158."""
    x = np.asarray(x, float)
    g = np.full(x.size, np.nan)
    for k in range(x.size):
        xp = x.copy()
        xm = x.copy()
        xp[np.unravel_index(k, x.shape)] += step / 2
        xm[np.unravel_index(k, x.shape)] += -step / 2
        g[k] = (log_p(xp) - log_p(xm)) / step
    return g.reshape(x.shape)


def hessian_fd(grad, x, step=1e-4):
    """Hessian of the NEGATIVE log-pdf by differencing the gradient.  ref: distribution/distribution.py:160-198."""
    x = np.asarray(x, float)
    n = x.size
    H = np.full((n, n), np.nan)
    for k in range(n):
        xp = x.copy()
        xm = x.copy()
        xp[np.unravel_index(k, x.shape)] += step / 2
        xm[np.unravel_index(k, x.shape)] += -step / 2
        H[:, k] = (grad(xm) - grad(xp)).ravel() / step
    return H


# ----------------------------------------------------------------------------- analytic derivatives
def normal_grad_response(x, mu, Q):
    """grad = -Q r, Hessian (of -log p) = Q.  ref: location_scale.py:222-232."""
    r = np.asarray(x, float) - np.asarray(mu, float)
    return -np.asarray(Q, float) @ r, np.asarray(Q, float)


def normal_grad_linear_mean(y, X, beta, Q):
    """param enters the mean linearly through X: grad = X'Q r, H = n_rep X'QX.  ref: location_scale.py:234-242."""
    y = np.asarray(y, float)
    r = np.sum(y - X @ beta, axis=1, keepdims=True)
    gtp = X.T @ Q
    return gtp @ r, y.shape[1] * gtp @ X


def lognormal_log_p(x, mu, Q):
    """LogNormal.log_p: MVN log-pdf at log(x) minus sum(log x).  ref: location_scale.py:296-303"""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    with np.errstate(all="ignore"):
        lx = np.log(x)
        return normal_log_p(lx, mu, Q) - float(np.sum(lx))


def lognormal_grad_response(x, mu, Q):
    """Response branch: grad = -(1/x)(1 + Q r), r = log x - mu (location_scale.py:340-343); Hessian of -log p =
    diag(1/x) Q diag(1/x) - diag((1 + Q r)/x^2) (location_scale.py:383-399, n = 1 replicate)."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    mu = np.asarray(mu, dtype=np.float64).reshape(-1, 1)
    Q = np.asarray(Q, dtype=np.float64)
    r = np.log(x) - mu
    rec = 1.0 / x
    one_qr = 1.0 + Q @ r
    g = -rec * one_qr
    H = (rec * Q * rec.T) - np.diagflat(rec ** 2 * one_qr)
    return g, H


def normal_linear_log_p(y, X, theta, Q, transform=False):
    """Normal.log_p with mean X f(theta), f = exp when transform.  ref: location_scale.py:145-167, parameter.py:255-281"""
    f = np.exp(theta) if transform else np.asarray(theta, dtype=np.float64)
    return normal_log_p(np.asarray(y, float).reshape(-1, 1), X @ f.reshape(-1, 1), Q)


def normal_linear_grad(y, X, theta, Q, transform=False):
    """Mean-parameter branch: grad = J Q r, H = n_rep J Q J' with J = mean.grad(state, param) = X' (times exp(theta)
    row-wise when transformed).  ref: location_scale.py:234-250, parameter.py:199-228, 283-297"""
    theta = np.asarray(theta, dtype=np.float64).reshape(-1, 1)
    f = np.exp(theta) if transform else theta
    J = (np.exp(theta) * X.T) if transform else X.T
    r = np.asarray(y, float).reshape(-1, 1) - X @ f
    return J @ Q @ r, J @ Q @ J.T


def gamma_grad_response(x, shape, rate):
    """d/dx log Gamma(x; a, b) = (a-1)/x - b ; -d2/dx2 = (a-1)/x^2 (diagonal).  (analytic counterpart of the FD default)"""
    x, shape, rate = np.broadcast_arrays(np.asarray(x, float), np.asarray(shape, float), np.asarray(rate, float))
    return (shape - 1.0) / x - rate, np.diagflat((shape - 1.0) / x ** 2)


def poisson_grad_rate(k, rate):
    """d/dmu log Poisson(k; mu) = k/mu - 1 ; -d2/dmu2 = k/mu^2 (diagonal)."""
    k, rate = np.broadcast_arrays(np.asarray(k, float), np.asarray(rate, float))
    return k / rate - 1.0, np.diagflat(k / rate ** 2)
