"""Oracle (test infrastructure): numpy restatement of the conjugate Gibbs samplers.

ref files are relative to /root/reference/src/openmcmc/.  Randomness is injected (z = the standard normals behind
scipy.stats.norm.rvs, g = the standard-gamma variates behind scipy.stats.gamma.rvs, i.e. gamma.rvs(a, scale=s)
== standard_gamma(a) * s), mirroring how the reference's own tests patch `rvs` (tests/test_sampler.py:211-215).
"""

import numpy as np

from oracle import gmrf


def regression_suffstats(X, y, w=None, beta=None):
    """G = X'WX, g = X'Wy, rss = (y-Xb)'W(y-Xb), cnt = #(w>0).

    ref: location_scale.py:234-242 (grad_times_prec @ grad_param.T with grad_param = X.T, parameter.py:228),
         sampler.py:190-192 (A.T @ Q_rsp @ (y - predictor_exclude)), sampler.py:275-284 (residual, quadratic form,
         count of positive diagonal entries).  The ScaledMatrix scalar (tau) is applied by the caller.
    """
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1, 1)
    n = X.shape[0]
    if w is None:
        XtW = X.T
        cnt = float(n)
    else:
        w = np.asarray(w, dtype=np.float64).reshape(-1)
        XtW = X.T * w
        cnt = float(np.sum(w > 0))
    G = XtW @ X
    g = XtW @ y
    if beta is None:
        r = y
    else:
        r = y - X @ np.asarray(beta, dtype=np.float64).reshape(-1, 1)
    rss = float((r.T @ (r if w is None else w.reshape(-1, 1) * r)).item())
    return G, g, rss, cnt


def normal_normal_dense(G, g, tau, P0, lam, mu0, z):
    """NormalNormal.sample for one Normal likelihood with linear mean and one Normal prior.

    ref: sampler.py:154-207: Q = lam*P0 + tau*G ; b = lam*P0 @ mu0 + tau*g ; gmrf.sample_normal_canonical(b, Q).
    Returns dict(Q, b, L, mu, x).
    """
    p = G.shape[0]
    P0 = np.asarray(P0, dtype=np.float64)
    if P0.ndim == 0:
        P0 = float(P0) * np.eye(p)
    elif P0.ndim == 1:
        P0 = np.diag(P0)
    mu0 = np.zeros((p, 1)) if mu0 is None else np.asarray(mu0, dtype=np.float64).reshape(p, 1)
    Q_prior = float(lam) * P0
    Q = Q_prior + float(tau) * G
    b = Q_prior @ mu0 + float(tau) * np.asarray(g, dtype=np.float64).reshape(p, 1)
    x, mu, L = gmrf.sample_normal_canonical(b, Q, np.asarray(z, dtype=np.float64).reshape(p, 1))
    return {"Q": Q, "b": b, "L": L, "mu": mu, "x": x}


def normal_gamma(a0, b0, ss, cnt, g):
    """NormalGamma.sample for a scalar precision.  ref: sampler.py:252-288.

    a* = a0 + cnt/2, b* = b0 + ss/2, sample = g / b*  (g ~ Gamma(a*, 1) injected); b* == 0 -> scale = inf (:285-286).
    Returns (sample, a*, b*).
    """
    a_post = float(a0) + cnt / 2.0
    b_post = float(b0) + ss / 2.0
    scale = np.inf if b_post == 0 else 1.0 / b_post
    return g * scale, a_post, b_post


def quadform(P, x, mu=None):
    """(x-mu)' P (x-mu) and #(diag P > 0) for scalar / diagonal / dense P.  ref: sampler.py:276-284."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    r = x if mu is None else x - np.asarray(mu, dtype=np.float64).reshape(-1, 1)
    P = np.asarray(P, dtype=np.float64)
    p = r.shape[0]
    if P.ndim == 0:
        return float(P) * float((r.T @ r).item()), float(p if P > 0 else 0)
    if P.ndim == 1:
        return float(np.sum(P * r[:, 0] ** 2)), float(np.sum(P > 0))
    return float((r.T @ P @ r).item()), float(np.sum(np.diag(P) > 0))


def gibbs_regression_sweep(X, y, state, z, g_tau, g_lam, P0=1.0, mu0=None, w=None, order=("beta", "tau", "lambda")):
    """One sweep of the example-3 Gibbs sampler (NormalNormal beta, NormalGamma tau, NormalGamma lambda).

    ref: examples/3_linear_regression.ipynb model; mcmc.py:98-100 sweep order = sampler list order.
    `state` holds beta (p,1), tau, lambda, a_tau, b_tau, a_lambda, b_lambda.  Returns the new state (copy).
    """
    s = dict(state)
    for name in order:
        if name == "beta":
            G, gv, _, _ = regression_suffstats(X, y, w)
            s["beta"] = normal_normal_dense(G, gv, s["tau"], P0, s["lambda"], mu0, z)["x"]
        elif name == "tau":
            _, _, rss, cnt = regression_suffstats(X, y, w, s["beta"])
            s["tau"], _, _ = normal_gamma(s["a_tau"], s["b_tau"], rss, cnt, g_tau)
        elif name == "lambda":
            ss, cnt = quadform(P0, s["beta"], mu0)
            s["lambda"], _, _ = normal_gamma(s["a_lambda"], s["b_lambda"], ss, cnt, g_lam)
    return s
