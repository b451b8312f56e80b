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


def normal_normal_dense_truncated(G, g, tau, P0, lam, mu0, x, lower, upper, u):
    """NormalNormal.sample with a truncated Normal prior: same (Q, b) as normal_normal_dense, then ONE coordinate-wise
    Gibbs scan from the current value x.  ref: sampler.py:196-205 -> gmrf.gibbs_canonical_truncated_normal."""
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
    return {"Q": Q, "b": b, "x": gmrf.gibbs_canonical_truncated_normal(b, Q, x, lower, upper, u)}


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


def gmrf_normal_normal(pd, pe, w, y, mu0, lam, tau, z):
    """NormalNormal.sample for a field with prior N(mu0, (lam P)^-1), P tridiagonal (main pd, off pe), and likelihood
    y ~ N(b, (tau W)^-1), W = diag(w).

    ref: sampler.py:154-207: Q = lam P + tau W (Identity-mean likelihood Hessian = Q_rsp, location_scale.py:234-242
    with grad = I), b = lam P mu0 + tau W y; gmrf.sample_normal_canonical on the sparse branch (gmrf.py:167-198,
    489-520).  Returns dict(x, mu, l, c, d, e, b).
    """
    pd, pe, w, y, mu0 = (np.asarray(v, dtype=np.float64).ravel() for v in (pd, pe, w, y, mu0))
    d = lam * pd + tau * w
    e = lam * pe
    Pmu = pd * mu0
    Pmu[:-1] += pe * mu0[1:]
    Pmu[1:] += pe * mu0[:-1]
    b = lam * Pmu + tau * w * y
    x, mu, l, c = gmrf.tridiag_sample_canonical(d, e, b, np.asarray(z, dtype=np.float64).ravel())
    return {"x": x, "mu": mu, "l": l, "c": c, "d": d, "e": e, "b": b}


def gibbs_gmrf_sweep(pd, pe, w, y, mu0, state, z, g_lam, g_tau, order=("b", "lambda", "tau")):
    """One sweep of the example-4 Gibbs sampler.  ref: examples/4_GMRF_smoother.ipynb; mcmc.py:98-100.
    state: b (n,), lambda, tau, a_lam, b_lam, a_tau, b_tau.  cnt = number of positive diagonal entries (sampler.py:283)."""
    s = dict(state)
    pd, pe, w, y, mu0 = (np.asarray(v, dtype=np.float64).ravel() for v in (pd, pe, w, y, mu0))
    for name in order:
        if name == "b":
            s["b"] = gmrf_normal_normal(pd, pe, w, y, mu0, s["lambda"], s["tau"], z)["x"]
        elif name == "lambda":
            ss = gmrf.tridiag_quadform(pd, pe, s["b"] - mu0)
            s["lambda"], _, _ = normal_gamma(s["a_lam"], s["b_lam"], ss, float(np.sum(pd > 0)), g_lam)
        elif name == "tau":
            r = y - s["b"]
            s["tau"], _, _ = normal_gamma(s["a_tau"], s["b_tau"], float(np.sum(w * r * r)), float(np.sum(w > 0)), g_tau)
    return s


def gmrf_log_post(pd, pe, w, y, mu0, s):
    """Model.log_p of the example-4 model.  ref: model.py:57-70, location_scale.py:145-167 -> gmrf.py:321-348,
    distribution.py:241-261.  log|lam P| = n log lam + log|P|."""
    from oracle import dist

    pd, pe, w, y, mu0 = (np.asarray(v, dtype=np.float64).ravel() for v in (pd, pe, w, y, mu0))
    n = pd.size
    r = y - s["b"]
    lp = dist.normal_log_p_from_ss(n, s["tau"], float(np.sum(np.log(w))), float(np.sum(w * r * r)))
    lp += dist.normal_log_p_from_ss(n, s["lambda"], gmrf.tridiag_logdet(pd, pe), gmrf.tridiag_quadform(pd, pe, s["b"] - mu0))
    lp += dist.gamma_log_p(s["lambda"], s["a_lam"], s["b_lam"]) + dist.gamma_log_p(s["tau"], s["a_tau"], s["b_tau"])
    return lp


# ----------------------------------------------------------------------------------------------- mixture model (f2)
def mixture_allocation(x, mu, tau, prob, u):
    """MixtureAllocation.sample: gam_ik = prob_k N(x_i; mu_k, 1/tau_k) normalised over k; z_i = #{k: u_i > cumsum_k}.
    ref: sampler.py:331-355.  x [n], mu / tau [K], prob [1 or n, K], u [n]."""
    from scipy.stats import norm

    x = np.asarray(x, dtype=np.float64).reshape(-1, 1)
    mu, tau = np.asarray(mu, float).ravel(), np.asarray(tau, float).ravel()
    prob = np.atleast_2d(np.asarray(prob, float))
    gam = np.empty((x.shape[0], mu.size))
    for k in range(mu.size):
        gam[:, [k]] = prob[:, [k]] * norm.pdf(x, loc=mu[k], scale=1 / np.sqrt(tau[k]))
    gam = gam / np.sum(gam, axis=1).reshape(-1, 1)
    return np.sum(np.asarray(u, float).reshape(-1, 1) > np.cumsum(gam, axis=1), axis=1).astype(np.float64)


def mixture_stats(x, mu, z, K):
    """Per component: n_k, sum x, sum (x - mu_k)^2 over the observations allocated to it (parameter.py:522-538)."""
    x, z = np.asarray(x, float).ravel(), np.asarray(z).ravel().astype(int)
    mu = np.asarray(mu, float).ravel()
    out = np.zeros((K, 3))
    for k in range(K):
        m = z == k
        out[k] = (m.sum(), x[m].sum(), np.sum((x[m] - mu[k]) ** 2))
    return out


def mixture_normal_gamma(x, mu, z, a0, b0, g):
    """NormalGamma K-loop: a*_k = a_k + n_k/2, b*_k = b_k + S2_k/2, tau_k = g_k / b*_k.  ref: sampler.py:272-288."""
    K = np.asarray(mu).size
    st = mixture_stats(x, mu, z, K)
    a_post = np.broadcast_to(np.asarray(a0, float).ravel(), (K,)) + st[:, 0] / 2
    b_post = np.broadcast_to(np.asarray(b0, float).ravel(), (K,)) + st[:, 2] / 2
    return np.asarray(g, float).ravel() / b_post, a_post, b_post


def mixture_normal_log_p(x, mu, tau, z):
    """Normal.log_p with MixtureParameterVector mean / MixtureParameterMatrix precision.  ref: location_scale.py:145-167"""
    x, z = np.asarray(x, float).ravel(), np.asarray(z).ravel().astype(int)
    m, t = np.asarray(mu, float).ravel()[z], np.asarray(tau, float).ravel()[z]
    return float(0.5 * (np.sum(np.log(t)) - x.size * np.log(2 * np.pi) - np.sum(t * (x - m) ** 2)))


def categorical_log_p(z, prob):
    """Categorical.log_p for a (p, 1) response: sum_i log prob[i or 0, z_i].  ref: distribution.py:318-345"""
    z = np.asarray(z).ravel().astype(int)
    prob = np.atleast_2d(np.asarray(prob, float))
    rows = np.arange(z.size) if prob.shape[0] > 1 else np.zeros(z.size, dtype=int)
    return float(np.sum(np.log(prob[rows, z])))


def gibbs_mixture_sweep(X, y, w, state, prob, a_tau, b_tau, z_beta, g, u):
    """One sweep of the standard mixture model (tests/test_sampler.py:113-147 form): NormalNormal(beta) with the
    mixture prior N(mu[z], diag(tau[z])^-1), NormalGamma(tau) K-loop, MixtureAllocation(z)."""
    s = dict(state)
    G, gv, _, _ = regression_suffstats(X, y, w)
    zi = np.asarray(s["z"]).ravel().astype(int)
    s["beta"] = normal_normal_dense(G, gv, 1.0, np.asarray(s["tau"], float).ravel()[zi], 1.0,
                                    np.asarray(s["mu"], float).ravel()[zi], z_beta)["x"]
    s["tau"], _, _ = mixture_normal_gamma(s["beta"], s["mu"], s["z"], a_tau, b_tau, g)
    s["z"] = mixture_allocation(s["beta"], s["mu"], s["tau"], prob, u)
    return s



def multi_likelihood_sweep(terms, state, P0, mu0, z, g, a0=1e-3, b0=1e-3):
    """One Gibbs sweep of a NormalNormal update with SEVERAL likelihood terms followed by the NormalGamma updates of
    every precision scalar (the samplers in the order beta, tau_1 .. tau_L, lambda).

    ref: sampler.py:179-192 -- Q = Q_prior + sum_l tau_l X_l' W_l X_l, b = Q_prior mu0 + sum_l tau_l X_l' W_l y_l (an
    Identity-mean term is X_l = I, sampler.py:187-188); sampler.py:252-288 for the Gamma updates.
    terms: list of (X, y, w or None); state: dict(beta, taus [L], lam); g: the L + 1 injected standard-gamma variates.
    Returns the new state."""
    p = state["beta"].size
    G, gv = np.zeros((p, p)), np.zeros((p, 1))
    for (X, y, w), tau in zip(terms, state["taus"]):
        Gl, gl, _, _ = regression_suffstats(X, y, w)
        G += float(tau) * Gl
        gv += float(tau) * gl
    beta = normal_normal_dense(G, gv, 1.0, P0, state["lam"], mu0, z)["x"]
    taus = []
    for (X, y, w), gk in zip(terms, g[:-1]):
        _, _, rss, cnt = regression_suffstats(X, y, w, beta)
        taus.append(normal_gamma(a0, b0, rss, cnt, gk)[0])
    ss, cnt = quadform(P0, beta, mu0)
    lam = normal_gamma(a0, b0, ss, cnt, g[-1])[0]
    return {"beta": beta, "taus": taus, "lam": lam}
