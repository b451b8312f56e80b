"""Oracle (test infrastructure): numpy restatement of the Metropolis-Hastings family.

ref files are relative to /root/reference/src/openmcmc/.  The conditional model of the sampled parameter is a list of
`Term`s mirroring the distributions the reference would hold in `sampler.model` (sampler.py:53-55); randomness is
injected (z = standard normals behind norm.rvs, tn_u = the uniforms behind truncnorm.rvs, u = accept uniforms), the way
the reference's tests patch scipy.stats.*.rvs (tests/test_sampler.py:211-215).
"""

from dataclasses import dataclass

import numpy as np

from oracle import dist, gmrf


@dataclass
class Term:
    """One member distribution of the conditional model, as a function of the sampled parameter theta (p, n)."""

    kind: str             # 'poisson_rate' | 'gamma_response' | 'normal_response' | 'uniform_response' |
                          # 'lognormal_response' (p1 = mean, Q) | 'normal_linear' (data = y, X, Q, transform)
    data: np.ndarray = None     # poisson counts
    p1: np.ndarray = None       # gamma shape | normal mean | uniform lower
    p2: np.ndarray = None       # gamma rate | uniform upper
    Q: np.ndarray = None        # normal precision (scaled, dense)
    dom_lo: float = -np.inf
    dom_hi: float = np.inf
    analytic_in_reference: bool = False   # Normal has analytic derivatives (location_scale.py:190-250)
    X: np.ndarray = None        # normal_linear: design matrix
    transform: bool = False     # normal_linear: mean = X exp(theta) (parameter.py:232-297)

    def log_p(self, theta):
        theta = np.asarray(theta, float)
        if self.kind == "poisson_rate":
            return dist.poisson_log_p(self.data.reshape(theta.shape), theta)          # distribution.py:490-508
        if self.kind == "gamma_response":
            return dist.gamma_log_p(theta, np.broadcast_to(self.p1, theta.shape), np.broadcast_to(self.p2, theta.shape))
        if self.kind == "normal_response":                                            # location_scale.py:145-167
            if np.any(theta < self.dom_lo) or np.any(theta > self.dom_hi):
                return -np.inf
            return dist.normal_log_p(theta.reshape(-1, 1), np.broadcast_to(self.p1, theta.shape).reshape(-1, 1), self.Q)
        if self.kind == "uniform_response":
            return dist.uniform_log_p(self.p1, self.p2, theta.shape[0], theta.shape[1])
        if self.kind == "lognormal_response":                                         # location_scale.py:296-303
            return dist.lognormal_log_p(theta.reshape(-1, 1), np.broadcast_to(self.p1, theta.shape).reshape(-1, 1), self.Q)
        if self.kind == "normal_linear":
            return dist.normal_linear_log_p(self.data, self.X, theta, self.Q, self.transform)
        raise ValueError(self.kind)

    def grad_hess_analytic(self, theta):
        theta = np.asarray(theta, float)
        n = theta.size
        if self.kind == "poisson_rate":
            g, H = dist.poisson_grad_rate(self.data.reshape(theta.shape), theta)
        elif self.kind == "gamma_response":
            g, H = dist.gamma_grad_response(theta, np.broadcast_to(self.p1, theta.shape), np.broadcast_to(self.p2, theta.shape))
        elif self.kind == "normal_response":
            g, H = dist.normal_grad_response(theta.reshape(-1, 1), np.broadcast_to(self.p1, theta.shape).reshape(-1, 1), self.Q)
        elif self.kind == "lognormal_response":
            g, H = dist.lognormal_grad_response(theta.reshape(-1, 1), np.broadcast_to(self.p1, theta.shape).reshape(-1, 1), self.Q)
        elif self.kind == "normal_linear":
            g, H = dist.normal_linear_grad(self.data, self.X, theta, self.Q, self.transform)
        else:
            g, H = np.zeros(n), np.zeros((n, n))
        return np.asarray(g, float).reshape(theta.shape), np.asarray(H, float).reshape(n, n)

    def grad_hess_reference(self, theta):
        """What the reference computes: finite differences unless the distribution overrides grad_log_p
        (distribution.py:90-198; Normal: location_scale.py:190-250)."""
        if self.kind in ("normal_response", "lognormal_response", "normal_linear"):
            return self.grad_hess_analytic(theta)
        g = dist.grad_fd(self.log_p, theta)
        H = dist.hessian_fd(lambda x: dist.grad_fd(self.log_p, x), theta)
        return g, H


def log_p(terms, theta):
    """Sum over the conditional model.  ref: metropolis_hastings.py:150-154"""
    return sum(t.log_p(theta) for t in terms)


def grad_hess(terms, theta, method="analytic"):
    """Model.grad_log_p: sum of the member gradients / Hessians.  ref: model.py:72-112"""
    theta = np.asarray(theta, float)
    g, H = np.zeros(theta.shape), np.zeros((theta.size, theta.size))
    for t in terms:
        gt, Ht = t.grad_hess_analytic(theta) if method == "analytic" else t.grad_hess_reference(theta)
        g = g + gt
        H = H + Ht
    return g, H


def accept(log_accept, u):
    """ref: metropolis_hastings.py:163-173 — strict '<'; NaN rejects."""
    with np.errstate(all="ignore"):
        return bool(np.log(u) < log_accept)


def random_walk_step(terms, theta, step, var, u, limits=None, col=None):
    """One RandomWalk proposal + accept/reject.  ref: metropolis_hastings.py:212-269, 127-161.

    theta (p, n); `var` are the N(0,1) variates (untruncated, shape of the proposed block) or the uniforms behind
    truncnorm.rvs (truncated); col = replicate column for RandomWalkLoop (None: all elements).
    Returns (new_theta, info) with info = dict(logp_cur, logp_prop, lq_fwd, lq_rev, accepted).
    """
    theta = np.asarray(theta, float)
    step = np.array(step, ndmin=2, dtype=float)
    prop = theta.copy()
    if col is None:
        mu, stp = theta, np.broadcast_to(step, theta.shape)
    else:
        mu = theta[:, col]
        stp = step.flatten() if step.shape[1] == 1 else step[:, col].flatten()
        stp = np.broadcast_to(stp, mu.shape)
    var = np.asarray(var, float).reshape(mu.shape)
    if limits is None:
        z = mu + stp * var
        lq_fwd = lq_rev = 0.0
    else:
        lim = np.asarray(limits, float).reshape(-1, 2)
        lb = lim[:, 0] if col is not None else lim[:, [0]]
        ub = lim[:, 1] if col is not None else lim[:, [1]]
        z = gmrf.truncated_normal_rv(mu, stp, lb, ub, var)
        lq_fwd = float(np.sum(gmrf.truncated_normal_log_pdf(z, mu, stp, lb, ub)))
        lq_rev = float(np.sum(gmrf.truncated_normal_log_pdf(mu, z, stp, lb, ub)))
    if col is None:
        prop = z
    else:
        prop[:, col] = z
    lc, lp = log_p(terms, theta), log_p(terms, prop)
    with np.errstate(all="ignore"):
        log_accept = lp + lq_rev - (lc + lq_fwd)
    acc = accept(log_accept, u)
    return (prop if acc else theta), dict(logp_cur=lc, logp_prop=lp, lq_fwd=lq_fwd, lq_rev=lq_rev, accepted=acc)


def random_walk_loop_sweep(terms, theta, step, var, u, limits):
    """RandomWalkLoop.sample: column-at-a-time.  ref: metropolis_hastings.py:276-289.  var (n_rep, p), u (n_rep,)."""
    infos = []
    for col in range(theta.shape[1]):
        theta, info = random_walk_step(terms, theta, step, var[col], u[col], limits, col)
        infos.append(info)
    return theta, infos


def mmala_params(terms, theta, step, method):
    """(mu, L) of the mMALA proposal.  ref: metropolis_hastings.py:325-348"""
    g, H = grad_hess(terms, theta, method)
    L = gmrf.cholesky(H / step ** 2)
    mu = theta + 0.5 * gmrf.cho_solve(L, g.reshape(-1, 1)).reshape(g.shape)
    return mu, L


def mmala_log_density(x, mu, L):
    """ref: metropolis_hastings.py:350-373"""
    w = L.T @ (x - mu)
    return float(np.sum(np.log(np.diag(L))) - 0.5 * (w.T @ w).item())


def mmala_step(terms, theta, step, z, u, method="analytic"):
    """One ManifoldMALA proposal + accept/reject.  ref: metropolis_hastings.py:301-323, 127-161.  theta (p, 1)."""
    theta = np.asarray(theta, float).reshape(-1, 1)
    mu, L = mmala_params(terms, theta, step, method)
    prop = gmrf.sample_normal(mu, L, np.asarray(z, float).reshape(-1, 1))
    lq_fwd = mmala_log_density(prop, mu, L)
    info = dict(mu=mu, L=L, prop=prop, lq_fwd=lq_fwd)
    try:
        with np.errstate(all="ignore"):
            mu_r, L_r = mmala_params(terms, prop, step, method)
            if not (np.all(np.isfinite(L_r)) and np.all(np.isfinite(mu_r))):
                raise np.linalg.LinAlgError("non-finite reverse proposal")
    except np.linalg.LinAlgError:
        # the reference raises here; the device path rejects the move and flags the chain
        info.update(accepted=False, invalid=True)
        return theta, info
    lq_rev = mmala_log_density(theta, mu_r, L_r)
    lc, lp = log_p(terms, theta), log_p(terms, prop)
    with np.errstate(all="ignore"):
        log_accept = lp + lq_rev - (lc + lq_fwd)
    acc = accept(log_accept, u)
    info.update(logp_cur=lc, logp_prop=lp, lq_rev=lq_rev, log_accept=log_accept, accepted=acc, invalid=False)
    return (prop if acc else theta), info
