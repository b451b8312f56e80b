"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's ReversibleJump step for the Gaussian-kernel basis
model of its own tests (tests/test_reversible_jump.py:23-252), the model BASELINE configs[4] scales up.

Follows src/openmcmc/sampler/reversible_jump.py: proposal :76-94, birth_proposal :96-146, death_proposal :148-193,
matched_birth_transition :195-263, matched_death_transition :265-311, get_move_type :313-333,
get_move_probabilities :335-373, and MetropolisHastings._accept_reject_proposal (metropolis_hastings.py:127-161).
Pinned by tests/golden/rj_*.npz, recorded from the live reference with its rvs streams captured.

Quirks kept: the proposal log-density of the associated parameters is evaluated on the CURRENT state
and its LAST component; ridge 1e-10 in the matching solve; log(det F) without abs (negative determinant -> NaN ->
reject).  The matching solve G = (S + eps I)^-1 S[:, cols] is evaluated as (I - eps (S + eps I)^-1)[:, cols] (same
quantity; it needs one k x k buffer on the device and is far less sensitive to the conditioning of the Gram matrix
than the reference's own LU solve, whose error grows like cond(S) * 1e-16 — the tolerance the golden tests state).
"""

import numpy as np
from scipy import special, stats

EPS = 1e-10


def make_basis(X, theta, omega):
    """tests/test_reversible_jump.py:23-40: B[:, k] = norm.pdf(X, loc=theta_k, scale=omega_k)."""
    X = np.asarray(X, dtype=float).reshape(-1, 1)
    z = (X - theta.reshape(1, -1)) / omega.reshape(1, -1)
    return np.exp(-0.5 * z * z) / (np.sqrt(2 * np.pi) * omega.reshape(1, -1))


def move_probabilities(n, n_max, q, birth):
    """reversible_jump.py:335-373"""
    p_birth, p_death = q, 1.0 - q
    if n == n_max:
        p_death = 1.0
    if n == n_max - 1 and birth:
        p_death = 1.0
    if n == 1:
        p_birth = 1.0
    if n == 2 and not birth:
        p_birth = 1.0
    return p_birth, p_death


def truncnorm_logpdf(x, mean, scale, lower, upper):
    """gmrf.py:295-318"""
    a, b = (lower - mean) / scale, (upper - mean) / scale
    return float(stats.truncnorm.logpdf(x, a, b, loc=mean, scale=scale))


def model_log_p(m, n, theta, omega, beta, B):
    """Sum of the log-densities of every distribution in the model (metropolis_hastings.py:152-155)."""
    lp = 0.0
    if m["y"] is not None:    # Normal response, precision tau_y * I   (location_scale.py:145-167, gmrf.py:321-348)
        r = m["y"] - B @ beta
        nd = r.size
        lp += 0.5 * (nd * np.log(m["tau_y"]) - nd * np.log(2 * np.pi) - m["tau_y"] * float(r @ r))
    d = beta - m["mu_beta"]     # Normal prior, one mixture component: iid N(mu_beta, 1/tau_beta)
    lp += 0.5 * (n * np.log(m["tau_beta"]) - n * np.log(2 * np.pi) - m["tau_beta"] * float(d @ d))
    lp += n * np.log(m["rho"]) - special.gammaln(n + 1.0) - m["rho"]            # Poisson (distribution.py:490-508)
    lp += -n * np.log(m["theta_hi"] - m["theta_lo"])                             # Uniform (distribution.py:422-442)
    if m["a_omega"] is not None:                                                 # Gamma (distribution.py:241-261)
        lp += float(np.sum(stats.gamma.logpdf(omega, m["a_omega"], scale=1.0 / m["b_omega"])))
    return lp


def prop_density_last(m, theta, omega):
    """sum over the associated parameters of log_p(current_state, by_observation=True)[-1]  (:132, :176; F8)."""
    lp = -np.log(m["theta_hi"] - m["theta_lo"])
    if m["a_omega"] is not None:
        lp += float(stats.gamma.logpdf(omega[-1], m["a_omega"], scale=1.0 / m["b_omega"]))
    return lp


def _logdet(F):
    """np.log(np.linalg.det(F)) (:259, :298): NaN for a negative determinant."""
    sign, lad = np.linalg.slogdet(F)
    return lad if sign > 0 else np.nan


def rj_step(m, st, draws):
    """One ReversibleJump.sample().  st: dict(n, theta, omega, beta, B) with exact-size arrays; draws: dict(u_move,
    theta_new, omega_new, beta_new, del_index, u_accept) — the FINAL values of the reference's variates (beta_new None
    => the matched mean itself, the mocked samplers of the reference tests).  Returns (new_state, info)."""
    n, theta, omega, beta, B = st["n"], st["theta"], st["omega"], st["beta"], st["B"]
    n_max, q = m["n_max"], m["birth_probability"]
    if n == 0:
        raise ValueError("Reversible jump MCMC: Number of parameters cannot be zero.")
    birth = False if n == n_max else True if n == 1 else bool(draws["u_move"] <= q)     # :326-333
    lpd_last = prop_density_last(m, theta, omega)
    p_birth, p_death = move_probabilities(n, n_max, q, birth)
    if birth:
        th_p = np.append(theta, draws["theta_new"])
        om_p = np.append(omega, draws["omega_new"]) if m["a_omega"] is not None else np.append(omega, omega[-1])
        B_p = np.column_stack([B, make_basis(m["X"], th_p[-1:], om_p[-1:])[:, 0]])
        S = B_p.T @ B_p
        Z = np.linalg.inv(S + EPS * np.eye(n + 1))
        G = (np.eye(n + 1) - EPS * Z)[:, :n]                       # = solve(S + eps I, B_p' B)   (:239-242)
        mu_star = G @ beta
        beta_p = mu_star.copy()
        if draws.get("beta_new") is not None:
            beta_p[-1] = draws["beta_new"]
        elif draws.get("u_trunc") is not None:      # free-running: the uniform behind truncnorm.rvs (gmrf.py:269-292)
            if m["match_limits"] is not None:
                lo_, hi_ = m["match_limits"]
                a_, b_ = (lo_ - mu_star[-1]) / m["match_scale"], (hi_ - mu_star[-1]) / m["match_scale"]
                beta_p[-1] = float(stats.truncnorm.ppf(draws["u_trunc"], a_, b_, loc=mu_star[-1], scale=m["match_scale"]))
            else:
                beta_p[-1] = mu_star[-1] + m["match_scale"] * float(stats.norm.ppf(draws["u_trunc"]))
        if m["match_limits"] is not None:
            lq_f = truncnorm_logpdf(beta_p[-1], mu_star[-1], m["match_scale"], *m["match_limits"])
        else:
            lq_f = float(stats.norm.logpdf(beta_p[-1], mu_star[-1], m["match_scale"]))
        F = np.column_stack([G, np.eye(n + 1)[:, -1]])
        lq_r = _logdet(F)
        lq_f += np.log(p_birth) + lpd_last
        lq_r += np.log(p_death)
        prop = dict(n=n + 1, theta=th_p, omega=om_p, beta=beta_p, B=B_p)
        d = -1
    else:
        d = int(draws["del_index"])
        th_p, om_p = np.delete(theta, d), np.delete(omega, d)
        B_p = np.delete(B, d, axis=1)
        S = B.T @ B
        Z = np.linalg.inv(S + EPS * np.eye(n))
        Gfull = np.eye(n) - EPS * Z                               # columns != d = solve(S + eps I, B' B_p)  (:287-290)
        F = Gfull.copy()
        F[:, d] = np.eye(n)[:, d]                                  # np.insert(G, d, e_d)               (:291)
        mu_aug = np.linalg.solve(F, beta)
        param_del = mu_aug[d]
        beta_p = np.delete(mu_aug, d)
        lq_f = _logdet(F)
        if m["match_limits"] is not None:
            lq_r = truncnorm_logpdf(param_del, 0.0, m["match_scale"], *m["match_limits"])
        else:
            lq_r = float(stats.norm.logpdf(param_del, 0.0, m["match_scale"]))
        lq_f += np.log(p_death)
        lq_r += np.log(p_birth) + lpd_last
        prop = dict(n=n - 1, theta=th_p, omega=om_p, beta=beta_p, B=B_p)
    lp_c = model_log_p(m, n, theta, omega, beta, B)
    lp_p = model_log_p(m, prop["n"], prop["theta"], prop["omega"], prop["beta"], prop["B"])
    log_accept = lp_p + lq_r - (lp_c + lq_f)
    acc = bool(np.log(draws["u_accept"]) < log_accept)             # strict; NaN rejects (metropolis_hastings.py:173)
    info = dict(birth=birth, del_index=d, logp_cur=lp_c, logp_prop=lp_p, lq_fwd=lq_f, lq_rev=lq_r, log_accept=log_accept,
                accepted=acc, prop=prop)
    return (prop if acc else st), info


# ----------------------------------------------------------------------------------------------- companion samplers
def knot_walk_sweep(m, st, which, step, limits, tn_u, u):
    """RandomWalkLoop.sample over the knots (which = 'theta') or widths ('omega') with the basis rebuilt for every
    proposal (the reference tests' move_function = make_basis as state_update_function).

    ref: metropolis_hastings.py:276-289 (loop over columns), :212-269 (truncated proposal of one column, densities),
    :127-161 (accept on the FULL model: RandomWalk keeps it when a state_update_function is given, :201-210).
    tn_u / u: per component, the uniform behind truncnorm.rvs and the accept uniform.  Returns (state, n_accepted)."""
    from oracle import gmrf

    st = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in st.items()}
    n, lo, hi = st["n"], limits[0], limits[1]
    n_acc = 0
    for j in range(n):
        cur = st[which][j]
        z = float(gmrf.truncated_normal_rv(cur, step, lo, hi, tn_u[j]))
        lq_f = float(gmrf.truncated_normal_log_pdf(z, cur, step, lo, hi))
        lq_r = float(gmrf.truncated_normal_log_pdf(cur, z, step, lo, hi))
        prop = dict(st)
        prop[which] = st[which].copy()
        prop[which][j] = z
        prop["B"] = make_basis(m["X"], prop["theta"], prop["omega"])
        lp_c = model_log_p(m, n, st["theta"], st["omega"], st["beta"], st["B"])
        lp_p = model_log_p(m, n, prop["theta"], prop["omega"], prop["beta"], prop["B"])
        with np.errstate(all="ignore"):
            if bool(np.log(u[j]) < lp_p + lq_r - (lp_c + lq_f)):
                st = prop
                n_acc += 1
    return st, n_acc


def coef_mmala_step(m, st, step, z, u):
    """ManifoldMALA.sample on the live coefficients.  Conditional model: response Normal(y | B beta, (tau_y I)^-1) (or
    Null) and the iid Normal prior; gradient / Hessian by the mean-parameter and response branches of
    Normal.grad_log_p (location_scale.py:222-250); proposal N(beta + Hs^-1 g / 2, Hs^-1), Hs = H / step^2, forward and
    reverse (metropolis_hastings.py:301-373).  Returns (state, info)."""
    n, beta, B = st["n"], st["beta"], st["B"]

    def grad_hess(b):
        g = -m["tau_beta"] * (b - m["mu_beta"])
        H = m["tau_beta"] * np.eye(n)
        if m["y"] is not None:
            g = g + m["tau_y"] * (B.T @ (m["y"] - B @ b))
            H = H + m["tau_y"] * (B.T @ B)
        return g, H

    def params(b):
        g, H = grad_hess(b)
        L = np.linalg.cholesky(H / step ** 2)
        return b + 0.5 * np.linalg.solve(L.T, np.linalg.solve(L, g)), L

    def log_density(x, mean, L):      # up to the constant the reference drops as well (:350-373)
        w = L.T @ (x - mean)
        return float(np.sum(np.log(np.diag(L))) - 0.5 * w @ w)

    def cond_log_p(b):
        lp = 0.0
        if m["y"] is not None:
            r = m["y"] - B @ b
            lp += 0.5 * (r.size * np.log(m["tau_y"]) - r.size * np.log(2 * np.pi) - m["tau_y"] * float(r @ r))
        d = b - m["mu_beta"]
        return lp + 0.5 * (n * np.log(m["tau_beta"]) - n * np.log(2 * np.pi) - m["tau_beta"] * float(d @ d))

    mu_f, L = params(beta)
    prop = mu_f + np.linalg.solve(L.T, np.asarray(z, float)[:n])
    mu_r, L_r = params(prop)
    lq_f, lq_r = log_density(prop, mu_f, L), log_density(beta, mu_r, L_r)
    log_accept = cond_log_p(prop) + lq_r - (cond_log_p(beta) + lq_f)
    acc = bool(np.log(u) < log_accept)
    out = dict(st)
    if acc:
        out["beta"] = prop
    return out, dict(prop=prop, lq_fwd=lq_f, lq_rev=lq_r, log_accept=log_accept, accepted=acc)

