"""GPU parity of the one-call `openmcmc_b200.gmrf` functions against numpy / scipy restatements
of the reference's formulas (gmrf.py:29-61, 167-198, 269-348) with injected variates; tolerances of the north star
(deterministic quantities rel 1e-10, draws with injected z 1e-9)."""

import numpy as np
import pytest
from scipy import sparse, stats

pytestmark = pytest.mark.gpu


def _alg25(Q, b, z):
    """Rue & Held Alg. 2.5 in numpy: x = mu + L^-T z, mu = Q^-1 b (gmrf.py:167-198, 29-61)."""
    Qd = Q.toarray() if sparse.issparse(Q) else np.asarray(Q)
    L = np.linalg.cholesky(Qd)
    mu = np.linalg.solve(L.T, np.linalg.solve(L, b))
    return mu + np.linalg.solve(L.T, z)


@pytest.mark.parametrize("p,kind", [(7, "dense"), (64, "dense"), (1, "dense"), (5, "diag"), (300, "tridiag"), (5000, "tridiag"),
                                    (2500, "diag")])
def test_sample_normal_canonical_matches_algorithm_2_5(p, kind):
    from openmcmc_b200 import gmrf

    rng = np.random.default_rng(p)
    if kind == "dense":
        A = rng.standard_normal((p, p))
        Q = A @ A.T + p * np.eye(p)
    elif kind == "diag":
        Q = sparse.diags([rng.random(p) + 0.5], [0], format="csc")
    else:
        Q = (3.0 * gmrf.precision_irregular(np.cumsum(rng.exponential(size=p))) + sparse.identity(p) * 0.7).tocsc()
    b = rng.standard_normal((p, 1))
    z = rng.standard_normal((p, 1))
    x = gmrf.sample_normal_canonical(b, Q=Q, z=z)
    ref = _alg25(Q, b, z)
    assert x.shape == (p, 1)
    assert np.max(np.abs(x - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref)))
    # sample_normal: mu + L^-T z, column per draw
    Z = rng.standard_normal((p, 2))
    mu = rng.standard_normal((p, 1))
    xs = gmrf.sample_normal(mu, Q=Q, n=2, z=Z)
    ref2 = mu + np.linalg.solve(np.linalg.cholesky(Q.toarray() if sparse.issparse(Q) else Q).T, Z)
    assert xs.shape == (p, 2) and np.max(np.abs(xs - ref2)) <= 1e-9 * max(1.0, np.max(np.abs(ref2)))


def test_free_running_draws_are_fresh_and_standardise():
    from openmcmc_b200 import gmrf

    gmrf._calls = 1000      # the call counter keys the generator: pinned, so the statistics below do not depend on test order
    p = 4000
    Q = (2.0 * gmrf.precision_irregular(np.arange(p) * 0.5) + sparse.identity(p)).tocsc()
    b = np.zeros((p, 1))
    x1 = gmrf.sample_normal_canonical(b, Q=Q, seed=5)
    x2 = gmrf.sample_normal_canonical(b, Q=Q, seed=5)
    assert not np.array_equal(x1, x2)                    # successive calls advance the generator
    L = np.linalg.cholesky(Q.toarray())
    w = L.T @ x1                                          # L'x ~ N(0, I)
    assert stats.kstest(w.ravel(), "norm").pvalue > 1e-3 and abs(w.std() - 1.0) < 0.05


@pytest.mark.parametrize("sparse_q", [False, True])
def test_multivariate_normal_pdf_matches_scipy(sparse_q):
    from openmcmc_b200 import gmrf

    rng = np.random.default_rng(11)
    p, n = (6, 5) if not sparse_q else (400, 3)
    if sparse_q:
        Q = (1.5 * gmrf.precision_irregular(np.cumsum(rng.exponential(size=p))) + 0.3 * sparse.identity(p)).tocsc()
        Qd = Q.toarray()
    else:
        A = rng.standard_normal((p, p))
        Q = Qd = A @ A.T + p * np.eye(p)
    mu = rng.standard_normal((p, 1))
    x = mu + rng.standard_normal((p, n))
    ref = stats.multivariate_normal(mean=mu.ravel(), cov=np.linalg.inv(Qd)).logpdf(x.T)
    got = gmrf.multivariate_normal_pdf(x, mu, Q, by_observation=True)
    np.testing.assert_allclose(got, np.atleast_1d(ref), rtol=1e-10)
    np.testing.assert_allclose(gmrf.multivariate_normal_pdf(x, mu, Q), np.sum(ref), rtol=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        gmrf.multivariate_normal_pdf(x[:3], mu[:3], -np.eye(3))


def test_truncated_normal_functions_match_scipy():
    from openmcmc_b200 import gmrf

    gmrf._calls = 2000      # pinned generator state (see above)
    rng = np.random.default_rng(2)
    mean = rng.standard_normal(50) * 2
    scale = rng.random(50) + 0.2
    lower, upper = mean - rng.random(50) * 3, mean + rng.random(50) * 3
    u = rng.random(50)
    a, b = (lower - mean) / scale, (upper - mean) / scale
    x = gmrf.truncated_normal_rv(mean, scale, lower, upper, u=u)
    np.testing.assert_allclose(x, stats.truncnorm.ppf(u, a, b, loc=mean, scale=scale), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gmrf.truncated_normal_log_pdf(x, mean, scale, lower, upper),
                               stats.truncnorm.logpdf(x, a, b, loc=mean, scale=scale), rtol=1e-10, atol=1e-12)
    # one-sided (None = infinite), outside the support, free-running draws stay inside
    lp = gmrf.truncated_normal_log_pdf(np.array([-1.0, 0.5]), 0.0, 1.0, 0.0, None)
    assert lp[0] == -np.inf and np.isclose(lp[1], stats.truncnorm.logpdf(0.5, 0.0, np.inf))
    d = gmrf.truncated_normal_rv(np.zeros(2000), np.ones(2000), -0.5, 2.0, seed=3)
    assert d.shape == (2000,) and d.min() >= -0.5 and d.max() <= 2.0
    assert stats.kstest(d, stats.truncnorm(-0.5, 2.0).cdf).pvalue > 1e-3
    assert gmrf.truncated_normal_rv(0.0, 1.0, -1.0, 1.0, size=7, seed=1).shape == (7,)


def test_distribution_rvs_and_missing_initial_values():
    """dist.rvs(state, n) (p x n draws; location_scale.py:252-272, distribution.py:263-278, 354-374, 444-458, 510-523) and
    the start-up draw of a sampled parameter that is missing from the initial state (mcmc.py:78-80).  Free-running
    variates: checked in distribution (KS / moments)."""
    from openmcmc_b200.distribution.distribution import Categorical, Gamma, Poisson, Uniform
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal
    from openmcmc_b200 import devdist, gmrf, hostcalls

    hostcalls.set_seed(0)    # pinned generator state: the statistics below do not depend on test order
    devdist._rvs_calls, gmrf._calls = 0, 3000
    rng = np.random.default_rng(4)
    g = Gamma("tau", shape="a", rate="b")
    x = g.rvs({"a": np.array([[3.0]]), "b": np.array([[2.0]])}, n=4000)
    assert x.shape == (1, 4000) and stats.kstest(x.ravel(), stats.gamma(3.0, scale=0.5).cdf).pvalue > 1e-3
    xv = g.rvs({"a": np.array([[2.0], [5.0]]), "b": np.array([[1.0]])}, n=3000)
    assert xv.shape == (2, 3000) and abs(xv[0].mean() - 2.0) < 0.15 and abs(xv[1].mean() - 5.0) < 0.2
    nm = Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P", scalar="lam"))
    st = {"mu": np.array([[1.0], [-2.0], [0.5]]), "P": sparse.diags([[1.0, 4.0, 0.25]], [0], format="csc"), "lam": 4.0}
    b = nm.rvs(st, n=400)
    assert b.shape == (3, 400) and np.all(np.abs(b.mean(axis=1) - st["mu"].ravel()) < 0.25)
    assert np.allclose(b.std(axis=1), [0.5, 0.25, 1.0], rtol=0.2)
    k = Poisson("k", rate="r").rvs({"r": np.array([[2.0], [9.0]])}, n=3000)
    assert k.shape == (2, 3000) and np.all(k == np.floor(k)) and abs(k[0].mean() - 2.0) < 0.15 and abs(k[1].mean() - 9.0) < 0.3
    u = Uniform("x", domain_response_lower=np.array([[-1.0]]), domain_response_upper=np.array([[3.0]])).rvs(
        {"x": np.zeros((2, 1))}, n=2000)
    assert u.shape == (2, 2000) and u.min() >= -1.0 and u.max() <= 3.0 and abs(u.mean() - 1.0) < 0.1
    z = Categorical("z", prob="pr").rvs({"pr": np.array([[0.2, 0.8], [0.9, 0.1]])}, n=2000)
    assert z.shape == (2, 2000) and abs(z[0].mean() - 0.8) < 0.05 and abs(z[1].mean() - 0.1) < 0.05
    # a sampled parameter without an initial value is drawn from its prior (mcmc.py:78-80)
    n, p = 60, 3
    X = rng.standard_normal((n, p))
    y = X @ np.array([[1.0], [-1.0], [0.5]]) + 0.1 * rng.standard_normal((n, 1))
    mdl = Model([Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
                 Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
                 Gamma("tau", shape="a_tau", rate="b_tau")])
    state = {"y": y, "X": X, "P_tau": sparse.identity(n, format="csc"), "P_lambda": sparse.identity(p, format="csc"),
             "mu": np.zeros((p, 1)), "lambda": 0.01, "a_tau": 1.0, "b_tau": 1.0}
    M = MCMC(state, [NormalNormal("beta", mdl), NormalGamma("tau", mdl)], model=mdl, n_burn=20, n_iter=50, n_thin=2)
    assert M.state["beta"].shape == (p, 1) and M.state["tau"].shape == (1, 1) and M.state["tau"][0, 0] > 0
    M.run_mcmc()
    assert np.all(np.abs(M.store["beta"].mean(axis=1) - [1.0, -1.0, 0.5]) < 0.1)
    # several chains: one prior draw PER CHAIN (column = global chain id), so a shard starts where the whole run does
    def fresh(**kw):
        hostcalls.set_seed(0)
        devdist._rvs_calls, gmrf._calls = 0, 3000
        return MCMC(dict(state), [NormalNormal("beta", mdl), NormalGamma("tau", mdl)], model=mdl, n_burn=0, n_iter=3, **kw)
    M4 = fresh(n_chains=4)
    b0 = M4._chain_starts["beta"]
    assert b0.shape == (4, p, 1) and len({tuple(np.round(v.ravel(), 12)) for v in b0}) == 4
    assert np.array_equal(M4.state["beta"], b0[0])
    M2 = fresh(n_chains=2, chain_offset=2)
    np.testing.assert_array_equal(M2._chain_starts["beta"], b0[2:])
    np.testing.assert_array_equal(M2._chain_starts["tau"], M4._chain_starts["tau"][2:])
    M4.run_mcmc()
    M2.run_mcmc()
    np.testing.assert_array_equal(M2.store["beta"], M4.store["beta"][2:])


@pytest.mark.parametrize("n", [3, 64, 130, 300])
def test_cholesky_cho_solve_solve_dense(n):
    """ref tests/test_grmf.py:312-375 of the reference: cholesky == np.linalg.cholesky, lower-triangular, L L' = P;
    solve / cho_solve == np.linalg.solve -- here through omc_dense_factor (blocked Cholesky, DMMA trailing update)."""
    from openmcmc_b200 import gmrf

    rng = np.random.default_rng(n)
    A = rng.standard_normal((n, n))
    Q = A @ A.T / n + np.eye(n)
    L = gmrf.cholesky(Q)
    Lr = np.linalg.cholesky(Q)
    np.testing.assert_allclose(L, Lr, rtol=1e-10, atol=1e-13)
    assert not np.any(np.triu(L, 1))
    np.testing.assert_allclose(gmrf.cholesky(Q, lower=False), Lr.T, rtol=1e-10, atol=1e-13)
    b = rng.standard_normal((n, 2))
    ref = np.linalg.solve(Q, b)
    np.testing.assert_allclose(gmrf.cho_solve((L, True), b), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    np.testing.assert_allclose(gmrf.cho_solve((L.T, False), b), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    np.testing.assert_allclose(gmrf.solve(Q, b), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    z = rng.standard_normal((n, 1))
    for a in (L.T, L):                                     # the triangular solves of gmrf.py:61
        r = np.linalg.solve(a, z)
        np.testing.assert_allclose(gmrf.solve(a, z), r, rtol=1e-9, atol=1e-9 * np.abs(r).max())
    with pytest.raises(np.linalg.LinAlgError):
        gmrf.cholesky(Q - 3.0 * np.eye(n))
    # draws from a precomputed factor: mu + L^-T z without re-factorising
    mu = rng.standard_normal((n, 1))
    x = gmrf.sample_normal(mu, L=L, z=z)
    np.testing.assert_allclose(x, mu + np.linalg.solve(Lr.T, z), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("n", [2, 50, 4000])
def test_sparse_cholesky_and_solves_tridiagonal(n):
    from openmcmc_b200 import gmrf

    rng = np.random.default_rng(n + 1)
    Q = (2.0 * gmrf.precision_irregular(np.cumsum(rng.exponential(size=n) + 0.2)) + 0.5 * sparse.identity(n)).tocsc()
    L = gmrf.sparse_cholesky(Q)
    assert sparse.issparse(L)
    Lr = np.linalg.cholesky(Q.toarray())
    np.testing.assert_allclose(L.toarray(), Lr, rtol=1e-10, atol=1e-13)
    assert sparse.issparse(gmrf.cholesky(Q)) and sparse.triu(gmrf.cholesky(Q), 1).nnz == 0
    b = rng.standard_normal((n, 1))
    ref = np.linalg.solve(Q.toarray(), b)
    np.testing.assert_allclose(gmrf.cho_solve((L, True), b), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    np.testing.assert_allclose(gmrf.solve(Q, b), ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
    z = rng.standard_normal((n, 1))
    x = gmrf.sample_normal_canonical(b, L=L, z=z)
    np.testing.assert_allclose(x, _alg25(Q, b, z), rtol=1e-9, atol=1e-9 * np.abs(x).max())
    # a sparse matrix with a wider band goes through the dense factorisation
    if n <= 50:
        W = Q + sparse.diags([0.1 * np.ones(n - 1)], [1], shape=(n, n)).T @ sparse.diags([0.1 * np.ones(n - 1)], [1], shape=(n, n))
        Wd = (W + W.T).toarray() / 2 + np.eye(n)
        np.testing.assert_allclose(gmrf.sparse_cholesky(sparse.csc_matrix(Wd)).toarray(), np.linalg.cholesky(Wd),
                                   rtol=1e-10, atol=1e-13)


def test_gibbs_canonical_truncated_normal_and_truncated_sampling():
    """ref gmrf.py:201-266: the coordinate scan with the uniforms behind truncnorm.rvs injected equals a numpy / scipy
    restatement; ref tests/test_grmf.py:93-147: Gibbs and rejection samples stay inside the box and agree in mean."""
    from openmcmc_b200 import gmrf

    rng = np.random.default_rng(12)
    p = 9
    A = rng.standard_normal((p, p))
    Q = A @ A.T + p * np.eye(p)
    b = rng.standard_normal((p, 1))
    lower, upper = -0.3 * np.ones((p, 1)), 0.5 * np.ones((p, 1))
    x0 = np.zeros((p, 1))
    u = rng.random(p)
    x = gmrf.gibbs_canonical_truncated_normal(b, Q, x0.copy(), lower, upper, u=u)
    ref = x0.copy()
    for i in range(p):
        v = 1.0 / Q[i, i]
        m = v * (b[i, 0] - Q[i, :] @ ref[:, 0] + Q[i, i] * ref[i, 0])
        s = np.sqrt(v)
        ref[i, 0] = stats.truncnorm.ppf(u[i], (lower[i, 0] - m) / s, (upper[i, 0] - m) / s, loc=m, scale=s)
    np.testing.assert_allclose(x, ref, rtol=1e-9, atol=1e-12)
    # unbounded: a plain canonical draw
    z = rng.standard_normal((p, 1))
    assert gmrf.gibbs_canonical_truncated_normal(b, Q, x0.copy()).shape == (p, 1)
    mu = np.linalg.solve(Q, b)
    G = gmrf.sample_truncated_normal(mu, Q=Q, lower=lower, upper=upper, n=40, method="Gibbs", seed=3)
    R = gmrf.sample_truncated_normal(mu, Q=Q, lower=lower, upper=upper, n=200, method="Rejection", seed=4)
    for S in (G, R):
        assert np.all(S >= lower) and np.all(S <= upper)
    assert np.all(np.abs(G.mean(axis=1) - R.mean(axis=1)) < 0.25)
    with pytest.raises(TypeError):
        gmrf.sample_truncated_normal(mu, Q=Q, method="other")
    with pytest.raises(ValueError):
        gmrf.sample_truncated_normal_rejection(mu, Q=Q, lower=upper, upper=lower)
