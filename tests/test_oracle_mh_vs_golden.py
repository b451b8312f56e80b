"""CPU: the numpy oracle of the Metropolis-Hastings family (oracle/mh.py) reproduces golden chains recorded from the
live reference with its random streams replayed (tests/golden/make_golden.py)."""

import os

import numpy as np
import pytest

from oracle import gmrf, mh

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))


def poisson_gamma_terms(g):
    return [mh.Term("poisson_rate", data=g["y"]), mh.Term("gamma_response", p1=g["a"], p2=g["b"])]


def normal_terms(g):
    p = g["theta0"].shape[0]
    return [mh.Term("normal_response", p1=g["mu"], Q=g["lam"] * g["P"]),
            mh.Term("normal_response", p1=g["yobs"], Q=g["tau"] * np.diag(g["w"]))], p


@pytest.mark.parametrize("name", ["mmala_poisson_gamma_p6", "mmala_poisson_gamma_p32_vec"])
def test_mmala_fd_chain(name):
    """Finite-difference derivatives: the reference cannot reproduce its own FD Hessian beyond ~1e-6, so
    the chain replay is held to 1e-5 and the deterministic probes to the measured FD self-noise."""
    g = _load(name)
    terms = poisson_gamma_terms(g)
    g0, H0 = mh.grad_hess(terms, g["lam0"], "reference")
    np.testing.assert_allclose(g0, g["grad0"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(H0, g["hess0"], rtol=1e-4, atol=1e-4)
    ga, Ha = mh.grad_hess(terms, g["lam0"], "analytic")
    np.testing.assert_allclose(ga, g["grad0"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(Ha, g["hess0"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(mh.log_p(terms, g["lam0"]), g["logp0"], rtol=1e-13)
    theta = g["lam0"]
    n_acc = 0
    for it in range(g["store_lam"].shape[1]):
        theta, info = mh.mmala_step(terms, theta, float(g["step"]), g["z"][it], g["u"][it], "reference")
        n_acc += info["accepted"]
        np.testing.assert_allclose(theta.ravel(), g["store_lam"][:, it], rtol=1e-5)
        np.testing.assert_allclose(mh.log_p(terms, theta), g["store_log_post"][it, 0], rtol=1e-6)
    assert n_acc == g["accept"][0]


@pytest.mark.parametrize("name", ["mmala_normal_p7", "mmala_normal_p40"])
def test_mmala_analytic_chain(name):
    g = _load(name)
    terms, p = normal_terms(g)
    g0, H0 = mh.grad_hess(terms, g["theta0"], "reference")
    np.testing.assert_allclose(g0, g["grad0"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(H0, g["hess0"], rtol=1e-12, atol=1e-13)
    theta = g["theta0"]
    for it in range(g["store_theta"].shape[1]):
        theta, info = mh.mmala_step(terms, theta, float(g["step"]), g["z"][it], g["u"][it], "analytic")
        np.testing.assert_allclose(theta.ravel(), g["store_theta"][:, it], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(mh.log_p(terms, theta), g["store_log_post"][it, 0], rtol=1e-11)


@pytest.mark.parametrize("name", ["rwl_poisson_gamma_1x8", "rwl_poisson_gamma_1x32"])
def test_random_walk_loop_chain(name):
    g = _load(name)
    terms = poisson_gamma_terms(g)
    theta = g["lam0"]
    n_acc = 0
    for it in range(g["store_lam"].shape[2]):
        theta, infos = mh.random_walk_loop_sweep(terms, theta, g["step"], g["tn_u"][it], g["u"][it], g["limits"])
        n_acc += sum(i["accepted"] for i in infos)
        np.testing.assert_allclose(theta, g["store_lam"][:, :, it], rtol=1e-11)
        np.testing.assert_allclose(mh.log_p(terms, theta), g["store_log_post"][it, 0], rtol=1e-12)
    assert n_acc == g["accept"][0] and g["accept"][1] == g["store_lam"].shape[2] * g["store_lam"].shape[1]


@pytest.mark.parametrize("name", ["rw_poisson_gamma_p6", "rw_trunc_scalar"])
def test_random_walk_chain(name):
    g = _load(name)
    terms = poisson_gamma_terms(g)
    theta = g["lam0"]
    limits = g["limits"] if g["limits"].size else None
    n_acc = 0
    for it in range(g["store_lam"].shape[1]):
        theta, info = mh.random_walk_step(terms, theta, g["step"], g["z"][it].reshape(theta.shape), g["u"][it], limits)
        n_acc += info["accepted"]
        np.testing.assert_allclose(theta.ravel(), g["store_lam"][:, it], rtol=1e-11)
    assert n_acc == g["accept"][0]


def lognormal_terms(g):
    return [mh.Term("lognormal_response", p1=g["mu"], Q=g["lam"] * g["P"]),
            mh.Term("normal_response", p1=g["yobs"], Q=g["tau"] * np.diag(g["w"]))]


def mhreg_terms(g):
    return [mh.Term("normal_linear", data=g["y"], X=g["X"], Q=g["tau"] * np.diag(g["w"]), transform=bool(g["transform"])),
            mh.Term("normal_response", p1=np.zeros_like(g["beta0"]), Q=g["lam"] * g["P"])]


@pytest.mark.parametrize("name", ["lognormal_mmala_p5_dense", "lognormal_mmala_p24_diag", "lognormal_rw_p6_dense"])
def test_lognormal_chain(name):
    """LogNormal prior (location_scale.py:275-418), response branch; analytic in the reference."""
    g = _load(name)
    terms = lognormal_terms(g)
    g0, H0 = terms[0].grad_hess_analytic(g["theta0"])
    np.testing.assert_allclose(g0, g["grad0"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(H0, g["hess0"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(terms[0].log_p(g["theta0"]), g["logp0"], rtol=1e-13)
    theta = g["theta0"]
    n_acc = 0
    for it in range(g["store_theta"].shape[1]):
        if str(g["sampler"]) == "mmala":
            theta, info = mh.mmala_step(terms, theta, float(g["step"]), g["z"][it], g["u"][it], "analytic")
        else:
            theta, info = mh.random_walk_step(terms, theta, np.array([[float(g["step"])]]), g["z"][it].reshape(theta.shape),
                                              g["u"][it], None)
        n_acc += info["accepted"]
        np.testing.assert_allclose(theta.ravel(), g["store_theta"][:, it], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(mh.log_p(terms, theta), g["store_log_post"][it, 0], rtol=1e-11)
    assert n_acc == g["accept"][0]


@pytest.mark.parametrize("name", ["mhreg_mmala_n60_p6", "mhreg_mmala_n200_p30_eye", "mhreg_exp_mmala_n80_p5"])
def test_mh_regression_chain(name):
    """SURVEY a4 in an MH sampler / f4: linear (optionally exp-transformed) Normal mean, mean-parameter branch of
    Normal.grad_log_p (location_scale.py:234-250; parameter.py:199-228, 283-297)."""
    g = _load(name)
    terms = mhreg_terms(g)
    g0, H0 = terms[0].grad_hess_analytic(g["beta0"])
    np.testing.assert_allclose(g0, g["grad0"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(H0, g["hess0"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(terms[0].log_p(g["beta0"]), g["logp0"], rtol=1e-13)
    theta = g["beta0"]
    n_acc = 0
    for it in range(g["store_beta"].shape[1]):
        theta, info = mh.mmala_step(terms, theta, float(g["step"]), g["z"][it], g["u"][it], "analytic")
        n_acc += info["accepted"]
        np.testing.assert_allclose(theta.ravel(), g["store_beta"][:, it], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(mh.log_p(terms, theta), g["store_log_post"][it, 0], rtol=1e-10)
    assert n_acc == g["accept"][0]


@pytest.mark.parametrize("name", ["mhreg_lognormal_mmala_n70_p5", "mhreg_lognormal_mmala_n150_p20_eye"])
def test_mh_lognormal_regression_chain(name):
    """LogNormal response with a linear mean, mean-parameter branch (location_scale.py:296-303, 344-347, 401-404): the
    Normal-linear term on log(y); the Jacobian -sum(log y) only shifts log_p."""
    g = _load(name)
    logy = np.log(g["y"])
    terms = [mh.Term("normal_linear", data=logy, X=g["X"], Q=g["tau"] * np.diag(g["w"]), transform=False),
             mh.Term("normal_response", p1=np.zeros_like(g["beta0"]), Q=g["lam"] * g["P"])]
    jac = -np.sum(logy)
    g0, H0 = terms[0].grad_hess_analytic(g["beta0"])
    np.testing.assert_allclose(g0, g["grad0"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(H0, g["hess0"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(terms[0].log_p(g["beta0"]) + jac, g["logp0"], rtol=1e-13)
    theta = g["beta0"]
    n_acc = 0
    for it in range(g["store_beta"].shape[1]):
        theta, info = mh.mmala_step(terms, theta, float(g["step"]), g["z"][it], g["u"][it], "analytic")
        n_acc += info["accepted"]
        np.testing.assert_allclose(theta.ravel(), g["store_beta"][:, it], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(mh.log_p(terms, theta) + jac, g["store_log_post"][it, 0], rtol=1e-10)
    assert n_acc == g["accept"][0]


@pytest.mark.parametrize("name", ["replicated_regression_d40_p5_r6", "replicated_regression_d9_p3_r25"])
def test_mh_replicated_regression_chain(name):
    """Replicates in the columns of y with a LinearCombination mean (distribution.py:8-10; the n_rep factor of
    location_scale.py:238-241) == the single-column regression on the design stacked n_rep times."""
    g = _load(name)
    X, y, w = g["X"], g["y"], g["w"]
    dim, n_rep = y.shape
    Xs, ys, ws = np.tile(X, (n_rep, 1)), y.T.reshape(-1, 1), np.tile(w, n_rep)
    p = X.shape[1]
    terms = [mh.Term("normal_linear", data=ys, X=Xs, Q=float(g["tau"]) * np.diag(ws), transform=False),
             mh.Term("normal_response", p1=np.zeros((p, 1)), Q=float(g["lam"]) * np.eye(p))]
    g0, H0 = terms[0].grad_hess_analytic(g["beta0"])
    np.testing.assert_allclose(g0, g["grad0"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(H0, g["hess0"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(terms[0].log_p(g["beta0"]), g["logp0"], rtol=1e-12)
    theta, n_acc = g["beta0"], 0
    from scipy import stats
    a, b, tau = float(g["a"]), float(g["b"]), float(g["tau"])
    lp_tau = stats.gamma.logpdf(tau, a, scale=1 / b)                    # the Gamma prior on tau is part of log_post
    for it in range(g["store_beta"].shape[1]):
        theta, info = mh.mmala_step(terms, theta, 0.8, g["z"][it], g["u"][it], "analytic")
        n_acc += info["accepted"]
        np.testing.assert_allclose(theta.ravel(), g["store_beta"][:, it], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(mh.log_p(terms, theta) + lp_tau, g["store_log_post"][it, 0], rtol=1e-10)
    assert n_acc == g["accept"][0]


def test_truncnorm_grid():
    g = _load("truncnorm_grid")
    x = gmrf.truncated_normal_rv(g["mean"], g["scale"], g["lower"], g["upper"], g["u"])
    np.testing.assert_allclose(x, g["x"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(gmrf.truncated_normal_log_pdf(g["x"], g["mean"], g["scale"], g["lower"], g["upper"]),
                               g["logpdf"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(gmrf.truncated_normal_log_pdf(g["x_other"], g["mean"], g["scale"], g["lower"],
                                                             g["upper"]), g["logpdf_other"], rtol=1e-11, atol=1e-11)
