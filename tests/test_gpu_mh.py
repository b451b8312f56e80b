"""GPU parity of the Metropolis-Hastings family: truncated-normal special functions, model
log_p / gradient / Hessian, RandomWalk / RandomWalkLoop / ManifoldMALA chains replayed against goldens recorded from the
live reference (tests/golden/make_golden.py) and against the numpy oracle (oracle/mh.py), plus free-running statistics.

Tolerances (BASELINE.json north_star): deterministic quantities rel 1e-10, injected-draw chains 1e-9 — on the ANALYTIC
paths.  The reference's finite-difference derivatives cannot be reproduced beyond their own noise (gradient
abs ~1e-9, Hessian abs ~1e-5), so the FD parity mode is held to that noise and the FD chain to 1e-5."""

import os

import numpy as np
import pytest
from scipy import sparse, stats

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False))


def _pg_model():
    from openmcmc_b200.distribution.distribution import Gamma, Poisson
    from openmcmc_b200.model import Model

    return Model([Poisson("y", rate="lam"), Gamma("lam", shape="a", rate="b")])


def _pg_state(g):
    return {"y": g["y"], "lam": g["lam0"].copy(), "a": g["a"], "b": g["b"]}


def _pg_terms(g):
    from oracle import mh

    return [mh.Term("poisson_rate", data=g["y"]), mh.Term("gamma_response", p1=g["a"], p2=g["b"])]


# ------------------------------------------------------------------------------------------------ special functions
def test_truncnorm_kernels_match_scipy_golden():
    import torch

    from openmcmc_b200 import kernels as K

    K.init_device()
    g = _load("truncnorm_grid")
    d = {k: torch.as_tensor(v).cuda() for k, v in g.items()}
    out = torch.empty_like(d["u"])
    K.truncnorm_rv(d["mean"], d["scale"], d["lower"], d["upper"], d["u"], out)
    np.testing.assert_allclose(out.cpu().numpy(), g["x"], rtol=1e-9, atol=1e-9)
    K.truncnorm_logpdf(d["x"], d["mean"], d["scale"], d["lower"], d["upper"], out)
    np.testing.assert_allclose(out.cpu().numpy(), g["logpdf"], rtol=1e-10, atol=1e-10)
    K.truncnorm_logpdf(d["x_other"], d["mean"], d["scale"], d["lower"], d["upper"], out)
    np.testing.assert_allclose(out.cpu().numpy(), g["logpdf_other"], rtol=1e-10, atol=1e-10)


# ------------------------------------------------------------------------------------------------ log_p / derivatives
@pytest.mark.parametrize("name", ["mmala_poisson_gamma_p6", "mmala_poisson_gamma_p32_vec"])
def test_model_logp_grad_hess_poisson_gamma(name):
    """Model.log_p / grad_log_p through the reference's own call signatures (host dict state)."""
    from oracle import mh

    g = _load(name)
    mdl = _pg_model()
    state = _pg_state(g)
    np.testing.assert_allclose(mdl.log_p(state), g["logp0"], rtol=1e-12)
    # analytic derivatives == oracle analytic (rel 1e-10), and agree with the reference's FD to its truncation error
    ga, Ha = mdl["y"].grad_log_p(state, "lam", method="analytic")
    gb, Hb = mdl["lam"].grad_log_p(state, "lam", method="analytic")
    go, Ho = mh.grad_hess(_pg_terms(g), g["lam0"], "analytic")
    np.testing.assert_allclose(ga + gb, go, rtol=1e-10)
    np.testing.assert_allclose(Ha + Hb, Ho, rtol=1e-10)
    np.testing.assert_allclose(ga + gb, g["grad0"], rtol=1e-6, atol=1e-7)
    # the reference's FD stencil evaluated in-kernel: held to the reference's own FD noise
    gf, Hf = mdl.grad_log_p(state, "lam", hessian_required=True)
    assert gf.shape == g["grad0"].shape and Hf.shape == g["hess0"].shape
    np.testing.assert_allclose(gf, g["grad0"], rtol=1e-6, atol=5e-8)
    np.testing.assert_allclose(Hf, g["hess0"], rtol=1e-3, atol=5e-4)
    g_only = mdl.grad_log_p(state, "lam", hessian_required=False)
    np.testing.assert_allclose(g_only, gf, rtol=0, atol=0)


def _normal_model_state(g):
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import ScaledMatrix

    mdl = Model([Normal("theta", mean="mu", precision=ScaledMatrix(matrix="P", scalar="lam")),
                 Normal("yobs", mean="theta", precision=ScaledMatrix(matrix="W", scalar="tau"))])
    state = {"theta": g["theta0"].copy(), "mu": g["mu"], "P": g["P"], "lam": float(g["lam"]), "yobs": g["yobs"],
             "W": sparse.diags(g["w"], format="csc"), "tau": float(g["tau"])}
    return mdl, state


@pytest.mark.parametrize("name", ["mmala_normal_p7", "mmala_normal_p40"])
def test_model_grad_hess_normal_analytic(name):
    g = _load(name)
    mdl, state = _normal_model_state(g)
    gr, H = mdl.grad_log_p(state, "theta", hessian_required=True)
    np.testing.assert_allclose(gr, g["grad0"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(H, g["hess0"], rtol=1e-10, atol=1e-12)


# ------------------------------------------------------------------------------------------------ chains (injected draws)
@pytest.mark.parametrize("name", ["mmala_normal_p7", "mmala_normal_p40"])
def test_mmala_replays_reference_chain_analytic(name):
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA

    g = _load(name)
    mdl, state = _normal_model_state(g)
    smp = ManifoldMALA("theta", mdl, step=np.array([[float(g["step"])]]))
    n_iter = g["store_theta"].shape[1]
    M = MCMC(state, [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"theta": {"z": g["z"], "u": g["u"]}})
    M.run_mcmc()
    np.testing.assert_allclose(M.store["theta"], g["store_theta"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)
    assert smp.accept_rate.count == {"accept": int(g["accept"][0]), "proposal": int(g["accept"][1])}


@pytest.mark.parametrize("name", ["mmala_poisson_gamma_p6", "mmala_poisson_gamma_p32_vec"])
def test_mmala_poisson_gamma_chains(name):
    """FD mode replays the reference chain within FD noise; analytic mode replays the analytic oracle to 1e-9."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA
    from oracle import mh

    g = _load(name)
    mdl = _pg_model()
    n_iter = g["store_lam"].shape[1]
    dd = {"lam": {"z": g["z"], "u": g["u"]}}
    smp = ManifoldMALA("lam", mdl, step=np.array([[float(g["step"])]]), derivatives="fd")
    M = MCMC(_pg_state(g), [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws=dd)
    M.run_mcmc()
    np.testing.assert_allclose(M.store["lam"], g["store_lam"], rtol=1e-5)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-6)
    assert smp.accept_rate.count["accept"] == int(g["accept"][0])
    # analytic derivatives vs the oracle with analytic derivatives, probes of the last step included
    smp = ManifoldMALA("lam", mdl, step=np.array([[float(g["step"])]]))
    M = MCMC(_pg_state(g), [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws=dd, probes=True)
    M.run_mcmc()
    theta = g["lam0"]
    terms = _pg_terms(g)
    for it in range(n_iter):
        theta, info = mh.mmala_step(terms, theta, float(g["step"]), g["z"][it], g["u"][it], "analytic")
        np.testing.assert_allclose(M.store["lam"][:, it], theta.ravel(), rtol=1e-9)
    pr = M.plan.probes["lam"]
    np.testing.assert_allclose(pr["mu"].cpu().numpy()[0], info["mu"].ravel(), rtol=1e-10)
    np.testing.assert_allclose(pr["L"].cpu().numpy()[0], info["L"], rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(pr["prop"].cpu().numpy()[0], info["prop"].ravel(), rtol=1e-9)
    sc = pr["scalars"].cpu().numpy()[0]
    if not info["invalid"]:
        np.testing.assert_allclose(sc[:5], [info["logp_cur"], info["logp_prop"], info["lq_fwd"], info["lq_rev"],
                                            info["log_accept"]], rtol=1e-9)
        assert bool(sc[5]) == info["accepted"]


@pytest.mark.parametrize("name", ["rwl_poisson_gamma_1x8", "rwl_poisson_gamma_1x32"])
def test_random_walk_loop_replays_reference_chain(name):
    """(1, 32) is the BASELINE C4b layout: the column-parallel (lane = column) kernel at its full warp width."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import RandomWalkLoop

    g = _load(name)
    width = g["store_lam"].shape[1]
    mdl = _pg_model()
    smp = RandomWalkLoop("lam", mdl, step=np.array([[float(g["step"])]]), domain_limits=g["limits"],
                         max_variable_size=(1, width))
    n_iter = g["store_lam"].shape[2]
    M = MCMC(_pg_state(g), [smp], model=mdl, n_burn=0, n_iter=n_iter,
             debug_draws={"lam": {"tn_u": g["tn_u"], "u": g["u"]}})
    M.run_mcmc()
    assert M.store["lam"].shape == g["store_lam"].shape
    np.testing.assert_allclose(M.store["lam"], g["store_lam"], rtol=1e-9)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)
    assert smp.accept_rate.count == {"accept": int(g["accept"][0]), "proposal": int(g["accept"][1])}
    assert M.state["lam"].shape == (1, width)


@pytest.mark.parametrize("name", ["rw_poisson_gamma_p6", "rw_trunc_scalar"])
def test_random_walk_replays_reference_chain(name):
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import RandomWalk

    g = _load(name)
    mdl = _pg_model()
    smp = RandomWalk("lam", mdl, step=g["step"], domain_limits=g["limits"] if g["limits"].size else None)
    n_iter = g["store_lam"].shape[1]
    M = MCMC(_pg_state(g), [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"lam": {"z": g["z"], "u": g["u"]}})
    M.run_mcmc()
    np.testing.assert_allclose(M.store["lam"], g["store_lam"], rtol=1e-9)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)
    assert smp.accept_rate.count == {"accept": int(g["accept"][0]), "proposal": int(g["accept"][1])}


def test_random_walk_loop_without_limits_raises_like_reference():
    """RandomWalkLoop without domain_limits cannot work for n_rep > 1 (ValueError in the reference)."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import RandomWalkLoop

    g = _load("rwl_poisson_gamma_1x8")
    mdl = _pg_model()
    with pytest.raises(ValueError):
        MCMC(_pg_state(g), [RandomWalkLoop("lam", mdl)], model=mdl, n_burn=0, n_iter=1).run_mcmc()


# ------------------------------------------------------------------------------------------------ free-running chains
@pytest.mark.parametrize("kind,p", [("mmala", 8), ("rwl", 8), ("rwl", 32), ("mmala", 32)])
def test_free_running_poisson_gamma_posterior(kind, p):
    """Poisson counts with a Gamma(a, b) prior are conjugate: lam_j | y ~ Gamma(a + y_j, b + 1).  One draw per chain
    after burn-in is an independent posterior sample: KS p > 0.01 per coordinate (north_star), means within MC error."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA, RandomWalkLoop

    C = 2048
    rng = np.random.default_rng(3)
    y = rng.poisson(rng.gamma(5.0, 1.0, size=p)).astype(float)
    mdl = _pg_model()
    if kind == "mmala":
        state = {"y": y.reshape(p, 1), "lam": (y + 1.0).reshape(p, 1), "a": np.array([[2.0]]), "b": np.array([[0.5]])}
        # (the acceptance rate of a Langevin proposal falls with the dimension: 0.6 accepts 0.7 % of the moves at p = 32)
        smp = ManifoldMALA("lam", mdl, step=np.array([[0.6 if p <= 8 else 0.3]]))
    else:
        state = {"y": y.reshape(1, p), "lam": (y + 1.0).reshape(1, p), "a": np.array([[2.0]]), "b": np.array([[0.5]])}
        smp = RandomWalkLoop("lam", mdl, step=np.array([[2.0]]), domain_limits=np.array([[0.0, np.inf]]),
                             max_variable_size=(1, p))
    n_burn = 300 if p <= 8 else 1500
    M = MCMC(state, [smp], model=mdl, n_burn=n_burn, n_iter=2, n_chains=C, seed=11)
    M.run_mcmc()
    last = M.store["lam"].reshape(C, p, -1)[:, :, -1]
    assert np.all((M.status & 3) == 0)   # bit 4 (a rejected invalid proposal) is informational
    rate = smp.accept_rate.acceptance_rate
    assert 10 < rate < 99.9, rate
    for j in range(p):
        post = stats.gamma(a=2.0 + y[j], scale=1.0 / 1.5)
        # north star: KS p > 0.01 per coordinate; with 32 coordinates tested at once the bar is Bonferroni-scaled so that
        # the family-wise false-alarm rate stays that of 8 coordinates at 0.01
        assert stats.kstest(last[:, j], post.cdf).pvalue > 0.01 * min(1.0, 8.0 / p), (kind, j, rate, y[j], last[:, j].mean())
        assert abs(last[:, j].mean() - post.mean()) < 5 * post.std() / np.sqrt(C)
    # sharding invariance: the second half of the chains computed alone with chain_offset reproduces the same draws
    M2 = MCMC(state, [type(smp)(**{k: getattr(smp, k) for k in ("param", "step", "max_variable_size")}, model=mdl,
                                **({"domain_limits": smp.domain_limits} if kind == "rwl" else {}))],
              model=mdl, n_burn=n_burn, n_iter=2, n_chains=C // 2, seed=11, chain_offset=C // 2)
    M2.run_mcmc()
    np.testing.assert_array_equal(M2.store["lam"], M.store["lam"][C // 2:])


def test_mmala_invalid_proposal_rejects_and_flags():
    """a proposal outside the support makes the reference crash; the device path rejects and flags."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA

    g = _load("mmala_poisson_gamma_p6")
    mdl = _pg_model()
    z = np.full((1, 6), -50.0)   # drives every coordinate far below zero
    smp = ManifoldMALA("lam", mdl, step=np.array([[1.0]]))
    M = MCMC(_pg_state(g), [smp], model=mdl, n_burn=0, n_iter=1, debug_draws={"lam": {"z": z, "u": np.array([0.5])}})
    M.run_mcmc()
    np.testing.assert_array_equal(M.store["lam"][:, 0], g["lam0"].ravel())
    assert M.status[0] == 4
    assert smp.accept_rate.count == {"accept": 0, "proposal": 1}


# ------------------------------------------------------------------------------------------------ SURVEY f4 / a4 in MH
def _lognormal_model_state(g):
    from openmcmc_b200.distribution.location_scale import LogNormal, Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import ScaledMatrix

    mdl = Model([LogNormal("theta", mean="mu", precision=ScaledMatrix(matrix="P", scalar="lam")),
                 Normal("yobs", mean="theta", precision=ScaledMatrix(matrix="W", scalar="tau"))])
    P = g["P"] if str(g["prior"]) == "dense" else sparse.csc_matrix(g["P"])
    state = {"theta": g["theta0"].copy(), "mu": g["mu"], "P": P, "lam": float(g["lam"]), "yobs": g["yobs"],
             "W": sparse.diags(g["w"], format="csc"), "tau": float(g["tau"])}
    return mdl, state


@pytest.mark.parametrize("name", ["lognormal_mmala_p5_dense", "lognormal_mmala_p24_diag", "lognormal_rw_p6_dense"])
def test_lognormal_logp_grad_hess_and_chain(name):
    """LogNormal (location_scale.py:275-418): log_p, response-branch gradient / Hessian to 1e-10 against the reference's
    values, then the mMALA / RandomWalk chain replayed with the reference's draws to 1e-9."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA, RandomWalk

    g = _load(name)
    mdl, state = _lognormal_model_state(g)
    np.testing.assert_allclose(mdl["theta"].log_p(state), g["logp0"], rtol=1e-10)
    gr, H = mdl["theta"].grad_log_p(state, "theta", hessian_required=True)
    np.testing.assert_allclose(gr, g["grad0"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(H, g["hess0"], rtol=1e-10, atol=1e-12)
    step = np.array([[float(g["step"])]])
    smp = ManifoldMALA("theta", mdl, step=step) if str(g["sampler"]) == "mmala" else RandomWalk("theta", mdl, step=step)
    n_iter = g["store_theta"].shape[1]
    M = MCMC(state, [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"theta": {"z": g["z"], "u": g["u"]}})
    M.run_mcmc()
    np.testing.assert_allclose(M.store["theta"], g["store_theta"], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)
    assert smp.accept_rate.count == {"accept": int(g["accept"][0]), "proposal": int(g["accept"][1])}


def _mhreg_model_state(g):
    from openmcmc_b200.distribution.location_scale import LogNormal, Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, LinearCombinationWithTransform, ScaledMatrix

    mean = (LinearCombinationWithTransform(form={"beta": "X"}, transform={"beta": True}) if bool(g["transform"])
            else LinearCombination(form={"beta": "X"}))
    lik = LogNormal if ("lognormal" in g and bool(g["lognormal"])) else Normal
    mdl = Model([lik("y", mean=mean, precision=ScaledMatrix(matrix="W", scalar="tau")),
                 Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P", scalar="lam"))])
    n = g["y"].shape[0]
    W = sparse.diags(g["w"], format="csc") if bool(g["weighted"]) else sparse.identity(n, format="csc")
    state = {"y": g["y"], "X": g["X"], "beta": g["beta0"].copy(), "W": W, "tau": float(g["tau"]),
             "mu": np.zeros_like(g["beta0"]), "P": g["P"], "lam": float(g["lam"])}
    return mdl, state


@pytest.mark.parametrize("name", ["mhreg_mmala_n60_p6", "mhreg_mmala_n200_p30_eye", "mhreg_exp_mmala_n80_p5",
                                  "mhreg_lognormal_mmala_n70_p5", "mhreg_lognormal_mmala_n150_p20_eye"])
def test_mmala_on_regression_coefficients(name):
    """Mean-parameter branch of Normal.grad_log_p inside ManifoldMALA (location_scale.py:234-250), also through the exp
    transform of LinearCombinationWithTransform (parameter.py:232-297): evaluated from the data-only regression record.
    The LogNormal cases are the same branch of LogNormal.grad_log_p (location_scale.py:344-347, 401-404): the record is
    built on log(y) and log_p carries the Jacobian -sum(log y)."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA

    g = _load(name)
    mdl, state = _mhreg_model_state(g)
    np.testing.assert_allclose(mdl["y"].log_p(state), g["logp0"], rtol=1e-10)
    gr, H = mdl["y"].grad_log_p(state, "beta", hessian_required=True)
    np.testing.assert_allclose(gr, g["grad0"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(H, g["hess0"], rtol=1e-10, atol=1e-12)
    if bool(g["transform"]):
        np.testing.assert_allclose(mdl["y"].mean.predictor(state), g["X"] @ np.exp(g["beta0"]), rtol=1e-12)
    smp = ManifoldMALA("beta", mdl, step=np.array([[float(g["step"])]]))
    n_iter = g["store_beta"].shape[1]
    M = MCMC(state, [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"beta": {"z": g["z"], "u": g["u"]}})
    M.run_mcmc()
    np.testing.assert_allclose(M.store["beta"], g["store_beta"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)
    assert smp.accept_rate.count == {"accept": int(g["accept"][0]), "proposal": int(g["accept"][1])}


@pytest.mark.parametrize("name", ["replicated_regression_d40_p5_r6", "replicated_regression_d9_p3_r25"])
def test_mmala_on_replicated_regression(name):
    """y of shape (dim, n_rep) with a LinearCombination mean (distribution.py:8-10, location_scale.py:238-241): compiled
    as the single-column regression on the stacked design (engine.unreplicate); the caller's state comes back unchanged."""
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA

    g = _load(name)
    p = g["X"].shape[1]
    mdl = Model([Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="W", scalar="tau")),
                 Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P", scalar="lam")),
                 Gamma("tau", shape="a", rate="b")])
    state = {"y": g["y"].copy(), "X": g["X"], "beta": g["beta0"].copy(), "W": sparse.diags(g["w"], format="csc"),
             "tau": float(g["tau"]), "mu": np.zeros((p, 1)), "P": sparse.identity(p, format="csc"), "lam": float(g["lam"]),
             "a": float(g["a"]), "b": float(g["b"])}
    np.testing.assert_allclose(mdl["y"].log_p(state), g["logp0"], rtol=1e-10)
    gr, H = mdl["y"].grad_log_p(state, "beta", hessian_required=True)
    np.testing.assert_allclose(gr, g["grad0"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(H, g["hess0"], rtol=1e-10, atol=1e-11)
    smp = ManifoldMALA("beta", mdl, step=np.array([[0.8]]))
    n_iter = g["store_beta"].shape[1]
    M = MCMC(state, [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"beta": {"z": g["z"], "u": g["u"]}})
    M.run_mcmc()
    np.testing.assert_allclose(M.store["beta"], g["store_beta"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)
    assert smp.accept_rate.count == {"accept": int(g["accept"][0]), "proposal": int(g["accept"][1])}
    assert M.state["y"].shape == g["y"].shape and np.array_equal(M.state["y"], g["y"])


def test_mmala_regression_free_running_matches_conjugate_posterior():
    """Free-running mMALA on regression coefficients (fixed tau, lambda): the posterior is Gaussian in closed form, so
    per-coordinate means / variances over many chains must agree within Monte-Carlo error."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA

    g = _load("mhreg_mmala_n60_p6")
    mdl, state = _mhreg_model_state(g)
    C = 256
    smp = ManifoldMALA("beta", mdl, step=np.array([[0.9]]))
    M = MCMC(state, [smp], model=mdl, n_burn=60, n_iter=40, n_thin=2, n_chains=C, seed=2)
    M.run_mcmc()
    b = M.store["beta"]                                               # (C, p, n_iter)
    X, y, w = g["X"], g["y"], g["w"]
    Q = float(g["tau"]) * X.T @ (w[:, None] * X) + float(g["lam"]) * g["P"]
    mean = np.linalg.solve(Q, float(g["tau"]) * X.T @ (w[:, None] * y)).ravel()
    sd = np.sqrt(np.diag(np.linalg.inv(Q)))
    draws = b[:, :, ::8].transpose(1, 0, 2).reshape(b.shape[1], -1)      # thinned further: ~independent
    se = sd / np.sqrt(draws.shape[1] / 2)
    assert np.all(np.abs(draws.mean(axis=1) - mean) < 5 * se), (draws.mean(axis=1), mean, se)
    assert np.all(np.abs(draws.std(axis=1) / sd - 1) < 0.15)


# ------------------------------------------------------------------------------------------------ replicated responses
def _replicated_model_state(g):
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import ScaledMatrix

    dim = g["y"].shape[0]
    if int(g["scaled"]):
        prec = ScaledMatrix(matrix="P", scalar="tau")
        extra = {"P": sparse.diags([g["pdiag"]], [0], format="csc"), "tau": g["tau"].copy()}
    else:
        prec, extra = "tau", {"tau": g["tau"].copy()}
    mdl = Model([Normal("y", mean="h", precision=prec), Normal("h", mean="mu", precision="lambda")])
    state = {"y": g["y"].copy(), "h": g["h0"].copy(), "mu": g["mu"].copy(), "lambda": g["lam"].copy(), **extra}
    assert state["y"].shape[0] == dim
    return mdl, state


@pytest.mark.parametrize("name", ["replicated_d1_r5", "replicated_d3_r7_scaled", "replicated_d2_r4_diag"])
def test_replicated_response_matches_reference(name):
    """y of shape (dim, n_rep) with replicates in its columns and an Identity mean (the reference's examples 1 and 2;
    distribution.py:8-10): log_p / gradient / Hessian at the start state (1e-10) and RandomWalk / NormalNormal chains on
    the mean with the reference's variates injected (1e-9), recorded from the live reference.  The device path compiles
    the single-column form (engine.unreplicate); the caller's model and state come back as they were."""
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.sampler.metropolis_hastings import RandomWalk
    from openmcmc_b200.sampler.sampler import NormalNormal

    g = _load(name)
    dim, n_rep = g["y"].shape
    mdl, state = _replicated_model_state(g)
    np.testing.assert_allclose(mdl.log_p(state), float(g["logp0"]), rtol=1e-10)
    gr, H = mdl.grad_log_p(state, "h")
    np.testing.assert_allclose(gr, g["grad0"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(H, g["hess0"], rtol=1e-10, atol=1e-14)
    n_iter = g["rw_store_h"].shape[1]
    smp = RandomWalk("h", mdl, step=np.array([[4.0]]))
    M = MCMC(state, [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"h": {"z": g["rw_z"], "u": g["rw_u"]}})
    M.run_mcmc()
    np.testing.assert_allclose(M.store["h"], g["rw_store_h"], rtol=1e-9)
    np.testing.assert_allclose(M.store["log_post"], g["rw_store_log_post"], rtol=1e-10)
    assert smp.accept_rate.count == {"accept": int(g["rw_accept"][0]), "proposal": int(g["rw_accept"][1])}
    assert M.state["y"].shape == (dim, n_rep) and not any(k.startswith("__replicates__") for k in M.state)
    assert type(mdl["y"].mean).__name__ == "Identity" and smp.model["y"] is mdl["y"]
    mdl, state = _replicated_model_state(g)
    M = MCMC(state, [NormalNormal("h", mdl)], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"h": {"z": g["nn_z"]}})
    M.run_mcmc()
    np.testing.assert_allclose(M.store["h"], g["nn_store_h"], rtol=1e-9)
    np.testing.assert_allclose(M.store["log_post"], g["nn_store_log_post"], rtol=1e-10)
    # many chains at once share the replicated data
    M = MCMC(state, [NormalNormal("h", mdl)], model=mdl, n_burn=50, n_iter=200, n_chains=64, seed=3)
    M.run_mcmc()
    prec = np.asarray(g["lam"]).reshape(-1)[0] + n_rep * (float(np.ravel(g["tau"])[0]) * (g["pdiag"][0] if int(g["scaled"]) else 1.0)
                                                          if dim == 1 or int(g["scaled"]) else np.ravel(g["tau"])[0])
    post_mean0 = (np.asarray(g["lam"]).reshape(-1)[0] * g["mu"][0, 0] + (prec - np.asarray(g["lam"]).reshape(-1)[0]) * g["y"][0].mean()) / prec
    assert abs(M.store["h"][:, 0, :].mean() - post_mean0) < 5 * prec ** -0.5 / np.sqrt(64 * 200 / 4)


def test_log_p_by_observation_matches_scipy():
    """log_p(state, by_observation=True): one value per replicate column (distribution.py:241-261, 422-442, 490-508,
    location_scale.py:145-167 -> gmrf.py:321-348), against scipy on seeded inputs (rel 1e-10)."""
    from scipy import sparse

    from openmcmc_b200.distribution.distribution import Gamma, Poisson, Uniform
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.parameter import ScaledMatrix

    rng = np.random.default_rng(8)
    p, n = 5, 7
    lam = rng.gamma(3.0, 1.0, size=(p, 1))
    state = {"k": rng.poisson(lam, size=(p, n)).astype(float), "lam": lam, "x": rng.gamma(2.0, 1.0, size=(p, n)),
             "a": rng.random((p, 1)) + 1.0, "b": np.array([[1.7]]), "y": rng.standard_normal((p, n)), "mu": rng.standard_normal((p, 1)),
             "P": sparse.diags(rng.random(p) + 0.5, format="csc"), "tau": np.array([[2.0]]), "u": rng.random((p, n))}
    got = Poisson("k", rate="lam").log_p(state, by_observation=True)
    np.testing.assert_allclose(got, stats.poisson.logpmf(state["k"], lam).sum(axis=0), rtol=1e-10)
    got = Gamma("x", shape="a", rate="b").log_p(state, by_observation=True)
    np.testing.assert_allclose(got, stats.gamma.logpdf(state["x"], state["a"], scale=1 / 1.7).sum(axis=0), rtol=1e-10)
    nrm = Normal("y", mean="mu", precision=ScaledMatrix(matrix="P", scalar="tau"))
    got = nrm.log_p(state, by_observation=True)
    cov = np.linalg.inv(2.0 * state["P"].toarray())
    ref = np.array([stats.multivariate_normal.logpdf(state["y"][:, j], state["mu"][:, 0], cov) for j in range(n)])
    np.testing.assert_allclose(got, ref, rtol=1e-10)
    np.testing.assert_allclose(nrm.log_p(state), ref.sum(), rtol=1e-10)
    uni = Uniform("u", domain_response_lower=np.zeros((p, 1)), domain_response_upper=2.0 * np.ones((p, 1)))
    np.testing.assert_allclose(uni.log_p(state, by_observation=True), np.full(n, -p * np.log(2.0)), rtol=1e-14)
