"""GPU parity of the temporal-GMRF path (BASELINE configs[2]): omc_tridiag_nn_draw against the
numpy oracle (oracle/gmrf.py, oracle/conjugate.py) on seeded inputs, the MCMC driver against golden chains recorded
from the live reference (dense notebook form and sparse form), and size-independent properties at n = 1e6.

Tolerances: factor / mean / log-det rel 1e-10, injected-z draws 1e-9 (BASELINE.json north_star); on irregular grids
the reference's own dense-vs-sparse disagreement is kappa*eps, still inside these bounds for the cases
here."""

import glob
import os

import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "gmrf_*.npz")))


def _rw1(n, rng, irregular):
    from oracle import gmrf

    s = np.cumsum(0.2 + rng.random(n)) if irregular else np.arange(n) * (60.0 / 99.0)
    pd, pe = gmrf.precision_irregular_diagonals(s)
    pd = pd.copy()
    pd[0] += 1e-3
    return pd, pe


def _dev(a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("n,C,irregular,weighted,with_mu", [
    (1, 2, False, False, False), (2, 3, False, True, True), (7, 2, True, False, False), (2047, 2, False, False, True),
    (2048, 3, True, True, False), (2049, 2, False, False, False), (10000, 3, True, True, True),
    (70001, 2, False, False, False), (5000, 3, True, False, False), (4609, 5, False, False, False)])
def test_tridiag_draw_matches_oracle(n, C, irregular, weighted, with_mu):
    import torch

    from openmcmc_b200 import kernels as K
    from oracle import conjugate, gmrf

    K.init_device()
    rng = np.random.default_rng(n + C)
    pd, pe = _rw1(n, rng, irregular) if n > 1 else (np.array([1.3]), np.zeros(0))
    w = rng.random((C, n)) + 0.5 if weighted else np.ones((C, n))
    y = rng.standard_normal((C, n)) + 2
    mu0 = rng.standard_normal(n) if with_mu else np.zeros(n)
    lam = rng.random(C) * 100 + 1
    tau = rng.random(C) * 2 + 0.3
    z = rng.standard_normal((C, n))
    d_pd, d_pe, d_y, d_z, d_lam, d_tau = _dev(pd), _dev(pe), _dev(y), _dev(z), _dev(lam), _dev(tau)
    d_w = _dev(w) if weighted else None
    d_mu = _dev(mu0) if with_mu else None
    h = None
    if with_mu:
        h = torch.empty(n, dtype=torch.float64, device="cuda")
        K.tridiag_matvec(d_pd, d_pe, K.vec(d_mu), 1, n, h)
    ws = torch.zeros(K.tridiag_workspace(C, n), dtype=torch.uint8, device="cuda")
    x = torch.empty(C, n, dtype=torch.float64, device="cuda")
    out = {k: torch.zeros(C, dtype=torch.float64, device="cuda") for k in ("ss_prior", "ss_lik", "logdet")}
    pl = torch.empty(C, n, dtype=torch.float64, device="cuda")
    pc = torch.empty(C, max(n - 1, 1), dtype=torch.float64, device="cuda")
    status = torch.zeros(C, dtype=torch.int32, device="cuda")
    common = dict(lam=K.vec(d_lam, 1), tau=K.vec(d_tau, 1), w=K.vec(d_w, n) if weighted else None, y=K.vec(d_y, n),
                  h=K.vec(h) if with_mu else None, mu0=K.vec(d_mu) if with_mu else None)
    for rep in range(2):   # twice: the workspace (tickets, epochs, counters) must re-arm itself
        K.tridiag_nn_draw(K.tridiag_args(C, n, d_pd, d_pe, ws, x=x, debug_z=d_z, probe_l=pl, probe_c=pc, status=status,
                                         **out, **common))
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    xs = x.cpu().numpy()
    mean = torch.empty_like(x)
    d_zero = torch.zeros_like(d_z)
    K.tridiag_nn_draw(K.tridiag_args(C, n, d_pd, d_pe, ws, x=mean, debug_z=d_zero, **common))
    ss2 = {k: torch.zeros(C, dtype=torch.float64, device="cuda") for k in ("ss_prior", "ss_lik")}
    K.tridiag_quadforms(K.tridiag_args(C, n, d_pd, d_pe, ws, x=x, **ss2, **common))
    torch.cuda.synchronize()
    for c in range(C):
        o = conjugate.gmrf_normal_normal(pd, pe, w[c], y[c], mu0, lam[c], tau[c], z[c])
        np.testing.assert_allclose(pl[c].cpu().numpy(), o["l"], rtol=1e-10)
        if n > 1:
            np.testing.assert_allclose(pc[c, : n - 1].cpu().numpy(), o["c"], rtol=1e-10)
        np.testing.assert_allclose(mean[c].cpu().numpy(), o["mu"], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(xs[c], o["x"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(out["logdet"][c].item(), 2 * np.sum(np.log(o["l"])), rtol=1e-11)
        ssp = gmrf.tridiag_quadform(pd, pe, o["x"] - mu0) if n > 1 else pd[0] * (o["x"][0] - mu0[0]) ** 2
        np.testing.assert_allclose(out["ss_prior"][c].item(), ssp, rtol=1e-9)
        np.testing.assert_allclose(out["ss_lik"][c].item(), np.sum(w[c] * (y[c] - o["x"]) ** 2), rtol=1e-10)
        np.testing.assert_allclose(ss2["ss_prior"][c].item(), out["ss_prior"][c].item(), rtol=1e-11)
        np.testing.assert_allclose(ss2["ss_lik"][c].item(), out["ss_lik"][c].item(), rtol=1e-11)


def test_tridiag_not_positive_definite_sets_status():
    import torch

    from openmcmc_b200 import kernels as K

    K.init_device()
    n, C = 5000, 2
    pd = np.full(n, 2.0)
    pe = np.full(n - 1, -1.0)
    pd[3000] = -5.0   # indefinite
    ws = torch.zeros(K.tridiag_workspace(C, n), dtype=torch.uint8, device="cuda")
    x = torch.empty(C, n, dtype=torch.float64, device="cuda")
    status = torch.zeros(C, dtype=torch.int32, device="cuda")
    tau = _dev(np.array([0.1, 100.0]))   # chain 1 is rescued by the likelihood term
    d_pd, d_pe, d_y, d_z = _dev(pd), _dev(pe), _dev(np.ones(n)), _dev(np.zeros((C, n)))
    K.tridiag_nn_draw(K.tridiag_args(C, n, d_pd, d_pe, ws, x=x, tau=K.vec(tau, 1), y=K.vec(d_y), debug_z=d_z,
                                     status=status))
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [1, 0]


def _build(g, n_chains=1):
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    n = g["y"].size
    P = sparse.diags([g["pe"], g["pd"], g["pe"]], offsets=[-1, 0, 1], format="csc")
    W = sparse.diags(g["w"], format="csc")
    mean = "b" if str(g["form"]) == "notebook" else LinearCombination(form={"b": "I"})
    mdl = Model([Normal("y", mean=mean, precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
                 Normal("b", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
                 Gamma("lambda", shape="a_lam", rate="b_lam"),
                 Gamma("tau", shape="a_tau", rate="b_tau")])
    state = {"y": g["y"].copy(), "b": g["y"].copy(), "mu": g["mu"], "lambda": 100, "P_lambda": P, "a_lam": 10, "b_lam": 1,
             "tau": 1, "P_tau": W, "a_tau": 1, "b_tau": 1, "I": sparse.identity(n, format="csc")}
    smap = {"b": NormalNormal("b", mdl), "lambda": NormalGamma("lambda", mdl), "tau": NormalGamma("tau", mdl)}
    return mdl, [smap[str(k)] for k in g["order"]], state


@pytest.mark.parametrize("name", NAMES)
def test_mcmc_replays_reference_gmrf_chain(name):
    from openmcmc_b200.mcmc import MCMC

    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    mdl, samplers, state = _build(g)
    n_iter = g["store_b"].shape[1]
    dd = {"b": {"z": g["z"]}, "lambda": {"g": g["g_lambda"]}, "tau": {"g": g["g_tau"]}}
    M = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=n_iter, debug_draws=dd)
    M.run_mcmc()
    assert M.store["b"].shape == g["store_b"].shape
    np.testing.assert_allclose(M.store["b"], g["store_b"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(M.store["lambda"], g["store_lambda"], rtol=1e-9)
    np.testing.assert_allclose(M.store["tau"], g["store_tau"], rtol=1e-9)
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)
    assert M.launches_per_sweep() <= 7   # aggregate + tile scan + solve + partials + two gamma draws + counter: the quadratic forms are fused


def test_gmrf_free_running_smoother_recovers_truth():
    """Free-running chains on the example-4 data: the posterior mean tracks the truth, tau is near its true value 1,
    chains differ, and chain sharding (chain_offset) reproduces the same draws bit for bit."""
    from openmcmc_b200.mcmc import MCMC

    g = dict(np.load(os.path.join(GOLD, "gmrf_n2500_sparse_weighted_mu.npz")))
    g["mu"] = np.zeros_like(g["mu"])
    g["w"] = np.ones_like(g["w"])
    mdl, samplers, state = _build(g)
    C = 16
    M = MCMC(state, samplers, model=mdl, n_burn=100, n_iter=100, n_chains=C, seed=5)
    M.run_mcmc()
    truth = np.sin(g["s"] / 20) + 2 * np.cos(g["s"] / 12) + 2
    post_mean = M.store["b"].mean(axis=(0, 2))
    assert np.sqrt(np.mean((post_mean - truth) ** 2)) < 0.4   # noise sd is 1: the smoother must beat y itself by > 2x
    # hyper-parameters against a free-running CPU oracle chain (numpy RNG): same posterior within Monte-Carlo error
    from oracle import conjugate

    rng = np.random.default_rng(0)
    n = g["y"].size
    s = {"b": g["y"].copy(), "lambda": 100.0, "tau": 1.0, "a_lam": 10.0, "b_lam": 1.0, "a_tau": 1.0, "b_tau": 1.0}
    taus, lams = [], []
    for it in range(160):
        s = conjugate.gibbs_gmrf_sweep(g["pd"], g["pe"], g["w"], g["y"], g["mu"], s, rng.standard_normal(n),
                                       rng.standard_gamma(10.0 + n / 2), rng.standard_gamma(1.0 + n / 2),
                                       tuple(str(k) for k in g["order"]))
        if it >= 100:
            taus.append(s["tau"])
            lams.append(s["lambda"])
    assert abs(M.store["tau"].mean() - np.mean(taus)) < 0.05 * np.mean(taus)
    assert abs(M.store["lambda"].mean() - np.mean(lams)) < 0.15 * np.mean(lams)
    assert np.std(M.store["b"][:, 100, -1]) > 0
    assert np.all(M.status == 0)
    M2 = MCMC(state, samplers, model=mdl, n_burn=100, n_iter=100, n_chains=C // 2, seed=5, chain_offset=C // 2)
    M2.run_mcmc()
    np.testing.assert_array_equal(M2.store["b"], M.store["b"][C // 2:])
    np.testing.assert_array_equal(M2.store["lambda"], M.store["lambda"][C // 2:])


def test_full_size_properties_n_1e6():
    """BASELINE configs[2] size (n = 1e6): size-independent properties instead of an oracle run —
    Q mu = b (posterior mean), L'(x - mu) = z (draw) with the probed factor, L L' = Q, and sampled normals ~ N(0,1)."""
    import torch
    from scipy import stats

    from openmcmc_b200 import kernels as K

    K.init_device()
    n, C = 1_000_000, 3
    rng = np.random.default_rng(0)
    s = np.arange(n) * (60.0 / 99.0)
    dr = 1.0 / np.diff(s)
    pd = np.append(np.append(dr[0], dr[:-1] + dr[1:]), dr[-1])
    pd[0] += 1e-3
    pe = -dr
    y = np.sin(s / 20) + 2 * np.cos(s / 12) + 2 + rng.standard_normal((C, n))
    lam, tau = np.array([100.0, 37.0, 400.0]), np.array([1.0, 0.6, 2.0])
    d_pd, d_pe, d_y = _dev(pd), _dev(pe), _dev(y)
    ws = torch.zeros(K.tridiag_workspace(C, n), dtype=torch.uint8, device="cuda")
    mean = torch.empty(C, n, dtype=torch.float64, device="cuda")
    x = torch.empty_like(mean)
    pl, pc = torch.empty_like(mean), torch.empty(C, n - 1, dtype=torch.float64, device="cuda")
    d_lam, d_tau, d_zero = _dev(lam), _dev(tau), torch.zeros_like(mean)   # named: kernels hold raw pointers
    common = dict(lam=K.vec(d_lam, 1), tau=K.vec(d_tau, 1), y=K.vec(d_y, n))
    status = torch.zeros(C, dtype=torch.int32, device="cuda")
    K.tridiag_nn_draw(K.tridiag_args(C, n, d_pd, d_pe, ws, x=mean, debug_z=d_zero, probe_l=pl, probe_c=pc,
                                     status=status, **common))
    seedc = torch.zeros(1, dtype=torch.int64, device="cuda")
    K.tridiag_nn_draw(K.tridiag_args(C, n, d_pd, d_pe, ws, x=x, rng_=K.rng(seed=9, sweep=seedc, site=1), **common))
    torch.cuda.synchronize()
    assert int(status.max()) == 0
    mu, xs, l, c = mean.cpu().numpy(), x.cpu().numpy(), pl.cpu().numpy(), pc.cpu().numpy()
    for k in range(C):
        d = lam[k] * pd + tau[k]
        e = lam[k] * pe
        # L L' = Q
        np.testing.assert_allclose(l[k] ** 2 + np.append(0.0, c[k] ** 2), d, rtol=1e-12)
        np.testing.assert_allclose(l[k][:-1] * c[k], e, rtol=1e-12)
        # Q mu = b
        Qmu = d * mu[k]
        Qmu[:-1] += e * mu[k][1:]
        Qmu[1:] += e * mu[k][:-1]
        np.testing.assert_allclose(Qmu, tau[k] * y[k], rtol=1e-9, atol=1e-9)
        # z = L'(x - mu) is standard normal
        v = xs[k] - mu[k]
        z = l[k] * v
        z[:-1] += c[k] * v[1:]
        assert abs(z.mean()) < 5e-3 and abs(z.std() - 1) < 5e-3
        assert stats.kstest(z[::97], "norm").pvalue > 0.01
