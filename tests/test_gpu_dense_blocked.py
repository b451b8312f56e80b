"""GPU parity: the blocked Cholesky (DMMA trailing update) entry points and the re-centred sufficient statistics.

  * omc_dense_factor (gmrf.cholesky / cho_solve / solve / sample_normal_canonical / log-det; gmrf.py:414-486, 167-198)
    against numpy / scipy on seeded SPD matrices, n from 1 to 300 (shared-memory and global-workspace storage);
  * omc_nn_dense_draw's rss epilogue: rss(beta) = rss0 - 2 d'c0 + d'G d against the explicit residual of the oracle
    (sampler.py:275-284), including a high signal-to-noise case where y'y - 2 g'beta + beta'G beta would cancel;
  * MCMC runs with engine.RECENTER on / off give the same chains (draws 1e-9, log_post 1e-10).
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spd(rng, m, n, cond=1e3):
    A = rng.standard_normal((m, n, n))
    Q = A @ A.transpose(0, 2, 1) / n + np.eye(n) * (1.0 / cond)
    return Q


@pytest.mark.parametrize("n", [1, 2, 7, 16, 17, 33, 64, 65, 100, 128, 136, 137, 200, 256, 300])
def test_dense_factor_matches_numpy(n):
    import torch
    from scipy.linalg import cho_solve, solve_triangular

    from openmcmc_b200 import kernels as K

    K.init_device(0)
    rng = np.random.default_rng(100 + n)
    m = 3
    Q = _spd(rng, m, n)
    b = rng.standard_normal((m, n))
    z = rng.standard_normal((m, n))
    t = lambda v: torch.tensor(v, device="cuda")
    dQ, db, dz = t(Q), t(b), t(z)
    L = torch.empty(m, n, n, dtype=torch.float64, device="cuda")
    ld = torch.empty(m, dtype=torch.float64, device="cuda")
    mean = torch.empty(m, n, dtype=torch.float64, device="cuda")
    x = torch.empty(m, n, dtype=torch.float64, device="cuda")
    status = torch.zeros(m, dtype=torch.int32, device="cuda")
    ws = K.nn_dense_workspace(m, n)
    dws = torch.empty(ws, dtype=torch.float64, device="cuda") if ws else None
    K.dense_factor(dQ, n, b=db, z=dz, L=L, logdet=ld, mean=mean, x=x, status=status, workspace=dws)
    torch.cuda.synchronize()
    assert int(status.sum()) == 0
    for c in range(m):
        Lr = np.linalg.cholesky(Q[c])
        np.testing.assert_allclose(L[c].cpu().numpy(), Lr, rtol=1e-10, atol=1e-13)
        assert abs(ld[c].item() - 2 * np.log(np.diag(Lr)).sum()) <= 1e-11 * max(1.0, abs(ld[c].item()))
        mu = cho_solve((Lr, True), b[c])
        np.testing.assert_allclose(mean[c].cpu().numpy(), mu, rtol=1e-9, atol=1e-9 * np.abs(mu).max())
        xr = mu + solve_triangular(Lr.T, z[c], lower=False)
        np.testing.assert_allclose(x[c].cpu().numpy(), xr, rtol=1e-9, atol=1e-9 * np.abs(xr).max())
    # precomputed factor: solve(L', z) (gmrf.py:61) and cho_solve((L, True), b) (gmrf.py:462)
    x2 = torch.empty_like(x)
    mean2 = torch.empty_like(x)
    K.dense_factor(L, n, b=db, z=dz, mean=mean2, x=x2, factored=True, workspace=dws)
    torch.cuda.synchronize()
    np.testing.assert_allclose(mean2.cpu().numpy(), mean.cpu().numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(x2.cpu().numpy(), x.cpu().numpy(), rtol=1e-9, atol=1e-12)
    K.dense_factor(L, n, z=dz, x=x2, factored=True, workspace=dws)
    torch.cuda.synchronize()
    for c in range(m):
        v = solve_triangular(np.linalg.cholesky(Q[c]).T, z[c], lower=False)
        np.testing.assert_allclose(x2[c].cpu().numpy(), v, rtol=1e-9, atol=1e-9 * np.abs(v).max())


def test_dense_factor_flags_non_pd():
    import torch

    from openmcmc_b200 import kernels as K

    K.init_device(0)
    n = 40
    Q = np.stack([np.eye(n), np.eye(n)])
    Q[0, 20, 20] = -1.0
    dQ = torch.tensor(Q, device="cuda")
    ld = torch.zeros(2, dtype=torch.float64, device="cuda")
    status = torch.zeros(2, dtype=torch.int32, device="cuda")
    K.dense_factor(dQ, n, logdet=ld, status=status)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [1, 0]
    assert np.isnan(ld[0].item()) and abs(ld[1].item()) < 1e-14


@pytest.mark.parametrize("C,n,p,noise", [(3, 400, 3, 0.1), (2, 300, 31, 0.1), (3, 500, 64, 0.1), (2, 700, 128, 0.1),
                                         (2, 900, 200, 0.1), (3, 600, 24, 1e-4), (2, 800, 64, 1e-5)])
def test_recentred_rss_matches_explicit_residual(C, n, p, noise):
    """Prologue (solve-only centre, residual stream, fused pass) + draw epilogue, against the oracle's explicit
    residual at the drawn beta.  With noise 1e-5 the raw form y'y - 2 g'b + b'Gb keeps ~6 digits; the re-centred one
    keeps the 1e-10 bar."""
    import torch

    from openmcmc_b200 import kernels as K
    from oracle import conjugate

    K.init_device(0)
    rng = np.random.default_rng(17 + p)
    X = rng.standard_normal((C, n, p))
    X[:, :, 0] = 1.0
    y = (X @ rng.standard_normal((C, p, 1)))[:, :, 0] + noise * rng.standard_normal((C, n))
    tau = rng.random(C) * 0.5 / noise ** 2 + 0.5
    lam = rng.random(C) + 0.01
    z = rng.standard_normal((C, p))
    t = lambda v: torch.tensor(v, device="cuda")
    dX, dy, dtau, dlam, dz = t(X), t(y), t(tau), t(lam), t(z)
    rec = p * p + p + 2
    stats = torch.zeros((C, rec), dtype=torch.float64, device="cuda")
    ns, wsz = K.reg_pass_workspace(C, n, p)
    work = torch.empty(max(wsz, 1), dtype=torch.float64, device="cuda")
    K.reg_pass(dX, dy, None, None, stats, work, C, n, p)
    dws_n = K.nn_dense_workspace(C, p)
    dws = torch.empty(dws_n, dtype=torch.float64, device="cuda") if dws_n else None
    zero = torch.zeros(1, dtype=torch.float64, device="cuda")
    bhat = torch.empty((C, p), dtype=torch.float64, device="cuda")
    K.nn_dense_draw(C, p, stats, K.vec(None), K.MAT_EYE, K.vec(None), K.vec(zero), K.vec(None), bhat, K.rng(),
                    solve_only=True, ridge_rel=1e-12, workspace=dws)
    r0 = torch.empty((C, n), dtype=torch.float64, device="cuda")
    K.linear_predictor(C, n, [(K.vec(dX, n * p), K.vec(bhat, p), p, False)], r0, residual_of=K.vec(dy, n))
    scratch = torch.empty_like(stats)
    K.reg_pass(dX, r0, None, None, scratch, work, C, n, p)
    center = torch.cat([bhat, scratch[:, p * p:]], dim=1).contiguous()
    beta = torch.empty((C, p), dtype=torch.float64, device="cuda")
    K.nn_dense_draw(C, p, stats, K.vec(dtau, 1), K.MAT_EYE, K.vec(None), K.vec(dlam, 1), K.vec(None), beta,
                    K.rng(seed=1, site=2), debug_z=dz, center=center, rss_out=stats.data_ptr() + 8 * (p * p + p),
                    workspace=dws)
    torch.cuda.synchronize()
    out, b = stats.cpu().numpy(), beta.cpu().numpy()
    for c in range(C):
        # the centre is the least-squares point up to the 1e-12 jitter
        ls = np.linalg.lstsq(X[c], y[c], rcond=None)[0]
        np.testing.assert_allclose(bhat[c].cpu().numpy(), ls, rtol=1e-7, atol=1e-9)
        _, _, rss, _ = conjugate.regression_suffstats(X[c], y[c], None, b[c])
        assert abs(out[c, p * p + p] - rss) <= 1e-10 * rss, (out[c, p * p + p], rss)


def test_mcmc_recentred_equals_explicit_residual_chain():
    """The same seeded run with engine.RECENTER on / off: identical draws (the rss values differ in the last digits
    only, the Gamma draws that consume them therefore agree to ~1e-12)."""
    from scipy import sparse

    from openmcmc_b200 import engine
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    rng = np.random.default_rng(5)
    C, n, p = 6, 700, 80
    X = rng.standard_normal((C, n, p))
    y = X @ rng.standard_normal((C, p, 1)) + 0.1 * rng.standard_normal((C, n, 1))
    w = rng.random(n) + 0.1

    def run(flag):
        mdl = Model([Normal("y", mean=LinearCombination(form={"beta": "X"}),
                            precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
                     Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
                     Gamma("tau", shape="a", rate="b"), Gamma("lambda", shape="a", rate="b")], response={"y": "mean"})
        samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
        state = {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.diags(w, format="csc"), "tau": 1.0,
                 "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": 0.01, "a": 1e-3,
                 "b": 1e-3}
        old = engine.RECENTER
        engine.RECENTER = flag
        try:
            M = MCMC(state, samplers, model=mdl, n_burn=3, n_iter=8, n_thin=2, n_chains=C, seed=3)
            M.run_mcmc()
        finally:
            engine.RECENTER = old
        labels = [label for label, _ in M._ops["sweep"]]
        return M, labels

    Ma, la = run(True)
    Mb, lb = run(False)
    assert not any(x.startswith("reg_") for x in la), la           # no pass over X in the steady-state sweep
    assert any(x.startswith("reg_rss") for x in lb), lb
    for key in ("beta", "tau", "lambda", "y"):
        np.testing.assert_allclose(Ma.store[key], Mb.store[key], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(Ma.store["log_post"], Mb.store["log_post"], rtol=1e-10)
