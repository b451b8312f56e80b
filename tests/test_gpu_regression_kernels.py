"""GPU parity: conjugate-regression kernels (omc_reg_pass / omc_nn_dense_draw / omc_quadform / omc_ng_draw)
against the numpy oracle on identical seeded inputs, through the C-ABI.  Tolerance: rel 1e-10 (BASELINE.json)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _rel(a, b):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _make(C, n, p, seed, weighted):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((C, n, p))
    X[:, :, 0] = 1.0
    beta_true = rng.standard_normal((C, p, 1))
    y = X @ beta_true + 0.1 * rng.standard_normal((C, n, 1))
    w = rng.random((C, n)) + 0.1 if weighted else None
    if weighted:
        w[:, ::7] = 0.0
    beta = rng.standard_normal((C, p))
    return X, y[:, :, 0], w, beta


@pytest.mark.parametrize(
    "C,n,p,weighted",
    [(3, 1000, 3, False), (2, 257, 64, False), (5, 64, 8, True), (2, 1031, 17, True), (1, 5, 2, False),
     (4, 300, 33, False), (300, 130, 64, False), (1, 20000, 64, True), (2, 1, 1, False), (2, 4096, 40, True),
     # p > 64: 64-column panel pairs (gram_wide_kernel) + the streaming rss kernel
     (2, 500, 128, False), (3, 333, 100, True), (1, 3001, 256, True), (2, 70, 65, False), (150, 96, 72, False),
     (1, 700, 300, False)],
)
def test_reg_pass_matches_oracle(C, n, p, weighted):
    import torch

    from openmcmc_b200 import kernels as K
    from oracle import conjugate

    K.init_device(0)
    X, y, w, beta = _make(C, n, p, 1234 + n + p, weighted)
    dX, dy, db = (torch.tensor(v, device="cuda") for v in (X, y, beta))
    dw = torch.tensor(w, device="cuda") if weighted else None
    rec = p * p + p + 2
    stats = torch.full((C, rec), float("nan"), dtype=torch.float64, device="cuda")
    ns, ws = K.reg_pass_workspace(C, n, p)
    work = torch.empty(max(ws, 1), dtype=torch.float64, device="cuda")
    K.reg_pass(dX, dy, dw, db, stats, work, C, n, p)
    torch.cuda.synchronize()
    out = stats.cpu().numpy()
    for c in range(C):
        G, g, rss, cnt = conjugate.regression_suffstats(X[c], y[c], None if w is None else w[c], beta[c])
        assert _rel(out[c, : p * p].reshape(p, p), G) < RTOL
        assert _rel(out[c, p * p : p * p + p], g.ravel()) < RTOL
        assert abs(out[c, p * p + p] - rss) <= RTOL * abs(rss) + 1e-300
        assert out[c, p * p + p + 1] == cnt
    # residual-only pass (omc_reg_rss) with another beta: rss | cnt refreshed, G | g left untouched bit for bit
    beta2 = beta + 0.25
    stats[:, p * p + p:] = float("nan")
    K.reg_rss(dX, dy, dw, torch.tensor(beta2, device="cuda"), stats, work, C, n, p)
    torch.cuda.synchronize()
    out2 = stats.cpu().numpy()
    assert np.array_equal(out2[:, : p * p + p], out[:, : p * p + p])
    for c in range(C):
        _, _, rss, cnt = conjugate.regression_suffstats(X[c], y[c], None if w is None else w[c], beta2[c])
        assert abs(out2[c, p * p + p] - rss) <= RTOL * abs(rss) + 1e-300
        assert out2[c, p * p + p + 1] == cnt


@pytest.mark.parametrize("prior", ["eye", "diag", "dense"])
@pytest.mark.parametrize("C,n,p", [(4, 500, 3), (3, 400, 64), (2, 100, 31), (2, 300, 45), (3, 200, 17), (5, 60, 8),
                                   (2, 50, 1), (300, 130, 64), (2, 120, 33), (2, 300, 100), (3, 400, 128),
                                   (2, 500, 136), (2, 500, 200), (1, 700, 256), (1, 900, 300)])
def test_nn_dense_draw_injected_z(C, n, p, prior):
    import torch

    from openmcmc_b200 import kernels as K
    from oracle import conjugate

    K.init_device(0)
    rng = np.random.default_rng(7 + p)
    X, y, _, _ = _make(C, n, p, 99 + p, False)
    tau = rng.random(C) + 0.5
    lam = rng.random(C) + 0.01
    mu0 = rng.standard_normal((C, p))
    z = rng.standard_normal((C, p))
    if prior == "eye":
        P0, kind = None, K.MAT_EYE
    elif prior == "diag":
        P0, kind = rng.random((C, p)) + 0.2, K.MAT_DIAG
    else:
        A = rng.standard_normal((C, p, p))
        P0, kind = A @ A.transpose(0, 2, 1) + p * np.eye(p), K.MAT_DENSE
    t = lambda v: None if v is None else torch.tensor(v, device="cuda")
    dX, dy, dtau, dlam, dmu0, dz, dP0 = map(t, (X, y, tau, lam, mu0, z, P0))
    rec = p * p + p + 2
    stats = torch.empty((C, rec), dtype=torch.float64, device="cuda")
    ns, ws = K.reg_pass_workspace(C, n, p)
    work = torch.empty(max(ws, 1), dtype=torch.float64, device="cuda")
    K.reg_pass(dX, dy, None, None, stats, work, C, n, p)
    beta = torch.empty((C, p), dtype=torch.float64, device="cuda")
    pQ = torch.empty((C, p, p), dtype=torch.float64, device="cuda")
    pL = torch.empty_like(pQ)
    pb = torch.empty((C, p), dtype=torch.float64, device="cuda")
    pmu = torch.empty_like(pb)
    status = torch.zeros(C, dtype=torch.int32, device="cuda")
    stride = 0 if P0 is None else int(np.prod(P0.shape[1:]))
    ws = K.nn_dense_workspace(C, p)
    dws = torch.empty(ws, dtype=torch.float64, device="cuda") if ws else None
    K.nn_dense_draw(C, p, stats, K.vec(dtau, 1), kind, K.vec(dP0, stride), K.vec(dlam, 1), K.vec(dmu0, p), beta,
                    K.rng(seed=1, site=3), debug_z=dz, probe_Q=pQ, probe_b=pb, probe_L=pL, probe_mu=pmu, status=status,
                    workspace=dws)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    for c in range(C):
        G, g, _, _ = conjugate.regression_suffstats(X[c], y[c])
        P0c = 1.0 if P0 is None else P0[c]
        ref = conjugate.normal_normal_dense(G, g, tau[c], P0c, lam[c], mu0[c], z[c])
        assert _rel(pQ[c].cpu().numpy(), ref["Q"]) < RTOL
        assert _rel(pb[c].cpu().numpy(), ref["b"].ravel()) < RTOL
        assert _rel(pL[c].cpu().numpy(), ref["L"]) < RTOL
        assert _rel(pmu[c].cpu().numpy(), ref["mu"].ravel()) < 1e-9
        assert _rel(beta[c].cpu().numpy(), ref["x"].ravel()) < 1e-9


def test_nn_dense_draw_not_pd_sets_status():
    import torch

    from openmcmc_b200 import kernels as K

    K.init_device(0)
    C, p = 2, 4
    rec = p * p + p + 2
    stats = torch.zeros((C, rec), dtype=torch.float64, device="cuda")
    stats[1, : p * p] = torch.eye(p, dtype=torch.float64).reshape(-1)
    lam = torch.tensor([-1.0, 1.0], dtype=torch.float64, device="cuda")
    beta = torch.zeros((C, p), dtype=torch.float64, device="cuda")
    status = torch.zeros(C, dtype=torch.int32, device="cuda")
    K.nn_dense_draw(C, p, stats, K.vec(None), K.MAT_EYE, K.vec(None), K.vec(lam, 1), K.vec(None), beta,
                    K.rng(seed=1), status=status)
    torch.cuda.synchronize()
    assert status.cpu().tolist() == [1, 0]
    assert torch.isnan(beta[0]).all() and torch.isfinite(beta[1]).all()


def test_quadform_and_ng_draw_injected():
    import torch

    from openmcmc_b200 import kernels as K
    from oracle import conjugate

    K.init_device(0)
    rng = np.random.default_rng(5)
    C, p = 6, 19
    x = rng.standard_normal((C, p))
    mu = rng.standard_normal((C, p))
    A = rng.standard_normal((C, p, p))
    Pd = A @ A.transpose(0, 2, 1)
    Pdiag = rng.random((C, p))
    Pdiag[:, 3] = 0.0
    a0 = rng.random(C) + 0.1
    b0 = rng.random(C)
    b0[0] = 0.0
    g = rng.gamma(3.0, size=C)
    t = lambda v: torch.tensor(v, device="cuda")
    ss = torch.empty(C, dtype=torch.float64, device="cuda")
    cnt = torch.empty_like(ss)
    dx, dmu, da0, db0, dg = t(x), t(mu), t(a0), t(b0), t(g)  # keep alive: omc_vec_t holds raw pointers
    for kind, P, stride in ((K.MAT_DENSE, Pd, p * p), (K.MAT_DIAG, Pdiag, p), (K.MAT_EYE, None, 0)):
        dP = None if P is None else t(P)
        K.quadform(C, p, K.vec(dx, p), K.vec(dmu, p), kind, K.vec(dP, stride), ss, cnt)
        torch.cuda.synchronize()
        for c in range(C):
            s_ref, c_ref = conjugate.quadform(1.0 if P is None else P[c], x[c], mu[c])
            assert abs(ss[c].item() - s_ref) <= 1e-11 * abs(s_ref)
            assert cnt[c].item() == c_ref
    # all-zero residual with b0 = 0 -> infinite sample (reference guard sampler.py:285-286)
    ss0 = torch.zeros(C, dtype=torch.float64, device="cuda")
    out = torch.empty(C, dtype=torch.float64, device="cuda")
    pa = torch.empty_like(out)
    pb = torch.empty_like(out)
    out2 = torch.empty_like(out)
    K.ng_draw(C, K.vec(da0, 1), K.vec(db0, 1), K.vec(ss0, 1), K.vec(cnt, 1), out, K.rng(seed=3), debug_g=dg,
              probe_a=pa, probe_b=pb)
    K.ng_draw(C, K.vec(da0, 1), K.vec(db0, 1), K.vec(ss, 1), K.vec(cnt, 1), out2, K.rng(seed=3), debug_g=dg)
    torch.cuda.synchronize()
    assert np.isinf(out[0].item())
    for c in range(1, C):
        ref, a_ref, b_ref = conjugate.normal_gamma(a0[c], b0[c], 0.0, cnt[c].item(), g[c])
        assert abs(out[c].item() - ref) <= 1e-12 * abs(ref)
        assert abs(pa[c].item() - a_ref) <= 1e-15 * a_ref and abs(pb[c].item() - b_ref) <= 1e-15 * max(b_ref, 1e-300)
    for c in range(C):
        ref, _, _ = conjugate.normal_gamma(a0[c], b0[c], conjugate.quadform(1.0, x[c], mu[c])[0], cnt[c].item(), g[c])
        assert abs(out2[c].item() - ref) <= 1e-11 * abs(ref)


def test_philox_normals_and_gammas_moments():
    """Free-running draws: N(0,1) and Gamma(a,1) moments within Monte-Carlo error (KS p > 0.01)."""
    import torch
    from scipy import stats

    from openmcmc_b200 import kernels as K

    K.init_device(0)
    C, p = 4096, 16
    rec = p * p + p + 2
    st = torch.zeros((C, rec), dtype=torch.float64, device="cuda")
    st[:, : p * p] = torch.eye(p, dtype=torch.float64, device="cuda").reshape(-1)
    beta = torch.empty((C, p), dtype=torch.float64, device="cuda")
    sweep = torch.zeros(1, dtype=torch.int64, device="cuda")
    draws = []
    for it in range(3):
        K.nn_dense_draw(C, p, st, K.vec(None), K.MAT_EYE, K.vec(None), K.vec(None), K.vec(None), beta,
                        K.rng(seed=11, sweep=sweep, site=1))
        K.counter_add(sweep, 1)
        draws.append(beta.cpu().numpy().copy())
    # Q = 2I -> draws ~ N(0, 1/2)
    z = np.concatenate(draws).ravel() * np.sqrt(2.0)
    assert stats.kstest(z, "norm").pvalue > 0.01
    assert not np.allclose(draws[0], draws[1])  # sweep counter advances the stream
    for shape in (0.3, 1.0, 7.5, 5001.0):
        a0 = torch.full((C,), shape, dtype=torch.float64, device="cuda")
        one = torch.ones(C, dtype=torch.float64, device="cuda")
        zero = torch.zeros(C, dtype=torch.float64, device="cuda")
        out = torch.empty(C, dtype=torch.float64, device="cuda")
        K.ng_draw(C, K.vec(a0, 1), K.vec(one, 1), K.vec(zero, 1), K.vec(zero, 1), out, K.rng(seed=5, sweep=sweep, site=2))
        torch.cuda.synchronize()
        assert stats.kstest(out.cpu().numpy(), "gamma", args=(shape,)).pvalue > 0.01


@pytest.mark.parametrize("prior", ["eye", "diag", "dense"])
@pytest.mark.parametrize("C,n,p", [(5, 40, 1), (4, 300, 3), (3, 90, 8), (3, 90, 9), (4, 200, 17), (2, 150, 31),
                                   (2, 150, 32), (3, 200, 33), (3, 300, 40), (2, 300, 45), (2, 300, 57), (2, 300, 63),
                                   (301, 130, 64)])
def test_nn_dense_draw_without_probes_matches_oracle(C, n, p, prior):
    """The production path of omc_nn_dense_draw at p <= 64 (one warp per chain, Q in registers, dense_warp.cu; no
    probe outputs): the draw itself against the oracle's mu + L^-T z, and the solve-only mode against Q^-1 b."""
    import torch

    from openmcmc_b200 import kernels as K
    from oracle import conjugate

    K.init_device(0)
    rng = np.random.default_rng(70 + p)
    X, y, _, _ = _make(C, n, p, 199 + p, False)
    tau = rng.random(C) + 0.5
    lam = rng.random(C) + 0.01
    mu0 = rng.standard_normal((C, p))
    z = rng.standard_normal((C, p))
    if prior == "eye":
        P0, kind = None, K.MAT_EYE
    elif prior == "diag":
        P0, kind = rng.random((C, p)) + 0.2, K.MAT_DIAG
    else:
        A = rng.standard_normal((C, p, p))
        P0, kind = A @ A.transpose(0, 2, 1) + p * np.eye(p), K.MAT_DENSE
    t = lambda v: None if v is None else torch.tensor(v, device="cuda")
    dX, dy, dtau, dlam, dmu0, dz, dP0 = map(t, (X, y, tau, lam, mu0, z, P0))
    rec = p * p + p + 2
    stats = torch.empty((C, rec), dtype=torch.float64, device="cuda")
    ns, ws = K.reg_pass_workspace(C, n, p)
    work = torch.empty(max(ws, 1), dtype=torch.float64, device="cuda")
    K.reg_pass(dX, dy, None, None, stats, work, C, n, p)
    beta = torch.empty((C, p), dtype=torch.float64, device="cuda")
    mean = torch.empty((C, p), dtype=torch.float64, device="cuda")
    status = torch.zeros(C, dtype=torch.int32, device="cuda")
    stride = 0 if P0 is None else int(np.prod(P0.shape[1:]))
    args = (C, p, stats, K.vec(dtau, 1), kind, K.vec(dP0, stride), K.vec(dlam, 1), K.vec(dmu0, p))
    K.nn_dense_draw(*args, beta, K.rng(seed=1, site=3), debug_z=dz, status=status)
    K.nn_dense_draw(*args, mean, K.rng(seed=1, site=3), solve_only=True, ridge_rel=0.0, status=status)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0
    for c in range(0, C, max(1, C // 7)):
        G, g, _, _ = conjugate.regression_suffstats(X[c], y[c])
        ref = conjugate.normal_normal_dense(G, g, tau[c], 1.0 if P0 is None else P0[c], lam[c], mu0[c], z[c])
        scale = np.abs(ref["x"]).max()
        np.testing.assert_allclose(beta[c].cpu().numpy(), ref["x"].ravel(), rtol=1e-9, atol=1e-9 * scale)
        np.testing.assert_allclose(mean[c].cpu().numpy(), ref["mu"].ravel(), rtol=1e-9, atol=1e-9 * scale)
