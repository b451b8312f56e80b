"""GPU: the streamed sample store (device ring -> pinned staging -> host arrays, stream_store.py) gives bit-identical
stores to the resident one (ref: mcmc.py:105-111, sampler.py:89-118), a second run restarts the store, and the
diagnostics work on runs in chain blocks and on streamed runs."""

import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu


def _regression(C, n, p, seed):
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    rng = np.random.default_rng(seed)
    X = rng.standard_normal((C, n, p))
    y = X @ rng.standard_normal((C, p, 1)) + 0.1 * rng.standard_normal((C, n, 1))
    mdl = Model([Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
                 Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
                 Gamma("tau", shape="a", rate="b"), Gamma("lambda", shape="a", rate="b")], response={"y": "mean"})
    samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
    state = {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.identity(n, format="csc"), "tau": 1.0,
             "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": 0.01, "a": 1e-3, "b": 1e-3}
    return mdl, samplers, state


def _gmrf(C, n, seed):
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    rng = np.random.default_rng(seed)
    s = np.arange(n) * (60.0 / 99.0)
    dr = 1.0 / np.diff(s)
    pd = np.append(np.append(dr[0], dr[:-1] + dr[1:]), dr[-1])
    pd[0] += 1e-3
    P = sparse.diags([-dr, pd, -dr], offsets=[-1, 0, 1], format="csc")
    y = (np.sin(s / 20) + 2)[None, :, None] + rng.standard_normal((C, n, 1))
    mdl = Model([Normal("y", mean=LinearCombination(form={"b": "I"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
                 Normal("b", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
                 Gamma("lambda", shape="a_lam", rate="b_lam"), Gamma("tau", shape="a_tau", rate="b_tau")])
    samplers = [NormalNormal("b", mdl), NormalGamma("lambda", mdl), NormalGamma("tau", mdl)]
    state = {"y": y, "b": y.copy(), "mu": np.zeros(n), "lambda": 100, "P_lambda": P, "a_lam": 10, "b_lam": 1, "tau": 1,
             "P_tau": sparse.identity(n, format="csc"), "I": sparse.identity(n, format="csc"), "a_tau": 1, "b_tau": 1}
    return mdl, samplers, state


@pytest.mark.parametrize("kind", ["regression", "gmrf"])
def test_streamed_store_is_bit_identical(kind, monkeypatch):
    from openmcmc_b200 import mcmc as M

    if kind == "regression":
        C, build = 6, (lambda: _regression(6, 400, 9, 3))
        kw = dict(n_burn=3, n_iter=11, n_thin=2)
    else:
        C, build = 3, (lambda: _gmrf(3, 200_000, 4))          # 4.8 MB per stored iteration and entry: several chunks
        kw = dict(n_burn=1, n_iter=7, n_thin=1)
    runs = {}
    for mode in (False, True):
        mdl, samplers, state = build()
        if mode:                                               # two slabs only: the ring wraps several times
            monkeypatch.setattr(M, "RING_BYTES", 1)
        run = M.MCMC(state, samplers, model=mdl, n_chains=C, seed=5, stream_store=mode, **kw)
        run.run_mcmc()
        assert run._streamed == mode
        if mode:
            assert run._ring == 2 and run.timing["streamed_d2h_bytes"] == run._slab_bytes * kw["n_iter"]
        runs[mode] = run
    for key in runs[False].store:
        assert np.array_equal(runs[True].store[key], runs[False].store[key]), key
    for name in ("tau", "lambda"):
        assert np.array_equal(runs[True].state[name], runs[False].state[name])


def test_second_run_restarts_the_store_and_store_exists_before_the_run():
    from openmcmc_b200.mcmc import MCMC

    mdl, samplers, state = _regression(1, 120, 4, 8)
    state = dict(state, X=state["X"][0], y=state["y"][0])
    run = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=5, seed=1)
    # ref: mcmc.py:81-85 -- NaN arrays of the reference's shapes exist before run_mcmc()
    assert run.store["beta"].shape == (4, 5) and np.isnan(run.store["beta"]).all()
    assert run.store["log_post"].shape == (5, 1) and run.store["y"].shape == (120, 5)
    run.run_mcmc()
    first = {k: v.copy() for k, v in run.store.items()}
    run.run_device()
    run.stream.synchronize()
    run.collect()
    assert np.isfinite(run.store["beta"]).all()
    assert not np.array_equal(run.store["beta"], first["beta"])     # the chain went on; the store holds the NEW draws


def test_summarize_on_blocked_and_streamed_runs():
    from openmcmc_b200 import diagnostics as G
    from openmcmc_b200.mcmc import MCMC

    out = {}
    for mode, kw in (("plain", {}), ("blocked", dict(upload_blocks=2)), ("streamed", dict(stream_store=True))):
        mdl, samplers, state = _regression(8, 300, 5, 11)
        run = MCMC(state, samplers, model=mdl, n_burn=5, n_iter=40, n_chains=8, seed=2, **kw)
        run.run_mcmc()
        out[mode] = G.summarize(run)
    for mode in ("blocked", "streamed"):
        for prm in ("beta", "tau", "lambda"):
            for key in ("ess", "rhat", "mean"):
                np.testing.assert_allclose(out[mode][prm][key].cpu().numpy(), out["plain"][prm][key].cpu().numpy(),
                                           rtol=1e-12, atol=0)
