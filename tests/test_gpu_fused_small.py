"""GPU: omc_fused_small (several per-chain small operations in one launch) gives the chains of the one-kernel-per-
operation plan: bit for bit on the Gibbs regression (sums of one term, the quadratic form in the stand-alone kernel's
warp order), to rounding on an MH model whose log-density sums run as a warp tree instead of a block reduction.  ref: mcmc.py:98-111, sampler.py:252-288."""

import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu


def _run(flag, build, **kw):
    from openmcmc_b200 import engine
    from openmcmc_b200.mcmc import MCMC

    old = engine.FUSE_SMALL
    engine.FUSE_SMALL = flag
    try:
        mdl, samplers, state = build()
        M = MCMC(state, samplers, model=mdl, **kw)
        M.run_mcmc()
    finally:
        engine.FUSE_SMALL = old
    return M


def test_fused_regression_sweep_is_bit_identical():
    from test_gpu_stream_store import _regression

    kw = dict(n_burn=2, n_iter=9, n_thin=2, n_chains=7, seed=4)
    a = _run(True, lambda: _regression(7, 300, 11, 21), **kw)
    b = _run(False, lambda: _regression(7, 300, 11, 21), **kw)
    la = [label for label, _ in a._ops["sweep"]] + [label for label, _ in a._ops["store"]]
    lb = [label for label, _ in b._ops["sweep"]] + [label for label, _ in b._ops["store"]]
    assert sum(x.startswith("fused[") for x in la) == 2 and not any(x.startswith("fused[") for x in lb), (la, lb)
    assert a._sweep_graph.num_kernels() + a._store_graph.num_kernels() < b._sweep_graph.num_kernels() + b._store_graph.num_kernels() - 6
    for key in b.store:      # (the fused quadratic form runs in the stand-alone kernel's order: the same bits)
        assert np.array_equal(a.store[key], b.store[key]), key


def test_fused_mh_store_matches_unfused():
    from openmcmc_b200.distribution.distribution import Gamma, Poisson
    from openmcmc_b200.model import Model
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA

    p = 16
    rng = np.random.default_rng(0)
    y = rng.poisson(rng.gamma(5.0, 1.0, size=(p, 1))).astype(float)

    def build():
        mdl = Model([Poisson("y", rate="lam"), Gamma("lam", shape="a", rate="b")])
        state = {"y": y, "lam": y + 1.0, "a": np.array([[2.0]]), "b": np.array([[0.5]])}
        return mdl, [ManifoldMALA("lam", mdl, step=np.array([[0.4]]))], state

    kw = dict(n_burn=5, n_iter=30, n_chains=50, seed=9)
    a, b = _run(True, build, **kw), _run(False, build, **kw)
    assert np.array_equal(a.store["lam"], b.store["lam"])
    np.testing.assert_allclose(a.store["log_post"], b.store["log_post"], rtol=1e-13)


def test_stored_sweep_graph_is_bit_identical_and_three_launches():
    """n_thin = 1: sweep and store epilogue replay as ONE graph whose small ops fuse across the boundary (draw, fused
    small ops, both counters) -- the same chains, stores and counters as the two-graph schedule."""
    from openmcmc_b200 import mcmc
    from test_gpu_stream_store import _regression

    def run(flag):
        old = mcmc.FUSE_STORED_SWEEP
        mcmc.FUSE_STORED_SWEEP = flag
        try:
            mdl, samplers, state = _regression(9, 250, 10, 33)
            mdl.response = None                                # no fitted values: nothing unfusable in the store epilogue
            M = mcmc.MCMC(state, samplers, model=mdl, n_burn=3, n_iter=12, n_thin=1, n_chains=9, seed=5)
            M.run_mcmc()
        finally:
            mcmc.FUSE_STORED_SWEEP = old
        return M

    a, b = run(True), run(False)
    assert a._stored_sweep_graph is not None and b._stored_sweep_graph is None
    assert a._stored_sweep_graph.num_kernels() == 3, [label for label, _ in a._ops["stored_sweep"]]
    assert a.launches_of(3, 12, 1) == 3 * 3 + 12 * 3 and b.launches_of(3, 12, 1) == 15 * 3 + 12 * b._store_graph.num_kernels()
    for key in b.store:
        assert np.array_equal(a.store[key], b.store[key]), key
    for key in ("beta", "tau", "lambda"):
        assert np.array_equal(a.state[key], b.state[key]), key
    assert int(a.plan.sweep_counter.item()) == int(b.plan.sweep_counter.item()) == 15
    assert int(a.plan.iter_counter.item()) == int(b.plan.iter_counter.item()) == 12


def test_run_split_into_calls_continues_the_store():
    """run_device(restart_store=False): a run replayed one stored sweep per call (what bench.py does around its per-sweep
    L2 flushes) fills the same store as one call."""
    from openmcmc_b200 import mcmc
    from test_gpu_stream_store import _regression

    def make():
        mdl, samplers, state = _regression(5, 120, 6, 41)
        mdl.response = None
        return mcmc.MCMC(state, samplers, model=mdl, n_burn=0, n_iter=7, n_thin=1, n_chains=5, seed=3)

    a = make()
    a.run_mcmc()
    b = make()
    b.prepare()
    for k in range(7):
        b.run_device(n_burn=0, n_iter=1, n_thin=1, restart_store=(k == 0))
    b.collect()
    for key in a.store:
        assert np.array_equal(a.store[key], b.store[key]), key
