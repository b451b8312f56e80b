"""GPU: omc_chain_stats / omc_rhat_combine against the numpy restatement (oracle/diagnostics.py) element for element,
and the diagnostics summary of a real MCMC run.  No reference counterpart exists (parity unpinned)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ar1(rng, N, C, size, rho):
    x = np.zeros((N, C, size))
    e = rng.standard_normal((N, C, size))
    x[0] = e[0]
    for t in range(1, N):
        x[t] = rho * x[t - 1] + np.sqrt(1 - rho ** 2) * e[t]
    return x + rng.standard_normal((1, C, size)) * 0.3 + 5.0


@pytest.mark.parametrize("N,C,size,stride,max_lag", [(5, 3, 2, 1, 127), (64, 37, 3, 2, 127), (500, 40, 5, 1, 127),
                                                     (1000, 8, 64, 7, 40), (3, 2, 1, 1, 10), (2, 33, 1, 1, 5)])
def test_chain_stats_and_rhat_match_oracle(N, C, size, stride, max_lag):
    import torch

    from openmcmc_b200 import diagnostics as G
    from openmcmc_b200 import kernels as K
    from oracle import diagnostics as D

    K.init_device()
    rng = np.random.default_rng(N + C)
    x = _ar1(rng, N, C, size, 0.7)
    d = torch.as_tensor(x).cuda()
    rec = G.chain_stats(d, elem_stride=stride, max_lag=max_lag)
    ref = D.chain_stats(x, elem_stride=stride, max_lag=max_lag)
    torch.cuda.synchronize()
    np.testing.assert_allclose(rec.cpu().numpy(), ref, rtol=1e-9, atol=1e-12)
    if N >= 4:
        comb = G.rhat_combine(rec).cpu().numpy()
        np.testing.assert_allclose(comb, D.rhat_combine(ref), rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("N,C,size,stride,pooled", [(7, 3, 2, 1, False), (64, 5, 3, 2, True), (500, 12, 4, 1, False),
                                                    (1000, 6, 9, 4, True), (3000, 2, 1, 1, True), (1, 2, 1, 1, False)])
def test_rank_normalize_matches_oracle(N, C, size, stride, pooled):
    """omc_rank_normalize against scipy.stats.rankdata(method="average") + norm.ppf, ties and heavy tails included; the
    bulk-ESS / rank-normalised split-R-hat records then follow from the kernels above."""
    import torch

    from openmcmc_b200 import diagnostics as G
    from openmcmc_b200 import kernels as K
    from oracle import diagnostics as D

    K.init_device()
    rng = np.random.default_rng(N + C + 1)
    x = _ar1(rng, N, C, size, 0.6)
    x[:, 0, 0] = np.round(x[:, 0, 0], 1)                       # ties
    if C > 1:
        x[:, 1, 0] = rng.standard_cauchy(N)                   # no variance: the case rank normalisation exists for
    d = torch.as_tensor(x).cuda()
    z = G.rank_normalize(d, elem_stride=stride, pooled=pooled)
    zr = D.rank_normalize(x, elem_stride=stride, pooled=pooled)
    torch.cuda.synchronize()
    np.testing.assert_allclose(z.cpu().numpy(), zr, rtol=1e-12, atol=1e-13)
    if N >= 4:
        rec = G.chain_stats(z)
        np.testing.assert_allclose(rec.cpu().numpy(), D.chain_stats(zr), rtol=1e-9, atol=1e-12)


def test_summary_of_a_gibbs_run():
    """64 chains of the small Gibbs regression: R-hat near 1 for every coefficient, ESS close to the number of stored
    draws (conjugate Gibbs mixes in one sweep), and the per-chain minimum ESS feeds the ESS/s metric."""
    import torch
    from scipy import sparse

    from openmcmc_b200 import diagnostics as G
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.mcmc import MCMC
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    rng = np.random.default_rng(3)
    C, n, p = 64, 400, 5
    X = rng.standard_normal((n, p))
    y = X @ rng.standard_normal((p, 1)) + 0.3 * rng.standard_normal((n, 1))
    mdl = Model([
        Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
        Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
        Gamma("tau", shape="a_tau", rate="b_tau"), Gamma("lambda", shape="a_lambda", rate="b_lambda")])
    samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
    state = {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.identity(n, format="csc"), "tau": 1.0,
             "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": 0.01,
             "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3, "b_lambda": 1e-3}
    M = MCMC(state, samplers, model=mdl, n_burn=50, n_iter=400, n_chains=C, seed=11)
    M.run_mcmc()
    S = G.summarize(M)
    assert set(S) == {"beta", "tau", "lambda"}
    assert S["beta"]["records"].shape == (C, p, 8) and S["beta"]["n_chains_total"] == C
    rhat = S["beta"]["rhat"].cpu().numpy()
    assert np.all(np.abs(rhat - 1) < 0.02), rhat
    ess = S["beta"]["ess"].cpu().numpy()
    assert np.all(ess > 0.5 * C * 400), ess
    # rank-normalised forms (Vehtari et al. 2021): pooled ranks -> rank-normalised split-R-hat and bulk-ESS
    Sp = G.summarize(M, rank_normalized="pooled")
    assert np.all(np.abs(Sp["beta"]["rhat"].cpu().numpy() - 1) < 0.02)
    assert np.all(Sp["beta"]["ess"].cpu().numpy() > 0.5 * C * 400)
    Sc = G.summarize(M, rank_normalized="chain")
    bulk = G.min_ess_per_chain(Sc).cpu().numpy()
    assert bulk.shape == (C,) and np.all(bulk > 100)
    per_chain = G.min_ess_per_chain(S)
    assert per_chain.shape == (C,) and float(per_chain.min()) > 50
    # the device mean agrees with the host store
    np.testing.assert_allclose(S["tau"]["mean"].cpu().numpy()[0], M.store["tau"].mean(), rtol=1e-9)
