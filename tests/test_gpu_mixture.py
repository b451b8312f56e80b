"""GPU parity of the mixture-model conjugate updates: MixtureAllocation, the NormalGamma K-loop over a
MixtureParameterMatrix precision, NormalNormal with a mixture prior, Categorical / mixture-Normal log-densities.
Kernels against the numpy oracle (oracle/conjugate.py) and the full MCMC driver against chains recorded from the live
reference (tests/golden/mixture_*.npz).  Allocations (integers) must match exactly; the rest to 1e-9 / 1e-10."""

import glob
import os

import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
NAMES = sorted(os.path.basename(f)[:-4] for f in glob.glob(os.path.join(GOLD, "mixture_*.npz")))


def _dev(a):
    import torch

    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()


@pytest.mark.parametrize("n,K,C,prob_rows", [(1, 1, 2, 1), (37, 3, 4, 1), (1000, 7, 3, 1000), (5000, 64, 2, 1)])
def test_mixture_kernels_match_oracle(n, K, C, prob_rows):
    import torch

    from openmcmc_b200 import kernels as Kn
    from oracle import conjugate

    Kn.init_device()
    rng = np.random.default_rng(n + K)
    mu = np.sort(rng.standard_normal((C, K)) * 3, axis=1)
    tau = rng.random((C, K)) * 3 + 0.3
    z0 = rng.integers(0, K, size=(C, n)).astype(float)
    x = np.take_along_axis(mu, z0.astype(int), axis=1) + rng.standard_normal((C, n)) / np.sqrt(
        np.take_along_axis(tau, z0.astype(int), axis=1))
    prob = rng.random((prob_rows, K)) + 0.1
    prob /= prob.sum(axis=1, keepdims=True)
    u = rng.random((C, n))
    d_x, d_mu, d_tau, d_prob, d_u, d_z = _dev(x), _dev(mu), _dev(tau), _dev(prob), _dev(u), _dev(z0)
    # ---- statistics / record / gathers / log-density of the CURRENT allocation
    stats = torch.zeros(C, K, 4, dtype=torch.float64, device="cuda")
    rec = torch.zeros(C, K * K + K + 2, dtype=torch.float64, device="cuda")
    gmu, gtau = torch.zeros(C, n, dtype=torch.float64, device="cuda"), torch.zeros(C, n, dtype=torch.float64, device="cuda")
    lp = torch.zeros(C, dtype=torch.float64, device="cuda")
    Kn.mixture_stats(C, n, K, Kn.vec(d_x, n), Kn.vec(d_mu, K), Kn.vec(d_tau, K), d_z, stats, record=rec, gather_mu=gmu,
                     gather_tau=gtau, logp=lp)
    lc = torch.zeros(C, dtype=torch.float64, device="cuda")
    Kn.logp_categorical(C, n, K, d_z, Kn.vec(d_prob), prob_rows, lc, False)
    # ---- new allocation with injected uniforms
    z_new = torch.empty(C, n, dtype=torch.float64, device="cuda")
    Kn.mixture_allocation(C, n, K, Kn.vec(d_x, n), Kn.vec(d_mu, K), Kn.vec(d_tau, K), Kn.vec(d_prob), prob_rows, z_new,
                          Kn.rng(seed=1), debug_u=d_u)
    torch.cuda.synchronize()
    for c in range(C):
        so = conjugate.mixture_stats(x[c], mu[c], z0[c], K)
        np.testing.assert_allclose(stats[c, :, :3].cpu().numpy(), so, rtol=1e-11, atol=1e-12)
        r = rec[c].cpu().numpy()
        np.testing.assert_allclose(r[: K * K].reshape(K, K), np.diag(tau[c] * so[:, 0]), rtol=1e-12)
        np.testing.assert_allclose(r[K * K: K * K + K], tau[c] * so[:, 1], rtol=1e-11, atol=1e-12)
        np.testing.assert_allclose(gmu[c].cpu().numpy(), mu[c][z0[c].astype(int)], rtol=0)
        np.testing.assert_allclose(gtau[c].cpu().numpy(), tau[c][z0[c].astype(int)], rtol=0)
        np.testing.assert_allclose(lp[c].item(), conjugate.mixture_normal_log_p(x[c], mu[c], tau[c], z0[c]), rtol=1e-10)
        np.testing.assert_allclose(lc[c].item(), conjugate.categorical_log_p(z0[c], prob), rtol=1e-11)
        zo = conjugate.mixture_allocation(x[c], mu[c], tau[c], prob, u[c])
        assert np.mean(z_new[c].cpu().numpy() != zo) <= (2.0 / n if n > 500 else 0)   # a tie at 1e-16 may flip one


def _build(g):
    from openmcmc_b200.distribution.distribution import Categorical, Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import Identity, LinearCombination, MixtureParameterMatrix, MixtureParameterVector
    from openmcmc_b200.sampler.sampler import MixtureAllocation, NormalGamma, NormalNormal

    mdl = Model([Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=Identity("W")),
                 Normal("beta", mean=MixtureParameterVector(param="mu", allocation="z"),
                        precision=MixtureParameterMatrix(param="tau", allocation="z")),
                 Gamma("tau", shape="a_tau", rate="b_tau"),
                 Categorical("z", prob="prob")])
    samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), MixtureAllocation("z", mdl, response_param="beta")]
    state = {"y": g["y"], "X": g["X"], "W": sparse.diags(g["w"], format="csc"), "beta": g["beta0"].copy(),
             "mu": g["mu0"].copy(), "tau": g["tau0"].copy(), "z": g["z0"].copy(), "prob": g["prob"],
             "a_tau": g["a_tau"], "b_tau": g["b_tau"]}
    return mdl, samplers, state


@pytest.mark.parametrize("name", NAMES)
def test_mcmc_replays_reference_mixture_chain(name):
    from openmcmc_b200.mcmc import MCMC

    g = dict(np.load(os.path.join(GOLD, name + ".npz")))
    mdl, samplers, state = _build(g)
    n_iter = g["store_beta"].shape[1]
    dd = {"beta": {"z": g["z_beta"]}, "tau": {"g": g["g"]}, "z": {"u": g["u"]}}
    M = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=n_iter, debug_draws=dd)
    M.run_mcmc()
    np.testing.assert_allclose(M.store["beta"], g["store_beta"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(M.store["tau"], g["store_tau"], rtol=1e-9)
    np.testing.assert_array_equal(M.store["z"], g["store_z"])
    np.testing.assert_allclose(M.store["log_post"], g["store_log_post"], rtol=1e-10)


def test_mixture_free_running_recovers_clusters_and_shards():
    """Free-running (Philox) chains on well-separated clusters: the allocation of every coefficient settles on its true
    component in most stored draws; a shard of the chains run alone (chain_offset) reproduces its draws bit for bit."""
    from openmcmc_b200.mcmc import MCMC

    rng = np.random.default_rng(5)
    n, p, K = 400, 30, 3
    X = rng.standard_normal((n, p))
    z_true = rng.integers(0, K, size=p)
    mu = np.array([[-4.0], [0.0], [4.0]])
    beta = mu[z_true] + 0.2 * rng.standard_normal((p, 1))
    y = X @ beta + 0.3 * rng.standard_normal((n, 1))
    g = {"y": y, "X": X, "w": np.ones(n), "beta0": np.zeros((p, 1)), "mu0": mu, "tau0": np.ones((K, 1)),
         "z0": rng.integers(0, K, size=(p, 1)).astype(float), "prob": np.full((1, K), 1.0 / K),
         "a_tau": 2.0 * np.ones((K, 1)), "b_tau": 0.5 * np.ones((K, 1))}
    mdl, samplers, state = _build(g)
    C = 16
    M = MCMC(state, samplers, model=mdl, n_burn=50, n_iter=50, n_chains=C, seed=9)
    M.run_mcmc()
    zs = M.store["z"]                                     # (C, p, n_iter)
    assert np.mean(zs == z_true.reshape(1, p, 1)) > 0.97
    assert np.all(M.store["tau"] > 0) and np.std(M.store["tau"][:, 0, -1]) > 0
    M2 = MCMC(state, samplers, model=mdl, n_burn=50, n_iter=50, n_chains=8, seed=9, chain_offset=8)
    M2.run_mcmc()
    np.testing.assert_array_equal(M2.store["z"], zs[8:])
    np.testing.assert_array_equal(M2.store["beta"], M.store["beta"][8:])
