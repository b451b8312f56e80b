"""Generate golden vectors from the LIVE, unmodified reference (run in the build container only).

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference (/root/reference/src/openmcmc) is imported as-is; its random streams are replaced by recorded ones the
same way its own tests do it (monkeypatching scipy.stats.*.rvs, tests/test_sampler.py:211-215), so the very same
numbers can be injected into the oracle and into the CUDA kernels (debug_draws).  Nothing here runs on the GPU box:
the .npz files are committed.
"""

import os
import sys

import numpy as np
from scipy import sparse, stats

REF = "/root/reference/src"
sys.path.insert(0, REF)
OUT = os.path.dirname(os.path.abspath(__file__))

from openmcmc import gmrf  # noqa: E402
from openmcmc.distribution.distribution import Gamma, Poisson, Uniform  # noqa: E402
from openmcmc.distribution.location_scale import Normal  # noqa: E402
from openmcmc.mcmc import MCMC  # noqa: E402
from openmcmc.model import Model  # noqa: E402
from openmcmc.parameter import Identity, LinearCombination, ScaledMatrix  # noqa: E402
from openmcmc.sampler.metropolis_hastings import ManifoldMALA, RandomWalk, RandomWalkLoop  # noqa: E402
from openmcmc.sampler.sampler import NormalGamma, NormalNormal  # noqa: E402


class Streams:
    """Replay/record wrapper for the scipy.stats rvs entry points the reference uses."""

    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.log = {"z": [], "g": [], "u": [], "tn_u": [], "randint": []}
        self._saved = {}

    def norm_rvs(self, size=1, scale=1, loc=0):
        z = self.rng.standard_normal(size)
        self.log["z"].append(np.array(z, dtype=float).ravel())
        return loc + scale * z

    def gamma_rvs(self, a, scale=1, size=None):
        a = np.asarray(a, dtype=float)
        shp = np.broadcast(a, np.asarray(scale)).shape if size is None else size
        g = self.rng.standard_gamma(np.broadcast_to(a, shp))
        self.log["g"].append(np.array(g, dtype=float).ravel())
        return g * scale

    def uniform_rvs(self, size=None, loc=0, scale=1):
        u = self.rng.random(size)
        self.log["u"].append(np.array(u, dtype=float).ravel())
        return loc + scale * u

    def truncnorm_rvs(self, a, b, loc=0, scale=1, size=1):
        u = self.rng.random(size)
        self.log["tn_u"].append(np.array(u, dtype=float).ravel())
        return stats.truncnorm.ppf(u, a, b, loc=loc, scale=scale)

    def __enter__(self):
        for dist, name, fn in ((stats.norm, "rvs", self.norm_rvs), (stats.gamma, "rvs", self.gamma_rvs),
                               (stats.uniform, "rvs", self.uniform_rvs), (stats.truncnorm, "rvs", self.truncnorm_rvs)):
            self._saved[(dist, name)] = getattr(dist, name)
            setattr(dist, name, fn)
        return self

    def __exit__(self, *exc):
        for (dist, name), fn in self._saved.items():
            setattr(dist, name, fn)

    def stack(self, key):
        return np.array(self.log[key]) if self.log[key] else np.zeros((0,))


# ----------------------------------------------------------------------------------------------- regression (C1 shape)
def regression_case(n, p, seed, n_iter, weighted=False, order=("beta", "tau", "lambda"), prior="eye"):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, p))
    X[:, 0] = 1.0
    beta_true = rng.standard_normal((p, 1))
    y = X @ beta_true + 0.1 * rng.standard_normal((n, 1))
    if weighted:
        w = rng.random(n) + 0.2  # strictly positive: a singular P_tau breaks the reference's own log_p
        P_tau = sparse.diags(w, format="csc")
    else:
        P_tau = sparse.csc_matrix(np.eye(n))
    if prior == "eye":
        P_lambda = sparse.csc_matrix(np.eye(p))
    elif prior == "diag":
        P_lambda = sparse.diags(rng.random(p) + 0.5, format="csc")
    else:
        A = rng.standard_normal((p, p))
        P_lambda = A @ A.T + p * np.eye(p)
    mu = rng.standard_normal((p, 1)) * 0.1
    mdl = Model(
        [Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
         Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
         Gamma("tau", shape="a_tau", rate="b_tau"),
         Gamma("lambda", shape="a_lambda", rate="b_lambda")],
        response={"y": "mean"},
    )
    smap = {"beta": NormalNormal("beta", mdl), "tau": NormalGamma("tau", mdl), "lambda": NormalGamma("lambda", mdl)}
    samplers = [smap[k] for k in order]
    state = {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": P_tau, "tau": 1.0, "P_lambda": P_lambda, "mu": mu,
             "lambda": 0.01, "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3, "b_lambda": 1e-3}
    with Streams(seed + 1) as s:
        import openmcmc.mcmc as m

        m.tqdm = lambda it: it
        M = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=n_iter)
        M.run_mcmc()
    z = s.stack("z")                      # [n_iter, p]
    g = s.stack("g")                      # [2*n_iter, 1] in sampler order (tau/lambda interleaved as in `order`)
    gorder = [k for k in order if k != "beta"]
    out = {
        "X": X, "y": y, "mu": mu, "w": (np.asarray(P_tau.diagonal()) if weighted else np.zeros(0)),
        "P_lambda": (P_lambda.toarray() if sparse.issparse(P_lambda) else P_lambda),
        "order": np.array(order), "z": z, "weighted": weighted, "prior": prior,
        "g_" + gorder[0]: g[0::2, 0], "g_" + gorder[1]: g[1::2, 0],
        "store_beta": M.store["beta"], "store_tau": M.store["tau"], "store_lambda": M.store["lambda"],
        "store_log_post": M.store["log_post"], "store_y": M.store["y"],
    }
    return out


def main():
    cases = {
        "regression_n50_p3": regression_case(50, 3, 0, 6),
        "regression_n200_p8_weighted_diag": regression_case(200, 8, 1, 4, weighted=True, prior="diag"),
        "regression_n300_p17_dense_reordered": regression_case(300, 17, 2, 4, order=("tau", "lambda", "beta"),
                                                               prior="dense"),
        "regression_n1000_p64": regression_case(1000, 64, 3, 3),
    }
    for name, d in cases.items():
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print("wrote", name, {k: np.shape(v) for k, v in d.items() if k.startswith("store")})


if __name__ == "__main__":
    main()
