"""The reference's examples/3_linear_regression.ipynb with the import root changed to openmcmc_b200 (same model, same
samplers, same MCMC call); prints the posterior summaries the notebook plots.  Needs a B200 (no CPU path).

    python examples/3_linear_regression.py [n_chains]
"""
import os
import sys

import numpy as np
from scipy import sparse
from scipy.stats import norm

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmcmc_b200.distribution.distribution import Gamma
from openmcmc_b200.distribution.location_scale import Normal
from openmcmc_b200.mcmc import MCMC
from openmcmc_b200.model import Model
from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1
np.random.seed(0)
N = 100
true_beta = np.array([2, 0.5])
x = np.sort(np.random.rand(N))
X = np.stack([np.ones(N), x], 1)
true_tau = 100.0
y = X @ true_beta + norm.rvs(loc=0, scale=np.sqrt(1 / true_tau), size=N)

mean_form = LinearCombination(form={"beta": "X"})
tau_predictor = ScaledMatrix(matrix="P_tau", scalar="tau")
lambda_predictor = ScaledMatrix(matrix="P_lambda", scalar="lambda")
mdl = Model([Normal("y", mean=mean_form, precision=tau_predictor),
             Normal("beta", mean="mu", precision=lambda_predictor),
             Gamma("tau", shape="a_tau", rate="b_tau"),
             Gamma("lambda", shape="a_lambda", rate="b_lambda")], response={"y": "mean"})
sampler = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
initial_state = {"y": y, "X": X, "beta": [0, 0],
                 "P_tau": sparse.csc_matrix(np.eye(N)), "tau": 1,
                 "P_lambda": sparse.csc_matrix(np.eye(2)), "mu": [0, 0], "lambda": 0.01,
                 "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3, "b_lambda": 1e-3}

M = MCMC(initial_state, sampler, model=mdl, n_burn=1000, n_iter=1000, n_chains=n_chains)
M.run_mcmc()
beta = M.store["beta"].reshape(-1, 2, 1000)      # (chains, 2, n_iter); a single chain has the reference's (2, n_iter)
tau = M.store["tau"].reshape(-1, 1000)
print("posterior mean of beta:", beta.mean(axis=(0, 2)), " truth:", true_beta)
print("posterior mean of tau :", tau.mean(), " truth:", true_tau)
fit = M.store["y"].reshape(-1, N, 1000)
q = np.quantile(fit[0], [0.025, 0.5, 0.975], axis=1)
print("fitted line, max |median - truth|:", np.abs(q[1] - X @ true_beta).max())
assert np.all(np.abs(beta.mean(axis=(0, 2)) - true_beta) < 0.1) and 50 < tau.mean() < 200
