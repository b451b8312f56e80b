"""The reference's examples/4_GMRF_smoother.ipynb with the import root changed to openmcmc_b200: a first-order random-walk
(GMRF) prior on a time series, NormalNormal for the smooth b, NormalGamma for the smoothing precision lambda and the
noise precision tau.  The notebook runs n_time = 100; the size is a command-line argument here (the tridiagonal path
runs 10^6 points per chain).  Needs a B200 (no CPU path).

    python examples/4_GMRF_smoother.py [n_time] [n_chains]
"""
import os
import sys

import numpy as np
import pandas as pd
from scipy import sparse
from scipy.stats import norm

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmcmc_b200 import gmrf
from openmcmc_b200.distribution.distribution import Gamma
from openmcmc_b200.distribution.location_scale import Normal
from openmcmc_b200.mcmc import MCMC
from openmcmc_b200.model import Model
from openmcmc_b200.parameter import ScaledMatrix
from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

n_time = int(sys.argv[1]) if len(sys.argv) > 1 else 100
n_chains = int(sys.argv[2]) if len(sys.argv) > 2 else 1
np.random.seed(1)
# generate GMRF temporal precision matrix
TIME = pd.date_range(start="2022-04-01T01:00:00", end="2022-04-01T01:01:00", periods=n_time)
P_lambda = gmrf.precision_temporal(time=TIME)
P_lambda[0, 0] = P_lambda[0, 0] + 0.001   # make full rank
s = np.asarray((TIME - TIME.min()).total_seconds())
truth = np.sin(s / 20) + 2 * np.cos(s / 12) + 2
y = truth + norm.rvs(loc=0, scale=1.0, size=n_time)

tau_predictor = ScaledMatrix(matrix="P_tau", scalar="tau")
lambda_predictor = ScaledMatrix(matrix="P_lambda", scalar="lambda")
mdl = Model([Normal("y", mean="b", precision=tau_predictor),
             Normal("b", mean="mu", precision=lambda_predictor),
             Gamma("lambda", shape="a_lam", rate="b_lam"),
             Gamma("tau", shape="a_tau", rate="b_tau")])
initial_state = {"y": y, "b": y, "mu": np.zeros(n_time), "lambda": 100, "P_lambda": P_lambda, "a_lam": 10, "b_lam": 1,
                 "tau": 1, "P_tau": sparse.identity(n_time, format="csc"), "a_tau": 1, "b_tau": 1}
samplers = [NormalNormal("b", mdl), NormalGamma("lambda", mdl), NormalGamma("tau", mdl)]

M = MCMC(initial_state, samplers, model=mdl, n_burn=200, n_iter=500 if n_time <= 10_000 else 20, n_chains=n_chains)
M.run_mcmc()
b = M.store["b"].reshape(n_chains, n_time, -1)
smooth = b.mean(axis=(0, 2))
print("rms(y - truth)      :", np.sqrt(np.mean((y - truth) ** 2)))
print("rms(smooth - truth) :", np.sqrt(np.mean((smooth - truth) ** 2)))
print("posterior mean tau  :", M.store["tau"].mean(), " lambda:", M.store["lambda"].mean())
assert np.sqrt(np.mean((smooth - truth) ** 2)) < 0.6 * np.sqrt(np.mean((y - truth) ** 2))
