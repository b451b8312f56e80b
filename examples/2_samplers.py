"""The reference's examples/2_samplers.ipynb with the import root changed: five replicated observations of one height,
RandomWalk and then NormalNormal on the mean.  Needs a B200 (no CPU path)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmcmc_b200.distribution.location_scale import Normal
from openmcmc_b200.mcmc import MCMC
from openmcmc_b200.model import Model
from openmcmc_b200.sampler.metropolis_hastings import RandomWalk
from openmcmc_b200.sampler.sampler import NormalNormal

mdl = Model([Normal("y", mean="h", precision="tau"),
             Normal("h", mean="mu", precision="lambda")])
state = {}
state["y"] = np.array([150, 155, 190, 160, 173], ndmin=2)
state["h"] = np.array(200, ndmin=2)
state["tau"] = np.array(1 / 200, ndmin=2)
state["mu"] = np.array(160, ndmin=2)
state["lambda"] = np.array(1 / 100, ndmin=2)

# closed-form posterior of h: precision lambda + 5 tau, mean (lambda mu + tau sum y) / precision
prec = 1 / 100 + 5 / 200
post_mean = (160 / 100 + state["y"].sum() / 200) / prec

sampler = [RandomWalk("h", model=mdl, step=5.0)]
m = MCMC(state, sampler, model=mdl, n_burn=0, n_iter=1000)
m.run_mcmc()
print("RandomWalk   h: mean %.2f sd %.2f" % (m.store["h"][:, 200:].mean(), m.store["h"][:, 200:].std()))

sampler = [NormalNormal("h", model=mdl)]
m = MCMC(state, sampler, model=mdl, n_burn=0, n_iter=1000)
m.run_mcmc()
print("NormalNormal h: mean %.2f sd %.2f" % (m.store["h"].mean(), m.store["h"].std()))
print("closed form   : mean %.2f sd %.2f" % (post_mean, prec ** -0.5))
assert abs(m.store["h"].mean() - post_mean) < 1.0 and abs(m.store["h"].std() - prec ** -0.5) < 0.5
