"""The reference's examples/1_model_distributions.ipynb with the import root changed: a two-level Normal model, its
log-density and its gradient / Hessian with respect to the mean.  Needs a B200 (no CPU path)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from openmcmc_b200.distribution.location_scale import Normal
from openmcmc_b200.model import Model

my_dist = Normal("y", mean="h", precision="tau")
mdl = Model([Normal("y", mean="h", precision="tau"),
             Normal("h", mean="mu", precision="lambda")])
state = {}
state["y"] = np.array([150, 155, 190, 160, 173], ndmin=2)
state["h"] = np.array(180, ndmin=2)
state["tau"] = np.array(1 / 200, ndmin=2)
state["mu"] = np.array(160, ndmin=2)
state["lambda"] = np.array(1 / 100, ndmin=2)

print("log_p   :", mdl.log_p(state))
gradient, hessian = mdl.grad_log_p(state, param="h")
print("gradient:", gradient)
print("hessian :", hessian)
# closed forms: sum of the six Normal log-densities; d/dh = tau sum(y - h) - lambda (h - mu); Hessian = 5 tau + lambda
y, h = state["y"].ravel(), 180.0
lp = np.sum(-0.5 * np.log(2 * np.pi * 200) - (y - h) ** 2 / 400) - 0.5 * np.log(2 * np.pi * 100) - (h - 160) ** 2 / 200
assert np.isclose(float(np.ravel(mdl.log_p(state))[0]), lp, rtol=1e-10)
assert np.isclose(float(np.ravel(gradient)[0]), np.sum(y - h) / 200 - (h - 160) / 100, rtol=1e-10)
assert np.isclose(float(np.ravel(hessian)[0]), 5 / 200 + 1 / 100, rtol=1e-10)
