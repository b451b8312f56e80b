"""Run the sweep ops of one golden regression case eagerly with a sync after each op (find a failing launch)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
from test_gpu_mcmc_regression import build, GOLD
from openmcmc_b200.mcmc import MCMC
from openmcmc_b200 import kernels as K

name = sys.argv[1]
g = dict(np.load(os.path.join(GOLD, name + ".npz")))
mdl, samplers, state = build(g)
dd = {"beta": {"z": g["z"]}, "tau": {"g": g["g_tau"]}, "lambda": {"g": g["g_lambda"]}}
M = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=g["store_beta"].shape[1], debug_draws=dd)
orig = K.Graph.capture
def eager(fn):
    return None
ops_seen = {}
class FakeGraph:
    @staticmethod
    def capture(fn):
        return None
K.Graph.capture = staticmethod(lambda fn: None)
M.prepare()
for phase in ("prologue", "sweep", "store"):
    for label, fn in M._ops[phase]:
        try:
            with torch.cuda.stream(M.stream):
                fn()
            torch.cuda.synchronize()
            print(phase, label, "ok")
        except Exception as e:
            print(phase, label, "FAILED", e)
            raise
