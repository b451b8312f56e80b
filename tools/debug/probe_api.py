import sys, traceback
sys.path.insert(0, "/root/repo")
import numpy as np
from scipy import sparse
from openmcmc_b200.distribution.distribution import Gamma, Poisson, Uniform, Categorical
from openmcmc_b200.distribution.location_scale import Normal, LogNormal
from openmcmc_b200.model import Model
from openmcmc_b200.mcmc import MCMC
from openmcmc_b200.parameter import LinearCombination, ScaledMatrix, Identity
from openmcmc_b200.sampler.sampler import NormalNormal, NormalGamma
from openmcmc_b200.sampler.metropolis_hastings import RandomWalk, RandomWalkLoop, ManifoldMALA

def tryit(name, fn):
    try:
        r = fn()
        print("OK  ", name, "->", (np.shape(r) if hasattr(r, "shape") else r) if r is not None else "")
    except Exception as e:
        print("FAIL", name, type(e).__name__, str(e)[:150])

rng = np.random.default_rng(0)
n, p = 50, 3
X = rng.standard_normal((n, p)); beta = rng.standard_normal((p, 1)); y = X @ beta + 0.1 * rng.standard_normal((n, 1))
mdl = Model([Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
             Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
             Gamma("tau", shape="a_tau", rate="b_tau"), Gamma("lambda", shape="a_lambda", rate="b_lambda")])
state = {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.identity(n, format="csc"), "tau": np.array([[1.0]]),
         "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": np.array([[0.01]]),
         "a_tau": np.array([[1e-3]]), "b_tau": np.array([[1e-3]]), "a_lambda": np.array([[1e-3]]), "b_lambda": np.array([[1e-3]])}
tryit("NormalNormal.sample(state)", lambda: NormalNormal("beta", mdl).sample(dict(state))["beta"])
tryit("NormalGamma.sample(state)", lambda: NormalGamma("tau", mdl).sample(dict(state))["tau"])
tryit("model.log_p", lambda: mdl.log_p(state))
tryit("Normal y log_p", lambda: mdl["y"].log_p(state))
tryit("Gamma log_p", lambda: mdl["tau"].log_p(state))
tryit("grad beta", lambda: mdl.grad_log_p(state, "beta")[1])
tryit("Normal.rvs n=4", lambda: mdl["beta"].rvs(state, n=4))
tryit("Gamma.rvs", lambda: mdl["tau"].rvs(state, n=3))
tryit("predictor", lambda: mdl["y"].mean.predictor(state))
st2 = {k: v for k, v in state.items() if k != "beta"}
def run_missing():
    M = MCMC(st2, [NormalNormal("beta", mdl), NormalGamma("tau", mdl)], model=mdl, n_burn=5, n_iter=10, n_thin=2)
    M.run_mcmc(); return M.store["beta"]
tryit("MCMC missing initial beta + thin", run_missing)
pm = Model([Poisson("k", rate="lam"), Gamma("lam", shape="a", rate="b")])
ps = {"k": rng.poisson(5.0, (6, 1)).astype(float), "lam": 5 * np.ones((6, 1)), "a": np.array([[2.0]]), "b": np.array([[0.5]])}
tryit("Poisson log_p", lambda: pm["k"].log_p(ps))
tryit("Poisson rvs", lambda: pm["k"].rvs(ps, n=2))
tryit("fd grad", lambda: pm.grad_log_p(ps, "lam")[0])
tryit("RandomWalk.sample", lambda: RandomWalk("lam", pm, step=np.array([[0.3]])).sample(dict(ps))["lam"])
tryit("ManifoldMALA.sample", lambda: ManifoldMALA("lam", pm, step=np.array([[0.5]])).sample(dict(ps))["lam"])
tryit("RandomWalkLoop no limits", lambda: RandomWalkLoop("lam", pm, step=np.array([[0.3]])).sample(dict(ps)))
um = Uniform("x", domain_response_lower=np.array([[0.0]]), domain_response_upper=np.array([[2.0]]))
tryit("Uniform log_p/rvs", lambda: (um.log_p({"x": np.array([[1.0]])}), um.rvs({"x": np.array([[1.0]])}, n=3).shape))
cm = Categorical("z", prob="pr")
tryit("Categorical log_p", lambda: cm.log_p({"z": np.array([[0], [1], [1]]), "pr": np.array([[0.3, 0.7]])}))
tryit("bad mean type", lambda: Normal("y", mean=3, precision="tau"))
tryit("not PD single chain", lambda: NormalNormal("beta", mdl).sample(dict(state, **{"lambda": np.array([[-1e9]])})))
