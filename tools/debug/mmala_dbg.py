import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from scipy import sparse
from oracle import mh
from openmcmc_b200.mcmc import MCMC
from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA
from openmcmc_b200.distribution.location_scale import Normal
from openmcmc_b200.model import Model
from openmcmc_b200.parameter import ScaledMatrix
np.set_printoptions(precision=6, linewidth=200)
g = dict(np.load("tests/golden/mmala_normal_p7.npz"))
mdl = Model([Normal("theta", mean="mu", precision=ScaledMatrix(matrix="P", scalar="lam")),
             Normal("yobs", mean="theta", precision=ScaledMatrix(matrix="W", scalar="tau"))])
terms = [mh.Term("normal_response", p1=g["mu"], Q=g["lam"] * g["P"]),
         mh.Term("normal_response", p1=g["yobs"], Q=g["tau"] * np.diag(g["w"]))]
for n_iter in (1, 2):
    state = {"theta": g["theta0"].copy(), "mu": g["mu"], "P": g["P"], "lam": float(g["lam"]), "yobs": g["yobs"],
             "W": sparse.diags(g["w"], format="csc"), "tau": float(g["tau"])}
    smp = ManifoldMALA("theta", mdl, step=np.array([[float(g["step"])]]))
    M = MCMC(state, [smp], model=mdl, n_burn=0, n_iter=n_iter, debug_draws={"theta": {"z": g["z"], "u": g["u"]}}, probes=True)
    M.run_mcmc()
    th = g["theta0"]
    for it in range(n_iter):
        th, info = mh.mmala_step(terms, th, float(g["step"]), g["z"][it], g["u"][it])
    pr = M.plan.probes["theta"]
    print("n_iter", n_iter)
    print(" store", M.store["theta"][:, -1])
    print(" oracle", th.ravel())
    print(" mu dev", pr["mu"].cpu().numpy()[0]); print(" mu ora", info["mu"].ravel())
    print(" prop dev", pr["prop"].cpu().numpy()[0]); print(" prop ora", info["prop"].ravel())
    print(" L diff", np.abs(pr["L"].cpu().numpy()[0] - info["L"]).max())
    print(" scal dev", pr["scalars"].cpu().numpy()[0]); print(" scal ora", info["logp_cur"], info["logp_prop"], info["lq_fwd"], info["lq_rev"], info["log_accept"])
    print(" sweep counter", M.plan.sweep_counter.item())
