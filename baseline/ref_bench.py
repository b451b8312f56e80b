"""CPU timing of the UNMODIFIED reference (openmcmc 1.0.7 installed into baseline/_ref by baseline/install_ref.sh).

Used by `bench.py --impl reference` and bench.py's `cpu_baseline` leg when baseline/_ref exists (`kind: "reference"`);
otherwise bench.py falls back to the numpy port in oracle/cpu_bench.py (`kind: "port"`).  Not product code: nothing
under openmcmc_b200/ imports it.

One worker process per host core, one BLAS thread each (the reference is single-threaded; SURVEY B.3: 8 BLAS threads
are slower than 1 at p = 64); each worker builds its chains' synthetic data (the shapes / priors of SURVEY §8d, as
bench.py builds them for the GPU arm), then times `MCMC(state, samplers, model, n_burn=0, n_iter=sweeps).run_mcmc()`
per chain — the reference's own public API and stock code path, with tqdm replaced by the identity and the acceptance
print silenced.  Model construction is outside the timed region.

    python baseline/ref_bench.py --workload c2 --chains-per-worker 1 --sweeps 5 --workers 8
"""

import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "openmcmc"))


def _import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import openmcmc.mcmc as ref_mcmc

    ref_mcmc.tqdm = lambda it, *a, **k: it        # progress bar replaced by the identity
    return ref_mcmc


def _regression(rng, n, p, response):
    import numpy as np
    from scipy import sparse

    from openmcmc.distribution.distribution import Gamma
    from openmcmc.distribution.location_scale import Normal
    from openmcmc.model import Model
    from openmcmc.parameter import LinearCombination, ScaledMatrix
    from openmcmc.sampler.sampler import NormalGamma, NormalNormal

    X = rng.standard_normal((n, p))
    X[:, 0] = 1.0
    y = X @ rng.standard_normal((p, 1)) + 0.1 * rng.standard_normal((n, 1))
    mdl = Model([
        Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
        Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
        Gamma("tau", shape="a_tau", rate="b_tau"),
        Gamma("lambda", shape="a_lambda", rate="b_lambda")], response={"y": "mean"} if response else None)
    samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
    state = {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.identity(n, format="csc"), "tau": 1.0,
             "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": 0.01,
             "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3, "b_lambda": 1e-3}
    return mdl, samplers, state


def _gmrf(rng, n):
    """Example 4 in its sparse-enabled form (mean = LinearCombination({'b': 'I'})), regular grid."""
    import numpy as np
    from scipy import sparse

    from openmcmc import gmrf
    from openmcmc.distribution.distribution import Gamma
    from openmcmc.distribution.location_scale import Normal
    from openmcmc.model import Model
    from openmcmc.parameter import LinearCombination, ScaledMatrix
    from openmcmc.sampler.sampler import NormalGamma, NormalNormal

    s = np.arange(n) * (60.0 / 99.0)
    P = gmrf.precision_irregular(s).tolil()
    P[0, 0] += 1e-3
    P = P.tocsc()
    y = (np.sin(s / 20) + 2 * np.cos(s / 12) + 2 + rng.standard_normal(n)).reshape(-1, 1)
    mdl = Model([Normal("y", mean=LinearCombination(form={"b": "I"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
                 Normal("b", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
                 Gamma("lambda", shape="a_lam", rate="b_lam"),
                 Gamma("tau", shape="a_tau", rate="b_tau")])
    samplers = [NormalNormal("b", mdl), NormalGamma("lambda", mdl), NormalGamma("tau", mdl)]
    state = {"y": y, "b": y.copy(), "mu": np.zeros((n, 1)), "lambda": 100.0, "P_lambda": P, "a_lam": 10.0, "b_lam": 1.0,
             "tau": 1.0, "P_tau": sparse.identity(n, format="csc"), "I": sparse.identity(n, format="csc"), "a_tau": 1.0,
             "b_tau": 1.0}
    return mdl, samplers, state


def _mh(rng, p, loop):
    import numpy as np

    from openmcmc.distribution.distribution import Gamma, Poisson
    from openmcmc.model import Model
    from openmcmc.sampler.metropolis_hastings import ManifoldMALA, RandomWalkLoop

    shape = (1, p) if loop else (p, 1)
    y = rng.poisson(rng.gamma(5.0, 1.0, size=shape)).astype(np.float64)
    mdl = Model([Poisson("y", rate="lam"), Gamma("lam", shape="a", rate="b")])
    if loop:
        smp = RandomWalkLoop("lam", mdl, step=np.array([[0.5]]), domain_limits=np.array([[0.0, np.inf]]),
                             max_variable_size=(1, p))
    else:
        smp = ManifoldMALA("lam", mdl, step=np.array([[0.5]]))
    state = {"y": y, "lam": y + 1.0, "a": np.array([[2.0]]), "b": np.array([[0.5]])}
    return mdl, [smp], state


def _rj(rng, n_data, n_max, full):
    """The reference's RJ test model (tests/test_reversible_jump.py:137-252: Gaussian-kernel basis, mixture prior on
    the coefficients, Poisson prior on the count) at the C5 shape, Normal response; the three callbacks restate what
    that test installs (basis rebuilt from knots / widths; allocation vector grown / shrunk)."""
    import numpy as np
    from scipy import sparse
    from scipy.stats import norm

    from openmcmc import parameter
    from openmcmc.distribution.distribution import Gamma, Poisson, Uniform
    from openmcmc.distribution.location_scale import Normal
    from openmcmc.model import Model
    from openmcmc.sampler.metropolis_hastings import ManifoldMALA, RandomWalkLoop
    from openmcmc.sampler.reversible_jump import ReversibleJump

    def basis(X, knots, widths):
        return np.hstack([norm.pdf(X, loc=knots[:, k], scale=widths[:, k]) for k in range(knots.shape[1])])

    def moved(state, _column):
        state["B"] = basis(state["X"], state["theta"], state["omega"])
        return state, 0.0, 0.0

    def born(_current, prop):
        prop["B"] = basis(prop["X"], prop["theta"], prop["omega"])
        prop["alloc_beta"] = np.vstack([prop["alloc_beta"], np.zeros((1, 1), dtype=int)])
        return prop, 0.0, 0.0

    def died(_current, prop, index):
        prop["B"] = np.delete(prop["B"], obj=index, axis=1)
        prop["alloc_beta"] = np.delete(prop["alloc_beta"], obj=index, axis=0)
        return prop, 0.0, 0.0

    rho = n_max / 4.0
    k0 = int(rho)
    X = np.sort(rng.uniform(-10, 10, n_data)).reshape(-1, 1)
    theta = rng.uniform(-10, 10, (1, k0))
    omega = rng.uniform(0.8, 1.6, (1, k0))
    beta = rng.standard_normal((k0, 1))
    B = basis(X, theta, omega)
    y = B @ beta + 0.1 * rng.standard_normal((n_data, 1))
    mdl = Model([
        Normal("y", mean=parameter.LinearCombination(form={"beta": "B"}),
               precision=parameter.ScaledMatrix(matrix="P", scalar="tau_y")),
        Normal("beta", mean=parameter.MixtureParameterVector(param="mu_beta", allocation="alloc_beta"),
               precision=parameter.MixtureParameterMatrix(param="tau_beta", allocation="alloc_beta")),
        Poisson("n_basis", rate="rho"),
        Uniform("theta", domain_response_lower=np.array([[-10.0]]), domain_response_upper=np.array([[10.0]])),
        Gamma("omega", shape="a_omega", rate="b_omega")])
    rj = ReversibleJump(param="n_basis", model=mdl, associated_params=["theta", "omega"], n_max=n_max,
                        state_birth_function=born, state_death_function=died,
                        matching_params={"variable": "beta", "matrix": "B", "scale": 1.0, "limits": [-10.0, 10.0]})
    samplers = [rj]
    if full:
        samplers = [ManifoldMALA(param="beta", model=mdl, step=np.array(0.8), max_variable_size=n_max),
                    RandomWalkLoop(param="theta", model=mdl, step=np.array(0.3), max_variable_size=n_max,
                                   domain_limits=np.array([-10.0, 10.0], ndmin=2), state_update_function=moved),
                    RandomWalkLoop(param="omega", model=mdl, step=np.array(0.1), max_variable_size=n_max,
                                   domain_limits=np.array([0.5, 2.0], ndmin=2), state_update_function=moved),
                    rj]
    state = {"y": y, "beta": beta, "tau_y": 100.0, "P": sparse.eye(n_data), "B": B, "n_basis": k0, "X": X,
             "theta": theta, "omega": omega, "mu_beta": np.zeros((1, 1)), "tau_beta": 0.25 * np.ones((1, 1)), "rho": rho,
             "alloc_beta": np.zeros((k0, 1), dtype=int), "a_omega": 3.0 * np.ones((1, 1)), "b_omega": 2.0 * np.ones((1, 1))}
    return mdl, samplers, state


def _worker(workload, chains, sweeps, seed, n, p, n_thin=1, response=False):
    import numpy as np

    ref_mcmc = _import_reference()
    rng = np.random.default_rng(seed)
    np.random.seed(seed)                          # the reference draws from numpy's global RandomState
    runs = []
    for _ in range(chains):
        if workload in ("c1", "c2"):
            mdl, samplers, state = _regression(rng, n, p, response)
        elif workload == "c3":
            mdl, samplers, state = _gmrf(rng, n)
        elif workload in ("c4a", "c4b"):
            mdl, samplers, state = _mh(rng, p, loop=workload == "c4b")
        elif workload in ("c5", "c5full"):
            mdl, samplers, state = _rj(rng, n, p, full=workload == "c5full")
        else:
            raise ValueError(workload)
        n_iter = max(sweeps // n_thin, 1)
        runs.append(ref_mcmc.MCMC(state, samplers, mdl, n_burn=0, n_iter=n_iter, n_thin=n_thin))
    failed = 0
    t0 = time.perf_counter()
    for M in runs:
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                M.run_mcmc()
        except Exception:                          # the reference raises on an invalid mMALA proposal
            failed += 1
    return time.perf_counter() - t0, failed


def run_parallel(workload="c2", workers=None, chains_per_worker=1, sweeps=3, n=10000, p=64, seed=0, n_thin=1,
                 response=False):
    """`workers` single-threaded processes of the unmodified reference; returns dict(value=chain-iterations/s, ...)."""
    workers = workers or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    n_iter = max(sweeps // n_thin, 1)
    sweeps = n_iter * n_thin
    cmd = [sys.executable, os.path.abspath(__file__), "--worker", "--workload", workload, "--chains-per-worker",
           str(chains_per_worker), "--sweeps", str(sweeps), "--n", str(n), "--p", str(p), "--n-thin", str(n_thin)]
    if response:
        cmd.append("--response")
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd + ["--seed", str(seed + i)], env=env, stdout=subprocess.PIPE) for i in range(workers)]
    inner, failed = [], 0
    for pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("ref_bench worker failed")
        secs, bad = out.decode().strip().splitlines()[-1].split()
        inner.append(float(secs))
        failed += int(bad)
    wall = time.perf_counter() - t0
    total = workers * chains_per_worker * sweeps
    return {"value": total / max(inner), "unit": "chain-iterations/s", "cores": workers, "seconds": max(inner),
            "wall": wall, "chains_failed": failed, "kind": "reference",
            "sample": f"{workers} workers x {chains_per_worker} chains x {sweeps} sweeps of {workload} (n={n}, p={p}, "
                      f"n_thin={n_thin}): unmodified openmcmc 1.0.7 MCMC.run_mcmc() from baseline/_ref incl. its "
                      f"per-iteration store / log_post{' / fitted values' if response else ''}, 1 BLAS thread per worker"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--chains-per-worker", type=int, default=1)
    ap.add_argument("--sweeps", type=int, default=3)
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--p", type=int, default=64)
    ap.add_argument("--n-thin", type=int, default=1)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--response", action="store_true")
    a = ap.parse_args()
    if a.worker:
        secs, failed = _worker(a.workload, a.chains_per_worker, a.sweeps, a.seed, a.n, a.p, a.n_thin, a.response)
        print(secs, failed)
    else:
        print(json.dumps(run_parallel(a.workload, a.workers, a.chains_per_worker, a.sweeps, a.n, a.p, a.seed, a.n_thin,
                                      a.response)))


if __name__ == "__main__":
    main()
