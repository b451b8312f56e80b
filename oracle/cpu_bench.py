"""Oracle (test infrastructure): CPU timing leg used by bench.py (`cpu_baseline` and `--impl reference`).

Times the numpy restatement of the reference's Gibbs sweep on the host cores: one worker process per core, one BLAS
thread each (the reference is single-threaded; SURVEY B.3 measured that 8 BLAS threads are slower than 1 at p=64),
chains looped inside each worker exactly as a user of the reference would loop `MCMC.run_mcmc()` over chains.

    python -m oracle.cpu_bench --workload c2 --chains-per-worker 2 --sweeps 3 --workers 16
"""

import argparse
import json
import os
import subprocess
import sys
import time


def _worker(workload, chains, sweeps, seed, n, p):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    import numpy as np

    rng = np.random.default_rng(seed)
    if workload in ("c1", "c2"):
        return _worker_regression(rng, chains, sweeps, n, p)
    if workload == "c3":
        return _worker_gmrf(rng, chains, sweeps, n)
    if workload in ("c4a", "c4b"):
        return _worker_mh(rng, chains, sweeps, p, workload)
    if workload in ("c5", "c5full"):
        return _worker_rj(rng, chains, sweeps, n, p, full=workload == "c5full")
    raise ValueError(workload)


def _worker_rj(rng, chains, sweeps, n_data, n_max, full=False):
    """C5: ReversibleJump steps on the Gaussian-kernel basis model (n_data points, rho = n_max / 4 expected knots) with a
    Normal response, plus the per-iteration log_post (mcmc.py:108)."""
    import numpy as np

    from oracle import rj

    rho = n_max / 4.0
    X = np.sort(rng.uniform(-10, 10, n_data))
    data = []
    for _ in range(chains):
        k = int(rho)
        th, om = rng.uniform(-10, 10, k), rng.uniform(0.8, 1.6, k)
        B = rj.make_basis(X, th, om)
        be = rng.standard_normal(k)
        y = B @ be + 0.1 * rng.standard_normal(n_data)
        m = dict(X=X, y=y, tau_y=100.0, tau_beta=0.25, mu_beta=0.0, rho=rho, a_omega=3.0, b_omega=2.0, theta_lo=-10.0,
                 theta_hi=10.0, n_max=n_max, birth_probability=0.5, match_scale=1.0, match_limits=(-10.0, 10.0))
        data.append((m, dict(n=k, theta=th, omega=om, beta=be, B=B)))
    t0 = time.perf_counter()
    for m, st in data:
        for _ in range(sweeps):
            if full:   # the other three samplers of the RJ model: ManifoldMALA(beta), RandomWalkLoop(theta / omega)
                k = st["n"]
                st, _ = rj.coef_mmala_step(m, st, 0.8, rng.standard_normal(k), rng.random())
                st, _ = rj.knot_walk_sweep(m, st, "theta", 0.3, (-10.0, 10.0), rng.random(k), rng.random(k))
                st, _ = rj.knot_walk_sweep(m, st, "omega", 0.1, (0.5, 2.0), rng.random(k), rng.random(k))
            d = dict(u_move=rng.random(), theta_new=rng.uniform(-10, 10), omega_new=rng.gamma(3.0) / 2.0, beta_new=None,
                     u_trunc=rng.random(), del_index=float(rng.integers(0, st["n"])), u_accept=rng.random())
            st, _ = rj.rj_step(m, st, d)
            _ = rj.model_log_p(m, st["n"], st["theta"], st["omega"], st["beta"], st["B"])
    return time.perf_counter() - t0


def _worker_regression(rng, chains, sweeps, n, p):
    import numpy as np

    from oracle import conjugate, dist

    data = []
    for _ in range(chains):
        X = rng.standard_normal((n, p))
        X[:, 0] = 1.0
        y = X @ rng.standard_normal((p, 1)) + 0.1 * rng.standard_normal((n, 1))
        st = {"beta": np.zeros((p, 1)), "tau": 1.0, "lambda": 0.01, "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3,
              "b_lambda": 1e-3}
        data.append((X, y, st))
    t0 = time.perf_counter()
    for X, y, st in data:
        for _ in range(sweeps):
            st = conjugate.gibbs_regression_sweep(X, y, st, rng.standard_normal(p), rng.standard_gamma(n / 2),
                                                  rng.standard_gamma(p / 2))
            # per-iteration log-posterior as mcmc.py:108 (rss with the new beta, both Normal terms, both Gamma terms)
            _, _, rss, _ = conjugate.regression_suffstats(X, y, None, st["beta"])
            ssb, _ = conjugate.quadform(1.0, st["beta"], None)
            _ = (dist.normal_log_p_from_ss(n, st["tau"], 0.0, rss) + dist.normal_log_p_from_ss(p, st["lambda"], 0.0, ssb)
                 + dist.gamma_log_p(st["tau"], 1e-3, 1e-3) + dist.gamma_log_p(st["lambda"], 1e-3, 1e-3))
    return time.perf_counter() - t0


def _worker_gmrf(rng, chains, sweeps, n):
    """Example-4 sweep with the reference's own sparse call stack (gmrf.py:167-198 sparse branch: one splu with natural
    ordering and no pivoting + three spsolve), the NormalGamma updates and the per-iteration log_post, which
    re-factorises both Normal precisions (gmrf.py:339)."""
    import numpy as np

    from oracle import conjugate, dist, gmrf

    s = np.arange(n) * (60.0 / 99.0)
    pd, pe = gmrf.precision_irregular_diagonals(s)
    pd = pd.copy()
    pd[0] += 1e-3
    truth = np.sin(s / 20) + 2 * np.cos(s / 12) + 2
    data = [(truth + rng.standard_normal(n), {"lambda": 100.0, "tau": 1.0}) for _ in range(chains)]
    t0 = time.perf_counter()
    for y, st in data:
        for _ in range(sweeps):
            x, L = gmrf.sparse_sample_normal_canonical(st["lambda"] * pd + st["tau"], st["lambda"] * pe, st["tau"] * y,
                                                       rng.standard_normal(n))
            ssp = gmrf.tridiag_quadform(pd, pe, x)
            st["lambda"], _, _ = conjugate.normal_gamma(10.0, 1.0, ssp, float(n), rng.standard_gamma(10.0 + n / 2))
            r = y - x
            ssl = float(r @ r)
            st["tau"], _, _ = conjugate.normal_gamma(1.0, 1.0, ssl, float(n), rng.standard_gamma(1.0 + n / 2))
            # log_post: Normal.log_p factorises lambda*P and tau*I again (location_scale.py:145-167 -> gmrf.py:321-348)
            ld_p = gmrf.sparse_logdet(st["lambda"] * pd, st["lambda"] * pe)
            ld_w = gmrf.sparse_logdet(np.full(n, st["tau"]), np.zeros(n - 1))
            _ = (0.5 * (ld_p - n * np.log(2 * np.pi) - st["lambda"] * ssp) + 0.5 * (ld_w - n * np.log(2 * np.pi) - st["tau"] * ssl)
                 + dist.gamma_log_p(st["lambda"], 10.0, 1.0) + dist.gamma_log_p(st["tau"], 1.0, 1.0))
    return time.perf_counter() - t0


def _worker_mh(rng, chains, sweeps, p, workload):
    """C4: Poisson counts with a Gamma(2, 0.5) prior; c4a ManifoldMALA with the reference's finite-difference
    derivatives (distribution.py:124-198), c4b RandomWalkLoop over a (1, p) parameter with truncated proposals."""
    import numpy as np

    from oracle import mh

    data = []
    for _ in range(chains):
        y = rng.poisson(rng.gamma(5.0, 1.0, size=p)).astype(float)
        shape = (p, 1) if workload == "c4a" else (1, p)
        terms = [mh.Term("poisson_rate", data=y.reshape(shape)), mh.Term("gamma_response", p1=np.array([[2.0]]),
                                                                        p2=np.array([[0.5]]))]
        data.append((terms, (y + 1.0).reshape(shape)))
    t0 = time.perf_counter()
    for terms, theta in data:
        for _ in range(sweeps):
            if workload == "c4a":
                theta, _ = mh.mmala_step(terms, theta, 0.5, rng.standard_normal(p), rng.random(), "reference")
            else:
                theta, _ = mh.random_walk_loop_sweep(terms, theta, np.array([[0.5]]), rng.random((p, 1)), rng.random(p),
                                                     np.array([[0.0, np.inf]]))
            _ = mh.log_p(terms, theta)   # per-iteration log_post (mcmc.py:108)
    return time.perf_counter() - t0


def run_parallel(workload="c2", workers=None, chains_per_worker=2, sweeps=3, n=10000, p=64, seed=0):
    """Launch `workers` single-threaded processes; returns dict(value=chain-iterations/s, seconds, cores, sample)."""
    workers = workers or os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "oracle.cpu_bench", "--worker", "--workload", workload, "--chains-per-worker",
           str(chains_per_worker), "--sweeps", str(sweeps), "--n", str(n), "--p", str(p)]
    t0 = time.perf_counter()
    procs = [subprocess.Popen(cmd + ["--seed", str(seed + i)], env=env, stdout=subprocess.PIPE, cwd=root)
             for i in range(workers)]
    inner = []
    for pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError("cpu_bench worker failed")
        inner.append(float(out.decode().strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    total = workers * chains_per_worker * sweeps
    # throughput of the timed loops themselves (process start-up and data generation excluded)
    value = total / max(inner)
    return {"value": value, "unit": "chain-iterations/s", "cores": workers, "seconds": max(inner), "wall": wall,
            "sample": f"{workers} workers x {chains_per_worker} chains x {sweeps} sweeps of {workload} (n={n}, p={p}; "
                      f"numpy/scipy port of the reference sweep incl. per-iteration log_post), 1 BLAS thread per worker"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", action="store_true")
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--workers", type=int, default=None)
    ap.add_argument("--chains-per-worker", type=int, default=2)
    ap.add_argument("--sweeps", type=int, default=3)
    ap.add_argument("--n", type=int, default=10000)
    ap.add_argument("--p", type=int, default=64)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    if a.worker:
        print(_worker(a.workload, a.chains_per_worker, a.sweeps, a.seed, a.n, a.p))
    else:
        print(json.dumps(run_parallel(a.workload, a.workers, a.chains_per_worker, a.sweeps, a.n, a.p, a.seed)))


if __name__ == "__main__":
    main()
