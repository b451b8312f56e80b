#!/usr/bin/env python
"""bench.py — chain-iterations/s of the openMCMC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c3|c4a|c4b]

A "step" is one sweep (every sampler once) over all chains of the workload; `value` = chain-iterations/s with the
inputs resident in HBM; `e2e` = the same metric through the public API (`MCMC(...).run_mcmc()`) with HOST inputs, the
host->device upload and the device->host sample download inside the timed region.  N > 1: one process per GPU
(torchrun), chains sharded by rank (weak scaling, no data-path collective), time = max over ranks.
`--impl reference` times the CPU path (numpy/scipy port of the reference sweep, oracle/cpu_bench.py) on the host cores.

Default workload = BASELINE.json configs[1] (batched Bayesian linear regression), the configuration the metric is
quoted on that fits one GPU; the other configs are selectable for the per-config numbers in DESIGN.md / profiles/.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "chain-iterations/sec"
UNIT = "chain-iterations/s"

WORKLOADS = {
    # BASELINE.json configs[1]: batched Bayesian linear regression (the config the metric is quoted on; fits 1 GPU)
    "c2": dict(name="batched Bayesian linear regression: 4096 chains/GPU, n=10000, p=64, NormalNormal+NormalGamma Gibbs",
               kind="regression", chains=4096, n=10000, p=64, thin=1, dominant="nn_dense_draw",
               cpu=dict(chains_per_worker=8, sweeps=100), ref=dict(chains_per_worker=2, sweeps=5)),
    # BASELINE.json configs[0]: example-3 regression, single chain (latency bound)
    "c1": dict(name="examples/3_linear_regression: 1 chain, n=1000, p=3", kind="regression", chains=1, n=1000, p=3,
               thin=1, dominant="nn_dense_draw", cpu=dict(chains_per_worker=1, sweeps=4000),
               ref=dict(chains_per_worker=1, sweeps=500)),
    # BASELINE.json configs[2]: example-4 GMRF smoother scaled up (sparse-enabled form)
    "c3": dict(name="temporal GMRF smoother: 64 chains/GPU, n=1e6 grid points, tridiagonal NormalNormal + 2x NormalGamma",
               kind="gmrf", chains=64, n=1_000_000, p=0, thin=10, dominant="tridiag_nn_draw",
               cpu=dict(chains_per_worker=1, sweeps=3), ref=dict(chains_per_worker=1, sweeps=1)),
    # BASELINE.json configs[3]: 65,536 chains x 32 params over 8 GPUs = 8192 chains per GPU
    "c4a": dict(name="ManifoldMALA, Poisson counts + Gamma prior: 8192 chains/GPU x 32 params", kind="mh", chains=8192,
                n=0, p=32, thin=1, dominant="mmala", cpu=dict(chains_per_worker=1, sweeps=30),
                ref=dict(chains_per_worker=1, sweeps=4)),
    "c4b": dict(name="RandomWalkLoop (truncated proposals), Poisson counts + Gamma prior: 8192 chains/GPU x (1,32) params",
                kind="mh", chains=8192, n=0, p=32, thin=1, dominant="random_walk_loop",
                cpu=dict(chains_per_worker=4, sweeps=300), ref=dict(chains_per_worker=2, sweeps=50)),
    # BASELINE.json configs[4]: ReversibleJump on the Gaussian-kernel basis model, 8192 chains, capacity 128 components
    "c5": dict(name="ReversibleJump birth/death, Gaussian-kernel basis model: 8192 chains/GPU, n_data=512, n_max=128, "
                    "rho=32, Normal response, matched transitions", kind="rj", chains=8192, n=512, p=128, thin=1,
               dominant="reversible_jump", cpu=dict(chains_per_worker=2, sweeps=300), ref=dict(chains_per_worker=1, sweeps=100)),
    # the same model with all four samplers of the reference's RJ model: ManifoldMALA on the coefficients, RandomWalkLoop
    # on knots and widths (basis rebuilt per proposal), ReversibleJump (the CPU port times the RJ step only)
    "c5full": dict(name="full RJ source model: ManifoldMALA(beta) + RandomWalkLoop(theta) + RandomWalkLoop(omega) + "
                        "ReversibleJump, 8192 chains/GPU, n_data=512, n_max=128, rho=32", kind="rj", full=True, chains=8192,
                   n=512, p=128, thin=1, dominant="reversible_jump", cpu=dict(chains_per_worker=1, sweeps=40),
                   ref=dict(chains_per_worker=1, sweeps=10)),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=None, help="override chains per GPU (debug only)")
    ap.add_argument("--n", type=int, default=None, help="override n (debug only)")
    ap.add_argument("--upload-blocks", type=int, default=None,
                    help="chain blocks of the e2e run (upload of block k+1 under the sweeps of block k); default: "
                         "MCMC's automatic choice, one block per 4 GB of per-chain host input")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def config_of(args, wl):
    return {"workload": wl["name"], "chains_per_gpu": args.chains or wl["chains"], "n_obs": args.n or wl["n"],
            "p": wl["p"], "n_thin": wl["thin"]}


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args, wl, key):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_bench

    cores = os.cpu_count() or 1
    per_step = wl["ref"]   # each step = a bounded sample of the workload on every core
    n, p = args.n or wl["n"], wl["p"]
    for _ in range(1 if args.warmup > 0 else 0):
        cpu_bench.run_parallel(workload=key, workers=cores, n=n, p=p, **per_step)
    secs, steps_done = 0.0, 0
    for k in range(min(args.steps, 8)):
        r = cpu_bench.run_parallel(workload=key, workers=cores, n=n, p=p, seed=100 + k, **per_step)
        secs += r["seconds"]
        steps_done += 1
    total_its = cores * per_step["chains_per_worker"] * per_step["sweeps"] * steps_done
    value = total_its / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
        "warmup": 1 if args.warmup > 0 else 0, "ms_per_step": secs / steps_done * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_of(args, wl),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps_done} steps x ({cores} workers x {per_step['chains_per_worker']} chains x "
                                   f"{per_step['sweeps']} sweeps), numpy/scipy port of the reference sweep (oracle/), "
                                   "1 BLAS thread per worker"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []      # (host time, csv line)
        self.proc = None
        self.t_mark = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Start of the timed region: the sampler itself starts before the warm-up (nvidia-smi needs a second or more
        to come up on an 8-GPU node), samples from here on are the ones reported."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        timed = [r for t, r in self.rows if self.t_mark is None or t >= self.t_mark]
        window = "timed region + dominant-op timing"
        if not timed:     # the timed region was shorter than one sampling period: report the warm-up samples instead
            timed, window = [r for _, r in self.rows], "warm-up + timed region (timed region shorter than one sample)"
        for r in timed:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm), "window": window}


# --------------------------------------------------------------------------------------------- workloads (B200 arm)
def build_regression(C, n, p, dev, rank, host):
    import numpy as np
    import torch
    from scipy import sparse

    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    # synthetic data generated on the device (X = [1, N(0,1)...], y = X beta* + 0.1 eps), seed = rank
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    X = torch.randn(C, n, p, dtype=torch.float64, device=dev, generator=gen)
    X[:, :, 0] = 1.0
    beta_true = torch.randn(C, p, 1, dtype=torch.float64, device=dev, generator=gen)
    y = torch.bmm(X, beta_true) + 0.1 * torch.randn(C, n, 1, dtype=torch.float64, device=dev, generator=gen)
    if host:
        X, y = _pinned(X), _pinned(y)
    mdl = Model([
        Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
        Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
        Gamma("tau", shape="a_tau", rate="b_tau"),
        Gamma("lambda", shape="a_lambda", rate="b_lambda")])
    samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
    state = {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.identity(n, format="csc"), "tau": 1.0,
             "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": 0.01,
             "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3, "b_lambda": 1e-3}
    return mdl, samplers, state


def build_gmrf(C, n, p, dev, rank, host):
    """regular grid s_i = i*(60/99), P = precision_irregular(s), P[0,0] += 1e-3,
    y = sin(s/20) + 2 cos(s/12) + 2 + eps; lambda0 = 100, a_lam = 10, b_lam = 1, tau0 = 1, a_tau = b_tau = 1."""
    import numpy as np
    import torch
    from scipy import sparse

    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    s = np.arange(n) * (60.0 / 99.0)
    dr = 1.0 / np.diff(s)
    pd = np.append(np.append(dr[0], dr[:-1] + dr[1:]), dr[-1])
    pd[0] += 1e-3
    P = sparse.diags([-dr, pd, -dr], offsets=[-1, 0, 1], format="csc")
    gen = torch.Generator(device=dev)
    gen.manual_seed(4321 + rank)
    truth = torch.as_tensor(np.sin(s / 20) + 2 * np.cos(s / 12) + 2).to(dev)
    y = truth.reshape(1, n, 1) + torch.randn(C, n, 1, dtype=torch.float64, device=dev, generator=gen)
    if host:
        y = _pinned(y)
    mdl = Model([Normal("y", mean=LinearCombination(form={"b": "I"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
                 Normal("b", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
                 Gamma("lambda", shape="a_lam", rate="b_lam"),
                 Gamma("tau", shape="a_tau", rate="b_tau")])
    samplers = [NormalNormal("b", mdl), NormalGamma("lambda", mdl), NormalGamma("tau", mdl)]
    state = {"y": y, "b": y, "mu": np.zeros(n), "lambda": 100, "P_lambda": P, "a_lam": 10, "b_lam": 1, "tau": 1,
             "P_tau": sparse.identity(n, format="csc"), "I": sparse.identity(n, format="csc"), "a_tau": 1, "b_tau": 1}
    return mdl, samplers, state


def build_mh(C, n, p, dev, rank, host, loop):
    """y_j ~ Poisson(lambda*_j), lambda* ~ Gamma(5,1), prior Gamma(2, 0.5); step 0.5."""
    import numpy as np
    import torch

    from openmcmc_b200.distribution.distribution import Gamma, Poisson
    from openmcmc_b200.model import Model
    from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA, RandomWalkLoop

    rng = np.random.default_rng(99 + rank)
    shape = (C, 1, p) if loop else (C, p, 1)
    y = rng.poisson(rng.gamma(5.0, 1.0, size=shape)).astype(np.float64)
    yt = torch.as_tensor(y)
    yt = yt.pin_memory() if host else yt.to(dev)
    mdl = Model([Poisson("y", rate="lam"), Gamma("lam", shape="a", rate="b")])
    if loop:
        smp = RandomWalkLoop("lam", mdl, step=np.array([[0.5]]), domain_limits=np.array([[0.0, np.inf]]),
                             max_variable_size=(1, p))
    else:
        smp = ManifoldMALA("lam", mdl, step=np.array([[0.5]]))
    state = {"y": yt, "lam": torch.as_tensor(y + 1.0), "a": np.array([[2.0]]), "b": np.array([[0.5]])}
    return mdl, [smp], state


def build_rj(C, n_data, n_max, dev, rank, host, full=False):
    """the reference RJ test model scaled — n_data points on [-10, 10], capacity n_max, rho = n_max / 4,
    Gaussian-kernel basis, truncated matching [-10, 10] with scale 1; every chain has its own response."""
    import numpy as np
    import torch
    from scipy import sparse

    from openmcmc_b200.distribution.distribution import Gamma, Poisson, Uniform
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, MixtureParameterMatrix, MixtureParameterVector, ScaledMatrix
    from openmcmc_b200.sampler.reversible_jump import GaussianKernelBasis, ReversibleJump

    rng = np.random.default_rng(77 + rank)
    rho = n_max / 4.0
    k0 = int(rho)
    X = np.sort(rng.uniform(-10, 10, n_data))
    theta = rng.uniform(-10, 10, (C, 1, k0))
    omega = rng.uniform(0.8, 1.6, (C, 1, k0))
    beta = rng.standard_normal((C, k0, 1))
    z = (X.reshape(1, -1, 1) - theta) / omega
    Bt = np.exp(-0.5 * z * z) / (np.sqrt(2 * np.pi) * omega)                # [C, n_data, k0]
    y = Bt @ beta + 0.1 * rng.standard_normal((C, n_data, 1))
    yt = torch.as_tensor(y)
    yt = yt.pin_memory() if host else yt.to(dev)
    mdl = Model([Normal("y", mean=LinearCombination(form={"beta": "B"}), precision=ScaledMatrix(matrix="P", scalar="tau_y")),
                 Normal("beta", mean=MixtureParameterVector(param="mu_beta", allocation="alloc_beta"),
                        precision=MixtureParameterMatrix(param="tau_beta", allocation="alloc_beta")),
                 Poisson("n_basis", rate="rho"),
                 Uniform("theta", domain_response_lower=np.array([[-10.0]]), domain_response_upper=np.array([[10.0]])),
                 Gamma("omega", shape="a_omega", rate="b_omega")])
    smp = ReversibleJump(param="n_basis", model=mdl, associated_params=["theta", "omega"], n_max=n_max,
                         matching_params={"variable": "beta", "matrix": "B", "scale": 1.0, "limits": [-10.0, 10.0]},
                         basis=GaussianKernelBasis(matrix="B", locations="X", knots="theta", widths="omega"))
    state = {"y": yt, "beta": beta, "tau_y": 100.0, "P": sparse.eye(n_data), "B": np.zeros((n_data, k0)), "n_basis": k0,
             "X": X.reshape(-1, 1), "theta": theta, "omega": omega, "mu_beta": np.zeros((1, 1)),
             "tau_beta": 0.25 * np.ones((1, 1)), "rho": rho, "alloc_beta": np.zeros((k0, 1)),
             "a_omega": 3.0 * np.ones((1, 1)), "b_omega": 2.0 * np.ones((1, 1))}
    if full:
        from openmcmc_b200.sampler.metropolis_hastings import ManifoldMALA, RandomWalkLoop

        samplers = [ManifoldMALA(param="beta", model=mdl, step=np.array(0.8), max_variable_size=n_max),
                    RandomWalkLoop(param="theta", model=mdl, step=np.array(0.3), max_variable_size=n_max,
                                   domain_limits=np.array([-10.0, 10.0], ndmin=2)),
                    RandomWalkLoop(param="omega", model=mdl, step=np.array(0.1), max_variable_size=n_max,
                                   domain_limits=np.array([0.5, 2.0], ndmin=2)),
                    smp]
        return mdl, samplers, state
    return mdl, [smp], state


def _pinned(t):
    import torch

    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t)
    return h


def build(wl, C, n, dev, rank, host=False):
    if wl["kind"] == "regression":
        return build_regression(C, n, wl["p"], dev, rank, host)
    if wl["kind"] == "gmrf":
        return build_gmrf(C, n, wl["p"], dev, rank, host)
    if wl["kind"] == "rj":
        return build_rj(C, n, wl["p"], dev, rank, host, full=wl.get("full", False))
    return build_mh(C, n, wl["p"], dev, rank, host, loop=wl["dominant"] == "random_walk_loop")


def roofline_of(wl, C, n, p, op_ms, step_ms, peaks, fp64_peak, syrk_ms=None):
    """Algorithmic work of the dominant op per launch over its measured duration."""
    hbm_peak, hbm_src = peaks
    sec = op_ms * 1e-3
    if wl["kind"] == "regression":
        # per sweep: ONE stream over X, y for the residual (omc_reg_rss; SURVEY §8d bytes = 8 n (p+1) per chain-iteration);
        # G = X'X and g = X'y depend on the data alone and come from one omc_reg_pass (DMMA SYRK) in the prologue
        byts = C * 8 * n * (p + 1)
        roof = {"bound": "hbm", "kernel": "reg_pass_kernel<SYRK=false> (omc_reg_rss: residual pass over X, y)",
                "achieved": byts / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": byts / sec / 1e9 / hbm_peak,
                "peak_source": hbm_src, "traffic": None, "kernel_ms": op_ms, "share_of_step": op_ms / step_ms,
                "peak_note": "MEASURED_PEAKS.json's hbm_gbs is a COPY (half reads, half writes, bus turnarounds); this "
                             "kernel only reads, and a read-only stream runs faster than a copy on HBM3e, so frac can "
                             "exceed 1 against the copy figure (ncu: dram__bytes_read = 21.30 GB per launch, "
                             "profiles/r01b_ncu_c2.txt); against the 8 TB/s nominal it is achieved / 8000"}
        if syrk_ms:
            flops = C * (n * p * (p + 1) + 4 * n * p)   # SYRK + X'y + residual per chain
            tf = flops / (syrk_ms * 1e-3) / 1e12
            roof["prologue_syrk"] = {
                "bound": "tensor", "kernel": "reg_pass_kernel (FP64 DMMA SYRK + X'y + rss), once per run (prologue)",
                "achieved": tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": tf / fp64_peak if fp64_peak else None,
                "peak_source": "cuBLAS DGEMM fp64 6144^3 measured in this run (MEASURED_PEAKS.json has no fp64 figure)",
                "kernel_ms": syrk_ms}
        return roof
    if wl["kind"] == "gmrf":
        byts = C * 32 * n                                 # read y, P diag + off, write b
        return {"bound": "hbm", "kernel": "omc_tridiag_nn_draw (tg_aggregate_kernel + tg_tilescan_kernel + tg_solve_kernel)",
                "achieved": byts / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": byts / sec / 1e9 / hbm_peak,
                "peak_source": hbm_src, "traffic": None, "kernel_ms": op_ms, "share_of_step": op_ms / step_ms}
    if wl["kind"] == "rj":
        k = p / 4.0                                       # expected live components (rho = n_max / 4)
        byts = C * 2 * 8 * n * k                          # two passes over the live basis columns (8 n_data k)
        flops = C * (2 * n * k * k + 4.0 / 3.0 * k ** 3)
        return {"bound": "hbm", "kernel": "rj_kernel (Gram + in-place inverse + LU per chain; latency / FP64-ALU bound)",
                "achieved": byts / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": byts / sec / 1e9 / hbm_peak,
                "peak_source": hbm_src, "traffic": None, "kernel_ms": op_ms, "share_of_step": op_ms / step_ms,
                "fp64": {"achieved_tflops": flops / sec / 1e12, "flops_per_chain_step": flops / C}}
    byts = C * 32 * p                                     # read theta, y; write theta, sample
    return {"bound": "hbm", "kernel": f"{wl['dominant']}_kernel (latency / FP64-ALU bound: 1 KB per chain-iteration)",
            "achieved": byts / sec / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": byts / sec / 1e9 / hbm_peak,
            "peak_source": hbm_src, "traffic": None, "kernel_ms": op_ms, "share_of_step": op_ms / step_ms}


def run_b200(args, wl, key):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    C, n, p, thin = args.chains or wl["chains"], args.n or wl["n"], wl["p"], wl["thin"]
    thin = max(1, min(thin, args.steps))
    # CPU baseline first (before CUDA is initialised in this process), rank 0 at N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu_bench

        r = cpu_bench.run_parallel(workload=key, workers=os.cpu_count() or 1, n=n, p=p, **wl["cpu"])
        cpu_baseline = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]}
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from openmcmc_b200 import kernels as K
    from openmcmc_b200.mcmc import MCMC

    K.init_device(local)
    dev = torch.device("cuda", local)
    mdl, samplers, state = build(wl, C, n, dev, rank)
    n_iter = max(args.steps // thin, 1)
    M = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=n_iter, n_thin=thin, n_chains=C, seed=7, device=local,
             chain_offset=rank * C)
    M.prepare()
    launches_per_sweep = M.launches_per_sweep()
    store_launches = M._store_graph.num_kernels()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # warm-up (untimed), then exactly K timed sweeps; every n_thin-th sweep is followed by the store graph
    clocks = ClockSampler(local)
    clocks.start()
    M.run_device(n_burn=args.warmup, n_iter=0, n_thin=1)
    barrier()
    clocks.mark()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(M.stream):
        e0.record()
    M.run_device(n_burn=args.steps % thin if args.steps >= thin else 0, n_iter=n_iter, n_thin=thin)
    with torch.cuda.stream(M.stream):
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # dominant op alone: the very launch closure of the sweep plan, on the engine's stream, events on that stream
    op = next(fn for label, fn in M._ops["sweep"] if label.startswith(wl["dominant"]))
    reps = 10
    with torch.cuda.stream(M.stream):
        op_graph = K.Graph.capture(op)      # replayed from a graph like the sweep itself: no host launch gaps in the timing
        k0 = torch.cuda.Event(enable_timing=True)
        k1 = torch.cuda.Event(enable_timing=True)
        op_graph.launch(1)
        k0.record()
        op_graph.launch(reps)
        k1.record()
    barrier()
    op_ms = k0.elapsed_time(k1) / reps
    syrk_ms = None
    if wl["kind"] == "regression":   # the data-only SYRK pass of the prologue, timed the same way (it runs once per run)
        op2 = next((fn for label, fn in M._ops["prologue"] if label.startswith("reg_pass")), None)
        if op2 is not None:
            with torch.cuda.stream(M.stream):
                g2 = K.Graph.capture(op2)
                q0 = torch.cuda.Event(enable_timing=True)
                q1 = torch.cuda.Event(enable_timing=True)
                g2.launch(1)
                q0.record()
                g2.launch(reps)
                q1.record()
            barrier()
            syrk_ms = q0.elapsed_time(q1) / reps
    clk = clocks.stop()
    # the reference's form of the sweep (A'QA recomputed by every NormalNormal.sample): the full fused pass, DMMA SYRK
    # included, every sweep -- timed on the same resident inputs for comparison with the shipped data-only caching
    syrk_sweep = None
    if wl["kind"] == "regression":
        from openmcmc_b200 import engine

        engine.CACHE_DATA_ONLY = False
        try:
            M3 = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=n_iter, n_thin=thin, n_chains=C, seed=7, device=local,
                      chain_offset=rank * C)
            M3.prepare()
        finally:
            engine.CACHE_DATA_ONLY = True
        M3.run_device(n_burn=args.warmup, n_iter=0, n_thin=1)
        barrier()
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(M3.stream):
            s0.record()
        M3.run_device(n_burn=args.steps % thin if args.steps >= thin else 0, n_iter=n_iter, n_thin=thin)
        with torch.cuda.stream(M3.stream):
            s1.record()
        barrier()
        t3 = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t3, op=dist.ReduceOp.MAX)
        ms3 = float(t3.item())
        syrk_sweep = {"value": C * world * args.steps / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3 / args.steps,
                      "note": "same workload with G = X'X and g = X'y recomputed by the fused DMMA pass in EVERY sweep "
                              "(engine.CACHE_DATA_ONLY = False), as the reference does; the shipped plan keeps them "
                              "from the prologue because they depend on the data alone"}
        del M3
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # ---- diagnostics of the timed draws: per-chain ESS on the device store, all-gather of the per-chain records over
    #      the process group (NCCL over NVLink when N > 1), split-R-hat / ESS on every rank
    from openmcmc_b200 import diagnostics as G

    strides = {"b": max(1, n // 64)} if wl["kind"] == "gmrf" else None
    d0 = torch.cuda.Event(enable_timing=True)
    d1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(M.stream):
        d0.record()
    summ = G.summarize(M, elem_stride=strides)
    with torch.cuda.stream(M.stream):
        d1.record()
    torch.cuda.synchronize()
    ess_total = float(G.min_ess_per_chain(summ, all_ranks=True).sum().item())
    rhat_max = max(float(torch.nan_to_num(v["rhat"], nan=1.0).max().item()) for v in summ.values())
    ess = {"value": ess_total / (ms_max * 1e-3), "unit": "ESS/s", "n_stored": int(M.plan.iter_counter.item()),
           "ess_total": ess_total, "rhat_max": rhat_max, "params": sorted(summ), "diag_ms": d0.elapsed_time(d1),
           "gathered_chains": int(next(iter(summ.values()))["n_chains_total"]),
           "note": "sum over all chains of the minimum-over-parameters autocorrelation ESS (Geyer) of the draws stored "
                   "in the timed region, divided by the timed seconds; per-chain records all-gathered over the "
                   "process group; rhat_max is taken ACROSS chains and is only meaningful when chains share their data (the "
                   "c2/c4 workloads give every chain its own synthetic data set, so it is large by construction)"}
    M.collect()
    status_bad = int(((M.status & 3) != 0).sum())
    accept = {s.param: s.accept_rate.get_acceptance_rate() for s in samplers if hasattr(s, "accept_rate")}
    value = C * world * args.steps / (ms_max * 1e-3)

    try:
        peaks = (float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]),
                 "MEASURED_PEAKS.json hbm_gbs (of measured)")
    except Exception:
        peaks = (6650.0, "fallback 6.65 TB/s (of fallback)")
    fp64_peak = None
    if rank == 0 and wl["kind"] == "regression":
        # FP64 peak (cuBLAS DGEMM) measured live: the roofline denominator for the DMMA SYRK pass
        a = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
        b = torch.randn(6144, 6144, dtype=torch.float64, device=dev)
        best = 0.0
        for _ in range(4):
            g0 = torch.cuda.Event(enable_timing=True)
            g1 = torch.cuda.Event(enable_timing=True)
            g0.record()
            _ = a @ b
            g1.record()
            torch.cuda.synchronize()
            best = max(best, 2 * 6144 ** 3 / (g0.elapsed_time(g1) * 1e-3) / 1e12)
        fp64_peak = best
        del a, b

    # ---- e2e: public API with HOST (pinned) inputs, upload + K sweeps + sample download inside the timed region
    e2e = None
    if not args.no_e2e:
        del M, op, op_graph, state, summ
        op2 = g2 = None
        # device memory goes back to torch's caching allocator and STAYS there (no empty_cache): the e2e run takes its
        # buffers from the pool the way a long-running process would; on a fresh box the first cudaMalloc of 21 GB
        # cost up to 0.4 s (observed 0.03-0.09 s per 4.3 GB block), which is the driver's page-table work, not the path
        # host-memory guard: every rank of the node pins its own copy of the inputs (c2: 21 GB per rank); when the
        # node cannot hold them all, the e2e leg runs on the largest chain count per rank that fits and says so
        Ce, e2e_note = C, ""
        per_chain_host = 8.0 * n * (p + 1) if wl["kind"] == "regression" else 8.0 * n * 2 if wl["kind"] == "gmrf" else 0.0
        try:
            import psutil

            avail = psutil.virtual_memory().available
            local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            budget = 0.6 * avail / max(local_world, 1)
            if per_chain_host * C > budget:
                Ce = max(64, int(budget / per_chain_host) // 64 * 64)
                e2e_note = (f" [e2e on {Ce} of {C} chains per GPU: {local_world} ranks x {per_chain_host * C / 1e9:.1f} GB of "
                            f"pinned host input exceed 60% of the node's {avail / 1e9:.0f} GB of free host memory]")
        except Exception:
            pass
        mdl, samplers2, hstate = build(wl, Ce, n, dev, rank, host=True)
        import contextlib
        import io

        def e2e_run(blocks, first=True):
            if first:
                barrier()
            else:               # a retry on one rank must not wait on a collective the other ranks have passed
                torch.cuda.synchronize()
            t_start = time.perf_counter()
            run = MCMC(hstate, samplers2, model=mdl, n_burn=args.steps % thin if args.steps >= thin else 0, n_iter=n_iter,
                       n_thin=thin, n_chains=Ce, seed=7, device=local, chain_offset=rank * C, upload_blocks=blocks)
            with contextlib.redirect_stdout(io.StringIO()):
                run.run_mcmc()
            return run, t_start

        try:
            M2, t0 = e2e_run(args.upload_blocks)
        except Exception as exc:   # never lose the whole line to the e2e leg: one more try as a single block, and say so
            e2e_note += f" [first e2e attempt failed ({type(exc).__name__}: {str(exc)[:120]}); re-run with upload_blocks=1]"
            M2, t0 = e2e_run(1, first=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": Ce * world * args.steps / dt, "unit": UNIT, "chains_per_gpu": Ce,
               "upload_blocks": M2.timing.get("upload_blocks", 1),
               "h2d_bytes_per_step": M2.timing["h2d_bytes"] / args.steps,
               "d2h_bytes_per_step": M2.timing["d2h_bytes"] / args.steps, "seconds": dt,
               "phases_s": {k: round(M2.timing[k], 4) for k in ("prepare_s", "sweeps_s", "collect_s") if k in M2.timing},
               "blocks": M2.timing.get("blocks"),
               "note": "MCMC(...).run_mcmc() with pinned host inputs: upload + plan compile + graph capture + "
                       f"{args.steps} sweeps + download of all stored samples; upload_blocks > 1: the chains run as "
                       "chain blocks, block k+1 uploading while block k sweeps" + e2e_note}

    if rank == 0:
        roof = roofline_of(wl, C, n, p, op_ms, ms_max / args.steps, peaks, fp64_peak, syrk_ms)
        try:
            roof["traffic"] = json.load(open(os.path.join(ROOT, "profiles", f"{key}_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        cfg = config_of(args, wl)
        cfg.update({"l2": "per-GPU inputs are far larger than the 126 MB L2 (c2: 21 GB of X, c3: 512 MB of y + 1 GB of "
                          "scratch); c1/c4 working sets are L2-resident by nature of the workload",
                    "per_step": "1 sweep = every sampler once over all chains; every n_thin-th sweep also stores the "
                                "samples and log_post", "chains_failed": status_bad, "accept": accept})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clk, "e2e": e2e,
            "gpu_launches": launches_per_sweep * args.steps + store_launches * n_iter,
            "roofline": roof, "cpu_baseline": cpu_baseline, "ess": ess,
        }
        if syrk_sweep is not None:
            line["syrk_every_sweep"] = syrk_sweep
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        run_b200(args, wl, args.workload)


if __name__ == "__main__":
    main()
