#!/usr/bin/env python
"""bench.py — chain-iterations/s of the openMCMC hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c1|c3]

A "step" is one sweep (every sampler once) over all chains of the workload; `value` = chain-iterations/s with the
inputs resident in HBM; `e2e` = the same metric through the public API (`MCMC(...).run_mcmc()`) with HOST inputs, the
host->device upload and the device->host sample download inside the timed region.  N > 1: one process per GPU
(torchrun), chains sharded by rank (weak scaling, no data-path collective), time = max over ranks.
`--impl reference` times the CPU path (numpy port of the reference sweep, oracle/cpu_bench.py) on the host cores.
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
               chains=4096, n=10000, p=64),
    # BASELINE.json configs[0]: example-3 regression, single chain (latency bound)
    "c1": dict(name="examples/3_linear_regression: 1 chain, n=1000, p=3", chains=1, n=1000, p=3),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--chains", type=int, default=None, help="override chains per GPU (debug only)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_bench

    cores = os.cpu_count() or 1
    # each step = a bounded sample of the workload: every core runs 2 chains for 5 sweeps
    per_step = dict(chains_per_worker=2, sweeps=5)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_bench.run_parallel(workers=cores, n=wl["n"], p=wl["p"], **per_step)
    vals, secs = [], 0.0
    for k in range(args.steps if args.steps <= 8 else 8):
        r = cpu_bench.run_parallel(workers=cores, n=wl["n"], p=wl["p"], seed=100 + k, **per_step)
        vals.append(r["value"])
        secs += r["seconds"]
    steps_done = len(vals)
    total_its = cores * per_step["chains_per_worker"] * per_step["sweeps"] * steps_done
    value = total_its / secs
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_done,
        "warmup": 1, "ms_per_step": secs / steps_done * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "chains_per_gpu": wl["chains"], "n_obs": wl["n"], "p": wl["p"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{steps_done} steps x ({cores} workers x {per_step['chains_per_worker']} chains x "
                                   f"{per_step['sweeps']} sweeps), numpy port of the reference sweep (oracle/), "
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
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

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
        for r in self.rows:
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
        # median over the samples taken under load (upper half of the observed clocks)
        med = sm[len(sm) // 2] if sm else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------- B200 arm
def build_model():
    from openmcmc_b200.distribution.distribution import Gamma
    from openmcmc_b200.distribution.location_scale import Normal
    from openmcmc_b200.model import Model
    from openmcmc_b200.parameter import LinearCombination, ScaledMatrix
    from openmcmc_b200.sampler.sampler import NormalGamma, NormalNormal

    mdl = Model([
        Normal("y", mean=LinearCombination(form={"beta": "X"}), precision=ScaledMatrix(matrix="P_tau", scalar="tau")),
        Normal("beta", mean="mu", precision=ScaledMatrix(matrix="P_lambda", scalar="lambda")),
        Gamma("tau", shape="a_tau", rate="b_tau"),
        Gamma("lambda", shape="a_lambda", rate="b_lambda")])
    samplers = [NormalNormal("beta", mdl), NormalGamma("tau", mdl), NormalGamma("lambda", mdl)]
    return mdl, samplers


def make_state(X, y, n, p):
    from scipy import sparse
    import numpy as np

    return {"y": y, "X": X, "beta": np.zeros((p, 1)), "P_tau": sparse.identity(n, format="csc"), "tau": 1.0,
            "P_lambda": sparse.identity(p, format="csc"), "mu": np.zeros((p, 1)), "lambda": 0.01,
            "a_tau": 1e-3, "b_tau": 1e-3, "a_lambda": 1e-3, "b_lambda": 1e-3}


def run_b200(args, wl):
    import numpy as np
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # CPU baseline first (before CUDA is initialised in this process), rank 0 at N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import cpu_bench

        cpu_baseline = cpu_bench.run_parallel(workers=os.cpu_count() or 1, chains_per_worker=8, sweeps=100,
                                              n=wl["n"], p=wl["p"])
        cpu_baseline = {"value": cpu_baseline["value"], "unit": UNIT, "cores": cpu_baseline["cores"], "kind": "port",
                        "sample": cpu_baseline["sample"]}
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from openmcmc_b200 import kernels as K
    from openmcmc_b200.mcmc import MCMC

    K.init_device(local)
    C, n, p = args.chains or wl["chains"], wl["n"], wl["p"]
    dev = torch.device("cuda", local)
    # synthetic data generated on the device (X = [1, N(0,1)...], y = X beta* + 0.1 eps), seed = chain id
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    X = torch.randn(C, n, p, dtype=torch.float64, device=dev, generator=gen)
    X[:, :, 0] = 1.0
    beta_true = torch.randn(C, p, 1, dtype=torch.float64, device=dev, generator=gen)
    y = torch.bmm(X, beta_true) + 0.1 * torch.randn(C, n, 1, dtype=torch.float64, device=dev, generator=gen)
    mdl, samplers = build_model()
    M = MCMC(make_state(X, y, n, p), samplers, model=mdl, n_burn=0, n_iter=args.steps, n_chains=C, seed=7,
             device=local, chain_offset=rank * C)
    M.prepare()
    launches_per_sweep = M.launches_per_sweep()
    store_launches = M._store_graph.num_kernels()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # warm-up (untimed), then exactly K timed sweeps: every timed step = sweep graph + store graph (samples + log_post)
    M.run_device(n_burn=args.warmup, n_iter=0, n_thin=1)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(M.stream):
        e0.record()
    M.run_device(n_burn=0, n_iter=args.steps, n_thin=1)
    with torch.cuda.stream(M.stream):
        e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # dominant kernel alone, same buffers, on the engine's stream (events on that stream)
    rl = M.plan._regressions["y"]
    reps = 10
    with torch.cuda.stream(M.stream):
        k0 = torch.cuda.Event(enable_timing=True)
        k1 = torch.cuda.Event(enable_timing=True)
        K.reg_pass(rl.X.data, rl.y.data, None, rl.beta.data, rl.stats, rl.work, C, n, p)
        k0.record()
        for _ in range(reps):
            K.reg_pass(rl.X.data, rl.y.data, None, rl.beta.data, rl.stats, rl.work, C, n, p)
        k1.record()
    barrier()
    pass_ms = k0.elapsed_time(k1) / reps
    clk = clocks.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    M.collect()
    status_bad = int((M.status != 0).sum())
    value = C * world * args.steps / (ms_max * 1e-3)

    # FP64 peak (cuBLAS DGEMM) measured live: the roofline denominator for the DMMA SYRK pass
    fp64_peak = None
    hbm_peak = None
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak = float(peaks["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback 6.65 TB/s (of fallback)"
    if rank == 0:
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
        del M
        Xh = torch.empty(X.shape, dtype=torch.float64, pin_memory=True)
        yh = torch.empty(y.shape, dtype=torch.float64, pin_memory=True)
        Xh.copy_(X)
        yh.copy_(y)
        del X, y, rl
        torch.cuda.empty_cache()
        barrier()
        t0 = time.perf_counter()
        M2 = MCMC(make_state(Xh, yh, n, p), samplers, model=mdl, n_burn=0, n_iter=args.steps, n_chains=C, seed=7,
                  device=local, chain_offset=rank * C)
        M2.run_mcmc()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        e2e = {"value": C * world * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": M2.timing["h2d_bytes"] / args.steps,
               "d2h_bytes_per_step": M2.timing["d2h_bytes"] / args.steps,
               "seconds": dt, "note": "MCMC(...).run_mcmc() with pinned host X,y: upload + compile + graph capture + "
                                      f"{args.steps} sweeps + download of all stored samples"}

    if rank == 0:
        flops_alg = n * p * (p + 1) + 4 * n * p            # SYRK + X'y + residual per chain-iteration
        bytes_alg = 8 * n * (p + 1)
        roof = {"bound": "tensor", "kernel": "reg_pass_kernel (FP64 DMMA SYRK + X'y + rss)",
                "achieved": C * flops_alg / (pass_ms * 1e-3) / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": (C * flops_alg / (pass_ms * 1e-3) / 1e12) / fp64_peak if fp64_peak else None,
                "peak_source": "cuBLAS DGEMM fp64 6144^3 measured in this run (MEASURED_PEAKS.json has no fp64 figure)",
                "traffic": None, "kernel_ms": pass_ms, "share_of_step": pass_ms / (ms_max / args.steps),
                "hbm": {"achieved": C * bytes_alg / (pass_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": C * bytes_alg / (pass_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src}}
        try:
            roof["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "reg_pass_traffic.json")))["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["name"], "chains_per_gpu": C, "n_obs": n, "p": p,
                       "l2": "inputs (21.3 GB of X per GPU) are far larger than the 126 MB L2; no flush needed",
                       "per_step": "1 sweep = NormalNormal(beta) + fused X pass + NormalGamma(tau) + NormalGamma(lambda)"
                                   " + store of beta/tau/lambda/log_post", "chains_failed": status_bad},
            "clocks": clk, "e2e": e2e, "gpu_launches": (launches_per_sweep + store_launches) * args.steps,
            "roofline": roof, "cpu_baseline": cpu_baseline,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
