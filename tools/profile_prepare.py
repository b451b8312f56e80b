"""Host-side profile of MCMC.run_mcmc() on the e2e path (where do the milliseconds of `prepare` go?).

    python tools/profile_prepare.py c2|c3 [chains]
"""
import cProfile
import io
import pstats
import sys
import time

sys.path.insert(0, ".")
import torch

import bench
from openmcmc_b200 import kernels as K
from openmcmc_b200.mcmc import MCMC

key = sys.argv[1] if len(sys.argv) > 1 else "c2"
wl = bench.WORKLOADS[key]
C = int(sys.argv[2]) if len(sys.argv) > 2 else wl["chains"]
K.init_device(0)
dev = torch.device("cuda", 0)
thin = wl["thin"]
steps = 20
for rep in range(2):
    mdl, samplers, state = bench.build(wl, C, wl["n"], dev, 0, host=True)
    torch.cuda.synchronize()
    M = MCMC(state, samplers, model=mdl, n_burn=0, n_iter=max(steps // thin, 1), n_thin=thin, n_chains=C, seed=7, device=0)
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    M.run_mcmc()
    pr.disable()
    dt = time.perf_counter() - t0
    print(f"== {key} rep {rep}: {dt:.3f} s  timing {dict((k, round(v, 4)) for k, v in M.timing.items() if k.endswith('_s'))}")
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
    print(s.getvalue()[:9000])
    del M, state
