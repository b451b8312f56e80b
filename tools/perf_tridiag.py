"""Quick device-time probe of the C3 kernels at full size (not the bench; used while tuning): times the whole
omc_tridiag_nn_draw (aggregate + tile scan + solve) with CUDA events."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from openmcmc_b200 import kernels as K

K.init_device(0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
s = np.arange(n) * (60.0 / 99.0)
dr = 1.0 / np.diff(s)
pd = np.append(np.append(dr[0], dr[:-1] + dr[1:]), dr[-1])
pd[0] += 1e-3
d_pd, d_pe = torch.as_tensor(pd).cuda(), torch.as_tensor(-dr).cuda()
y = torch.randn(C, n, dtype=torch.float64, device="cuda") + 2
lam = torch.full((C,), 100.0, dtype=torch.float64, device="cuda")
tau = torch.ones(C, dtype=torch.float64, device="cuda")
x = torch.empty_like(y)
ws = torch.zeros(K.tridiag_workspace(C, n), dtype=torch.uint8, device="cuda")
ssp = torch.zeros(C, dtype=torch.float64, device="cuda")
ssl = torch.zeros(C, dtype=torch.float64, device="cuda")
sweep = torch.zeros(1, dtype=torch.int64, device="cuda")
args = K.tridiag_args(C, n, d_pd, d_pe, ws, lam=K.vec(lam, 1), tau=K.vec(tau, 1), y=K.vec(y, n), x=x,
                      rng_=K.rng(seed=3, sweep=sweep, site=1), ss_prior=ssp, ss_lik=ssl)
for _ in range(3):
    K.tridiag_nn_draw(args)
torch.cuda.synchronize()
reps = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    K.tridiag_nn_draw(args)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(json.dumps({"C": C, "n": n, "ms": ms, "alg_gbs": C * 32 * n / ms * 1e-6, "frac_of_6468": C * 32 * n / ms * 1e-6 / 6468.6,
                  "ss_lik_mean": float(ssl.mean()) / n}))
