import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
from openmcmc_b200 import kernels as K
from test_gpu_rj import _dev, _pad
K.init_device()
def run(inject, sample_omega=False, C=512, n_max=20, nd=50, rho=8.0, sweeps=1200):
    rng = np.random.default_rng(0)
    X = _dev(np.sort(rng.uniform(-10, 10, nd)))
    n = _dev(np.full(C, 4.0))
    theta = _dev(_pad([rng.uniform(-10, 10, 4) for _ in range(C)], n_max))
    omega = _dev(_pad([np.ones(4) for _ in range(C)], n_max, 1.0))
    beta = _dev(_pad([np.zeros(4) for _ in range(C)], n_max))
    B = torch.zeros(C, nd, n_max, dtype=torch.float64, device="cuda")
    sc = {k: _dev([v]) for k, v in dict(tau_beta=0.25, mu_beta=0.0, rho=rho, a=3.0, b=2.0).items()}
    sweep = torch.zeros(1, dtype=torch.int64, device="cuda")
    cnt = torch.zeros(C, 2, dtype=torch.int64, device="cuda")
    probe = torch.zeros(C, 8, dtype=torch.float64, device="cuda")
    dbg = torch.zeros(C, 6, dtype=torch.float64, device="cuda")
    args = K.rj_args(C, nd, n_max, n, theta, omega, beta, B, X, -10.0, 10.0, 0.5, sample_omega=sample_omega,
                     omega_shape=K.vec(sc["a"]), omega_rate=K.vec(sc["b"]), mu_beta=K.vec(sc["mu_beta"]),
                     tau_beta=K.vec(sc["tau_beta"]), rho=K.vec(sc["rho"]), match_scale=1.0, match_limits=(-10.0, 10.0),
                     rng_=K.rng(seed=3, sweep=sweep, site=1), counters=cnt, probe=probe, debug=dbg if inject else None)
    K.rj_basis(args)
    tot = 0.0; cntm = 0; births = 0; bacc = 0; dacc = 0
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    for it in range(sweeps):
        if inject:
            u = torch.rand(C, 6, dtype=torch.float64, device="cuda", generator=g)
            nn = n.clone()
            dbg[:, 0] = u[:, 0]; dbg[:, 1] = -10 + 20 * u[:, 1]; dbg[:, 2] = 1.0
            dbg[:, 3] = float("nan"); dbg[:, 4] = torch.floor(u[:, 4] * nn); dbg[:, 5] = u[:, 5]
        K.reversible_jump(args)
        K.counter_add(sweep, 1)
        if it >= 300:
            tot += float(n.mean()); cntm += 1
            p = probe.cpu().numpy()
            births += p[:, 0].sum(); bacc += (p[:, 0] * p[:, 7]).sum(); dacc += ((1 - p[:, 0]) * p[:, 7]).sum()
    torch.cuda.synchronize()
    acc = cnt.cpu().numpy()
    print("inject", inject, "omega", sample_omega, "mean n", tot / cntm, "acc", acc[:, 0].sum() / acc[:, 1].sum(),
          "birth frac", births / (cntm * C), "birth acc", bacc / births, "death acc", dacc / (cntm * C - births))

def check(sample_omega=False, C=512, n_max=20, nd=50, rho=8.0, sweeps=600):
    from oracle import rj
    rng = np.random.default_rng(0)
    Xh = np.sort(rng.uniform(-10, 10, nd))
    X = _dev(Xh)
    n = _dev(np.full(C, 4.0))
    theta = _dev(_pad([rng.uniform(-10, 10, 4) for _ in range(C)], n_max))
    omega = _dev(_pad([np.ones(4) for _ in range(C)], n_max, 1.0))
    beta = _dev(_pad([np.zeros(4) for _ in range(C)], n_max))
    B = torch.zeros(C, nd, n_max, dtype=torch.float64, device="cuda")
    sc = {k: _dev([v]) for k, v in dict(tau_beta=0.25, mu_beta=0.0, rho=rho, a=3.0, b=2.0).items()}
    sweep = torch.zeros(1, dtype=torch.int64, device="cuda")
    probe = torch.zeros(C, 8, dtype=torch.float64, device="cuda")
    args = K.rj_args(C, nd, n_max, n, theta, omega, beta, B, X, -10.0, 10.0, 0.5, sample_omega=sample_omega,
                     omega_shape=K.vec(sc["a"]), omega_rate=K.vec(sc["b"]), mu_beta=K.vec(sc["mu_beta"]),
                     tau_beta=K.vec(sc["tau_beta"]), rho=K.vec(sc["rho"]), match_scale=1.0, match_limits=(-10.0, 10.0),
                     rng_=K.rng(seed=3, sweep=sweep, site=1), probe=probe)
    K.rj_basis(args)
    lq = []
    for it in range(sweeps):
        K.reversible_jump(args); K.counter_add(sweep, 1)
        if it > 300:
            p = probe.cpu().numpy(); lq.append(p)
    torch.cuda.synchronize()
    nn = n.cpu().numpy().astype(int); th = theta.cpu().numpy(); be = beta.cpu().numpy(); om = omega.cpu().numpy(); Bh = B.cpu().numpy()
    err = 0.0; allb = []; allt = []
    for c in range(C):
        k = nn[c]
        err = max(err, np.abs(Bh[c][:, :k] - rj.make_basis(Xh, th[c, :k], om[c, :k])).max())
        allb += list(be[c, :k]); allt += list(th[c, :k])
    print("B consistency err", err, "beta sd", np.std(allb), "beta mean", np.mean(allb), "theta mean/sd", np.mean(allt), np.std(allt), 20/np.sqrt(12))
    P = np.concatenate(lq)
    b = P[:, 0] == 1
    print("birth: mean lq_f", np.nanmean(P[b, 4]), "lq_r", np.nanmean(P[b, 5]), "nan frac", np.isnan(P[b, 6]).mean(), "logp diff", np.nanmean(P[b,3]-P[b,2]))
    print("death: mean lq_f", np.nanmean(P[~b, 4]), "lq_r", np.nanmean(P[~b, 5]), "nan frac", np.isnan(P[~b, 6]).mean(), "logp diff", np.nanmean(P[~b,3]-P[~b,2]))
for sw,b in ((1500,300),(6000,3000),(20000,10000)):
    import functools
    def r(sweeps=sw, burn=b):
        pass

def run2(sweeps, burn, sample_omega=False, C=512, n_max=20, nd=50, rho=8.0):
    rng = np.random.default_rng(0)
    X = _dev(np.sort(rng.uniform(-10, 10, nd)))
    n = _dev(np.full(C, 4.0))
    theta = _dev(_pad([rng.uniform(-10, 10, 4) for _ in range(C)], n_max))
    omega = _dev(_pad([np.ones(4) for _ in range(C)], n_max, 1.0))
    beta = _dev(_pad([np.zeros(4) for _ in range(C)], n_max))
    B = torch.zeros(C, nd, n_max, dtype=torch.float64, device="cuda")
    sc = {k: _dev([v]) for k, v in dict(tau_beta=0.25, mu_beta=0.0, rho=rho, a=3.0, b=2.0).items()}
    sweep = torch.zeros(1, dtype=torch.int64, device="cuda")
    args = K.rj_args(C, nd, n_max, n, theta, omega, beta, B, X, -10.0, 10.0, 0.5, sample_omega=sample_omega,
                     omega_shape=K.vec(sc["a"]), omega_rate=K.vec(sc["b"]), mu_beta=K.vec(sc["mu_beta"]),
                     tau_beta=K.vec(sc["tau_beta"]), rho=K.vec(sc["rho"]), match_scale=1.0, match_limits=(-10.0, 10.0),
                     rng_=K.rng(seed=3, sweep=sweep, site=1))
    K.rj_basis(args)
    acc = torch.zeros((), dtype=torch.float64, device="cuda"); m = 0
    for it in range(sweeps):
        K.reversible_jump(args); K.counter_add(sweep, 1)
        if it >= burn and it % 10 == 0:
            acc += n.mean(); m += 1
    torch.cuda.synchronize()
    print("sweeps", sweeps, "burn", burn, "omega", sample_omega, "mean n", float(acc) / m)
run2(1500, 300); run2(6000, 3000); run2(24000, 12000); run2(24000, 12000, True)
