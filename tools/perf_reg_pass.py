"""Quick device-time probe of the C2 kernels at full size (not the bench; used while tuning)."""
import sys, json, torch
sys.path.insert(0, ".")
from openmcmc_b200 import kernels as K

K.init_device(0)
C, n, p = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 10000, 64
X = torch.randn(C, n, p, dtype=torch.float64, device="cuda")
y = torch.randn(C, n, dtype=torch.float64, device="cuda")
beta = torch.randn(C, p, dtype=torch.float64, device="cuda")
rec = p * p + p + 2
stats = torch.empty(C, rec, dtype=torch.float64, device="cuda")
ns, ws = K.reg_pass_workspace(C, n, p)
work = torch.empty(max(ws, 1), dtype=torch.float64, device="cuda")
tau = torch.ones(C, dtype=torch.float64, device="cuda"); lam = torch.ones(C, dtype=torch.float64, device="cuda")
out = torch.empty(C, dtype=torch.float64, device="cuda"); ss = torch.empty_like(out); cnt = torch.empty_like(out)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = t(lambda: K.reg_pass(X, y, None, beta, stats, work, C, n, p))
flops = C * (n * p * (p + 1) + 4 * n * p)
byts = C * 8 * n * (p + 1)
r = {"C": C, "n_split": ns, "reg_pass_ms": ms, "tflops_alg": flops / ms * 1e-9, "gbs": byts / ms * 1e-6,
     "chain_it_per_s": C / ms * 1e3}
ms2 = t(lambda: K.nn_dense_draw(C, p, stats, K.vec(tau, 1), K.MAT_EYE, K.vec(None), K.vec(lam, 1), K.vec(None), beta, K.rng(seed=1)))
r["nn_dense_draw_ms"] = ms2
ms3 = t(lambda: (K.quadform(C, p, K.vec(beta, p), K.vec(None), K.MAT_EYE, K.vec(None), ss, cnt),
                 K.ng_draw(C, K.vec(tau, 1), K.vec(tau, 1), K.vec(ss, 1), K.vec(cnt, 1), out, K.rng(seed=2))))
r["quad_ng_ms"] = ms3
print(json.dumps(r))
