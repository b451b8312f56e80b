# cuBLAS DGEMM / DSYRK-equivalent peak on this GPU (roofline denominator for the FP64 SYRK pass).
import torch, time, json
torch.backends.cuda.matmul.allow_tf32 = False
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
best = 0
for i in range(6):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    best = max(best, 2 * n**3 / t / 1e12)
# sustained
torch.cuda.synchronize(); t0 = time.time(); k = 0
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 3.0:
    c = a @ b; k += 1
    if k % 4 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * n**3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
# copy bw f64
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
bw = 0
for i in range(5):
    e0.record(); y.copy_(x); e1.record(); torch.cuda.synchronize()
    bw = max(bw, 2 * x.numel() * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9)
print(json.dumps({"fp64_dgemm_tflops_burst": best, "fp64_dgemm_tflops_sustained": sust, "copy_gbs": bw, "gpu": torch.cuda.get_device_name(0)}))
