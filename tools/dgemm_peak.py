"""cuBLAS DGEMM burst/sustained figure (FP64 roofline denominator; MEASURED_PEAKS.json has none)."""
import json, time, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda")
b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(3):
    c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
burst = 2 * n**3 / best * 1e-9
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); k = 0; t0 = time.time()
while time.time() - t0 < 3.0:
    c = a @ b; k += 1
    if k % 8 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sust = 2 * n**3 * k / e0.elapsed_time(e1) * 1e-9
print(json.dumps({"dgemm_tflops_burst": burst, "dgemm_tflops_sustained": sust, "n": n}))
