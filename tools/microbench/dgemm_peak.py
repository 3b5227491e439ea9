# cuBLAS DGEMM ceiling via torch.matmul(float64) -- the FP64 "tensor" roofline denominator.
import torch, json, time
torch.backends.cuda.matmul.allow_tf32 = False
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): (a @ b)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
# sustained
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): c = a @ b
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / 20
print(json.dumps({"dgemm_n": n, "best_ms": best, "tflops_burst": 2 * n**3 / best * 1e-9, "sustained_ms": sus, "tflops_sustained": 2 * n**3 / sus * 1e-9}))
