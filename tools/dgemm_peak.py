"""cuBLAS DGEMM throughput (the FP64 GEMM roofline denominator); prints one JSON line."""
import json, torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2): c = a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): c = a @ b
e1.record(); torch.cuda.synchronize()
sus = e0.elapsed_time(e1) / 20
print(json.dumps({"dgemm_n": n, "dgemm_tflops_burst": 2 * n**3 / best * 1e-9, "dgemm_tflops_sustained": 2 * n**3 / sus * 1e-9}))
