#!/usr/bin/env python
"""FP64 roofline denominator for the coefficient solve: cuBLAS DGEMM via torch.matmul (best of 10)."""
import torch
n = 8192
a = torch.randn(n, n, dtype=torch.float64, device="cuda"); b = torch.randn(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    c = a @ b
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c = a @ b; e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"fp64 dgemm {n}^3: {2 * n ** 3 / best / 1e9:.2f} TFLOP/s (best of 10, {best:.2f} ms)")
