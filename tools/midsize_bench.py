#!/usr/bin/env python
"""Mid-size host batches (1e4..1e7 rows): per-call time vs chunk size of the stream ring."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
from tools.perf_sweep import field_rows
obj = tricubic(field_rows((128,) * 3, torch.device("cuda", 0)), "quiet", mode="norm")
rng = np.random.default_rng(0)
lo = np.array(obj._geo.int_min); hi = np.array(obj._geo.int_max)
for N in (20_000, 100_000, 400_000, 1_600_000, 6_400_000):
    q = lo + rng.uniform(0, 1, (N, 3)) * (hi - lo) * 0.999
    line = f"[mid] N={N:8d}:"
    for chunk in (0, 16384, 65536, 262144):
        os.environ["ARB_HOST_CHUNK_ROWS"] = str(chunk)
        for _ in range(3):
            obj.Query(q)
        reps = 20 if N <= 400_000 else 5
        t0 = time.perf_counter()
        for _ in range(reps):
            obj.Query(q)
        dt = (time.perf_counter() - t0) / reps
        line += f"  chunk={chunk or 'default'}: {dt * 1e6:8.0f} us ({N / dt:.2e} q/s)"
    print(line, flush=True)
