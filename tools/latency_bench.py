#!/usr/bin/env python
"""Per-call latency of Query for small batches (the reference's single-point / line-query use)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
n = 48
ax = np.linspace(-1, 1, n); Z, Y, X = [a.ravel() for a in np.meshgrid(ax, ax, ax, indexing="ij")]
f = np.stack([X, Y, Z, np.sin(X) * np.cos(Y) * np.exp(-Z)], 1)
T = tricubic(f, "quiet")
rng = np.random.default_rng(0)
ref = {1: 78.9, 20: 278.3, 1000: 2317.2, 100000: 461446.2}     # reference numpy, us/call, measured in the build container
for N in (1, 20, 1000, 8192, 100000):
    q = rng.uniform(-0.9, 0.9, (N, 3))
    for _ in range(20):
        T.Query(q[0].copy() if N == 1 else q.copy())
    reps = 2000 if N <= 1000 else 100
    t0 = time.perf_counter()
    for _ in range(reps):
        T.Query(q[0].copy() if N == 1 else q.copy())
    dt = (time.perf_counter() - t0) / reps
    r = ref.get(N)
    print(f"[latency] N={N}: {dt * 1e6:.1f} us/call ({N / dt:.3e} q/s)" + (f"; reference numpy {r} us -> x{r / (dt * 1e6):.1f}" if r else ""), flush=True)
