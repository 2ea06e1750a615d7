#!/usr/bin/env python
"""Small, fixed workload for ncu captures: one 256^3 build (norm or both) + a few query launches."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import quadcubic, tricubic  # noqa: E402
from tools.perf_sweep import field_rows, time_query  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--d", type=int, default=3)
ap.add_argument("--mode", default="norm")
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--queries", type=int, default=1 << 24)
ap.add_argument("--launches", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
shape = (a.grid,) * 3 if a.d == 3 else (48, 48, 48, 32)
obj = (tricubic if a.d == 3 else quadcubic)(field_rows(shape, dev), "quiet", mode=a.mode)
g = torch.Generator(device=dev); g.manual_seed(1)
q = torch.rand(a.queries, a.d, generator=g, dtype=torch.float64, device=dev)
lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=dev)
hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=dev)
q = lo + q * (hi - lo) * (1 - 1e-12)
rate = time_query(obj, q, 0, steps=a.launches, warmup=0)
print(f"profile target d={a.d} mode={a.mode}: {rate:.3e} q/s (under profiler: not a bench value)")
