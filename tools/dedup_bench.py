#!/usr/bin/env python
"""What the warp-level de-duplication (lanes wanting the same block fetch it once) buys on clustered batches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
from tools.perf_sweep import field_rows, time_query
dev = torch.device("cuda", 0)
obj = tricubic(field_rows((256,) * 3, dev), "quiet", mode="norm")
n = 1 << 25
g = torch.Generator(device=dev); g.manual_seed(1)
lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=dev); hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=dev)
h = torch.tensor([float(obj.hx), float(obj.hy), float(obj.hz)], dtype=torch.float64, device=dev)
cases = {}
q = lo + torch.rand(n, 3, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)
cases["uniform random"] = q
obj.Query(q); cell = obj._last_cells
cases["sorted by cell (4 queries/cell adjacent)"] = q[torch.argsort(cell)].contiguous()
centre = 0.5 * (lo + hi)
cases["bunch, sigma = 3 cells, unsorted"] = (centre + torch.randn(n, 3, generator=g, dtype=torch.float64, device=dev) * 3 * h).contiguous()
p = centre + torch.randn(n // 64, 3, generator=g, dtype=torch.float64, device=dev) * 20 * h
cases["64 samples around each of 2^19 particles (0.1 cell spread)"] = (p[:, None, :] + 0.1 * h * torch.randn(n // 64, 64, 3, generator=g, dtype=torch.float64, device=dev)).reshape(-1, 3).contiguous()
for name, qq in cases.items():
    a = time_query(obj, qq, 0); b = time_query(obj, qq, 20)
    print(f"[dedup] {name}: with {a:.3e} q/s, without {b:.3e} q/s -> x{a / b:.2f}", flush=True)
