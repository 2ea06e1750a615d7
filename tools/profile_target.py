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
ap.add_argument("--order", default="random", choices=["random", "sorted"], help="row order of the batch (sorted = by cell)")
ap.add_argument("--table", default="cells", choices=["cells", "nodes", "free"], help="free = table=False (interleaved grid where it exists)")
ap.add_argument("--variant", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
shape = (a.grid,) * 3 if a.d == 3 else (48, 48, 48, 32)
kw = {"table": "nodes"} if a.table == "nodes" else ({"table": False, "interleave": True} if a.table == "free" and a.mode != "norm" else ({"table": False} if a.table == "free" else {}))
obj = (tricubic if a.d == 3 else quadcubic)(field_rows(shape, dev), "quiet", mode=a.mode, **kw)
g = torch.Generator(device=dev); g.manual_seed(1)
q = torch.rand(a.queries, a.d, generator=g, dtype=torch.float64, device=dev)
lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=dev)
hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=dev)
q = lo + q * (hi - lo) * (1 - 1e-12)
if a.order == "sorted":
    obj._lib.arb_set_query_variant(a.variant)          # one kernel launch for the pre-pass too (ncu --launch-skip counts it)
    obj.Query(q.clone())
    q = q[torch.argsort(obj._last_cells)].contiguous()
if a.table == "free":
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(a.launches):
        obj.Query(q)
    torch.cuda.synchronize()
    rate = a.queries * a.launches / (time.perf_counter() - t0)
elif a.table == "nodes":
    from tools.nodes_bench import rate as node_rate
    rate = node_rate(obj, q, a.variant, steps=a.launches, warmup=0)[0]
else:
    rate = time_query(obj, q, a.variant, steps=a.launches, warmup=0)
print(f"profile target d={a.d} mode={a.mode}: {rate:.3e} q/s (under profiler: not a bench value)")
