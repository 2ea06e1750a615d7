#!/usr/bin/env python
"""Upper bound for query binning: rate of the existing query kernel on batches pre-sorted by cell /
by z-layer (table blocks then come from L2), plus the cost of torch's own sort as a reference."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
from tools.perf_sweep import field_rows, time_query
dev = torch.device("cuda", 0)
obj = tricubic(field_rows((256,) * 3, dev), "quiet", mode="norm")
n = 1 << 26
g = torch.Generator(device=dev); g.manual_seed(1)
lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=dev); hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=dev)
q = lo + torch.rand(n, 3, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)
print(f"[bin] random order: {time_query(obj, q, 0):.4e} q/s", flush=True)
obj.Query(q[:1024])
# z-layer bins (253 bins), random inside a bin
layer = torch.floor((q[:, 2] - lo[2]) / float(obj.hz)).to(torch.int32)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); order = torch.argsort(layer, stable=True); e1.record(); torch.cuda.synchronize()
print(f"[bin] torch argsort of 2^26 int32 layer keys: {e0.elapsed_time(e1):.2f} ms", flush=True)
qz = q[order].contiguous()
print(f"[bin] sorted by z-layer (253 bins, 32.8 MB of table each): {time_query(obj, qz, 0):.4e} q/s", flush=True)
pair = layer // 2
qz2 = q[torch.argsort(pair, stable=True)].contiguous()
print(f"[bin] sorted by z-layer pairs (127 bins, 65.6 MB each): {time_query(obj, qz2, 0):.4e} q/s", flush=True)
quad = layer // 4
qz4 = q[torch.argsort(quad, stable=True)].contiguous()
print(f"[bin] sorted by 4 z-layers (64 bins, 131 MB each): {time_query(obj, qz4, 0):.4e} q/s", flush=True)
del qz, qz2, qz4
cells = torch.empty(n, dtype=torch.int64, device=dev)
res = obj.Query(q); cell = obj._last_cells
qc = q[torch.argsort(cell)].contiguous()
print(f"[bin] fully sorted by cell: {time_query(obj, qc, 0):.4e} q/s", flush=True)
