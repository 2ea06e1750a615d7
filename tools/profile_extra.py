#!/usr/bin/env python
"""ncu target for the table-free query kernel and the fused pusher (128^3 grid, small batches)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
from tools.perf_sweep import field_rows
dev = torch.device("cuda", 0)
rows = field_rows((256,) * 3, dev)
n = 1 << 23
g = torch.Generator(device=dev); g.manual_seed(1)
tf = tricubic(rows, "quiet", mode="norm", table=False)
lo = torch.tensor(tf._geo.int_min, dtype=torch.float64, device=dev); hi = torch.tensor(tf._geo.int_max, dtype=torch.float64, device=dev)
q = lo + torch.rand(n, 3, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)
tf.Query(q); tf.Query(q)
del tf
obj = tricubic(rows, "quiet", mode="norm")
h = float(obj.hx)
pos = lo + (0.25 + 0.5 * torch.rand(n, 3, generator=g, dtype=torch.float64, device=dev)) * (hi - lo)
vel = torch.randn(n, 3, generator=g, dtype=torch.float64, device=dev)
vel = vel / vel.norm(dim=1, keepdim=True) * 0.3 * h
obj.push(pos, vel, 1.0, 16, 1e-6)
torch.cuda.synchronize()
print("profile_extra done")
# 4-D table-free kernel (quadcubic(table=False)) on the 48^3 x 32 bench grid
from arbinterp_b200 import quadcubic
del obj
torch.cuda.empty_cache()
rows4 = field_rows((48, 48, 48, 32), dev)
tf4 = quadcubic(rows4[:, :5].contiguous(), "quiet", table=False)
lo4 = torch.tensor(tf4._geo.int_min, dtype=torch.float64, device=dev); hi4 = torch.tensor(tf4._geo.int_max, dtype=torch.float64, device=dev)
q4 = lo4 + torch.rand(n, 4, generator=g, dtype=torch.float64, device=dev) * (hi4 - lo4) * (1 - 1e-12)
tf4.Query(q4); tf4.Query(q4)
torch.cuda.synchronize()
print("profile_extra 4-D table-free done")
