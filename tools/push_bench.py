#!/usr/bin/env python
"""Particle-steps/s of the fused query+push kernel on the 256^3 bench field, against a Python loop of
Query (device tensors) + torch updates -- the per-step round trip it replaces."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
from tools.perf_sweep import field_rows
dev = torch.device("cuda", 0)
obj = tricubic(field_rows((256,) * 3, dev), "quiet", mode="norm")
nod = tricubic(field_rows((256,) * 3, dev), "quiet", mode="norm", table="nodes")      # the same push on the node table
n = 1 << 24
g = torch.Generator(device=dev); g.manual_seed(3)
lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=dev); hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=dev)
h = float(obj.hx)
for label, speed in (("0.01 cell/step", 0.01), ("0.3 cell/step", 0.3), ("3 cells/step", 3.0)):
    pos0 = lo + (0.25 + 0.5 * torch.rand(n, 3, generator=g, dtype=torch.float64, device=dev)) * (hi - lo)
    vel0 = torch.randn(n, 3, generator=g, dtype=torch.float64, device=dev)
    vel0 = vel0 / vel0.norm(dim=1, keepdim=True) * speed * h          # |v| dt = speed * h with dt = 1
    nsteps = 32 if speed < 1 else 8
    p, v = pos0.clone(), vel0.clone()
    obj.push(p, v, 1.0, 2, 1e-6)                                        # warm-up
    p, v = pos0.clone(), vel0.clone()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lost = obj.push(p, v, 1.0, nsteps, 1e-6); e1.record(); torch.cuda.synchronize()
    t_fused = e0.elapsed_time(e1) / 1e3
    # unfused: Python loop, everything on the device, one Query launch + a few torch kernels per step
    p2, v2 = pos0.clone(), vel0.clone()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    a = 1e-6 * obj.Query(p2)[1]
    for _ in range(nsteps):
        v2 += 0.5 * a; p2 += v2
        a = 1e-6 * obj.Query(p2)[1]
        v2 += 0.5 * a
    torch.cuda.synchronize(); t_loop = time.perf_counter() - t0
    ok = ~torch.isnan(p2[:, 0])
    err = float((p[ok] - p2[ok]).abs().max())
    pn, vn = pos0.clone(), vel0.clone()
    nod.push(pn, vn, 1.0, 2, 1e-6)
    pn, vn = pos0.clone(), vel0.clone()
    torch.cuda.synchronize(); e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(); nod.push(pn, vn, 1.0, nsteps, 1e-6); e3.record(); torch.cuda.synchronize()
    t_nodes = e2.elapsed_time(e3) / 1e3
    print(f"[push] {label}: node table {n * nsteps / t_nodes:.3e} particle-steps/s (x{t_fused / t_nodes:.2f} of the cell table), "
          f"max |dx| vs cell table {float((p[ok] - pn[ok]).abs().max()):.2e}", flush=True)
    print(f"[push] {label}: fused {n * nsteps / t_fused:.3e} particle-steps/s ({t_fused * 1e3:.1f} ms for {nsteps} steps of {n}), "
          f"Query loop {n * nsteps / t_loop:.3e} -> x{t_loop / t_fused:.1f}; lost {lost}; max |dx| fused vs loop {err:.2e}", flush=True)
