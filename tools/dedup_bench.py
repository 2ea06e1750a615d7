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

# ---- the same question for the table-free kernels (variant 20 = no de-duplication there too)
import ctypes
from arbinterp_b200 import _lib, quadcubic


def time_grid(tf, qq, variant, steps=5):
    lib = tf._lib
    d = tf._d
    nq = qq.shape[0]
    norm = torch.empty(nq, 1, dtype=torch.float64, device=dev)
    grad = torch.empty(nq, d, dtype=torch.float64, device=dev)
    cells = torch.empty(nq, dtype=torch.int64, device=dev)
    st = torch.cuda.current_stream()
    old = lib.arb_set_query_variant(variant)
    try:
        def launch():
            _lib.check(lib.arb_query_grid(ctypes.byref(tf._cgeom), tf._planes.data_ptr(), tf._pitch, tf._mode_code,
                                          qq.data_ptr(), nq, qq.shape[1], None, norm.data_ptr(), grad.data_ptr(),
                                          cells.data_ptr(), None, None, st.cuda_stream), "grid")
        launch(); launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(st)
        for _ in range(steps):
            launch()
        e1.record(st); torch.cuda.synchronize()
        return nq * steps / (e0.elapsed_time(e1) / 1e3)
    finally:
        lib.arb_set_query_variant(old)


del obj
torch.cuda.empty_cache()
tf = tricubic(field_rows((256,) * 3, dev), "quiet", mode="norm", table=False)
for name, qq in cases.items():
    a = time_grid(tf, qq[: n // 2], 0); b = time_grid(tf, qq[: n // 2], 20)
    print(f"[dedup table-free 3-D] {name}: with {a:.3e} q/s, without {b:.3e} q/s -> x{a / b:.2f}", flush=True)
del tf
tf4 = quadcubic(field_rows((48, 48, 48, 32), dev)[:, :5].contiguous(), "quiet", table=False)
lo4 = torch.tensor(tf4._geo.int_min, dtype=torch.float64, device=dev); hi4 = torch.tensor(tf4._geo.int_max, dtype=torch.float64, device=dev)
h4 = torch.tensor(tf4._geo.h, dtype=torch.float64, device=dev)
n4 = 1 << 23
q4 = lo4 + torch.rand(n4, 4, generator=g, dtype=torch.float64, device=dev) * (hi4 - lo4) * (1 - 1e-12)
p4 = 0.5 * (lo4 + hi4) + torch.randn(n4 // 64, 4, generator=g, dtype=torch.float64, device=dev) * 5 * h4
traj = (p4[:, None, :] + 0.1 * h4 * torch.randn(n4 // 64, 64, 4, generator=g, dtype=torch.float64, device=dev)).reshape(-1, 4).contiguous()
for name, qq in (("uniform random", q4), ("64 samples around each of 2^17 particles (0.1 cell spread)", traj)):
    a = time_grid(tf4, qq, 0); b = time_grid(tf4, qq, 20)
    print(f"[dedup table-free 4-D] {name}: with {a:.3e} q/s, without {b:.3e} q/s -> x{a / b:.2f}", flush=True)
