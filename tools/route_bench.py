#!/usr/bin/env python
"""The route kernel of slab-sharded queries (arb_route_rows, csrc/arb_route.cu) on ONE GPU with W virtual ranks whose
inboxes all live in local HBM: what the kernel costs without the links (the N-GPU figure in bench.py's `sharded` object
adds NVLink), and the inbox form of the query kernel (arb_query_inbox) against the plain kernel on the same rows.
4-D 'both' on 48^3 x 32, trajectory-like rows spread over W slabs.  `--launches K` for ncu."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import _lib, quadcubic  # noqa: E402
from tools.perf_sweep import field_rows  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ranks", type=int, default=8)
ap.add_argument("--rows", type=int, default=1 << 22)
ap.add_argument("--launches", type=int, default=10)
a = ap.parse_args()
dev = torch.device("cuda", 0)
W, n, d = a.ranks, a.rows, 4
whole = quadcubic(field_rows((48, 48, 48, 32), dev), "quiet", mode="both")
lib = whole._lib
ns = whole._geo.ncell[3]
his = [ns * (r + 1) // W for r in range(W)]
g = torch.Generator(device=dev); g.manual_seed(2)
lo = torch.tensor(whole._geo.int_min, dtype=torch.float64, device=dev)
hi = torch.tensor(whole._geo.int_max, dtype=torch.float64, device=dev)
q = (lo + torch.rand(n, d, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)).contiguous()
cap, ld_in, ld = n + 1024, 6, 10
inbox = [torch.empty((W * cap, ld_in), dtype=torch.float64, device=dev) for _ in range(W)]
counts = [torch.zeros(16, dtype=torch.int64, device=dev) for _ in range(W)]
vp = ctypes.c_void_p
ib = (vp * W)(*[t.data_ptr() for t in inbox]); cb = (vp * W)(*[t.data_ptr() for t in counts])
hi_arr = (ctypes.c_int64 * W)(*his)
cursor = torch.zeros(W + 2, dtype=torch.int64, device=dev)
outside = torch.empty(n, dtype=torch.bool, device=dev)
st = torch.cuda.current_stream(dev)


def route():
    cursor.zero_()
    _lib.check(lib.arb_route_rows(ctypes.byref(whole._cgeom), q.data_ptr(), n, d, hi_arr, W, 0, ib, cb, cap, cursor.data_ptr(),
                                  cursor[W + 1:].data_ptr(), outside.data_ptr(), st.cuda_stream), "arb_route_rows")


def timed(fn, k):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(k):
        fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


ms = timed(route, a.launches)
sent = torch.stack([c[0] for c in counts]).tolist()
print(f"[route] {n} rows to {W} virtual ranks in local HBM: {ms:.3f} ms per launch ({n * 48 / ms / 1e6:.0f} GB/s of inbox rows written), "
      f"rows per owner {sent}", flush=True)
# the owner side: rank 0's inbox (rows of slab 0 from sender 0) through the inbox form, against the plain kernel on the same rows
part = quadcubic(field_rows((48, 48, 48, 32), dev), "quiet", mode="both", slab=(0, his[0]))
res = torch.empty((n, ld), dtype=torch.float64, device=dev)
rb = (vp * W)(*([res.data_ptr()] * W))
m = int(counts[0][0])
rows0 = inbox[0][:m, :4].contiguous()


def inbox_query():
    _lib.check(lib.arb_query_inbox(ctypes.byref(part._cgeom), part.table.data_ptr(), part._mode_code, inbox[0].data_ptr(),
                                   counts[0].data_ptr(), cap, rb, W, ld, st.cuda_stream), "arb_query_inbox")


t_in = timed(inbox_query, a.launches)
t_plain = timed(lambda: part.Query(rows0), a.launches)
print(f"[route] owner side, {m} rows: inbox form {t_in:.3f} ms ({m / t_in / 1e6:.1f} M rows/ms ... {m / t_in * 1e3:.3e} q/s), "
      f"plain kernel + output allocation {t_plain:.3f} ms ({m / t_plain * 1e3:.3e} q/s)", flush=True)
