#!/usr/bin/env python
"""Node (Hermite) table against the cell coefficient table: memory, build time and query rate on uniformly random and
cell-sorted batches, 3-D 256^3 (norm / vector / both) and 4-D 48^3 x 32 (norm / both), plus a parity sample."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import _lib, quadcubic, tricubic  # noqa: E402
from tools.perf_sweep import field_rows  # noqa: E402

dev = torch.device("cuda", 0)


def rate(obj, q, variant=0, steps=5, warmup=2):
    lib = obj._lib
    d, mode = obj._d, obj._mode
    n = q.shape[0]
    kw = dict(dtype=torch.float64, device=q.device)
    comps = torch.empty(n, 3, **kw) if mode in ("vector", "both") else None
    norm = torch.empty(n, 1, **kw) if mode in ("norm", "both") else None
    grad = torch.empty(n, d, **kw) if mode in ("norm", "both") else None
    cells = torch.empty(n, dtype=torch.int64, device=q.device)
    ptr = lambda t: None if t is None else t.data_ptr()
    st = torch.cuda.current_stream()
    nodes = obj._nodes is not None
    fn = lib.arb_query_nodes if nodes else lib.arb_query
    tab = obj._nodes if nodes else obj.table
    old = lib.arb_set_query_variant(variant)
    try:
        def launch():
            _lib.check(fn(ctypes.byref(obj._cgeom), tab.data_ptr(), obj._mode_code, q.data_ptr(), n, q.shape[1],
                          ptr(comps), ptr(norm), ptr(grad), cells.data_ptr(), None, None, st.cuda_stream), "query")
        for _ in range(warmup):
            launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(st)
        for _ in range(steps):
            launch()
        e1.record(st)
        torch.cuda.synchronize()
        return n * steps / (e0.elapsed_time(e1) / 1e3), (comps, norm, grad)
    finally:
        lib.arb_set_query_variant(old)


def build_ms(obj, reps=4):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        obj._build_nodes() if obj._nodes is not None else obj._build_table()
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


def one(cls, shape, mode, n, g, scalar=False):
    rows = field_rows(shape, dev)
    d = len(shape)
    if scalar:
        rows = rows[:, :d + 1].contiguous()
    kw = {} if scalar else {"mode": mode}
    cell = cls(rows, "quiet", **kw)
    node = cls(rows, "quiet", table="nodes", **kw)
    del rows
    lo = torch.tensor(cell._geo.int_min, dtype=torch.float64, device=dev)
    hi = torch.tensor(cell._geo.int_max, dtype=torch.float64, device=dev)
    q = lo + torch.rand(n, d, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)
    cell.Query(q[:1024].clone())
    cell.Query(q.clone())
    qs = q[torch.argsort(cell._last_cells)].contiguous()
    tag = f"{d}-D {'x'.join(map(str, shape))} {mode}"
    print(f"[nodes] {tag}: cell table {cell.table.numel() * 8 / 1e9:.2f} GB built in {build_ms(cell):.2f} ms, "
          f"node table {node.nodes.numel() * 8 / 1e9:.3f} GB built in {build_ms(node):.2f} ms", flush=True)
    for name, qq in (("uniform random", q), ("cell-sorted", qs)):
        rc, oc = rate(cell, qq)
        line = [f"cells {rc:.3e}"]
        for v in (0, 71, 72, 73):
            rn, on = rate(node, qq, v)
            line.append(f"nodes v{v} {rn:.3e} (x{rn / rc:.2f})")
        err = 0.0
        for a, b in zip(oc, on):
            if a is not None:
                err = max(err, float(((a - b).abs() / (b.abs() + 1.0)).max()))
        print(f"[nodes] {tag} {name}: " + " | ".join(line) + f" | max |cells - nodes| / (|.| + 1) = {err:.2e}", flush=True)
    del cell, node
    torch.cuda.empty_cache()


def main():
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    n = int(os.environ.get("ARB_N", str(1 << 25)))
    for mode in ("norm", "vector", "both"):
        one(tricubic, (256,) * 3, mode, n if mode == "norm" else n // 2, g)
    one(tricubic, (128,) * 3, "norm", n, g)
    one(quadcubic, (48, 48, 48, 32), "norm", n // 4, g, scalar=True)
    one(quadcubic, (48, 48, 48, 32), "both", n // 8, g)


if __name__ == "__main__":
    main()
