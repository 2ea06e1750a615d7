#!/usr/bin/env python
"""Clustered-query fast path (compact slot rings, coordinate prefetch, sortedness probe): rate of every query-kernel
variant on uniformly random, cell-sorted, bunched and trajectory-like batches, 3-D and 4-D, each variant's outputs
compared bit for bit with the default kernel's."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import _lib, quadcubic, tricubic  # noqa: E402
from tools.perf_sweep import field_rows  # noqa: E402

dev = torch.device("cuda", 0)
VARIANTS = [int(v) for v in os.environ.get("ARB_VARIANTS", "0,25,24,40,43,60,61,62,63").split(",")]


def run(obj, q, variant, steps=5, warmup=2):
    lib = obj._lib
    d, mode = obj._d, obj._mode
    n = q.shape[0]
    kw = dict(dtype=torch.float64, device=q.device)
    comps = torch.zeros(n, 3, **kw) if mode in ("vector", "both") else None
    norm = torch.zeros(n, 1, **kw) if mode in ("norm", "both") else None
    grad = torch.zeros(n, d, **kw) if mode in ("norm", "both") else None
    cells = torch.zeros(n, dtype=torch.int64, device=q.device)
    ptr = lambda t: None if t is None else t.data_ptr()
    stream = torch.cuda.current_stream()
    old = lib.arb_set_query_variant(variant)
    try:
        def launch():
            _lib.check(lib.arb_query(ctypes.byref(obj._cgeom), obj.table.data_ptr(), obj._mode_code, q.data_ptr(), n,
                                     q.shape[1], ptr(comps), ptr(norm), ptr(grad), cells.data_ptr(), None, None,
                                     stream.cuda_stream), "query")
        for _ in range(warmup):
            launch()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(steps):
            launch()
        e1.record(stream)
        torch.cuda.synchronize()
        return n * steps / (e0.elapsed_time(e1) / 1e3), [t for t in (comps, norm, grad, cells) if t is not None]
    finally:
        lib.arb_set_query_variant(old)


def sweep(tag, obj, cases):
    for name, qq in cases.items():
        base_rate, base_out = run(obj, qq, 0)
        line = [f"v0 {base_rate:.3e}"]
        for v in VARIANTS:
            if v == 0:
                continue
            rate, out = run(obj, qq, v)
            same = all(torch.equal(a.view(torch.int64), b.view(torch.int64)) for a, b in zip(out, base_out))
            line.append(f"v{v} {rate:.3e} x{rate / base_rate:.2f}{'' if same else ' MISMATCH'}")
        print(f"[cluster {tag}] {name}: " + " | ".join(line), flush=True)


def cases3(obj, n, g):
    lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=dev)
    hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=dev)
    h = torch.tensor(obj._geo.h, dtype=torch.float64, device=dev)
    q = lo + torch.rand(n, 3, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)
    q[::1001] = 7.0                                   # some rows outside the volume
    out = {"uniform random": q}
    obj.Query(q.clone())
    cell = obj._last_cells
    out["sorted by cell"] = q[torch.argsort(cell)].contiguous()
    layer = torch.floor((q[:, 2] - lo[2]) / h[2])
    out["sorted by z layer"] = q[torch.argsort(layer)].contiguous()
    centre = 0.5 * (lo + hi)
    p = centre + torch.randn(n // 64, 3, generator=g, dtype=torch.float64, device=dev) * 20 * h
    out["64 samples around each particle (0.1 cell)"] = (
        p[:, None, :] + 0.1 * h * torch.randn(n // 64, 64, 3, generator=g, dtype=torch.float64, device=dev)).reshape(-1, 3).contiguous()
    p8 = lo + torch.rand(n // 8, 3, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * 0.99
    out["8 samples around each of n/8 random points (0.1 cell)"] = (
        p8[:, None, :] + 0.1 * h * torch.rand(n // 8, 8, 3, generator=g, dtype=torch.float64, device=dev)).reshape(-1, 3).contiguous()
    return out


def main():
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    n = int(os.environ.get("ARB_N", str(1 << 25)))
    for mode in os.environ.get("ARB_MODES", "norm,both").split(","):
        obj = tricubic(field_rows((256,) * 3, dev), "quiet", mode=mode)
        nn = n if mode == "norm" else n // 2
        sweep(f"3-D {mode}", obj, cases3(obj, nn, g))
        del obj
        torch.cuda.empty_cache()
    o4 = quadcubic(field_rows((48, 48, 48, 32), dev)[:, :5].contiguous(), "quiet")
    lo = torch.tensor(o4._geo.int_min, dtype=torch.float64, device=dev)
    hi = torch.tensor(o4._geo.int_max, dtype=torch.float64, device=dev)
    h = torch.tensor(o4._geo.h, dtype=torch.float64, device=dev)
    n4 = n // 4
    q4 = lo + torch.rand(n4, 4, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)
    o4.Query(q4.clone())
    cases = {"uniform random": q4, "sorted by cell": q4[torch.argsort(o4._last_cells)].contiguous()}
    p4 = lo + torch.rand(n4 // 64, 4, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * 0.98
    cases["64 samples around each particle (0.1 cell)"] = (
        p4[:, None, :] + 0.1 * h * torch.rand(n4 // 64, 64, 4, generator=g, dtype=torch.float64, device=dev)).reshape(-1, 4).contiguous()
    sweep("4-D norm", o4, cases)


if __name__ == "__main__":
    main()
