#!/usr/bin/env python
"""Variant / mode / dimension sweep of the query kernels and timing of the build kernels (GPU box).
Prints one line per configuration; used to pick defaults and to fill DESIGN.md tables."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import _lib, quadcubic, tricubic  # noqa: E402
from bench import ALG_BYTES  # noqa: E402


def field_rows(shape, device, vector=True):
    d = len(shape)
    lo_hi = [(-1.0, 1.0), (-1.0, 1.0), (-1.0, 1.0), (0.0, 1.0)]
    ax = [torch.linspace(lo_hi[a][0], lo_hi[a][1], shape[a], dtype=torch.float64, device=device) for a in range(d)]
    grids = torch.meshgrid(*reversed(ax), indexing="ij")
    c = [g.reshape(-1) for g in reversed(grids)]
    X, Y, Z = c[0], c[1], c[2]
    T = c[3] if d == 4 else torch.zeros_like(X)
    bx = torch.sin(2 * np.pi * X) * torch.cos(np.pi * Y) * torch.exp(-Z) * torch.cos(2 * T)
    by = X * X * Y + Z * (1 + T)
    bz = torch.cos(X + Y + Z + T)
    return torch.stack(c + [bx, by, bz], dim=1)


def time_build(obj, reps=3):
    d = obj._d
    geo = obj._geo
    lo, hi = obj._slab
    sub = obj._planes[:, lo:hi + 3].contiguous()
    n = (ctypes.c_int64 * 4)(*([geo.npts[a] for a in range(d - 1)] + [hi - lo + 3] + [1] * (4 - d)))
    stream = torch.cuda.current_stream()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(obj._lib.arb_build_coeffs(d, sub.data_ptr(), sub.shape[0], ctypes.byref(n), obj.table.data_ptr(),
                                             1, stream.cuda_stream), "build")
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def time_query(obj, q, variant, steps=5, warmup=2):
    lib = obj._lib
    d, mode = obj._d, obj._mode
    n = q.shape[0]
    kw = dict(dtype=torch.float64, device=q.device)
    comps = torch.empty(n, 3, **kw) if mode in ("vector", "both") else None
    norm = torch.empty(n, 1, **kw) if mode in ("norm", "both") else None
    grad = torch.empty(n, d, **kw) if mode in ("norm", "both") else None
    cells = torch.empty(n, dtype=torch.int64, device=q.device)
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
        return n * steps / (e0.elapsed_time(e1) / 1e3)
    finally:
        lib.arb_set_query_variant(old)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--d", default="3,4")
    ap.add_argument("--modes", default="norm,vector,both")
    ap.add_argument("--variants", default="0,20,21,22,23,1,10")
    ap.add_argument("--grid3", type=int, default=256)
    ap.add_argument("--grid4", default="48,48,48,32")
    ap.add_argument("--queries", type=int, default=1 << 25)
    ap.add_argument("--build-variants", default="0,9,1")
    ap.add_argument("--table-free", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    variants = [int(v) for v in args.variants.split(",")]
    for d in [int(x) for x in args.d.split(",")]:
        shape = (args.grid3,) * 3 if d == 3 else tuple(int(x) for x in args.grid4.split(","))
        for mode in args.modes.split(","):
            rows = field_rows(shape, dev)
            cls = tricubic if d == 3 else quadcubic
            obj = cls(rows, "quiet", mode=mode)
            del rows
            torch.cuda.empty_cache()
            ncomp = obj.table.shape[1]
            cells = obj.nc
            tab_gb = obj.table.numel() * 8 / 1e9
            flops = cells * ncomp * 2.0 * (4 ** d) ** 2
            for bv in [int(v) for v in args.build_variants.split(",")]:
                old = obj._lib.arb_set_build_variant(bv)
                tb = time_build(obj)
                obj._lib.arb_set_build_variant(old)
                print(f"[build] d={d} mode={mode} grid={shape} cells={cells} C={ncomp} table={tab_gb:.2f} GB cfg={bv}: "
                      f"{tb:.2f} ms  -> {cells * ncomp / tb * 1e3:.3e} cell-comps/s, {tab_gb / tb * 1e3:.0f} GB/s written, "
                      f"{flops / tb / 1e9:.1f} TFLOP/s dense-equivalent", flush=True)
            obj._build_table()      # leave the default configuration's table in place
            nq = args.queries if (d == 3 and mode == "norm") else args.queries // 2
            g = torch.Generator(device=dev); g.manual_seed(1)
            q = torch.rand(nq, d, generator=g, dtype=torch.float64, device=dev)
            lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=dev)
            hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=dev)
            q = lo + q * (hi - lo) * (1 - 1e-12)
            for v in variants:
                try:
                    rate = time_query(obj, q, v)
                    print(f"[query] d={d} mode={mode} variant={v}: {rate:.4e} q/s  "
                          f"{ALG_BYTES[(d, mode)] * rate / 1e9:.0f} GB/s algorithmic "
                          f"({ALG_BYTES[(d, mode)] * rate / 1e9 / 6458.4:.3f} of measured copy peak)", flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"[query] d={d} mode={mode} variant={v}: FAILED {e}", flush=True)
            if args.table_free:
                del obj
                torch.cuda.empty_cache()
                rows = field_rows(shape, dev)
                tf = (tricubic if d == 3 else quadcubic)(rows, "quiet", mode=mode, table=False)
                del rows
                comps = torch.empty(q.shape[0], 3, dtype=torch.float64, device=dev) if mode in ("vector", "both") else None
                norm = torch.empty(q.shape[0], 1, dtype=torch.float64, device=dev) if mode in ("norm", "both") else None
                grad = torch.empty(q.shape[0], d, dtype=torch.float64, device=dev) if mode in ("norm", "both") else None
                cells = torch.empty(q.shape[0], dtype=torch.int64, device=dev)
                ptr = lambda t: None if t is None else t.data_ptr()
                st = torch.cuda.current_stream()
                def launch():
                    _lib.check(tf._lib.arb_query_grid(ctypes.byref(tf._cgeom), tf._planes.data_ptr(), tf._pitch, tf._mode_code,
                                                      q.data_ptr(), q.shape[0], q.shape[1], ptr(comps), ptr(norm), ptr(grad),
                                                      cells.data_ptr(), None, None, st.cuda_stream), "grid")
                launch(); launch()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(st)
                for _ in range(5):
                    launch()
                e1.record(st); torch.cuda.synchronize()
                rate = q.shape[0] * 5 / (e0.elapsed_time(e1) / 1e3)
                print(f"[query] d={d} mode={mode} TABLE-FREE grid={shape}: {rate:.4e} q/s (grid {tf._planes.numel() * 8 / 1e6:.0f} MB)", flush=True)
                del tf
                obj = None
            del obj, q
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
