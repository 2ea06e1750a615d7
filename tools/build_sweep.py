#!/usr/bin/env python
"""Times every build-kernel variant on the bench grids and cross-checks the tables they write (GPU box).

    python tools/build_sweep.py [--variants 0,1,5,6,7,8] [--d 3,4] [--modes norm,both]
    python tools/build_sweep.py --profile 5 --d 3      # one build of one variant, for ncu captures

Variants: 0 and 5..8 separable FP64-pipe kernels, 9 Kronecker DMMA, 1 dense DMMA (arb_build.cu)."""
import argparse
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import _lib, quadcubic, tricubic  # noqa: E402
from tools.perf_sweep import field_rows  # noqa: E402

HBM = 6458.4


def build(obj, table, reps):
    d, geo = obj._d, obj._geo
    lo, hi = obj._slab
    sub = obj._planes[:, lo:hi + 3].contiguous()
    n = (ctypes.c_int64 * 4)(*([geo.npts[a] for a in range(d - 1)] + [hi - lo + 3] + [1] * (4 - d)))
    stream = torch.cuda.current_stream()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _lib.check(obj._lib.arb_build_coeffs(d, sub.data_ptr(), sub.shape[0], ctypes.byref(n), table.data_ptr(), 1,
                                             stream.cuda_stream), "build")
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0,5,6,7,9,1")
    ap.add_argument("--d", default="3,4")
    ap.add_argument("--modes", default="norm,both")
    ap.add_argument("--grid3", type=int, default=256)
    ap.add_argument("--grid4", default="48,48,48,32")
    ap.add_argument("--profile", type=int, default=-1)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    variants = [int(v) for v in a.variants.split(",")]
    for d in [int(x) for x in a.d.split(",")]:
        shape = (a.grid3,) * 3 if d == 3 else tuple(int(x) for x in a.grid4.split(","))
        for mode in a.modes.split(","):
            rows = field_rows(shape, dev)
            obj = (tricubic if d == 3 else quadcubic)(rows, "quiet", mode=mode)
            del rows
            torch.cuda.empty_cache()
            lib = obj._lib
            gb = obj.table.numel() * 8 / 1e9
            if a.profile >= 0:
                old = lib.arb_set_build_variant(a.profile)
                ms = build(obj, obj.table, 1)
                lib.arb_set_build_variant(old)
                print(f"[profile] d={d} mode={mode} variant={a.profile}: {ms:.3f} ms (under profiler: not a bench value)")
                return
            ref = obj.table.clone() if gb < 12 else None      # default-variant table from the constructor
            scale = float(ref[:-1].abs().max()) if ref is not None else 1.0
            for v in variants:
                old = lib.arb_set_build_variant(v)
                try:
                    obj.table.fill_(float("nan"))
                    ms = build(obj, obj.table, a.reps)
                    msg = ""
                    if ref is not None:
                        diff = float((obj.table[:-1] - ref[:-1]).abs().max())
                        bad = int(torch.isnan(obj.table[:-1]).sum())
                        msg = f"  max|table - default table| / max|table| = {diff / scale:.2e}, NaNs {bad}"
                    print(f"[build] d={d} mode={mode} grid={shape} table={gb:.2f} GB variant={v}: {ms:.3f} ms  "
                          f"{gb / ms * 1e3:.0f} GB/s written = {gb / ms * 1e3 / HBM:.3f} of measured HBM peak{msg}",
                          flush=True)
                except Exception as e:  # noqa: BLE001
                    print(f"[build] d={d} mode={mode} variant={v}: FAILED {e}", flush=True)
                finally:
                    lib.arb_set_build_variant(old)
            del obj, ref
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
