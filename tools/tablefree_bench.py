#!/usr/bin/env python
"""Table-free 'vector' / 'both': the component-interleaved grid (arb_query_gridil, LDGSTS gather, four lanes per
query; 4-D: query_gridil4_kernel, four t-plane passes) against the per-component TMA-box kernel (arb_query_grid), the
node table and the cell-table path; 3-D 256^3 and 128^3, 4-D 48^3x32 and 64^3x48; uniformly random and cell-sorted
batches.  ARB_TF_ONLY=4d runs the 4-D part only."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import quadcubic, tricubic  # noqa: E402
from tools.perf_sweep import field_rows  # noqa: E402

dev = torch.device("cuda", 0)


def rate(obj, q, steps=5, warmup=2):
    for _ in range(warmup):
        obj.Query(q)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        obj.Query(q)
    e1.record()
    torch.cuda.synchronize()
    return q.shape[0] * steps / (e0.elapsed_time(e1) / 1e3)


g = torch.Generator(device=dev)
g.manual_seed(3)
n = int(os.environ.get("ARB_N", str(1 << 24)))
cases = [] if os.environ.get("ARB_TF_ONLY") == "4d" else [((256,) * 3, n), ((128,) * 3, n)]
if os.environ.get("ARB_TF_ONLY") != "3d":
    cases += [((48, 48, 48, 32), n // 4), ((64, 64, 64, 48), n // 4)]
for shape, nq in cases:
    d = len(shape)
    cls = tricubic if d == 3 else quadcubic
    grid = "x".join(str(v) for v in shape) if d == 4 else f"{shape[0]}^3"
    for mode in ("vector", "both", "norm"):
        if d == 4 and mode == "norm":
            continue
        rows = field_rows(shape, dev)
        cells = cls(rows, "quiet", mode=mode)
        lo = torch.tensor(cells._geo.int_min, dtype=torch.float64, device=dev)
        hi = torch.tensor(cells._geo.int_max, dtype=torch.float64, device=dev)
        q = lo + torch.rand(nq, d, generator=g, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-12)
        cells.Query(q)
        qs = q[torch.argsort(cells._last_cells)].contiguous()
        base = {"uniform random": rate(cells, q), "cell-sorted": rate(cells, qs)}
        tgb = cells.table.numel() * 8 / 1e9
        cells.release()                        # `del` alone waits for the garbage collector (reference cycle)
        del cells
        torch.cuda.empty_cache()
        forms = [("planes + TMA boxes", dict(table=False, interleave=False))]
        if mode != "norm":
            forms.append(("interleaved grid", dict(table=False, interleave=True)))
            if d == 4:
                forms.append(("interleaved grid, variant 81 (forced: 2 CTAs/SM, ~250 registers)", dict(table=False, interleave=True)))
                forms.append(("interleaved grid, variant 82 (forced: 3 CTAs/SM, 168 registers)", dict(table=False, interleave=True)))
        forms.append(("node table", dict(table="nodes")))
        for name, kw in forms:
            obj = cls(rows, "quiet", mode=mode, **kw)
            obj._lib.arb_set_query_variant(81 if "variant 81" in name else (82 if "variant 82" in name else 0))
            mem = (obj._nodes if obj._nodes is not None else obj._packed if obj._packed is not None else obj._planes).numel() * 8 / 1e9
            line = [f"{k}: {rate(obj, qq):.3e} q/s (x{rate(obj, qq) / base[k]:.2f} of the cell table)"
                    for k, qq in (("uniform random", q), ("cell-sorted", qs))]
            print(f"[tablefree] {grid} {mode} {name} ({mem:.3f} GB vs {tgb:.2f} GB of cell table): " + " | ".join(line), flush=True)
            obj._lib.arb_set_query_variant(0)
            obj.release()
            del obj
            torch.cuda.empty_cache()
        print(f"[tablefree] {grid} {mode} cell table: " + " | ".join(f"{k}: {v:.3e} q/s" for k, v in base.items()), flush=True)
        del rows
