#!/usr/bin/env python
"""Small workload touching every kernel (3-D/4-D build, all query modes and variants, table-free path,
host pipeline) for compute-sanitizer runs.  Sizes are tiny: the sanitizer slows kernels 10-100x."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import _lib, quadcubic, tricubic

rng = np.random.default_rng(0)
x = np.linspace(-1, 1, 11); y = np.linspace(0, 1, 10); z = np.linspace(-1, 0, 9); t = np.linspace(0, 1, 7)
Z, Y, X = [a.ravel() for a in np.meshgrid(z, y, x, indexing="ij")]
f3 = np.stack([X, Y, Z, np.sin(X) * Y, X * Z, np.cos(Y + Z)], 1)
T, Z4, Y4, X4 = [a.ravel() for a in np.meshgrid(t, z[:7], y[:8], x[:9], indexing="ij")]
f4 = np.stack([X4, Y4, Z4, T, np.sin(X4) * T, X4 * Z4, np.cos(Y4 + T)], 1)
lib = _lib.load()
for mode in ("vector", "norm", "both"):
    for table in (True, False):
        o = tricubic(f3, "quiet", mode=mode, table=table)
        q = np.stack([rng.uniform(o.xIntMin, o.xIntMax, 700), rng.uniform(o.yIntMin, o.yIntMax * 1.1, 700),
                      rng.uniform(o.zIntMin, o.zIntMax, 700)], 1)
        for v in ((0, 1, 2, 10, 20) if table else (0,)):
            lib.arb_set_query_variant(v)
            o.Query(q.copy())
        lib.arb_set_query_variant(0)
    for bv in (5, 7, 8, 9):                       # every formulation of the build (default 0 is used below)
        lib.arb_set_build_variant(bv)
        tricubic(f3, "quiet", mode=mode); quadcubic(f4, "quiet", mode=mode)
    lib.arb_set_build_variant(0)
    o4 = quadcubic(f4, "quiet", mode=mode)
    q4 = np.stack([rng.uniform(o4.xIntMin, o4.xIntMax, 300), rng.uniform(o4.yIntMin, o4.yIntMax, 300),
                   rng.uniform(o4.zIntMin, o4.zIntMax, 300), rng.uniform(o4.tIntMin, o4.tIntMax * 1.1, 300)], 1)
    for v in (0, 10):
        lib.arb_set_query_variant(v)
        o4.Query(q4.copy())
    lib.arb_set_query_variant(0)
    for fixed in (False, True):
        quadcubic(f4, "quiet", mode=mode, table=False, fixed_d4=fixed).Query(q4.copy())
        if mode != "norm":                        # the interleaved 4-D grid is the default; the per-component boxes too
            quadcubic(f4, "quiet", mode=mode, table=False, fixed_d4=fixed, interleave=False).Query(q4.copy())
# round 2: compact slot rings / prefetch / probe / LDGSTS variants of the cell kernel, node tables (3-D paired and
# interleaved, 4-D with the A.py:860 term), interleaved table-free grid
for mode in ("vector", "norm", "both"):
    o = tricubic(f3, "quiet", mode=mode)
    o4 = quadcubic(f4, "quiet", mode=mode)
    qs = np.sort(q, axis=0)                        # clustered rows: the compact rings take one pass
    for v in (24, 25, 40, 42, 43, 50, 60, 62):
        lib.arb_set_query_variant(v)
        o.Query(q.copy()); o.Query(qs.copy()); o4.Query(q4.copy())
    lib.arb_set_query_variant(0)
    import torch as _t
    big = _t.from_numpy(np.tile(qs, (100, 1))).cuda()          # >= 2^16 rows: the probed launch
    o.Query(big)
    n3 = tricubic(f3, "quiet", mode=mode, table="nodes")
    n4 = quadcubic(f4, "quiet", mode=mode, table="nodes")
    n4f = quadcubic(f4, "quiet", mode=mode, table="nodes", fixed_d4=True)
    for v in (0, 71, 72, 73):
        lib.arb_set_query_variant(v)
        n3.Query(q.copy()); n4.Query(q4.copy()); n4f.Query(q4.copy())
    lib.arb_set_query_variant(0)
    n3.Query(q[:40].copy()); n3.update_values(f3[:, 3:] * 0.5)
# zero-copy path (<= 256 rows), single point, resumable push on slab tables, update_values
import torch
o = tricubic(f3, "quiet", mode="both")
o.Query(q[:40].copy()); o.Query(q[0].copy())
o4 = quadcubic(f4, "quiet", mode="both")
o4.Query(q4[:8].copy())
o4.update_values(f4[:, 4:] * 1.5)
for cls, f, d in ((tricubic, f3, 3), (quadcubic, f4, 4)):
    whole = cls(f[:, :d + 1].copy(), "quiet")
    ns = whole._geo.ncell[d - 1]
    parts = [cls(f[:, :d + 1].copy(), "quiet", slab=(0, ns // 2)), cls(f[:, :d + 1].copy(), "quiet", slab=(ns // 2, ns))]
    lo = np.array(whole._geo.int_min); hi = np.array(whole._geo.int_max)
    p = torch.from_numpy(lo + rng.uniform(0.1, 0.9, (500, d)) * (hi - lo)).cuda()
    v = torch.from_numpy(rng.normal(0, 0.3, (500, 3))).cuda()
    step = torch.zeros(500, dtype=torch.int64, device="cuda")
    for _ in range(6):
        for part in parts:
            part._push_local(p, v, step, 0.02, 12, -0.5, (0.0, 0.0, 0.2))
    whole.push(p.clone(), v.clone(), 0.02, 5, -0.5)
    # push on node tables: one component, 3-D interleaved ('both'), 4-D with and without the A.py:860 term
    cls(f[:, :d + 1].copy(), "quiet", table="nodes").push(p.clone(), v.clone(), 0.02, 5, -0.5)
    cls(f, "quiet", mode="both", table="nodes").push(p.clone(), v.clone(), 0.02, 5, -0.5)
    if d == 4:
        cls(f, "quiet", mode="both", table="nodes", fixed_d4=True).push(p.clone(), v.clone(), 0.02, 5, -0.5)
print("sanitize target done")
