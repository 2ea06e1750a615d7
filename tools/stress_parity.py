#!/usr/bin/env python
"""Randomised parity stress: random grid shapes (down to a single cell per axis), spacings, origins, modes,
kernel variants, build variants, table / table-free, against the CPU oracle.  Indices and NaN masks exact,
values within 1e-12 scaled.  (Checker use of oracle/ -- this is a test tool, not a product path.)"""
import argparse, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import _lib, quadcubic, tricubic
from oracle.arb_oracle import OracleInterp

ap = argparse.ArgumentParser(); ap.add_argument("--cases", type=int, default=150); ap.add_argument("--seed", type=int, default=0)
a = ap.parse_args()
rng = np.random.default_rng(a.seed)
lib = _lib.load()
worst, fails = 0.0, 0
for case in range(a.cases):
    d = int(rng.choice([3, 4]))
    shape = [int(rng.integers(4, 22 if d == 3 else 11)) for _ in range(d)]
    origin = rng.uniform(-3, 3, d); step = rng.choice([1e-3, 0.05, 0.37, 1.0, 40.0], d)
    axes = [origin[i] + step[i] * np.arange(shape[i]) for i in range(d)]
    mesh = np.meshgrid(*reversed(axes), indexing="ij"); coords = [m.ravel() for m in reversed(mesh)]
    scalar = bool(rng.integers(0, 2))
    phase = rng.uniform(0, 6, 3)
    arg = sum(c / (s * n) * 3.0 for c, s, n in zip(coords, step, shape))
    vals = [np.sin(arg + phase[0])] if scalar else [np.sin(arg + phase[0]), np.cos(1.7 * arg + phase[1]) + 0.1 * arg, np.sin(0.6 * arg + phase[2]) * arg]
    field = np.stack(coords + vals, axis=1)[rng.permutation(len(coords[0]))]
    mode = "scalar" if scalar else str(rng.choice(["vector", "norm", "both"]))
    form = int(rng.integers(0, 6))                 # 0-2 cell table, 3 table-free, 4 table-free planes (round-1 kernels), 5 node table
    table_free = form in (3, 4)
    bv = int(rng.integers(0, 10))
    qv = int(rng.choice([0, 1, 2, 10, 11, 20, 21, 22, 23, 30, 24, 25, 40, 41, 42, 43, 44, 45, 50, 51, 60, 61, 62, 63, 71, 72, 73, 81, 82]))
    kw = {} if scalar else {"mode": mode}
    if table_free:
        kw["table"] = False
        if form == 4:
            kw["interleave"] = False
    if form == 5:
        kw["table"] = "nodes"
        if d == 4 and rng.integers(0, 4) == 0:
            kw["fixed_d4"] = True
    lib.arb_set_build_variant(bv); lib.arb_set_query_variant(qv)
    try:
        obj = (tricubic if d == 3 else quadcubic)(field.copy(), "quiet", **kw)
        ora = OracleInterp(field, d, mode="vector" if scalar else mode, reference_quirk=not kw.get("fixed_d4", False))
        n = int(rng.integers(1, 20000))
        lo = np.array(ora.geo.int_min); hi = np.array(ora.geo.int_max)
        q = lo + rng.uniform(-0.05, 1.05, (n, d + int(rng.integers(0, 3)))) [:, :d + 2][:, :] * 1.0 if False else None
        ncol = d + int(rng.integers(0, 3))
        q = rng.uniform(0, 1, (n, ncol)); q[:, :d] = lo + rng.uniform(-0.03, 1.03, (n, d)) * (hi - lo)
        edge = np.isclose(q[:, :d], hi, rtol=0, atol=0).any(axis=1)          # exact upper edge: declared deviation
        q[edge, 0] = lo[0]
        if n > 10:
            q[rng.integers(0, n, 3), rng.integers(0, d, 3)] = [np.nan, np.inf, -np.inf]
        q_ref, q_gpu = q.copy(), q.copy()
        with np.errstate(invalid="ignore"):
            ref = ora.query(q_ref)
        got = obj.Query(q_gpu)
        ref = ref if isinstance(ref, tuple) else (ref,); got = got if isinstance(got, tuple) else (got,)
        s = max(float(np.abs(v).max()) for v in ora.values.values())
        ok = np.array_equal(q_ref, q_gpu, equal_nan=True) and np.array_equal(obj.queryInds, ora.query_inds)
        err = 0.0
        for i, (g_, r_) in enumerate(zip(got, ref)):
            ok &= np.array_equal(np.isnan(g_), np.isnan(r_))
            is_grad = (mode != "vector") and i == len(ref) - 1
            sc = s / np.array(ora.geo.h)[None, :] if is_grad else s
            m = ~np.isnan(r_)
            if m.any():
                err = max(err, float(np.max((np.abs(g_ - r_) / np.maximum(np.abs(r_), np.broadcast_to(sc, r_.shape)))[m])))
        worst = max(worst, err)
        if not ok or err > 1e-12:
            fails += 1
            print(f"[stress] FAIL case {case}: d={d} shape={shape} mode={mode} form={form} kw={kw} bv={bv} qv={qv} n={n} ok={ok} err={err:.3e}", flush=True)
    finally:
        lib.arb_set_build_variant(0); lib.arb_set_query_variant(0)
print(f"[stress] {a.cases} random cases, {fails} failures, worst scaled error {worst:.3e}")
sys.exit(1 if fails else 0)
