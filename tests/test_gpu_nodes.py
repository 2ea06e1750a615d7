"""GPU tier (-m gpu): the node (Hermite) table -- ``tricubic / quadcubic(..., table='nodes')``,
csrc/arb_nodes.cuh + the KIND_NODES / LDGSTS form of query_block_kernel -- against the golden vectors of the live
reference, the numpy oracle, and the cell-table path; and the new query-kernel variants (compact slot rings,
coordinate prefetch, LDGSTS fetch, sortedness probe) against the golden vectors.

Same bar as tests/test_gpu_parity.py: indices and NaN masks exact, values within 1e-12 scaled."""
import numpy as np
import pytest

from conftest import load_golden
from test_gpu_parity import (CASES, _analytic_field3, _analytic_field4, _check_outputs, _cls, _golden_ref, _scales,
                             _uniform_queries)

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.mark.parametrize("name,d,modes", CASES)
def test_node_table_golden(name, d, modes):
    """Golden vectors of the live reference (in-volume, out-of-volume, NaN, +-inf rows, extra query columns); the 4-D
    cases carry an xyzt term, so the A.py:860 correction of the node path is exercised."""
    g = load_golden(name)
    for mode in modes:
        kw = {} if mode == "scalar" else {"mode": mode}
        obj = _cls(d)(g["field"].copy(), "quiet", table="nodes", **kw)
        q = g[mode + "_q_in"].copy()
        res = obj.Query(q)
        _check_outputs(res, _golden_ref(g, mode), mode, g["field"], d, g["h"], f"{name}/{mode}/nodes")
        assert np.array_equal(q, g[mode + "_q_after"], equal_nan=True)
        assert np.array_equal(obj.queryInds, g[mode + "_inds"])
        with pytest.raises(AttributeError):
            obj.table
        npts = [int(v) + 3 for v in g["ncell_axis"]]
        ncomp = {"vector": 3, "norm": 1, "both": 4, "scalar": 1}[mode]
        if d == 3 and ncomp >= 3:
            assert list(obj.nodes.shape) == [npts[2] - 2, npts[1] - 2, npts[0] - 2, 4, 8]
        elif d == 3:
            assert list(obj.nodes.shape) == [ncomp, npts[2] - 2, npts[1] - 2, npts[0] - 3, 2, 8]
        else:
            assert list(obj.nodes.shape) == [ncomp] + [n - 2 for n in reversed(npts)] + [16]
    with pytest.raises(ValueError):
        _cls(d)(g["field"].copy(), "quiet", table="nodes", slab=(0, 1))
    with pytest.raises(ValueError):
        _cls(d)(g["field"].copy(), "quiet", table="cells")


def test_node_values_are_the_rows_of_D():
    """The stored node values are D applied at the node (A.py:129-173): f, fx = (f[+1] - f[-1]) / 2, fxy, ... in
    tau = tx + 2 ty + 4 tz order."""
    from arbinterp_b200 import tricubic
    rng = np.random.default_rng(5)
    nx, ny, nz = 9, 7, 8
    field = _analytic_field3(nx, ny, nz)
    vals = rng.standard_normal(nx * ny * nz)
    field = np.concatenate([field[:, :3], vals[:, None]], axis=1)
    obj = tricubic(field, "quiet", table="nodes")
    pairs = obj.nodes.cpu().numpy()[0]                     # [nz-2][ny-2][nx-3][2][8]: pair i = nodes i, i + 1
    assert np.array_equal(pairs[:, :, 1:, 0], pairs[:, :, :-1, 1])
    got = np.concatenate([pairs[:, :, :, 0], pairs[:, :, -1:, 1]], axis=2)          # [nz-2][ny-2][nx-2][8]
    f = vals.reshape(nz, ny, nx)
    c = (slice(1, -1),) * 3
    dx = lambda a: 0.5 * (a[:, :, 2:] - a[:, :, :-2])
    dy = lambda a: 0.5 * (a[:, 2:, :] - a[:, :-2, :])
    dz = lambda a: 0.5 * (a[2:, :, :] - a[:-2, :, :])
    want = {0: f[c], 1: dx(f)[1:-1, 1:-1], 2: dy(f)[1:-1, :, 1:-1], 3: dx(dy(f))[1:-1], 4: dz(f)[:, 1:-1, 1:-1],
            5: dx(dz(f))[:, 1:-1], 6: dy(dz(f))[:, :, 1:-1], 7: dx(dy(dz(f)))}
    for tau, w in want.items():
        assert np.allclose(got[..., tau], w, rtol=0, atol=1e-14), tau


@pytest.mark.parametrize("mode", ["vector", "norm", "both"])
@pytest.mark.parametrize("shape", [(37, 26, 23), (20, 41, 17)])
def test_node_table_oracle_parity_3d(mode, shape):
    from arbinterp_b200 import tricubic
    from oracle.arb_oracle import OracleInterp
    rng = np.random.default_rng(4321)
    field = _analytic_field3(*shape, rng=rng)
    obj = tricubic(field.copy(), "quiet", mode=mode, table="nodes")
    q = _uniform_queries(obj, 3, 100_000, rng, extra=2)
    q[::97, 1] = 5.0
    q[::1013, 2] = np.nan
    ora = OracleInterp(field, 3, mode=mode)
    q_ref = q.copy()
    ref = ora.query(q_ref)
    ref = ref if isinstance(ref, tuple) else (ref,)
    q_gpu = q.copy()
    res = obj.Query(q_gpu)
    worst = _check_outputs(res, ref, mode, field, 3, ora.geo.h, f"nodes 3d {shape} {mode}")
    assert np.array_equal(q_gpu, q_ref, equal_nan=True)
    assert np.array_equal(obj.queryInds, ora.query_inds)
    assert worst < 1e-13
    # device tensors and the host path agree bit for bit
    dev = obj.Query(torch.from_numpy(q.copy()).cuda())
    dev = dev if isinstance(dev, tuple) else (dev,)
    res = res if isinstance(res, tuple) else (res,)
    for x, y in zip(res, dev):
        assert np.array_equal(x, y.cpu().numpy(), equal_nan=True)


@pytest.mark.parametrize("fixed", [False, True])
@pytest.mark.parametrize("mode", ["vector", "norm", "both"])
def test_node_table_quadcubic_matches_cell_table(mode, fixed):
    """4-D node path against the cell-table path on a field with an xyzt monomial: quirk and fixed_d4 answers differ
    by ~1e-6, each must match its own table; and the quirk one must match the oracle."""
    from arbinterp_b200 import quadcubic
    rng = np.random.default_rng(31)
    field = _analytic_field4(13, 9, 8, 7, rng=rng)
    a = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=fixed)
    b = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=fixed, table="nodes")
    q = _uniform_queries(a, 4, 50_000, rng, extra=1)
    q[::53, 3] = -1.0
    qa, qb = q.copy(), q.copy()
    ra, rb = a.Query(qa), b.Query(qb)
    h = [a.hx, a.hy, a.hz, a.ht]
    ra = ra if isinstance(ra, tuple) else (ra,)
    _check_outputs(rb, ra, mode, field, 4, h, f"nodes vs cells 4d {mode} fixed={fixed}")
    assert np.array_equal(qa, qb, equal_nan=True) and np.array_equal(a.queryInds, b.queryInds)
    if not fixed:
        from oracle.arb_oracle import OracleInterp
        ora = OracleInterp(field, 4, mode=mode)
        ref = ora.query(q[:10_000].copy())
        ref = ref if isinstance(ref, tuple) else (ref,)
        got = b.Query(q[:10_000].copy())
        _check_outputs(got, ref, mode, field, 4, ora.geo.h, f"nodes vs oracle 4d {mode}")
        assert np.array_equal(b.queryInds, ora.query_inds)


@pytest.mark.parametrize("mode", ["vector", "both"])
def test_table_free_interleaved_grid(mode):
    """Table-free 3-D 'vector' / 'both' on the component-interleaved grid (arb_query_gridil, the default) against the
    per-component TMA-box kernel (interleave=False), the cell table and the oracle; odd nx, NaN / out-of-volume rows,
    device tensors, update_values."""
    from arbinterp_b200 import tricubic
    from oracle.arb_oracle import OracleInterp
    rng = np.random.default_rng(707)
    field = _analytic_field3(37, 26, 23, rng=rng)
    cells = tricubic(field.copy(), "quiet", mode=mode)
    il = tricubic(field.copy(), "quiet", mode=mode, table=False)
    pl = tricubic(field.copy(), "quiet", mode=mode, table=False, interleave=False)
    assert il._packed is not None and pl._packed is None
    q = _uniform_queries(cells, 3, 150_000, rng, extra=1)
    q[::101, 0] = -9.0
    q[::977, 2] = np.nan
    h = [cells.hx, cells.hy, cells.hz]
    outs = {}
    for name, obj in (("cells", cells), ("il", il), ("pl", pl)):
        qq = q.copy()
        r = obj.Query(qq)
        outs[name] = (r if isinstance(r, tuple) else (r,), qq, obj.queryInds)
    for name in ("il", "pl"):
        _check_outputs(outs[name][0], outs["cells"][0], mode, field, 3, h, f"table-free {name} vs cells {mode}")
        assert np.array_equal(outs[name][1], outs["cells"][1], equal_nan=True)
        assert np.array_equal(outs[name][2], outs["cells"][2])
    ora = OracleInterp(field, 3, mode=mode)
    ref = ora.query(q[:20_000].copy())
    ref = ref if isinstance(ref, tuple) else (ref,)
    _check_outputs(il.Query(q[:20_000].copy()), ref, mode, field, 3, ora.geo.h, f"table-free interleaved vs oracle {mode}")
    dev = il.Query(torch.from_numpy(q.copy()).cuda())
    dev = dev if isinstance(dev, tuple) else (dev,)
    for x, y in zip(outs["il"][0], dev):
        assert np.array_equal(x, y.cpu().numpy(), equal_nan=True)
    new_vals = field[:, 3:] * 0.5 - 0.125
    il.update_values(new_vals)
    fresh = tricubic(np.concatenate([field[:, :3], new_vals], axis=1), "quiet", mode=mode, table=False)
    a, b = il.Query(q.copy()), fresh.Query(q.copy())
    for x, y in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("fixed", [False, True])
@pytest.mark.parametrize("mode", ["vector", "both"])
def test_table_free_interleaved_grid_4d(mode, fixed):
    """Table-free 4-D 'vector' / 'both' on the component-interleaved grid [nt][nz][ny][nx][4] (query_gridil4_kernel:
    four lanes = the z-planes of a query, four passes = its t-planes, A.py:860 term from parity sums exchanged between
    the lanes) against the cell table, the per-component TMA-box kernel and -- with the quirk -- the oracle; the field
    has an xyzt monomial, so the quirk and fixed answers differ by ~1e-6.  Odd nx, clustered rows (shared slots),
    NaN / out-of-volume rows, device tensors, update_values."""
    from arbinterp_b200 import quadcubic
    rng = np.random.default_rng(4242)
    field = _analytic_field4(13, 9, 8, 7, rng=rng)
    cells = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=fixed)
    il = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=fixed, table=False, interleave=True)
    pl = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=fixed, table=False, interleave=False)
    assert il._packed is not None and pl._packed is None and tuple(il._packed.shape) == (7, 8, 9, 13, 4)
    assert quadcubic(field.copy(), "quiet", mode=mode, table=False)._packed is not None       # the default form
    q = _uniform_queries(cells, 4, 120_000, rng, extra=1)
    q[40_000:80_000] = q[40_000:40_400].repeat(100, axis=0)      # bunches of rows in the same cells: shared slots
    q[40_000:80_000, :4] += rng.uniform(0, 1e-4, (40_000, 4)) * np.array([cells.hx, cells.hy, cells.hz, cells.ht])
    q[::101, 0] = -9.0
    q[::977, 3] = np.nan
    q[::53, 3] = -1.0
    h = [cells.hx, cells.hy, cells.hz, cells.ht]
    outs = {}
    for name, obj in (("cells", cells), ("il", il), ("pl", pl)):
        qq = q.copy()
        r = obj.Query(qq)
        outs[name] = (r if isinstance(r, tuple) else (r,), qq, obj.queryInds)
    for name in ("il", "pl"):
        _check_outputs(outs[name][0], outs["cells"][0], mode, field, 4, h, f"4-D table-free {name} vs cells {mode} fixed={fixed}")
        assert np.array_equal(outs[name][1], outs["cells"][1], equal_nan=True)
        assert np.array_equal(outs[name][2], outs["cells"][2])
    if not fixed:
        from oracle.arb_oracle import OracleInterp
        ora = OracleInterp(field, 4, mode=mode)
        ref = ora.query(q[:10_000].copy())
        ref = ref if isinstance(ref, tuple) else (ref,)
        _check_outputs(il.Query(q[:10_000].copy()), ref, mode, field, 4, ora.geo.h, f"4-D interleaved vs oracle {mode}")
        assert np.array_equal(il.queryInds, ora.query_inds)
    dev = il.Query(torch.from_numpy(q.copy()).cuda())
    dev = dev if isinstance(dev, tuple) else (dev,)
    for x, y in zip(outs["il"][0], dev):
        assert np.array_equal(x, y.cpu().numpy(), equal_nan=True)
    new_vals = field[:, 4:] * 0.5 - 0.125
    il.update_values(new_vals)
    fresh = quadcubic(np.concatenate([field[:, :4], new_vals], axis=1), "quiet", mode=mode, fixed_d4=fixed, table=False,
                      interleave=True)
    a, b = il.Query(q.copy()), fresh.Query(q.copy())
    for x, y in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("name,d,mode", [("tri_12x10x9", 3, "both"), ("tri_12x10x9", 3, "norm"), ("quad_8x7x7x6", 4, "both")])
def test_node_table_save_and_load(name, d, mode, tmp_path):
    """save() / load() of a node table: the loaded interpolator answers bit-identically with no field, ingest or build;
    a file whose header does not match its node layout is refused."""
    g = load_golden(name)
    obj = _cls(d)(g["field"].copy(), "quiet", mode=mode, table="nodes")
    q = g[mode + "_q_in"].copy()
    ref = obj.Query(q.copy())
    path = tmp_path / "nodes.arb"
    obj.save(str(path), chunk_bytes=1 << 14)
    back = _cls(d).load(str(path), chunk_bytes=1 << 13)
    got = back.Query(q.copy())
    for a, b in zip(ref if isinstance(ref, tuple) else (ref,), got if isinstance(got, tuple) else (got,)):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(back.queryInds, obj.queryInds) and torch.equal(back.nodes, obj.nodes)
    with pytest.raises(AttributeError):
        back.table
    blob = bytearray(path.read_bytes())
    i = blob.find(b'"kind": "nodes"')
    blob[i:i + 15] = b'"kind": "cells"'
    bad = tmp_path / "bad.arb"
    bad.write_bytes(bytes(blob))
    with pytest.raises(ValueError):
        _cls(d).load(str(bad))


def test_node_table_quirk_is_visible():
    from arbinterp_b200 import quadcubic
    field = _analytic_field4(9, 8, 8, 7)[:, [0, 1, 2, 3, 5]]          # the column with the x*y*z*t monomial
    a = quadcubic(field.copy(), "quiet", table="nodes")
    b = quadcubic(field.copy(), "quiet", table="nodes", fixed_d4=True)
    rng = np.random.default_rng(3)
    q = _uniform_queries(a, 4, 5000, rng)
    va, _ = a.Query(q.copy())
    vb, _ = b.Query(q.copy())
    assert 1e-9 < np.abs(va - vb).max() < 1e-2


@pytest.mark.parametrize("d", [3, 4])
def test_node_table_update_values_and_small_batches(d):
    """update_values() rebuilds the node table in place; single points, the zero-copy path (<= 256 rows) and the
    latency path (<= 8192 rows) go through arb_query_nodes_host."""
    rng = np.random.default_rng(8)
    field = _analytic_field3(11, 9, 10, rng=rng) if d == 3 else _analytic_field4(8, 7, 6, 7, rng=rng)
    obj = _cls(d)(field.copy(), "quiet", mode="both", table="nodes")
    ref_obj = _cls(d)(field.copy(), "quiet", mode="both")
    q = _uniform_queries(obj, d, 3000, rng)
    h = [getattr(obj, "h" + c) for c in "xyzt"[:d]]
    for n in (1, 20, 300, 3000):
        _check_outputs(obj.Query(q[:n].copy()), ref_obj.Query(q[:n].copy()), "both", field, d, h, f"nodes small {n}")
    one = obj.Query(q[0].copy())
    want = ref_obj.Query(q[0].copy())
    assert np.allclose(one[0], want[0], rtol=1e-12) and np.isclose(one[1], want[1], rtol=1e-12)
    new_vals = field[:, d:] * 1.5 + 0.25
    obj.update_values(new_vals)
    fresh = _cls(d)(np.concatenate([field[:, :d], new_vals], axis=1), "quiet", mode="both", table="nodes")
    for x, y in zip(obj.Query(q.copy()), fresh.Query(q.copy())):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("variant", [24, 40, 41, 42, 43, 44, 45, 50, 51, 60, 61, 62, 63])
@pytest.mark.parametrize("name,d,modes", CASES)
def test_round2_kernel_variants_golden(name, d, modes, variant, cuda_lib):
    """Compact slot rings (16 / 12 / 8 slots per warp, multi-pass when a warp item needs more blocks), coordinate
    prefetch, the probed launch and the LDGSTS fetch answer like the default kernel."""
    g = load_golden(name)
    old = cuda_lib.arb_set_query_variant(variant)
    try:
        for mode in modes:
            kw = {} if mode == "scalar" else {"mode": mode}
            obj = _cls(d)(g["field"].copy(), "quiet", **kw)
            q = g[mode + "_q_in"].copy()
            res = obj.Query(q)
            _check_outputs(res, _golden_ref(g, mode), mode, g["field"], d, g["h"], f"{name}/{mode}/v{variant}")
            assert np.array_equal(q, g[mode + "_q_after"], equal_nan=True)
            assert np.array_equal(obj.queryInds, g[mode + "_inds"])
    finally:
        cuda_lib.arb_set_query_variant(old)


@pytest.mark.parametrize("variant", [40, 42, 50, 62, 71, 72, 73])
def test_round2_variants_bit_identical_on_large_batches(variant, cuda_lib):
    """2^20 rows (random, cell-sorted, bunched; masked and NaN rows mixed in): every variant of the cell-table kernel
    does the same arithmetic as the default one, so outputs and indices are equal bit for bit -- whichever launch the
    sortedness probe picks.  Variants 7x select the node-table kernel's forms and are compared with its default."""
    from arbinterp_b200 import tricubic
    rng = np.random.default_rng(17)
    field = _analytic_field3(45, 38, 41, rng=rng)
    obj = tricubic(field.copy(), "quiet", mode="both", **({"table": "nodes"} if variant >= 70 else {}))
    n = 1 << 20
    q = torch.from_numpy(_uniform_queries(obj, 3, n, rng)).cuda()
    q[::1001] = 9.0
    q[5::4099, 1] = float("nan")
    obj.Query(q.clone())
    order = torch.argsort(obj._last_cells)
    centre = q[:64].mean(dim=0)
    bunch = (centre + 0.05 * torch.randn(n, 3, dtype=torch.float64, device="cuda") * float(obj.hx)).contiguous()
    for name, qq in (("random", q), ("sorted", q[order].contiguous()), ("bunch", bunch)):
        base = obj.Query(qq.clone())
        base_cells = obj._last_cells.clone()
        old = cuda_lib.arb_set_query_variant(variant)
        try:
            got = obj.Query(qq.clone())
        finally:
            cuda_lib.arb_set_query_variant(old)
        assert torch.equal(obj._last_cells, base_cells), name
        for a, b in zip(got, base):
            assert torch.equal(a.view(torch.int64), b.view(torch.int64)), (name, variant)


@pytest.mark.parametrize("d", [3, 4])
def test_owner_keys_kernel_matches_torch(d, cuda_lib):
    """arb_owner_keys (owner rank + out-of-volume mask in one kernel) against sharding.owner_ranks and the torch mask,
    on rows inside, outside, exactly on the bounds and NaN / inf."""
    import ctypes
    from arbinterp_b200 import _lib
    from arbinterp_b200.sharding import owner_ranks, plan_slabs
    rng = np.random.default_rng(d)
    field = _analytic_field3(9, 8, 21, rng=rng) if d == 3 else _analytic_field4(7, 6, 5, 23, rng=rng)
    obj = _cls(d)(field, "quiet")
    g = obj._geo
    lo, hi = np.array(g.int_min), np.array(g.int_max)
    q = lo + rng.uniform(-0.1, 1.1, (50_000, d + 1))[:, :d] * (hi - lo)
    q[0], q[1] = lo, hi
    q[2, d - 1] = hi[d - 1] + 1e-300
    q[3:40:3, d - 1] = lo[d - 1] + np.arange(13) * g.h[d - 1]          # exactly on layer boundaries
    q[100, 0] = np.nan; q[101, d - 1] = np.nan; q[102, d - 1] = np.inf; q[103, 1] = -np.inf
    qt = torch.from_numpy(q).cuda()
    for world in (1, 2, 5, 8):
        slabs = plan_slabs(g.ncell[d - 1], world)
        # expected: the reference's cell location with a true division (A.py:1081-1086), as the query kernel computes it
        slow = q[:, d - 1]
        with np.errstate(invalid="ignore"):
            valid = (slow >= lo[d - 1]) & (slow <= hi[d - 1])
            layer = np.clip(np.where(valid, np.floor((slow - lo[d - 1]) / g.h[d - 1]), 0), 0, g.ncell[d - 1] - 1)
        want = np.where(valid, np.searchsorted([s[1] for s in slabs], layer, side="right"), 0)
        want_owner = torch.from_numpy(want).cuda()
        assert torch.equal(owner_ranks(qt[:, d - 1], g.int_min[d - 1], g.int_max[d - 1], g.h[d - 1], slabs).to(torch.int64),
                           want_owner.to(torch.int64)), world
        tl = torch.tensor(g.int_min, dtype=torch.float64, device="cuda")
        th = torch.tensor(g.int_max, dtype=torch.float64, device="cuda")
        want_out = ((qt < tl) | (qt > th)).any(dim=1)
        owner = torch.empty(len(q), dtype=torch.int16, device="cuda")
        outside = torch.empty(len(q), dtype=torch.bool, device="cuda")
        his = (ctypes.c_int64 * world)(*[s[1] for s in slabs])
        _lib.check(cuda_lib.arb_owner_keys(ctypes.byref(obj._cgeom), qt.data_ptr(), len(q), d, his, world, owner.data_ptr(),
                                           outside.data_ptr(), torch.cuda.current_stream().cuda_stream), "arb_owner_keys")
        assert torch.equal(owner.to(torch.int64), want_owner.to(torch.int64)), world
        assert torch.equal(outside, want_out), world


@pytest.mark.parametrize("d,mode,fixed", [(3, "norm", False), (3, "both", False), (4, "norm", False), (4, "both", False),
                                           (4, "both", True)])
def test_push_on_node_table_matches_cell_table(d, mode, fixed):
    """The fused query + push integrator on a node table (arb_push_nodes) follows the cell-table push to round-off:
    same lost particles, positions and velocities within 1e-9 of the box size after 40 steps; 3-D one-component and
    interleaved layouts, 4-D with the A.py:860 term and with fixed_d4."""
    rng = np.random.default_rng(50 + d)
    field = _analytic_field3(16, 14, 15, rng=rng) if d == 3 else _analytic_field4(10, 9, 8, 12, rng=rng)
    kw = {"fixed_d4": True} if fixed else {}
    cells = _cls(d)(field.copy(), "quiet", mode=mode, **kw)
    nodes = _cls(d)(field.copy(), "quiet", mode=mode, table="nodes", **kw)
    lo, hi = np.array(cells._geo.int_min), np.array(cells._geo.int_max)
    n = 5000
    pos = lo + rng.uniform(0.2, 0.8, (n, d)) * (hi - lo)
    pos[:50] = lo + rng.uniform(0.0, 0.02, (50, d)) * (hi - lo)          # start at the edge, moving out: lost in both
    vel = rng.normal(0, 0.05, (n, 3)) * (hi - lo)[:3]
    vel[:50] = -np.abs(vel[:50]) - 0.05 * (hi - lo)[:3]
    if d == 4:
        pos[:, 3] = lo[3] + rng.uniform(0.05, 0.3, n) * (hi[3] - lo[3])
    dt = 0.01 if d == 3 else 0.01 * (hi[3] - lo[3])
    pa, va = torch.from_numpy(pos.copy()).cuda(), torch.from_numpy(vel.copy()).cuda()
    pb, vb = pa.clone(), va.clone()
    la = cells.push(pa, va, dt, 40, -0.3, gravity=(0.0, 0.0, -0.05))
    lb = nodes.push(pb, vb, dt, 40, -0.3, gravity=(0.0, 0.0, -0.05))
    assert la == lb and la >= 50
    pa, pb, va, vb = pa.cpu().numpy(), pb.cpu().numpy(), va.cpu().numpy(), vb.cpu().numpy()
    assert np.array_equal(np.isnan(pa), np.isnan(pb))
    ok = ~np.isnan(pa[:, 0])
    span = (hi - lo)
    assert np.abs((pa[ok] - pb[ok]) / span).max() < 1e-9
    assert np.abs(va[ok] - vb[ok]).max() < 1e-7 * np.abs(va[ok]).max()
    # numpy arrays in, numpy arrays out
    p_np, v_np = pos.copy(), vel.copy()
    assert nodes.push(p_np, v_np, dt, 40, -0.3, gravity=(0.0, 0.0, -0.05)) == lb
    assert np.array_equal(p_np, pb, equal_nan=True)


def test_node_table_beyond_64_gb(cuda_lib):
    """The LDGSTS gather passes a slot's source between lanes as a 32-bit count of 128-byte units, so node tables may
    be larger than the 64 GB a count of 16-byte units could address (they fit a 180 GB GPU long before a cell table
    does).  C-ABI level, no field rows: an 824^3 one-component grid made on the device, its 71 GB node table, and
    queries concentrated in the last z-layers (table offsets above 64 GB) checked against the table-free path on the
    same grid and against the analytic field, which both forms reproduce to round-off (per-axis quadratics)."""
    import ctypes
    from arbinterp_b200 import _lib
    torch.cuda.empty_cache()
    free, total = torch.cuda.mem_get_info(0)
    if free < 100 * (1 << 30):
        pytest.skip(f"needs ~80 GB of free device memory, {free >> 30} GB free")
    dev = torch.device("cuda", 0)
    n = 840
    ax = torch.linspace(-1.0, 1.0, n, dtype=torch.float64, device=dev)
    X, Y, Z = ax.view(1, 1, n), ax.view(1, n, 1), ax.view(n, 1, 1)
    f = lambda x, y, z: 1.0 + x * x - 0.5 * y * y + 0.25 * z * z + 0.3 * x * y * z + 0.2 * y * z * z
    planes = f(X, Y, Z).contiguous().view(1, n, n, n)
    nodes = torch.empty((1, n - 2, n - 2, n - 3, 2, 8), dtype=torch.float64, device=dev)
    assert nodes.numel() * 8 > 68 * (1 << 30)          # well beyond 2^32 units of 16 bytes
    npts = (ctypes.c_int64 * 4)(n, n, n, 1)
    st = torch.cuda.current_stream(dev).cuda_stream
    _lib.check(cuda_lib.arb_build_nodes(3, planes.data_ptr(), 1, ctypes.byref(npts), n, nodes.data_ptr(), st), "arb_build_nodes")
    g = _lib.ArbGeom()
    g.d, g.ncomp = 3, 1
    h = float(ax[1] - ax[0])
    for a in range(4):
        g.ncell[a] = n - 3 if a < 3 else 1
        g.int_min[a] = float(ax[1]) if a < 3 else 0.0
        g.int_max[a] = float(ax[n - 2]) if a < 3 else 0.0
        g.h[a] = h if a < 3 else 1.0
    g.slab_lo, g.slab_hi, g.flags = 0, n - 3, 0
    gen = torch.Generator(device=dev).manual_seed(5)
    N = 400_000
    lo, hi = float(ax[1]), float(ax[n - 2])
    q = lo + torch.rand((N, 3), generator=gen, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-9)
    q[N // 2:, 2] = hi - torch.rand(N - N // 2, generator=gen, dtype=torch.float64, device=dev) * 60 * h   # last z-layers
    outs = {}
    for name in ("nodes", "grid"):
        norm = torch.empty(N, dtype=torch.float64, device=dev)
        grad = torch.empty((N, 3), dtype=torch.float64, device=dev)
        cell = torch.empty(N, dtype=torch.int64, device=dev)
        qq = q.clone()
        if name == "nodes":
            rc = cuda_lib.arb_query_nodes(ctypes.byref(g), nodes.data_ptr(), _lib.MODE_NORM, qq.data_ptr(), N, 3, None,
                                          norm.data_ptr(), grad.data_ptr(), cell.data_ptr(), None, None, st)
        else:
            rc = cuda_lib.arb_query_grid(ctypes.byref(g), planes.data_ptr(), n, _lib.MODE_NORM, qq.data_ptr(), N, 3, None,
                                         norm.data_ptr(), grad.data_ptr(), cell.data_ptr(), None, None, st)
        _lib.check(rc, name)
        outs[name] = (norm, grad, cell)
    torch.cuda.synchronize()
    assert torch.equal(outs["nodes"][2], outs["grid"][2])
    top = outs["nodes"][2].max().item() // ((n - 3) * (n - 3))
    assert (top * (n - 2) * (n - 3)) * 128 > 64 * (1 << 30), "no query reached table offsets beyond 64 GB"
    want = f(q[:, 0], q[:, 1], q[:, 2])
    S = float(planes.abs().max())
    for name in ("nodes", "grid"):
        assert float((outs[name][0] - want).abs().max()) < 1e-12 * S, name
    assert float((outs["nodes"][0] - outs["grid"][0]).abs().max()) < 1e-12 * S
    assert float((outs["nodes"][1] - outs["grid"][1]).abs().max()) < 1e-12 * S / h
    del nodes, planes
    torch.cuda.empty_cache()
