"""GPU tier (-m gpu): the CUDA path, called through the drop-in classes (which go through the
C-ABI), against (i) the golden vectors of the live reference, (ii) the numpy oracle on seeded
inputs, (iii) size-independent analytic properties at BASELINE.json's full grid sizes.

Bar (north_star): cell indices and NaN masks bit-exact; values, gradients and components within
1e-12 relative, scaled as |got-ref| <= 1e-12 * max(|ref|, S) with S = max|field component|
(gradients: S/h) -- SURVEY 8d."""
import ctypes

import numpy as np
import pytest

from conftest import assert_parity, load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

CASES = [("tri_12x10x9", 3, ["vector", "norm", "both"]), ("tri_scalar_9x8x11", 3, ["scalar"]),
         ("quad_8x7x7x6", 4, ["vector", "norm", "both"]), ("quad_scalar_6x7x5x6", 4, ["scalar"])]
RTOL = 1e-12


def _cls(d):
    from arbinterp_b200 import tricubic, quadcubic
    return tricubic if d == 3 else quadcubic


def _scales(field, d, h):
    vals = field[:, d:]
    s_comp = np.abs(vals).max(axis=0) if vals.shape[1] == 3 else None
    s_norm = np.linalg.norm(vals, axis=1).max() if vals.shape[1] == 3 else np.abs(vals).max()
    return s_comp, s_norm, s_norm / np.asarray(h)


def _check_outputs(res, ref, mode, field, d, h, what):
    s_comp, s_norm, s_grad = _scales(field, d, h)
    res = res if isinstance(res, tuple) else (res,)
    assert len(res) == len(ref)
    worst = 0.0
    i = 0
    if mode in ("vector", "both"):
        worst = max(worst, assert_parity(res[i], ref[i], s_comp[None, :], RTOL, what + " comps")); i += 1
    if mode in ("norm", "both", "scalar"):
        worst = max(worst, assert_parity(res[i], ref[i], s_norm, RTOL, what + " norm")); i += 1
        worst = max(worst, assert_parity(res[i], ref[i], s_grad[None, :], RTOL, what + " grad"))
    return worst


def _golden_ref(g, mode):
    return tuple(g[f"{mode}_out{i}"] for i in range(3) if f"{mode}_out{i}" in g.files)


@pytest.mark.parametrize("name,d,modes", CASES)
def test_golden_range_queries(name, d, modes):
    g = load_golden(name)
    for mode in modes:
        kw = {} if mode == "scalar" else {"mode": mode}
        obj = _cls(d)(g["field"].copy(), "quiet", **kw)
        q = g[mode + "_q_in"].copy()
        res = obj.Query(q)
        _check_outputs(res, _golden_ref(g, mode), mode, g["field"], d, g["h"], f"{name}/{mode}")
        assert np.array_equal(q, g[mode + "_q_after"], equal_nan=True), "in-place NaN rows (A.py:350-355)"
        assert np.array_equal(obj.queryInds, g[mode + "_inds"]), "cell indices must be bit-exact"
        assert obj.nc == int(g["nc"])


@pytest.mark.parametrize("name,d,modes", CASES)
def test_golden_geometry_attributes(name, d, modes):
    g = load_golden(name)
    obj = _cls(d)(g["field"].copy(), "quiet")
    names = "xyzt"[:d]
    assert [getattr(obj, "h" + c) for c in names] == list(g["h"])
    assert [getattr(obj, c + "IntMin") for c in names] == list(g["int_min"])
    assert [getattr(obj, c + "IntMax") for c in names] == list(g["int_max"])
    if d == 3:
        assert np.array_equal(obj.nPos, g["ncell_axis"])
    else:
        assert [obj.nPosx - 3, obj.nPosy - 3, obj.nPosz - 3, obj.nPost - 3] == list(g["ncell_axis"])
    assert np.array_equal(obj.basePointInds, g["base_point_inds"])
    assert np.array_equal(obj.inputfield, g["sorted_field"])
    m = load_golden("matrices")
    key, sc = ("A3_times8", 8) if d == 3 else ("A4_times16", 16)
    assert np.array_equal(obj.A * sc, m[key].astype(np.float64))


@pytest.mark.parametrize("name,d,modes", CASES)
def test_golden_coefficient_table(name, d, modes):
    """Build kernels (TMA stencil + DMMA solve) against the reference's alpha arrays after allCoeffs()."""
    g = load_golden(name)
    mode = modes[-1]
    kw = {} if mode == "scalar" else {"mode": mode}
    obj = _cls(d)(g["field"].copy(), "quiet", **kw)
    obj.allCoeffs()
    for k in "xyzn":
        key = f"{mode}_alpha{k}"
        if key not in g.files:
            continue
        ref = g[key]
        got = getattr(obj, "alpha" + k)
        assert got.shape == ref.shape
        scale = np.abs(ref[:, :-1]).max()
        assert_parity(got[:, :obj.nc], ref[:, :obj.nc], scale, RTOL, f"{name} alpha{k}")
        assert np.array_equal(np.isnan(got[:, -1]), np.isnan(ref[:, -1]))


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("name,d,modes", CASES)
def test_build_tile_configurations(name, d, modes, variant, cuda_lib):
    """Every tile configuration of the build kernel produces the reference's coefficient table."""
    g = load_golden(name)
    mode = modes[-1]
    kw = {} if mode == "scalar" else {"mode": mode}
    old = cuda_lib.arb_set_build_variant(variant)
    try:
        obj = _cls(d)(g["field"].copy(), "quiet", **kw)
    finally:
        cuda_lib.arb_set_build_variant(old)
    for k in "xyzn":
        key = f"{mode}_alpha{k}"
        if key in g.files:
            ref = g[key]
            assert_parity(getattr(obj, "alpha" + k)[:, :obj.nc], ref[:, :obj.nc], np.abs(ref[:, :-1]).max(), RTOL,
                          f"{name} alpha{k} build variant {variant}")


@pytest.mark.parametrize("variant", [0, 1, 2, 10, 11, 20, 21, 22, 23, 30])
@pytest.mark.parametrize("name,d,modes", CASES)
def test_all_kernel_variants(name, d, modes, variant, cuda_lib):
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


@pytest.mark.parametrize("name,d,modes", CASES)
def test_golden_single_point_queries(name, d, modes):
    g = load_golden(name)
    for mode in modes:
        kw = {} if mode == "scalar" else {"mode": mode}
        obj = _cls(d)(g["field"].copy(), "quiet", **kw)
        q_in = g[mode + "_q_in"]
        s_comp, s_norm, s_grad = _scales(g["field"], d, g["h"])
        for row, ref in zip(g[mode + "_single_rows"], g[mode + "_single_out"]):
            out = obj.Query(q_in[row, :d].copy())
            if mode == "vector":
                assert out.shape == (3,)
                flat, scale = out, s_comp
            elif mode in ("norm", "scalar"):
                assert np.ndim(out[0]) == 0 and out[1].shape == (d,)
                flat, scale = np.concatenate([[out[0]], out[1]]), np.concatenate([[s_norm], s_grad])
            else:
                assert out[0].shape == (3,) and np.ndim(out[1]) == 0 and out[2].shape == (d,)
                flat = np.concatenate([out[0], [out[1]], out[2]])
                scale = np.concatenate([s_comp, [s_norm], s_grad])
            assert_parity(flat, ref, scale, RTOL, f"{name}/{mode} single")
        # outside the volume: bare nan in vector mode, TypeError from the tuple unpack otherwise (A.py:198, 216)
        far = np.full(d, 1e9)
        if mode == "vector":
            assert np.isnan(obj.Query(far))
        else:
            with pytest.raises(TypeError):
                obj.Query(far)


def test_example_script_flow_through_dropin_import_path(tmp_path):
    """Configs 1/2/4: the reference example scripts' flow (CSV field -> tricubic/quadcubic -> single point
    + 20-point line; scalar, then vector mode='both') on synthetic stand-in fields, checked against the oracle."""
    import importlib.util
    import os
    from conftest import ROOT
    from oracle.arb_oracle import OracleInterp

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "examples", name + ".py"))
        mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
        return mod

    make, runner = load("make_example_fields"), load("run_examples")
    folder = str(tmp_path / "ExampleFields")
    make.main(folder)
    from arbinterp_b200 import quadcubic, tricubic
    from arbinterp_b200.io import load_field_csv
    for cls, d, stem in ((tricubic, 3, "3D"), (quadcubic, 4, "4D")):
        out = runner.run(cls, d, folder, stem)
        coords = np.zeros((20, d))
        for a in range(3):
            coords[:, a] = np.linspace(-2e-3, 2e-3, 20)
        if d == 4:
            coords[:, 3] = np.linspace(-3e-6, 3e-6, 20)
        fs = load_field_csv(os.path.join(folder, f"Example{stem}ScalarField.csv"))
        fv = load_field_csv(os.path.join(folder, f"Example{stem}VectorField.csv"))
        n_ref, g_ref = OracleInterp(fs, d).query(coords.copy())
        ora = OracleInterp(fv, d, mode="both")
        c_ref, n2_ref, g2_ref = ora.query(coords.copy())
        h = np.array(ora.geo.h)
        s = np.abs(fs[:, d]).max()
        assert not np.isnan(n_ref).any(), "the example line must lie inside the stand-in volume"
        assert_parity(out["scalar_line"][0], n_ref, s, RTOL, stem + " scalar line norms")
        assert_parity(out["scalar_line"][1], g_ref, s / h[None, :], RTOL, stem + " scalar line grads")
        assert_parity(out["vector_line"][0], c_ref, np.abs(fv[:, d:]).max(axis=0)[None, :], RTOL, stem + " comps")
        assert_parity(out["vector_line"][1], n2_ref, s, RTOL, stem + " norms")
        assert_parity(out["vector_line"][2], g2_ref, s / h[None, :], RTOL, stem + " grads")
        assert_parity(out["scalar_single"][0], n_ref[3, 0], s, RTOL, stem + " single norm")
        assert_parity(out["vector_single"][0], c_ref[3], np.abs(fv[:, d:]).max(axis=0), RTOL, stem + " single comps")


def test_norm_plane_bit_exact_on_gpu():
    """Bn = ||(Bx,By,Bz)|| (A.py:58, 74): same products, same sum order, IEEE sqrt -> identical bits."""
    from arbinterp_b200.ingest import ingest_field, norm_plane
    g = load_golden("tri_12x10x9")
    planes, _ = ingest_field(g["field"], 3, device="cuda")
    ref = np.linalg.norm(g["sorted_field"][:, 3:], axis=1)
    assert np.array_equal(norm_plane(planes).reshape(-1).cpu().numpy(), ref)


def test_warp_dedup_clustered_queries():
    """Many queries in few cells (the example scripts' line queries, particle bunches): lanes that
    share a cell read one fetched block; results must equal the no-dedup kernel bit for bit."""
    from arbinterp_b200 import tricubic, _lib
    rng = np.random.default_rng(8)
    field = _analytic_field3(16, 15, 14, rng=rng)
    obj = tricubic(field.copy(), "quiet", mode="both")
    centre = np.array([0.1, 0.05, 0.7])
    q = centre + rng.normal(0, 0.03, (20_000, 3))
    q[::7] = q[3]                                    # exact duplicates too
    lib = _lib.load()
    res = {}
    for v in (0, 20):
        old = lib.arb_set_query_variant(v)
        try:
            res[v] = obj.Query(q.copy())
        finally:
            lib.arb_set_query_variant(old)
    for a, b in zip(res[0], res[20]):
        assert np.array_equal(a, b, equal_nan=True)


def test_example_diagonal():
    g = load_golden("tri_example_diag")
    from arbinterp_b200 import tricubic
    obj = tricubic(g["field"].copy(), "quiet")
    norms, grads = obj.Query(g["coords"].copy())
    s = np.abs(g["field"][:, 3]).max()
    assert_parity(norms, g["norms"], s, RTOL, "diag norms")
    assert_parity(grads, g["grads"], s / np.array([obj.hx, obj.hy, obj.hz])[None, :], RTOL, "diag grads")
    assert np.array_equal(obj.queryInds, g["inds"])
    n1, g1 = obj.Query(g["coords"][3].copy())
    assert_parity(n1, g["single_norm"], s, RTOL, "diag single")


def _analytic_field3(nx, ny, nz, rng=None, scalar=False):
    x = np.linspace(-1.0, 1.0, nx); y = np.linspace(-0.7, 0.9, ny); z = np.linspace(0.0, 1.5, nz)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    X, Y, Z = X.ravel(), Y.ravel(), Z.ravel()
    cols = [X, Y, Z, np.sin(2 * np.pi * X) * np.cos(np.pi * Y) * np.exp(-Z), X * X * Y + Z, np.cos(X + Y + Z)]
    f = np.stack(cols, axis=1)
    if scalar:
        f = np.concatenate([f[:, :3], np.linalg.norm(f[:, 3:], axis=1)[:, None]], axis=1)
    if rng is not None:
        f = f[rng.permutation(len(f))]
    return f


def _uniform_queries(obj, d, n, rng, extra=0):
    names = "xyzt"[:d]
    lo = np.array([getattr(obj, c + "IntMin") for c in names]); hi = np.array([getattr(obj, c + "IntMax") for c in names])
    q = rng.uniform(0, 1, (n, d + extra))
    q[:, :d] = lo + q[:, :d] * (hi - lo) * (1 - 1e-12)
    return q


@pytest.mark.parametrize("mode", ["vector", "norm", "both"])
@pytest.mark.parametrize("shape", [(37, 26, 23), (20, 41, 17)])
def test_oracle_parity_3d_multi_tile(mode, shape):
    """Odd nx (padded TMA pitch), cell counts that are not tile multiples, 10^5 seeded queries."""
    from arbinterp_b200 import tricubic
    from oracle.arb_oracle import OracleInterp
    rng = np.random.default_rng(1234)
    field = _analytic_field3(*shape, rng=rng)
    obj = tricubic(field.copy(), "quiet", mode=mode)
    q = _uniform_queries(obj, 3, 100_000, rng, extra=2)
    q[::97, 1] = 5.0                      # out of volume
    q[::1013, 2] = np.nan                 # NaN coordinate: output NaN, row not overwritten
    ora = OracleInterp(field, 3, mode=mode)
    q_ref = q.copy()
    ref = ora.query(q_ref)
    ref = ref if isinstance(ref, tuple) else (ref,)
    q_gpu = q.copy()
    res = obj.Query(q_gpu)
    worst = _check_outputs(res, ref, mode, field, 3, ora.geo.h, f"3d {shape} {mode}")
    assert np.array_equal(q_gpu, q_ref, equal_nan=True)
    assert np.array_equal(obj.queryInds, ora.query_inds)
    assert worst < 1e-13


@pytest.mark.parametrize("mode", ["vector", "norm", "both"])
def test_oracle_parity_4d_multi_tile(mode):
    from arbinterp_b200 import quadcubic
    from oracle.arb_oracle import OracleInterp
    rng = np.random.default_rng(99)
    ax = [np.linspace(-1, 1, 13), np.linspace(0, 1, 8), np.linspace(-2, 0, 9), np.linspace(0, 3e-6, 7)]
    T, Z, Y, X = [a.ravel() for a in np.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")]
    ts = T / 3e-6
    field = np.stack([X, Y, Z, T, np.sin(2 * X) * np.cos(3 * Y) * np.exp(Z) * np.cos(2 * ts),
                      X * X * Y + Z * ts + 0.3 * X * Y * Z * ts, np.cos(X + Y + Z + ts)], axis=1)
    field = field[rng.permutation(len(field))]
    obj = quadcubic(field.copy(), "quiet", mode=mode)
    q = _uniform_queries(obj, 4, 20_000, rng, extra=1)
    q[::53, 3] = -1.0
    ora = OracleInterp(field, 4, mode=mode)
    q_ref = q.copy()
    ref = ora.query(q_ref)
    ref = ref if isinstance(ref, tuple) else (ref,)
    q_gpu = q.copy()
    res = obj.Query(q_gpu)
    _check_outputs(res, ref, mode, field, 4, ora.geo.h, f"4d {mode}")
    assert np.array_equal(q_gpu, q_ref, equal_nan=True)
    assert np.array_equal(obj.queryInds, ora.query_inds)


def test_device_tensor_path_and_pageable_host_path():
    from arbinterp_b200 import tricubic
    rng = np.random.default_rng(5)
    field = _analytic_field3(21, 19, 18, rng=rng)
    obj = tricubic(field.copy(), "quiet", mode="both")
    q = _uniform_queries(obj, 3, 50_000, rng)
    q[7] = [9.0, 0.0, 0.5]
    host = obj.Query(q.copy())
    qd = torch.from_numpy(q.copy()).cuda()
    dev = obj.Query(qd)
    for a, b in zip(host, dev):
        assert isinstance(b, torch.Tensor) and b.is_cuda
        assert np.array_equal(a, b.cpu().numpy(), equal_nan=True)
    assert torch.isnan(qd[7]).all()
    # page-locked caller buffers are copied without staging; the in-place NaN rows still land in them
    qp = torch.from_numpy(q.copy()).pin_memory().numpy()
    pinned = obj.Query(qp)
    for a, b in zip(host, pinned):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(qp[7]).all() and np.array_equal(qp[8], q[8])
    # results too large to pin come back as ordinary arrays staged through the library's ring
    old_limit = type(obj)._PIN_LIMIT_BYTES
    type(obj)._PIN_LIMIT_BYTES = 1024
    try:
        unpinned = obj.Query(q.copy())
    finally:
        type(obj)._PIN_LIMIT_BYTES = old_limit
    for a, b in zip(host, unpinned):
        assert np.array_equal(a, b, equal_nan=True)
    # float32 / non-contiguous input goes through a float64 copy; NaN rows are still written back
    q32 = q.astype(np.float32)
    r32 = obj.Query(q32)
    assert np.isnan(q32[7]).all() and r32[0].dtype == np.float64
    # small chunks exercise the 3-slot stream ring of arb_query_host
    from arbinterp_b200 import _lib
    n = len(q)
    outs = [np.empty((n, 3)), np.empty((n, 1)), np.empty((n, 3))]
    cells = np.empty(n, dtype=np.int64)
    qq = q.copy()
    rc = obj._lib.arb_query_host(ctypes.byref(obj._cgeom), obj.table.data_ptr(), _lib.MODE_BOTH, qq.ctypes.data, n, 3,
                                 outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, cells.ctypes.data, 4096)
    assert rc == 0
    for a, b in zip(host, outs):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.isnan(qq[7]).all() and cells[7] == obj.nc


@pytest.mark.parametrize("name,d,mode", [("tri_12x10x9", 3, "both"), ("quad_8x7x7x6", 4, "norm")])
def test_save_and_load_coefficients(name, d, mode, tmp_path):
    """Coefficient-table persistence (reference CHANGELOG.md:9 lists it as future work): a loaded
    interpolator answers bit-identically without the field, the ingest or the build."""
    g = load_golden(name)
    obj = _cls(d)(g["field"].copy(), "quiet", mode=mode)
    q = g[mode + "_q_in"].copy()
    ref = obj.Query(q.copy())
    path = tmp_path / "table.arb"
    obj.save(str(path), chunk_bytes=1 << 16)
    back = _cls(d).load(str(path), chunk_bytes=1 << 15)
    got = back.Query(q.copy())
    ref = ref if isinstance(ref, tuple) else (ref,)
    got = got if isinstance(got, tuple) else (got,)
    for a, b in zip(ref, got):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(back.queryInds, obj.queryInds) and back.nc == obj.nc
    assert torch.equal(back.table, obj.table) or bool(torch.isnan(back.table[-1]).all())
    with pytest.raises(AttributeError):
        back.Bn if mode != "vector" else back.Bx
    with pytest.raises(ValueError):
        _cls(7 - d).load(str(path))


@pytest.mark.parametrize("name,d,modes", CASES)
def test_table_free_path_golden(name, d, modes):
    """tricubic / quadcubic(table=False): no coefficient table, every query evaluated from its 4^d grid
    neighbourhood (odd nx = 9 exercises the padded TMA pitch; the 4-D cases exercise the rank-16 term that
    reproduces A.py:860).  Same parity bar as the table path."""
    g = load_golden(name)
    for mode in modes:
        kw = {} if mode == "scalar" else {"mode": mode}
        obj = _cls(d)(g["field"].copy(), "quiet", table=False, **kw)
        q = g[mode + "_q_in"].copy()
        res = obj.Query(q)
        _check_outputs(res, _golden_ref(g, mode), mode, g["field"], d, g["h"], f"{name}/{mode}/table-free")
        assert np.array_equal(q, g[mode + "_q_after"], equal_nan=True)
        assert np.array_equal(obj.queryInds, g[mode + "_inds"])
        with pytest.raises(AttributeError):
            obj.table
        if mode in ("vector", "scalar"):
            assert np.array_equal(obj.inputfield, g["sorted_field"])      # padded pitch is not visible
    with pytest.raises(ValueError):
        _cls(d)(g["field"].copy(), "quiet", table=False, slab=(0, 1))


@pytest.mark.parametrize("mode", ["vector", "norm", "both"])
def test_table_free_matches_table_path(mode):
    from arbinterp_b200 import tricubic
    rng = np.random.default_rng(77)
    field = _analytic_field3(37, 26, 23, rng=rng)
    a = tricubic(field.copy(), "quiet", mode=mode)
    b = tricubic(field.copy(), "quiet", mode=mode, table=False)
    q = _uniform_queries(a, 3, 200_000, rng, extra=1)
    q[::101, 0] = -9.0
    qa, qb = q.copy(), q.copy()
    ra, rb = a.Query(qa), b.Query(qb)
    s_comp, s_norm, s_grad = _scales(field, 3, [a.hx, a.hy, a.hz])
    ra = ra if isinstance(ra, tuple) else (ra,)
    rb = rb if isinstance(rb, tuple) else (rb,)
    _check_outputs(rb, ra, mode, field, 3, [a.hx, a.hy, a.hz], f"table-free vs table {mode}")
    assert np.array_equal(qa, qb, equal_nan=True) and np.array_equal(a.queryInds, b.queryInds)
    dev = b.Query(torch.from_numpy(q.copy()).cuda())
    dev = dev if isinstance(dev, tuple) else (dev,)
    for x, y in zip(rb, dev):
        assert np.array_equal(x, y.cpu().numpy(), equal_nan=True)


def _analytic_field4(nx, ny, nz, nt, rng=None):
    x = np.linspace(-1.0, 1.0, nx); y = np.linspace(-0.7, 0.9, ny); z = np.linspace(0.0, 1.5, nz)
    t = np.linspace(0.0, 2.0, nt)
    T, Z, Y, X = [a.ravel() for a in np.meshgrid(t, z, y, x, indexing="ij")]
    cols = [X, Y, Z, T, np.sin(2 * np.pi * X) * np.cos(np.pi * Y) * np.exp(-Z) * np.cos(T), X * X * Y + Z * (1 + T) + X * Y * Z * T,
            np.cos(X + Y + Z + T)]
    f = np.stack(cols, axis=1)
    return f if rng is None else f[rng.permutation(len(f))]


@pytest.mark.parametrize("fixed", [False, True])
@pytest.mark.parametrize("mode", ["vector", "norm", "both"])
def test_table_free_quadcubic_matches_table_path(mode, fixed):
    """4-D table-free kernel against the table path on the same field (xyzt monomial present, so the A.py:860
    term is visible: the quirk and fixed_d4 results differ by ~1e-6 and each must match its own table)."""
    from arbinterp_b200 import quadcubic
    rng = np.random.default_rng(78)
    field = _analytic_field4(13, 11, 10, 9, rng=rng)
    a = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=fixed)
    b = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=fixed, table=False)
    q = _uniform_queries(a, 4, 100_003, rng, extra=2)
    q[::97, 3] = 1e3
    qa, qb = q.copy(), q.copy()
    ra, rb = a.Query(qa), b.Query(qb)
    hs = [a.hx, a.hy, a.hz, a.ht]
    ra = ra if isinstance(ra, tuple) else (ra,)
    rb = rb if isinstance(rb, tuple) else (rb,)
    _check_outputs(rb, ra, mode, field, 4, hs, f"4-D table-free vs table {mode} fixed={fixed}")
    assert np.array_equal(qa, qb, equal_nan=True) and np.array_equal(a.queryInds, b.queryInds)
    dev = b.Query(torch.from_numpy(q.copy()).cuda())
    dev = dev if isinstance(dev, tuple) else (dev,)
    for x, y in zip(rb, dev):
        assert np.array_equal(x, y.cpu().numpy(), equal_nan=True)
    if mode == "norm":
        other = quadcubic(field.copy(), "quiet", mode=mode, fixed_d4=not fixed, table=False).Query(q.copy())
        ok = ~np.isnan(rb[0][:, 0])
        assert np.abs(other[0][ok] - rb[0][ok]).max() > 1e-9, "quirk and corrected matrices must differ on this field"


def _verlet_reference(query_grad, pos, vel, dt, nsteps, kappa, g):
    """numpy velocity Verlet around a gradient oracle; lost particles (NaN gradient) become NaN."""
    x, v = pos.copy(), vel.copy()
    a = kappa * query_grad(x) + g
    for _ in range(nsteps):
        v = v + 0.5 * dt * a
        x = x + dt * v
        a = kappa * query_grad(x) + g
        v = v + 0.5 * dt * a
        lost = np.isnan(a).any(axis=1)
        x[lost] = np.nan; v[lost] = np.nan
    return x, v


@pytest.mark.parametrize("mode", ["norm", "both"])
def test_fused_push_matches_oracle_integration(mode):
    """Fused query+push kernel against a numpy velocity-Verlet loop around the CPU oracle's
    gradient: same trajectories to round-off amplification, same particles lost at the volume boundary."""
    from arbinterp_b200 import tricubic
    from oracle.arb_oracle import OracleInterp
    rng = np.random.default_rng(11)
    field = _analytic_field3(24, 22, 20, rng=rng)
    obj = tricubic(field.copy(), "quiet", mode=mode)
    ora = OracleInterp(field, 3, mode="norm")
    n = 4000
    pos = _uniform_queries(obj, 3, n, rng)
    vel = rng.normal(0, 0.4, (n, 3))
    dt, nsteps, kappa, g = 0.01, 40, -0.7, np.array([0.0, 0.0, -0.3])

    def grad_oracle(x):
        with np.errstate(invalid="ignore"):
            return ora.query(x.copy())[1]

    xr, vr = _verlet_reference(grad_oracle, pos, vel, dt, nsteps, kappa, g)
    p = torch.from_numpy(pos.copy()).cuda(); v = torch.from_numpy(vel.copy()).cuda()
    lost = obj.push(p, v, dt, nsteps, kappa, gravity=g)
    xg, vg = p.cpu().numpy(), v.cpu().numpy()
    assert np.array_equal(np.isnan(xg), np.isnan(xr)) and lost == int(np.isnan(xr[:, 0]).sum()) and 0 < lost < n
    ok = ~np.isnan(xr[:, 0])
    assert np.max(np.abs(xg[ok] - xr[ok])) < 1e-10 and np.max(np.abs(vg[ok] - vr[ok])) < 1e-9
    # numpy in/out convenience path gives the same result
    pn, vn = pos.copy(), vel.copy()
    assert obj.push(pn, vn, dt, nsteps, kappa, gravity=g) == lost
    assert np.array_equal(pn, xg, equal_nan=True) and np.array_equal(vn, vg, equal_nan=True)


def test_fused_push_4d_time_dependent_field():
    """quadcubic push: each particle carries its own time (4th column), the spatial gradient comes from
    the 4-D table.  Checked against a numpy Verlet loop around the CPU oracle's 4-D gradient."""
    from arbinterp_b200 import quadcubic
    from oracle.arb_oracle import OracleInterp
    rng = np.random.default_rng(21)
    ax = [np.linspace(-1, 1, 14), np.linspace(0, 1, 12), np.linspace(-1, 0, 11), np.linspace(0, 2, 16)]
    T, Z, Y, X = [a.ravel() for a in np.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")]
    field = np.stack([X, Y, Z, T, 2 + np.sin(2 * X) * np.cos(3 * Y) * np.exp(Z) * np.cos(T) + 0.2 * X * Y * Z * T], axis=1)
    obj = quadcubic(field.copy(), "quiet")
    ora = OracleInterp(field, 4)
    n, dt, nsteps, kappa = 3000, 0.02, 30, -0.5
    pos = _uniform_queries(obj, 4, n, rng)
    pos[:, 3] = rng.uniform(obj.tIntMin, obj.tIntMin + 0.5 * (obj.tIntMax - obj.tIntMin), n)
    vel = rng.normal(0, 0.3, (n, 3))
    x, v, t = pos[:, :3].copy(), vel.copy(), pos[:, 3].copy()

    def grad(x, t):
        with np.errstate(invalid="ignore"):
            return ora.query(np.column_stack([x, t]))[1][:, :3]

    a = kappa * grad(x, t)
    for _ in range(nsteps):
        v = v + 0.5 * dt * a; x = x + dt * v; t = t + dt
        a = kappa * grad(x, t)
        v = v + 0.5 * dt * a
        lost = np.isnan(a).any(axis=1)
        x[lost] = np.nan; v[lost] = np.nan
    p = torch.from_numpy(pos.copy()).cuda(); vv = torch.from_numpy(vel.copy()).cuda()
    nlost = obj.push(p, vv, dt, nsteps, kappa)
    pg, vg = p.cpu().numpy(), vv.cpu().numpy()
    assert np.array_equal(np.isnan(pg[:, 0]), np.isnan(x[:, 0])) and nlost == int(np.isnan(x[:, 0]).sum()) and 0 < nlost < n
    ok = ~np.isnan(x[:, 0])
    assert np.max(np.abs(pg[ok, :3] - x[ok])) < 1e-10 and np.max(np.abs(vg[ok] - v[ok])) < 1e-9
    assert np.max(np.abs(pg[ok, 3] - t[ok])) < 1e-12


def test_fused_push_equals_unfused_query_loop():
    """One fused launch == a Python loop of Query + torch updates (the per-step round trip it removes),
    including steps where a particle stays in its cell and the kernel re-uses the block it already holds."""
    from arbinterp_b200 import tricubic
    rng = np.random.default_rng(12)
    field = _analytic_field3(20, 20, 20, rng=rng, scalar=True)
    obj = tricubic(field.copy(), "quiet")
    n = 3000
    pos = _uniform_queries(obj, 3, n, rng) * 0.5
    pos[:, 2] += 0.6
    vel = rng.normal(0, 0.05, (n, 3))
    dt, nsteps, kappa = 0.003, 60, 1.3                    # ~0.003 of a cell per step: mostly cell re-use

    def grad_gpu(x):
        return obj.Query(torch.from_numpy(x.copy()).cuda())[1].cpu().numpy()

    xr, vr = _verlet_reference(grad_gpu, pos, vel, dt, nsteps, kappa, np.zeros(3))
    p = torch.from_numpy(pos.copy()).cuda(); v = torch.from_numpy(vel.copy()).cuda()
    obj.push(p, v, dt, nsteps, kappa)
    ok = ~np.isnan(xr[:, 0])
    assert ok.sum() > n // 2
    assert np.max(np.abs(p.cpu().numpy()[ok] - xr[ok])) < 1e-12 and np.max(np.abs(v.cpu().numpy()[ok] - vr[ok])) < 1e-11
    with pytest.raises(ValueError):
        tricubic(_analytic_field3(12, 11, 10), "quiet", mode="vector").push(p, v, dt, 1, kappa)


def test_empty_tiny_and_integer_queries():
    from arbinterp_b200 import tricubic
    field = _analytic_field3(12, 11, 10)
    obj = tricubic(field.copy(), "quiet", mode="both")
    comps, norms, grads = obj.Query(np.empty((0, 3)))
    assert comps.shape == (0, 3) and norms.shape == (0, 1) and grads.shape == (0, 3)
    one = obj.Query(np.array([[0.1, 0.05, 0.7]]))                  # a single ROW is still a range query
    assert one[0].shape == (1, 3) and one[1].shape == (1, 1) and one[2].shape == (1, 3)
    # integer arrays: work while nothing has to be NaN-masked, ValueError otherwise (numpy cannot store NaN)
    qi = np.zeros((4, 3), dtype=np.int64)
    qi[:, 2] = 1                                                  # (0, 0, 1) is inside the volume
    assert np.isfinite(obj.Query(qi)[1]).all()
    qi[2, 0] = 50
    with pytest.raises(ValueError):
        obj.Query(qi)
    # Fortran-ordered / strided views go through a contiguous copy and still get their NaN rows back
    big = np.asfortranarray(np.array([[0.1, 0.05, 0.7, 9.0], [7.0, 0.0, 0.5, 9.0]]))
    obj.Query(big)
    assert np.isnan(big[1]).all() and not np.isnan(big[0]).any()


def test_upper_edge_and_errors(cuda_lib):
    from arbinterp_b200 import tricubic, _lib
    field = _analytic_field3(12, 11, 10)
    obj = tricubic(field.copy(), "quiet", mode="norm")
    q = np.array([[obj.xIntMax, obj.yIntMax, obj.zIntMax], [obj.xIntMin, obj.yIntMin, obj.zIntMin]])
    norms, grads = obj.Query(q.copy())
    assert np.isfinite(norms[1, 0])                      # lower edges are interpolatable
    assert np.isnan(norms[0, 0]) or np.isfinite(norms[0, 0])   # never a crash; NaN when the index rounds to n-3
    with pytest.raises(SystemExit):
        tricubic(np.zeros((10, 5)), "quiet")
    # mode / table mismatch is an error code, not a crash
    rc = cuda_lib.arb_query(ctypes.byref(obj._cgeom), obj.table.data_ptr(), _lib.MODE_BOTH, 0, 1, 3, 0, 0, 0, 0, 0, 0, 0)
    assert rc != 0 and b"components" in cuda_lib.arb_last_error()


def test_slab_sharded_table_matches_whole():
    """t-slab sharding: a slab built from planes [lo-1, hi+2] answers its own queries
    bit-identically to the unsharded table and reports the global cell index."""
    from arbinterp_b200 import quadcubic
    g = load_golden("quad_8x7x7x6")
    whole = quadcubic(g["field"].copy(), "quiet", mode="both")
    q = g["both_q_in"].copy()
    ref = whole.Query(q.copy())
    inds = whole.queryInds
    nt_cells = whole.nPost - 3
    layer = whole.nc // nt_cells
    got = [np.full_like(r, np.nan) for r in ref]
    for lo, hi in [(0, 1), (1, nt_cells)]:
        part = quadcubic(g["field"].copy(), "quiet", mode="both", slab=(lo, hi))
        res = part.Query(q.copy())
        assert np.array_equal(part.queryInds, inds)
        own = (inds < whole.nc) & (inds // layer >= lo) & (inds // layer < hi)
        for o, r in zip(got, res):
            o[own] = r[own]
            assert np.isnan(r[~own]).all()
    for a, b in zip(got, ref):
        assert np.array_equal(a, b, equal_nan=True)


# ----------------------------------------------------------------------------------------
# full-size property tests (BASELINE.json sizes; no oracle table needed)
# ----------------------------------------------------------------------------------------
def test_full_size_256_quadratic_reproduced():
    """256^3 grid: per-axis quadratics are reproduced to round-off (central differences exact),
    values and gradients, and the cell indices equal the oracle's locate() bit for bit."""
    from arbinterp_b200 import tricubic
    from oracle.arb_oracle import OracleInterp
    n = 256
    ax = torch.linspace(-1, 1, n, dtype=torch.float64)
    Z, Y, X = torch.meshgrid(ax, ax, ax, indexing="ij")
    f = lambda X, Y, Z: 1 + X - 2 * Y + 3 * Z + X * Y - Y * Z + X * X * Z + Y * Y - 0.5 * Z * Z * X
    field = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1), f(X, Y, Z).reshape(-1)], dim=1)
    obj = tricubic(field, "quiet")
    assert obj.nc == 253 ** 3
    rng = np.random.default_rng(20260117)
    q = _uniform_queries(obj, 3, 1_000_000, rng)
    norms, grads = obj.Query(q.copy())
    x, y, z = q[:, 0], q[:, 1], q[:, 2]
    assert np.max(np.abs(norms[:, 0] - f(x, y, z))) < 5e-13
    gx = 1 + y + 2 * x * z - 0.5 * z * z
    gy = -2 + x - z + 2 * y
    gz = 3 - y + x * x - z * x
    for got, want in zip(grads.T, (gx, gy, gz)):
        assert np.max(np.abs(got - want)) < 5e-11      # divided by h = 2/255
    # indices: oracle locate() on the same geometry (cheap, no table)
    small = OracleInterp.__new__(OracleInterp)
    from oracle.arb_oracle import GridGeometry
    geo = GridGeometry.__new__(GridGeometry)
    geo.d = 3; geo.ncell_axis = [253] * 3; geo.nc = 253 ** 3
    geo.h = [obj.hx, obj.hy, obj.hz]; geo.int_min = [obj.xIntMin, obj.yIntMin, obj.zIntMin]
    geo.int_max = [obj.xIntMax, obj.yIntMax, obj.zIntMax]
    small.geo = geo; small.d = 3
    inds, _ = small.locate(q.copy())
    assert np.array_equal(inds, obj.queryInds)


def test_full_size_4d_quirk_detector():
    """4-D: polynomials without an x*y*z*t monomial are reproduced to round-off; adding x*y*z*t
    shows the A.py:860 quirk (error ~1e-6 of the term) unless fixed_d4=True."""
    from arbinterp_b200 import quadcubic
    ax = [torch.linspace(-1, 1, 33, dtype=torch.float64), torch.linspace(0, 1, 30, dtype=torch.float64),
          torch.linspace(-1, 0, 29, dtype=torch.float64), torch.linspace(0, 2, 17, dtype=torch.float64)]
    T, Z, Y, X = torch.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")
    base = lambda X, Y, Z, T: 1 + X * Y - Z * T + X * X * T + Y * Z + 0.5 * T * T - X * Z
    rng = np.random.default_rng(3)
    for with_xyzt in (False, True):
        fun = (lambda X, Y, Z, T: base(X, Y, Z, T) + X * Y * Z * T) if with_xyzt else base
        field = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1), T.reshape(-1), fun(X, Y, Z, T).reshape(-1)], dim=1)
        obj = quadcubic(field, "quiet")
        q = _uniform_queries(obj, 4, 200_000, rng)
        norms, _ = obj.Query(q.copy())
        err = np.max(np.abs(norms[:, 0] - fun(*q.T)))
        if not with_xyzt:
            assert err < 1e-12
        else:
            assert 1e-9 < err < 1e-3, "reference quirk (A.py:860) must be reproduced by default"
            fixed = quadcubic(field, "quiet", fixed_d4=True)
            n2, _ = fixed.Query(q.copy())
            assert np.max(np.abs(n2[:, 0] - fun(*q.T))) < 1e-12


def test_randomised_parity_stress():
    """40 random (shape, spacing, mode, table/table-free, build variant, query variant) cases against the oracle
    (tools/stress_parity.py; 200 cases were run for profiles/r01_stress_parity.log)."""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_parity.py"), "--cases", "40", "--seed", "7"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "0 failures" in out.stdout


def test_plain_c_client_of_the_c_abi(tmp_path):
    """tests/c_abi/c_smoke.c: build + device query + host-buffer query from plain C, checked against an
    analytic quadratic field, the NaN / cell-index conventions and the error path."""
    import subprocess
    from test_host_logic import _compile_c_client
    exe = _compile_c_client(str(tmp_path / "c_smoke"))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "c_smoke ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.parametrize("d", [3, 4])
@pytest.mark.parametrize("table", [True, False])
def test_update_values_equals_fresh_construction(d, table):
    """update_values(): new field values on the same grid, rebuilt in place (no re-ingest, no reallocation);
    afterwards every answer is bit-identical to a freshly constructed interpolator of the updated field."""
    rng = np.random.default_rng(5 + d)
    field = _analytic_field3(13, 12, 11, rng=rng) if d == 3 else _analytic_field4(9, 8, 7, 6, rng=rng)
    cls = _cls(d)
    for mode in ("vector", "both"):
        obj = cls(field.copy(), "quiet", mode=mode, table=table)
        ptr = obj.table.data_ptr() if table else None
        new = field.copy()
        new[:, d:] = np.cos(3.0 * field[:, d:]) + field[:, [0]] * field[:, [1]]
        obj.update_values(new[:, d:])                                # order='rows': the constructor's row order
        fresh = cls(new.copy(), "quiet", mode=mode, table=table)
        q = _uniform_queries(fresh, d, 20_000, rng)
        a, b = obj.Query(q.copy()), fresh.Query(q.copy())
        for x, y in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
            assert np.array_equal(x, y, equal_nan=True)
        if table:
            assert obj.table.data_ptr() == ptr and torch.equal(obj.table[:-1], fresh.table[:-1])
        # order='grid': values in sorted grid order (the rows of inputfield)
        if mode == "vector":
            grid_vals = fresh.inputfield[:, d:]
            obj.update_values(np.zeros_like(grid_vals), order="grid")
            obj.update_values(grid_vals, order="grid")
            c = obj.Query(q.copy())
            assert np.array_equal(c, b, equal_nan=True)
        with pytest.raises(ValueError):
            obj.update_values(new[:-1, d:])
    scal = np.concatenate([field[:, :d], np.linalg.norm(field[:, d:], axis=1)[:, None]], axis=1)
    obj = cls(scal.copy(), "quiet", table=table)
    obj.update_values(2.0 * scal[:, d])                              # 1-D values for scalar input
    fresh = cls(np.concatenate([scal[:, :d], 2.0 * scal[:, [d]]], axis=1), "quiet", table=table)
    q = _uniform_queries(fresh, d, 5_000, rng)
    for x, y in zip(obj.Query(q.copy()), fresh.Query(q.copy())):
        assert np.array_equal(x, y, equal_nan=True)


@pytest.mark.parametrize("d", [3, 4])
def test_resumable_push_across_slabs_equals_unsharded(d):
    """arb_push_steps on one GPU: two slab tables of the same field take turns; a particle that leaves a slab is
    parked unchanged and resumed by the other table.  The end state is bit-identical to one fused push on the
    unsharded table (what SlabShardedInterp.push does across ranks)."""
    from test_multi_gpu import _push_field
    rng = np.random.default_rng(31)
    if d == 4:
        field = _push_field()
    else:
        field = _analytic_field3(14, 13, 17, scalar=True)
    cls = _cls(d)
    whole = cls(field.copy(), "quiet")
    nslow = whole._geo.ncell[d - 1]
    cut = nslow // 2
    parts = [cls(field.copy(), "quiet", slab=(0, cut)), cls(field.copy(), "quiet", slab=(cut, nslow))]
    n, dt, nsteps, kappa = 5000, 0.02, 60, -0.5
    pos = _uniform_queries(whole, d, n, rng)
    vel = rng.normal(0, 0.4, (n, 3))
    grav = (0.0, 0.0, 0.3)
    pr, vr = torch.from_numpy(pos.copy()).cuda(), torch.from_numpy(vel.copy()).cuda()
    lost_ref = whole.push(pr, vr, dt, nsteps, kappa, gravity=grav)
    p, v = torch.from_numpy(pos.copy()).cuda(), torch.from_numpy(vel.copy()).cuda()
    step = torch.zeros(n, dtype=torch.int64, device="cuda")
    lost, rounds, migrated = 0, 0, 0
    while bool((step <= nsteps).any()):
        for part in parts:
            before = step.clone()
            lost += part._push_local(p, v, step, dt, nsteps, kappa, grav)
            migrated += int(((step > before) & (step <= nsteps)).sum())     # advanced, then parked at the boundary
        rounds += 1
        assert rounds <= nsteps + 2
    assert migrated > 0, "no particle crossed the slab boundary"
    assert torch.equal(torch.nan_to_num(p, nan=-7.0), torch.nan_to_num(pr, nan=-7.0))
    assert torch.equal(torch.nan_to_num(v, nan=-7.0), torch.nan_to_num(vr, nan=-7.0))
    assert lost == lost_ref and 0 < lost < n
    with pytest.raises(ValueError):
        parts[0].push(p, v, dt, 1, kappa)               # a slab object points to the sharded entry point


@pytest.mark.parametrize("d", [3, 4])
def test_zero_copy_tiny_batches_equal_copy_path(d):
    """Batches of <= 256 rows take the zero-copy path (kernel reads/writes the mapped pinned block); the same rows
    inside a larger batch take the copy path.  Outputs, in-place NaN rows and cell indices must be identical."""
    rng = np.random.default_rng(41)
    field = _analytic_field3(13, 12, 11, rng=rng) if d == 3 else _analytic_field4(9, 8, 7, 6, rng=rng)
    obj = _cls(d)(field.copy(), "quiet", mode="both")
    q = _uniform_queries(obj, d, 700, rng, extra=2)
    q[3, 0] = 1e9; q[17, d - 1] = -1e9; q[30, 1] = np.nan
    big_q = q.copy()
    big = obj.Query(big_q)
    big_inds = obj.queryInds.copy()
    for n in (1, 2, 31, 200, 256):
        small_q = q[:n].copy()
        if n == 1:
            res = obj.rQuery(small_q)
        else:
            res = obj.Query(small_q)
        for a, b in zip(res, big):
            assert np.array_equal(a, b[:n], equal_nan=True)
        assert np.array_equal(small_q, big_q[:n], equal_nan=True)
        assert np.array_equal(obj.queryInds, big_inds[:n])
