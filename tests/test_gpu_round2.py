"""GPU tier, round 2: parity at BASELINE.json's configuration sizes (configs[1], configs[3]) against the numpy oracle,
the constructed upper-edge case, the constructor's mode-switch branches (A.py:20-104), stream ordering of the table-free
path, the coefficient-file header checks, and the single-process multi-device fan-out.

Bar (north_star): cell indices and NaN masks bit-exact; values, gradients and components within 1e-12 scaled,
|got - ref| <= 1e-12 * max(|ref|, S), S = max|field component| (S/h for gradients) -- SURVEY 8d."""
import numpy as np
import pytest

from conftest import assert_parity, load_golden

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
RTOL = 1e-12


def _quadrupole_field(n, half=3e-3, scalar=False):
    """configs[0]/[1] stand-in: mm-scale grid, quadrupole-like B plus a bias and a curvature term."""
    ax = np.linspace(-half, half, n)
    Z, Y, X = [a.ravel() for a in np.meshgrid(ax, ax, ax, indexing="ij")]
    g, b0 = 150.0, 1e-2
    cols = [X, Y, Z, g * X + b0 + 40.0 * X * Y / half, g * Y + 2e3 * Z * Z + 0.1 * np.sin(400 * X),
            -2 * g * Z + b0 * np.cos(300 * Y)]
    f = np.stack(cols, axis=1)
    if scalar:
        f = np.concatenate([f[:, :3], np.linalg.norm(f[:, 3:], axis=1)[:, None]], axis=1)
    return f


def _compare_chunks(obj, ora, q, d, mode, scales, chunk):
    """Every row of q through both paths, chunk by chunk; returns the worst scaled error."""
    s_comp, s_norm, s_grad = scales
    worst = 0.0
    for lo in range(0, len(q), chunk):
        part = q[lo:lo + chunk]
        q_ref, q_gpu = part.copy(), part.copy()
        ref = ora.query(q_ref)
        got = obj.Query(q_gpu)
        ref = ref if isinstance(ref, tuple) else (ref,)
        got = got if isinstance(got, tuple) else (got,)
        i = 0
        if mode in ("vector", "both"):
            worst = max(worst, assert_parity(got[i], ref[i], s_comp, RTOL, f"comps rows {lo}+")); i += 1
        if mode in ("norm", "both"):
            worst = max(worst, assert_parity(got[i], ref[i], s_norm, RTOL, f"norm rows {lo}+")); i += 1
            worst = max(worst, assert_parity(got[i], ref[i], s_grad, RTOL, f"grad rows {lo}+"))
        assert np.array_equal(q_gpu, q_ref, equal_nan=True), "in-place NaN rows"
        assert np.array_equal(obj.queryInds, ora.query_inds), "cell indices"
    return worst


def test_config2_101cubed_both_million_queries():
    """BASELINE configs[1]: 3-D vector field, mode='both', 10^6 random in-volume queries -- ALL compared to the oracle."""
    from arbinterp_b200 import tricubic
    from oracle.arb_oracle import OracleInterp
    field = _quadrupole_field(101)
    obj = tricubic(field.copy(), "quiet", mode="both")
    ora = OracleInterp(field, 3, mode="both", dense=False)
    rng = np.random.default_rng(20260117)
    lo = np.array([obj.xIntMin, obj.yIntMin, obj.zIntMin]); hi = np.array([obj.xIntMax, obj.yIntMax, obj.zIntMax])
    q = lo + rng.uniform(0, 1, (1_000_000, 3)) * (hi - lo) * (1 - 1e-12)
    vals = field[:, 3:]
    s_norm = np.linalg.norm(vals, axis=1).max()
    scales = (np.abs(vals).max(axis=0)[None, :], s_norm, (s_norm / np.array(ora.geo.h))[None, :])
    worst = _compare_chunks(obj, ora, q, 3, "both", scales, 125_000)
    assert worst < 1e-13


def test_config4_quadcubic_scalar_million_queries():
    """BASELINE configs[3]: 41^3 x 21 scalar field (us-scale t axis), quadcubic, 10^6 random (x,y,z,t) queries against
    the oracle; the field has an xyzt term, so the A.py:860 rank-16 term is part of what is compared."""
    from arbinterp_b200 import quadcubic
    from oracle.arb_oracle import OracleInterp
    ax = [np.linspace(-3e-3, 3e-3, 41)] * 3 + [np.linspace(-5e-6, 5e-6, 21)]
    T, Z, Y, X = [a.ravel() for a in np.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")]
    xs, ys, zs, ts = X / 3e-3, Y / 3e-3, Z / 3e-3, T / 5e-6
    u = np.sqrt((0.4 * xs) ** 2 + (0.4 * ys) ** 2 + (0.8 * zs) ** 2 + 0.05) * (1 + 0.3 * np.sin(2.2 * ts)) + 0.2 * xs * ys * zs * ts
    field = np.stack([X, Y, Z, T, u], axis=1)
    obj = quadcubic(field.copy(), "quiet")
    ora = OracleInterp(field, 4, dense=False)
    rng = np.random.default_rng(20260117)
    lo = np.array(ora.geo.int_min); hi = np.array(ora.geo.int_max)
    q = lo + rng.uniform(0, 1, (1_000_000, 4)) * (hi - lo) * (1 - 1e-12)
    s = np.abs(u).max()
    worst = _compare_chunks(obj, ora, q, 4, "norm", (None, s, (s / np.array(ora.geo.h))[None, :]), 40_000)
    assert worst < 1e-13


def test_upper_edge_constructed_rounding_cases():
    """The declared deviation, exercised for real: x and z axes are linspace(..., 9) grids on which
    (IntMax - IntMin) / h rounds to exactly n - 3 = 6 (the reference would index one cell past the end: wrap for x,
    IndexError for z), the y axis is a 12-point grid on which it rounds to 8.999999999999998 (last cell, as in the
    reference).  Here: NaN outputs and queryInds == nc for the first kind, oracle parity for the second."""
    from arbinterp_b200 import tricubic
    from oracle.arb_oracle import OracleInterp
    x = np.linspace(-1, 1, 9); y = np.linspace(-1, 1, 12); z = np.linspace(0, 2, 9)
    for ax, want in ((x, 6.0), (z, 6.0)):
        assert (ax[-2] - ax[1]) / abs(ax[0] - ax[1]) == want                  # rounds UP to n - 3
    assert np.floor((y[-2] - y[1]) / abs(y[0] - y[1])) == 8.0                 # stays in the last cell, n - 4
    Z, Y, X = [a.ravel() for a in np.meshgrid(z, y, x, indexing="ij")]
    field = np.stack([X, Y, Z, np.sin(X) + Y * Z, X * Y - Z, np.cos(X + Y) * Z], axis=1)
    for mode in ("vector", "norm", "both"):
        obj = tricubic(field.copy(), "quiet", mode=mode)
        ora = OracleInterp(field, 3, mode=mode)
        mid = np.array([0.1, 0.2, 0.9])
        q = np.array([[obj.xIntMax, mid[1], mid[2], 7.0],
                      [mid[0], obj.yIntMax, mid[2], 7.0],
                      [mid[0], mid[1], obj.zIntMax, 7.0],
                      [obj.xIntMax, obj.yIntMax, obj.zIntMax, 7.0],
                      [obj.xIntMin, obj.yIntMin, obj.zIntMin, 7.0]])
        q_in = q.copy()
        res = obj.Query(q_in)
        res = res if isinstance(res, tuple) else (res,)
        inds = obj.queryInds
        for r in res:
            assert np.isnan(r[[0, 2, 3]]).all() and np.isfinite(r[[1, 4]]).all()
        assert list(inds[[0, 2, 3]]) == [obj.nc] * 3 and inds[4] == 0
        assert np.array_equal(q_in, q)                                          # on the edge is not outside: rows stay
        # row 1 (y on its upper edge, index rounds down) and row 4 (all lower edges) agree with the oracle
        sub = q[[1, 4], :3].copy()
        ref = ora.query(sub)
        ref = ref if isinstance(ref, tuple) else (ref,)
        s = np.abs(field[:, 3:]).max() * 2
        for got, want_ in zip(res, ref):
            sc = s / np.array(ora.geo.h)[None, :] if (got is res[-1] and mode != "vector") else s
            assert_parity(got[[1, 4]], want_, sc, RTOL, f"upper edge {mode}")
        assert list(inds[[1, 4]]) == list(ora.query_inds)


def test_constructor_mode_switch_branches(capsys):
    """A.py:24-102 / 643-721: banner per branch, 'quiet' positional, quiet=True keyword, invalid mode -> vector,
    no mode -> vector, scalar input ignores the switch; A.py:104: wrong width exits."""
    from arbinterp_b200 import tricubic
    g = load_golden("tri_12x10x9")
    field = g["field"]
    expect = {
        "vector": "--- Vector field, interpolating for vector components --- ",
        "norm": "--- Vector field, interpolating for magnitude and gradient --- ",
        "both": "--- Vector field, interpolating vector components plus magnitude and gradient --- ",
        "bogus": "--- Vector field, invalid option, defaulting to interpolating for vector components --- ",
        None: "--- Vector field, no option selected, defaulting to interpolating for vector components --- ",
    }
    q = g["vector_q_in"]
    for mode, banner in expect.items():
        kw = {} if mode is None else {"mode": mode}
        obj = tricubic(field.copy(), **kw)
        assert capsys.readouterr().out.strip() == banner.strip()
        if mode in ("bogus", None, "vector"):                                  # all three answer like rQuery1
            out = obj.Query(q.copy())
            assert isinstance(out, np.ndarray) and out.shape == (len(q), 3)
            assert_parity(out, g["vector_out0"], np.abs(field[:, 3:]).max(), RTOL, f"mode={mode}")
            assert obj.Query.__func__ is type(obj).Query1
    tricubic(field.copy(), "quiet", mode="both")
    assert capsys.readouterr().out == ""
    tricubic(field.copy(), quiet=True)
    assert capsys.readouterr().out == ""
    gs = load_golden("tri_scalar_9x8x11")
    obj = tricubic(gs["field"].copy(), mode="vector")                           # README: switch ignored for scalar input
    assert capsys.readouterr().out.strip() == "--- Scalar field, ignoring switches, interpolating for magnitude and gradient ---"
    assert obj.Query.__func__ is type(obj).Query2
    with pytest.raises(SystemExit):
        tricubic(np.zeros((64, 5)), "quiet")


def test_explicit_vector_mode_stray_alpha_column():
    """A.py:42-45: mode='vector' given explicitly allocates alpha with nc + 2 columns, NaN in the last one only
    (column nc stays zero); the default / invalid-mode branches allocate nc + 1 (A.py:85, 99)."""
    from arbinterp_b200 import tricubic
    g = load_golden("tri_12x10x9")
    obj = tricubic(g["field"].copy(), "quiet", mode="vector")
    ref = g["vector_alphax"]
    got = obj.alphax
    assert got.shape == ref.shape == (64, obj.nc + 2)
    assert_parity(got[:, :obj.nc], ref[:, :obj.nc], np.abs(ref[:, :obj.nc]).max(), RTOL, "explicit vector alphax")
    assert np.array_equal(got[:, obj.nc:], ref[:, obj.nc:], equal_nan=True)
    assert tricubic(g["field"].copy(), "quiet").alphax.shape == (64, obj.nc + 1)
    assert tricubic(g["field"].copy(), "quiet", mode="nonsense").alphax.shape == (64, obj.nc + 1)


@pytest.mark.parametrize("name,d", [("tri_12x10x9", 3), ("quad_8x7x7x6", 4)])
@pytest.mark.parametrize("mode", ["vector", "norm", "both"])
def test_inputfield_kept_in_every_mode(name, d, mode):
    """A.py:14, 530-532: the sorted field is an attribute whatever the mode."""
    from arbinterp_b200 import tricubic, quadcubic
    g = load_golden(name)
    obj = (tricubic if d == 3 else quadcubic)(g["field"].copy(), "quiet", mode=mode)
    assert np.array_equal(obj.inputfield, g["sorted_field"])
    tf = (tricubic if d == 3 else quadcubic)(g["field"].copy(), "quiet", mode=mode, table=False)
    assert np.array_equal(tf.inputfield, g["sorted_field"])


def test_table_free_update_values_then_immediate_numpy_query():
    """update_values() writes the planes with torch ops on the current stream; a numpy Query right after
    runs on the library's own streams and must see the complete new planes (no construction in between)."""
    from arbinterp_b200 import tricubic
    n = 160
    ax = np.linspace(-1, 1, n)
    Z, Y, X = [a.ravel() for a in np.meshgrid(ax, ax, ax, indexing="ij")]
    base = np.stack([X, Y, Z, 1 + X * X + Y, 2 + Y * Z, 3 + Z * Z - X], axis=1)
    for mode in ("both", "vector"):
        obj = tricubic(base.copy(), "quiet", mode=mode, table=False)
        rng = np.random.default_rng(3)
        q = rng.uniform(-0.95, 0.95, (300_000, 3))
        for k in range(3):
            a = 1.0 + 0.5 * k
            vals = np.stack([a + X * X + Y, 2 * a + Y * Z, 3 * a + Z * Z - X], axis=1)
            obj.update_values(vals)
            res = obj.Query(q.copy())                                          # immediately, numpy path
            comps = res[0] if isinstance(res, tuple) else res
            want = np.stack([a + q[:, 0] ** 2 + q[:, 1], 2 * a + q[:, 1] * q[:, 2], 3 * a + q[:, 2] ** 2 - q[:, 0]], axis=1)
            assert np.abs(comps - want).max() < 1e-12                          # quadratics are reproduced to round-off


def test_load_rejects_inconsistent_headers(tmp_path):
    """a coefficient file's header is untrusted -- the table shape must follow from geometry, slab and mode,
    and the file must hold that many bytes, before anything is allocated."""
    import json
    from arbinterp_b200 import tricubic
    g = load_golden("tri_12x10x9")
    obj = tricubic(g["field"].copy(), "quiet", mode="norm")
    path = tmp_path / "t.arb"
    obj.save(str(path))
    raw = path.read_bytes()
    hlen = int(np.frombuffer(raw[8:16], dtype=np.uint64)[0])
    header = json.loads(raw[16:16 + hlen])

    def rewrite(mut, truncate=None):
        h = dict(header)
        mut(h)
        blob = json.dumps(h).encode()
        body = raw[(16 + hlen + 4095) // 4096 * 4096:]
        out = raw[:8] + np.uint64(len(blob)).tobytes() + blob + b"\0" * (-(16 + len(blob)) % 4096) + body
        p = tmp_path / "bad.arb"
        p.write_bytes(out if truncate is None else out[:truncate])
        return str(p)

    with pytest.raises(ValueError, match="table shape"):
        tricubic.load(rewrite(lambda h: h.update(table_shape=[1 << 40, 1, 64])))
    with pytest.raises(ValueError, match="table shape"):
        tricubic.load(rewrite(lambda h: h.update(mode="both")))
    with pytest.raises(ValueError, match="inconsistent geometry"):
        tricubic.load(rewrite(lambda h: h.update(slab=[0, 99])))
    with pytest.raises(ValueError, match="truncated"):
        tricubic.load(rewrite(lambda h: None, truncate=len(raw) - 4096))
    again = tricubic.load(rewrite(lambda h: None))                              # the untouched header still loads
    q = g["norm_q_in"].copy()
    a, b = again.Query(q.copy()), obj.Query(q.copy())
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(a, b))


def test_small_query_keeps_callers_current_device():
    """the latency path must not leave the process on the interpolator's device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from arbinterp_b200 import tricubic
    g = load_golden("tri_12x10x9")
    obj = tricubic(g["field"].copy(), "quiet", mode="norm", device="cuda:1")
    torch.cuda.set_device(0)
    obj.Query(g["norm_q_in"][:5].copy())
    obj.Query(g["norm_q_in"][0, :3].copy())
    assert torch.cuda.current_device() == 0


def test_single_process_devices_list():
    """tricubic(field, devices=[...]): one replica per listed GPU, numpy batches fanned out, results bit-identical to
    the single-GPU object (one GPU: the list form is accepted and equal; >= 2 GPUs: shares really run on both)."""
    from arbinterp_b200 import tricubic
    rng = np.random.default_rng(5)
    ax = np.linspace(-1, 1, 40)
    Z, Y, X = [a.ravel() for a in np.meshgrid(ax, ax, ax, indexing="ij")]
    field = np.stack([X, Y, Z, np.sin(2 * X) * np.cos(Y), X * Y + Z, np.cos(X + Y + Z)], axis=1)
    ndev = min(torch.cuda.device_count(), 4)
    one = tricubic(field.copy(), "quiet", mode="both")
    many = tricubic(field.copy(), "quiet", mode="both", devices=list(range(ndev)))
    assert len(many._replicas) == ndev
    q = rng.uniform(-1.02, 1.02, (700_001, 5))
    qa, qb = q.copy(), q.copy()
    ra, rb = one.Query(qa), many.Query(qb)
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(ra, rb))
    assert np.array_equal(qa, qb, equal_nan=True) and np.array_equal(one.queryInds, many.queryInds)
    vals = np.stack([X + 1, Y * Z, Z - X], axis=1)
    one.update_values(vals); many.update_values(vals)
    ra, rb = one.Query(q.copy()), many.Query(q.copy())
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(ra, rb))


def test_pageable_rows_staged_ahead(monkeypatch):
    """arb_query_host stages ordinary (pageable) numpy rows into its pinned ring on helper threads, chunks ahead of the
    one being issued (csrc/arb_host.cu, StageBuf / CopyPool).  A batch of several chunks with a ragged tail, extra
    query columns (README: ignored) and out-of-volume rows in the first / a middle / the last piece must come back
    bit-identical to the device-tensor path, with the NaN rows written into the caller's array (A.py:350-355)."""
    from arbinterp_b200 import tricubic
    rng = np.random.default_rng(11)
    ax = [np.linspace(-1, 1, 23), np.linspace(0, 1, 19), np.linspace(-2, 0, 21)]
    Z, Y, X = [a.ravel() for a in np.meshgrid(ax[2], ax[1], ax[0], indexing="ij")]
    field = np.stack([X, Y, Z, np.sin(2 * X) * np.cos(3 * Y) * np.exp(Z) + 2.0], axis=1)
    obj = tricubic(field, "quiet")
    n, ld = 2 * (1 << 20) + 777_777, 5
    lo = np.array([obj.xIntMin, obj.yIntMin, obj.zIntMin]); hi = np.array([obj.xIntMax, obj.yIntMax, obj.zIntMax])
    q = np.empty((n, ld))
    q[:, :3] = lo + rng.uniform(0, 1, (n, 3)) * (hi - lo) * (1 - 1e-9)
    q[:, 3:] = rng.normal(size=(n, 2))
    bad = [0, 5, 131_071, 131_072, (1 << 20) - 1, 1 << 20, (1 << 20) + 600_000, n - 2, n - 1]
    q[bad, 0] = 7.0
    dev_q = torch.from_numpy(q).cuda()
    ref = obj.Query(dev_q)
    ref_inds = obj.queryInds.copy()
    for chunk in ("0", "300000", "1048576"):
        monkeypatch.setenv("ARB_HOST_CHUNK_ROWS", chunk)
        qh = q.copy()
        got = obj.Query(qh)
        for a, b in zip(ref, got):
            assert np.array_equal(a.cpu().numpy(), b, equal_nan=True), f"chunk {chunk}"
        assert np.array_equal(obj.queryInds, ref_inds)
        assert np.isnan(qh[bad]).all() and np.array_equal(np.delete(qh, bad, axis=0), np.delete(q, bad, axis=0))
    assert torch.isnan(dev_q[bad]).all()


@pytest.mark.parametrize("d,mode", [(4, "both"), (3, "norm"), (3, "vector")])
def test_route_rows_and_inbox_query_virtual_ranks(d, mode):
    """Both legs of a slab-sharded query as kernels (arb_route_rows -> arb_query_inbox), exercised on ONE GPU with three
    virtual ranks whose inboxes, counts and result buffers all live on the same device: every sender's rows (one batch
    is empty, one has out-of-volume / NaN / layer-boundary rows) must come back at their home rows bit-identical to the
    unsharded table's answers, with the global cell index, and the out-of-volume mask must be the reference's bounds
    test (A.py:1069-1076).  The multi-process form over NVLink is tests/test_multi_gpu.py."""
    import ctypes
    from arbinterp_b200 import _lib, quadcubic, tricubic
    from test_gpu_parity import _analytic_field3, _analytic_field4, _uniform_queries
    rng = np.random.default_rng(90 + d)
    cls = tricubic if d == 3 else quadcubic
    field = _analytic_field3(14, 12, 16, rng=rng) if d == 3 else _analytic_field4(10, 9, 8, 13, rng=rng)
    whole = cls(field.copy(), "quiet", mode=mode)
    lib = whole._lib
    nslow = whole._geo.ncell[d - 1]
    W = 3
    his = [nslow // 3, 2 * nslow // 3, nslow]
    slabs = [(0 if r == 0 else his[r - 1], his[r]) for r in range(W)]
    parts = [cls(field.copy(), "quiet", mode=mode, slab=s) for s in slabs]
    sizes = [70_001, 0, 12_345]
    dev = torch.device("cuda", 0)
    batches = []
    for r, n in enumerate(sizes):
        q = _uniform_queries(whole, d, max(n, 1), rng)[:n]
        if n:
            q[::97, 0] = 9.0                                    # outside the volume in x: owner by t / z, NaN outputs
            q[5::211, d - 1] = np.nan                           # no layer: rank 0 answers NaN
            q[7::301, d - 1] = -50.0
            lo_s, h_s = whole._geo.int_min[d - 1], whole._geo.h[d - 1]
            for k, b in enumerate(his[:-1]):                    # exactly on the slab boundaries
                q[11 + k, d - 1] = lo_s + b * h_s
        batches.append(torch.from_numpy(q).to(dev))
    cap = max(sizes) + 100
    ld_in = (d + 2) // 2 * 2
    ncomp_out, ngrad = (0 if mode == "norm" else 3), (0 if mode == "vector" else 1 + d)
    ld = (ncomp_out + ngrad + 2) // 2 * 2
    inbox = [torch.full((W * cap, ld_in), float("nan"), dtype=torch.float64, device=dev) for _ in range(W)]
    counts = [torch.full((16,), -1, dtype=torch.int64, device=dev) for _ in range(W)]
    results = [torch.full((max(n, 1), ld), 123.0, dtype=torch.float64, device=dev) for n in sizes]
    vp = ctypes.c_void_p
    inbox_ptrs = (vp * W)(*[t.data_ptr() for t in inbox])
    count_ptrs = (vp * W)(*[t.data_ptr() for t in counts])
    result_ptrs = (vp * W)(*[t.data_ptr() for t in results])
    hi_arr = (ctypes.c_int64 * W)(*his)
    st = torch.cuda.current_stream(dev).cuda_stream
    outside = []
    for r in range(W):
        cursor = torch.zeros(W + 2, dtype=torch.int64, device=dev)
        out = torch.empty(sizes[r], dtype=torch.bool, device=dev)
        _lib.check(lib.arb_route_rows(ctypes.byref(whole._cgeom), batches[r].data_ptr(), sizes[r], d, hi_arr, W, r, inbox_ptrs,
                                      count_ptrs, cap, cursor.data_ptr(), cursor[W + 1:].data_ptr(), out.data_ptr(), st),
                   "arb_route_rows")
        outside.append(out)
    torch.cuda.synchronize()
    sent = torch.stack([c[:W] for c in counts])                 # [owner][sender]
    assert sent.sum(dim=0).tolist() == sizes and int(sent.min()) >= 0
    for o in range(W):
        _lib.check(lib.arb_query_inbox(ctypes.byref(parts[o]._cgeom), parts[o].table.data_ptr(), parts[o]._mode_code,
                                       inbox[o].data_ptr(), counts[o].data_ptr(), cap, result_ptrs, W, ld, st), "arb_query_inbox")
    torch.cuda.synchronize()
    for r, n in enumerate(sizes):
        if n == 0:
            assert bool((results[r] == 123.0).all())
            continue
        q = batches[r].clone()
        ref = whole.Query(q)
        ref = ref if isinstance(ref, tuple) else (ref,)
        flat = torch.cat([t.reshape(n, -1) for t in ref] + [whole._last_cells.view(torch.float64).unsqueeze(1)], dim=1)
        got = results[r][:, :flat.shape[1]]
        assert torch.equal(got.view(torch.int64), flat.view(torch.int64)), f"sender {r}"
        lo = torch.tensor(whole._geo.int_min, dtype=torch.float64, device=dev)
        hi = torch.tensor(whole._geo.int_max, dtype=torch.float64, device=dev)
        assert torch.equal(outside[r], ((batches[r] < lo) | (batches[r] > hi)).any(dim=1))
