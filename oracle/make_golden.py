"""Generate tests/golden/*.npz by running the LIVE reference (unmodified) in the build
container.  Run as ``python oracle/make_golden.py`` -- needs /root/reference; the tests
and the GPU box only ever read the committed .npz files.

Every fixture stores the exact input field (rows shuffled: the reference sorts them,
A.py:530-532), the query array before the call, and everything the reference returned or
left behind (outputs, the in-place NaN-masked query array, ``queryInds``, geometry, ``A``
and the coefficient arrays after ``allCoeffs``).
"""
import os
import sys
import warnings

import numpy as np

REF = os.environ.get("ARB_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))
from ARBInterp.ARBInterp import tricubic, quadcubic  # noqa: E402  (the reference itself)

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
warnings.filterwarnings("ignore")


def field3(nx, ny, nz, lo=(-1, -.5, 0), hi=(1, .5, 2), scalar=False, seed=0):
    x = np.linspace(lo[0], hi[0], nx); y = np.linspace(lo[1], hi[1], ny); z = np.linspace(lo[2], hi[2], nz)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    X, Y, Z = X.ravel(), Y.ravel(), Z.ravel()
    sx, sy, sz = [2.0 / (h - l) for l, h in zip(lo, hi)]
    bx = np.sin(2 * X * sx) * np.cos(3 * Y * sy) * np.exp(-Z * sz * 0.5)
    by = (X * sx) ** 2 * (Y * sy) + Z * sz
    bz = np.cos(X * sx + Y * sy + Z * sz)
    f = np.stack([X, Y, Z, bx, by, bz], axis=1)
    if scalar:
        f = np.concatenate([f[:, :3], np.linalg.norm(f[:, 3:], axis=1)[:, None]], axis=1)
    return f[np.random.default_rng(seed).permutation(len(f))]


def field4(nx, ny, nz, nt, lo=(-1, -.5, 0, 0), hi=(1, .5, 2, 1), scalar=False, seed=0):
    ax = [np.linspace(l, h, n) for l, h, n in zip(lo, hi, (nx, ny, nz, nt))]
    T, Z, Y, X = np.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")
    X, Y, Z, T = X.ravel(), Y.ravel(), Z.ravel(), T.ravel()
    s = [2.0 / (h - l) for l, h in zip(lo, hi)]
    xs, ys, zs, ts = X * s[0], Y * s[1], Z * s[2] * 0.5, T * s[3] * 0.5
    bx = np.sin(2 * xs) * np.cos(3 * ys) * np.exp(-zs) * np.cos(2 * ts)
    by = xs ** 2 * ys + zs * ts + 0.3 * xs * ys * zs * ts      # xyzt term exercises the A.py:860 quirk
    bz = np.cos(xs + ys + zs + ts)
    f = np.stack([X, Y, Z, T, bx, by, bz], axis=1)
    if scalar:
        f = np.concatenate([f[:, :4], np.linalg.norm(f[:, 4:], axis=1)[:, None]], axis=1)
    return f[np.random.default_rng(seed).permutation(len(f))]


def make_queries(obj, d, n, ncols, seed):
    """Uniform in-volume points, plus out-of-volume / NaN / inf rows and extra columns.
    Exact-upper-edge points are excluded (the reference's behaviour there is
    float-luck dependent)."""
    rng = np.random.default_rng(seed)
    names = "xyzt"[:d]
    lo = np.array([getattr(obj, c + "IntMin") for c in names]); hi = np.array([getattr(obj, c + "IntMax") for c in names])
    q = rng.uniform(0, 1, (n, ncols))
    q[:, :d] = lo + q[:, :d] * (hi - lo) * (1 - 1e-9)
    # exact lower edges are legal, exercise them
    q[0, :d] = lo
    q[1, 0] = lo[0]
    bad = rng.choice(np.arange(2, n), size=max(8, n // 10), replace=False)
    for j, r in enumerate(bad):
        a = j % d
        kind = j % 6
        q[r, a] = [lo[a] - 1e-6 * (hi[a] - lo[a]), hi[a] + 1e-6 * (hi[a] - lo[a]), np.nan, np.inf, -np.inf,
                   hi[a] + 10 * (hi[a] - lo[a])][kind]
    return q


def run_case(name, cls, d, field, modes, nq, ncols, seed, single_pts=3, store_alpha=("n",)):
    out = {"field": field, "d": d}
    for mode in modes:
        kw = {} if mode == "scalar" else {"mode": mode}
        obj = cls(field.copy(), "quiet", **kw)
        q0 = make_queries(obj, d, nq, ncols, seed)
        q = q0.copy()
        res = obj.Query(q)
        res = res if isinstance(res, tuple) else (res,)
        pre = f"{mode}_"
        out[pre + "q_in"] = q0
        out[pre + "q_after"] = q
        out[pre + "inds"] = obj.queryInds.astype(np.int64)
        for i, r in enumerate(res):
            out[pre + f"out{i}"] = r
        # single-point queries (sQuery path, A.py:213-342 / 916-1062) on in-volume rows
        good = np.where(~np.isnan(q[:, 0]))[0][:single_pts]
        sres = []
        for r in good:
            s = obj.Query(q0[r, :d].copy())
            s = s if isinstance(s, tuple) else (s,)
            sres.append(np.concatenate([np.atleast_1d(np.asarray(v, dtype=float)).ravel() for v in s]))
        out[pre + "single_rows"] = good
        out[pre + "single_out"] = np.array(sres)
        obj.allCoeffs()
        for k in store_alpha:
            if hasattr(obj, "alpha" + k):
                out[pre + "alpha" + k] = getattr(obj, "alpha" + k)
        if mode == modes[0]:
            names = "xyzt"[:d]
            out["h"] = np.array([getattr(obj, "h" + c) for c in names])
            out["int_min"] = np.array([getattr(obj, c + "IntMin") for c in names])
            out["int_max"] = np.array([getattr(obj, c + "IntMax") for c in names])
            out["ncell_axis"] = (np.array(obj.nPos) if d == 3 else
                                 np.array([obj.nPosx - 3, obj.nPosy - 3, obj.nPosz - 3, obj.nPost - 3]))
            out["nc"] = obj.nc
            out["base_point_inds"] = obj.basePointInds.astype(np.int64)
            out["sorted_field"] = obj.inputfield
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **out)
    print(name, os.path.getsize(path) // 1024, "KiB")


def main():
    os.makedirs(OUT, exist_ok=True)
    # constant matrices: the reference's CSV fixtures (bit-equal to B in makeAMatrix) and A.
    B3 = np.genfromtxt(os.path.join(REF, "examples", "B_Matrix_3D.csv"), delimiter=",")
    B4 = np.genfromtxt(os.path.join(REF, "examples", "B_Matrix_4D.csv"), delimiter=",")
    t = tricubic(field3(5, 5, 5), "quiet"); qd = quadcubic(field4(5, 5, 5, 5), "quiet")
    assert np.array_equal(B3, np.rint(B3)) and np.array_equal(B4, np.rint(B4))
    assert np.array_equal(t.A * 8, np.rint(t.A * 8)) and np.array_equal(qd.A * 16, np.rint(qd.A * 16))
    np.savez_compressed(os.path.join(OUT, "matrices.npz"),
                        B3=B3.astype(np.int8), B4=B4.astype(np.int8),
                        A3_times8=np.rint(t.A * 8).astype(np.int16), A4_times16=np.rint(qd.A * 16).astype(np.int16))
    run_case("tri_12x10x9", tricubic, 3, field3(12, 10, 9), ["vector", "norm", "both"], 400, 5, 11,
             store_alpha=("x", "n"))
    run_case("tri_scalar_9x8x11", tricubic, 3, field3(9, 8, 11, scalar=True, seed=3), ["scalar"], 300, 3, 12)
    # physical-units stand-in for Example3DScalarField: mm-scale coordinates,
    # queried with the example script's diagonal (examples/3D_ARBInterpExample.py:18-22)
    f = field3(15, 15, 15, lo=(-3e-3,) * 3, hi=(3e-3,) * 3, scalar=True, seed=5)
    obj = tricubic(f.copy(), "quiet")
    coords = np.zeros((20, 3))
    for a in range(3):
        coords[:, a] = np.linspace(-2e-3, 2e-3, 20)
    single = obj.Query(coords[3].copy())
    norms, grads = obj.Query(coords.copy())
    np.savez_compressed(os.path.join(OUT, "tri_example_diag.npz"), field=f, coords=coords, norms=norms, grads=grads,
                        single_norm=np.float64(single[0]), single_grad=single[1], inds=obj.queryInds.astype(np.int64))
    run_case("quad_8x7x7x6", quadcubic, 4, field4(8, 7, 7, 6), ["vector", "norm", "both"], 300, 6, 21,
             store_alpha=("n",))
    run_case("quad_scalar_6x7x5x6", quadcubic, 4, field4(6, 7, 5, 6, scalar=True, seed=4), ["scalar"], 200, 4, 22)


if __name__ == "__main__":
    main()
