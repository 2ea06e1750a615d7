"""CPU tier: the numpy oracle against the golden vectors generated from the live reference
(oracle/make_golden.py).  Bit-exact: the oracle evaluates the reference's expressions in the same
order, so outputs, in-place NaN masks and cell indices must be identical (==)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle.arb_oracle import (OracleInterp, a_matrix, derivative_subsets, difference_matrix, hermite_matrix)

CASES = [("tri_12x10x9", 3, ["vector", "norm", "both"]), ("tri_scalar_9x8x11", 3, ["scalar"]),
         ("quad_8x7x7x6", 4, ["vector", "norm", "both"]), ("quad_scalar_6x7x5x6", 4, ["scalar"])]


def test_matrices_match_reference_fixtures():
    m = load_golden("matrices")
    # B == examples/B_Matrix_{3,4}D.csv of the reference (README :74-76), stored as int8
    assert np.array_equal(hermite_matrix(3), m["B3"].astype(np.float64))
    assert np.array_equal(hermite_matrix(4), m["B4"].astype(np.float64))
    # A == reference's inv(B) @ D, exactly (multiples of 1/8 and 1/16)
    assert np.array_equal(a_matrix(3) * 8, m["A3_times8"].astype(np.float64))
    assert np.array_equal(a_matrix(4) * 16, m["A4_times16"].astype(np.float64))


def test_a3_is_catmull_rom_kronecker():
    M = np.array([[0, 1, 0, 0], [-.5, 0, .5, 0], [1, -2.5, 2, -.5], [-.5, 1.5, -1.5, .5]])
    assert np.array_equal(a_matrix(3), np.kron(M, np.kron(M, M)))


def test_a4_quirk_structure():
    """A.py:860: row 240 of D is empty; the reference A differs from M^(x4) by a rank-16 term."""
    M = np.array([[0, 1, 0, 0], [-.5, 0, .5, 0], [1, -2.5, 2, -.5], [-.5, 1.5, -1.5, .5]])
    k4 = np.kron(M, np.kron(M, np.kron(M, M)))
    D = difference_matrix(4)
    assert not D[240].any() and D[241:].any(axis=1).all()
    assert np.array_equal(a_matrix(4, reference_quirk=False), k4)
    assert np.linalg.matrix_rank(a_matrix(4) - k4) == 16


def test_derivative_order():
    assert derivative_subsets(3) == [(), (0,), (1,), (2,), (0, 1), (0, 2), (1, 2), (0, 1, 2)]
    assert len(derivative_subsets(4)) == 16 and derivative_subsets(4)[5:11] == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]


@pytest.mark.parametrize("name,d,modes", CASES)
def test_geometry(name, d, modes):
    g = load_golden(name)
    geo = OracleInterp(g["field"], d, mode="vector").geo
    assert np.array_equal(geo.h, g["h"]) and np.array_equal(geo.int_min, g["int_min"])
    assert np.array_equal(geo.int_max, g["int_max"]) and np.array_equal(geo.ncell_axis, g["ncell_axis"])
    assert geo.nc == int(g["nc"]) and np.array_equal(geo.base_point_inds, g["base_point_inds"])
    assert np.array_equal(geo.sorted, g["sorted_field"])


@pytest.mark.parametrize("name,d,modes", CASES)
def test_range_query_bit_exact(name, d, modes):
    g = load_golden(name)
    for mode in modes:
        o = OracleInterp(g["field"], d, mode="vector" if mode == "scalar" else mode)
        q = g[mode + "_q_in"].copy()
        res = o.query(q, exact_gemv=True)
        res = res if isinstance(res, tuple) else (res,)
        for i, r in enumerate(res):
            assert np.array_equal(r, g[f"{mode}_out{i}"], equal_nan=True), (name, mode, i)
        assert np.array_equal(q, g[mode + "_q_after"], equal_nan=True)      # in-place NaN rows (A.py:350-355)
        assert np.array_equal(o.query_inds, g[mode + "_inds"])


@pytest.mark.parametrize("name,d,modes", CASES)
def test_coefficient_table_bit_exact(name, d, modes):
    g = load_golden(name)
    mode = modes[-1]
    o = OracleInterp(g["field"], d, mode="vector" if mode == "scalar" else mode)
    o.all_coeffs(exact_gemv=True)
    for k in "xyzn":
        key = f"{mode}_alpha{k}"
        if key in g.files:
            assert np.array_equal(o.alpha[k], g[key], equal_nan=True), key


def test_batched_coefficients_close_to_gemv():
    """The fast (dgemm) coefficient path used for large oracle runs agrees with the per-cell dgemv."""
    g = load_golden("tri_12x10x9")
    a = OracleInterp(g["field"], 3, mode="norm"); a.all_coeffs(exact_gemv=True)
    b = OracleInterp(g["field"], 3, mode="norm"); b.all_coeffs()
    assert np.nanmax(np.abs(a.alpha["n"] - b.alpha["n"])) < 1e-12


def test_compact_store_matches_dense():
    g = load_golden("quad_8x7x7x6")
    q = g["both_q_in"]
    dense = OracleInterp(g["field"], 4, mode="both").query(q.copy())
    compact = OracleInterp(g["field"], 4, mode="both", dense=False).query(q.copy())
    for a, b in zip(dense, compact):
        assert np.array_equal(a, b, equal_nan=True)


def test_example_diagonal_fixture():
    """Config 1 stand-in: the example script's 20-point diagonal on a mm-scale scalar field."""
    g = load_golden("tri_example_diag")
    o = OracleInterp(g["field"], 3)
    norms, grads = o.query(g["coords"].copy(), exact_gemv=True)
    assert np.array_equal(norms, g["norms"]) and np.array_equal(grads, g["grads"])
    assert np.array_equal(o.query_inds, g["inds"])


def test_quadratic_field_reproduced():
    """Analytic tier: central differences are exact for per-axis quadratics, so the
    3-D interpolant reproduces them to round-off -- a check that needs no reference."""
    x = np.linspace(-1, 1, 9); y = np.linspace(0, 2, 8); z = np.linspace(-2, -1, 7)
    Z, Y, X = np.meshgrid(z, y, x, indexing="ij")
    f = lambda X, Y, Z: 1 + X - 2 * Y + 3 * Z + X * Y - Y * Z + X * X * Z + Y * Y - 0.5 * Z * Z * X
    field = np.stack([X.ravel(), Y.ravel(), Z.ravel(), f(X, Y, Z).ravel()], axis=1)
    o = OracleInterp(field, 3)
    rng = np.random.default_rng(0)
    q = np.stack([rng.uniform(x[1], x[-2], 200), rng.uniform(y[1], y[-2], 200), rng.uniform(z[1], z[-2], 200)], axis=1)
    norms, grads = o.query(q.copy())
    assert np.max(np.abs(norms[:, 0] - f(q[:, 0], q[:, 1], q[:, 2]))) < 1e-13
    gx = 1 + q[:, 1] + 2 * q[:, 0] * q[:, 2] - 0.5 * q[:, 2] ** 2
    assert np.max(np.abs(grads[:, 0] - gx)) < 1e-12
