"""CPU tier: property-based tests (hypothesis) of the host logic and of oracle invariants."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from arbinterp_b200.ingest import ingest_field, sorted_field
from arbinterp_b200.sharding import owner_ranks, plan_slabs
from oracle.arb_oracle import GridGeometry, OracleInterp


def _grid_rows(shape, origin, step, ncols, seed):
    rng = np.random.default_rng(seed)
    axes = [origin[a] + step[a] * np.arange(shape[a]) for a in range(len(shape))]
    mesh = np.meshgrid(*reversed(axes), indexing="ij")
    coords = [m.ravel() for m in reversed(mesh)]
    vals = [rng.normal(size=coords[0].size) for _ in range(ncols)]
    rows = np.stack(coords + vals, axis=1)
    return rows, rows[rng.permutation(len(rows))]


@settings(max_examples=40, deadline=None)
@given(d=st.sampled_from([3, 4]), data=st.data())
def test_ingest_is_permutation_invariant_and_matches_oracle_geometry(d, data):
    shape = [data.draw(st.integers(4, 7)) for _ in range(d)]
    origin = [data.draw(st.floats(-5, 5, allow_nan=False)) for _ in range(d)]
    step = [data.draw(st.sampled_from([1e-6, 0.125, 0.3, 1.0, 7.5])) for _ in range(d)]
    ncols = data.draw(st.sampled_from([1, 3]))
    rows, shuffled = _grid_rows(shape, origin, step, ncols, data.draw(st.integers(0, 2 ** 31)))
    planes, geo = ingest_field(shuffled, d)
    assert geo.npts == shape and planes.shape == tuple([ncols] + shape[::-1])
    assert np.array_equal(sorted_field(planes, geo).numpy(), rows)          # x-fastest sorted order (A.py:530-532)
    ref = GridGeometry(shuffled, d)
    assert geo.h == [float(v) for v in ref.h] and geo.int_min == [float(v) for v in ref.int_min]
    assert geo.int_max == [float(v) for v in ref.int_max] and geo.nc == ref.nc


@settings(max_examples=200, deadline=None)
@given(n_layers=st.integers(1, 500), world=st.integers(1, 16))
def test_plan_slabs_partitions_the_layers(n_layers, world):
    slabs = plan_slabs(n_layers, world)
    assert len(slabs) == world and slabs[0][0] == 0 and slabs[-1][1] == n_layers
    sizes = [hi - lo for lo, hi in slabs]
    assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:])) and max(sizes) - min(sizes) <= 1
    assert sizes == sorted(sizes, reverse=True)


@settings(max_examples=100, deadline=None)
@given(n_layers=st.integers(1, 60), world=st.integers(1, 8), seed=st.integers(0, 2 ** 31))
def test_owner_ranks_agree_with_brute_force(n_layers, world, seed):
    rng = np.random.default_rng(seed)
    slabs = plan_slabs(n_layers, world)
    t0, h = rng.uniform(-3, 3), rng.choice([0.01, 0.5, 2.0])
    t1 = t0 + n_layers * h
    t = rng.uniform(t0 - 2 * h, t1 + 2 * h, 300)
    t[:3] = [t0, t1, np.nan]
    own = owner_ranks(torch.from_numpy(t), t0, t1, h, slabs).numpy()
    for ti, oi in zip(t, own):
        if not (ti >= t0 and ti <= t1):
            assert oi == 0
            continue
        layer = min(int(np.floor((ti - t0) / h)), n_layers - 1)
        assert slabs[oi][0] <= layer < slabs[oi][1]


@settings(max_examples=15, deadline=None)
@given(seed=st.integers(0, 2 ** 31), d=st.sampled_from([3, 4]))
def test_oracle_interpolant_is_linear_in_the_field_and_exact_for_constants(seed, d):
    rng = np.random.default_rng(seed)
    shape = [5, 6, 5, 5][:d]
    rows_f, _ = _grid_rows(shape, [0.0] * d, [0.5] * d, 1, seed)
    rows_g, _ = _grid_rows(shape, [0.0] * d, [0.5] * d, 1, seed + 1)
    comb = rows_f.copy(); comb[:, d] = 2.0 * rows_f[:, d] - 3.0 * rows_g[:, d]
    const = rows_f.copy(); const[:, d] = 4.25
    lo = np.array([0.5] * d); hi = 0.5 * (np.array(shape) - 2)
    q = lo + rng.uniform(0, 1, (64, d)) * (hi - lo) * 0.999
    (nf, gf), (ng, gg), (nc, gc), (nk, gk) = [OracleInterp(r, d).query(q.copy()) for r in (rows_f, rows_g, comb, const)]
    # white-noise fields have coefficients of order 1e2 (|inv(B)| entries up to 27 / 81), hence the absolute slack
    assert np.allclose(nc, 2 * nf - 3 * ng, rtol=0, atol=1e-10) and np.allclose(gc, 2 * gf - 3 * gg, rtol=0, atol=1e-9)
    assert np.allclose(nk, 4.25, rtol=0, atol=1e-12) and np.allclose(gk, 0.0, rtol=0, atol=1e-11)
