"""CPU oracle for the ARBInterp interpolation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``arbinterp_b200/`` may import this
module; it is used by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` as the checker and
as the timed CPU arm -- never as the product path.

It is a numpy restatement of the reference algorithm
(``/root/reference/src/ARBInterp/ARBInterp.py``, abbreviated ``A.py`` below),
written dimension-generic (d = 3 tricubic, d = 4 quadcubic) instead of the
reference's two near-duplicate classes.  Each function cites the reference
lines it follows.

Parity pin: the reference's own tests hold no golden vectors for this path
(``tests/test_ARBInterp.py:1`` is a bare import), so the oracle is pinned against
(i) ``examples/B_Matrix_3D.csv`` / ``B_Matrix_4D.csv`` (the Hermite matrix, bit
equal) and (ii) outputs of the live reference run in the build container,
committed as ``tests/golden/*.npz`` by ``oracle/make_golden.py``.  The range
query below evaluates the same floating-point expressions in the same order
and array layouts as the reference, so it reproduces the golden outputs
bit-for-bit (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import itertools

import numpy as np

__all__ = [
    "hermite_matrix", "difference_matrix", "a_matrix", "GridGeometry",
    "OracleInterp", "derivative_subsets",
]


# --------------------------------------------------------------------------------------
# Constant matrices (A.py:107-175, A.py:726-878)
# --------------------------------------------------------------------------------------

def derivative_subsets(d: int):
    """Derivative types of the b-vector, in the reference's order.

    3D: f, fx, fy, fz, fxy, fxz, fyz, fxyz (A.py:118-125).
    4D: f, fx, fy, fz, ft, fxy, fxz, fxt, fyz, fyt, fzt, fxyz, fxyt, fxzt, fyzt,
    fxyzt (A.py:738-757).  Both are "subsets of the axes, by size then
    lexicographic".
    """
    out = []
    for size in range(d + 1):
        out.extend(itertools.combinations(range(d), size))
    return out


def hermite_matrix(d: int) -> np.ndarray:
    """Hermite constraint matrix ``B`` (4^d x 4^d): row ``r*2^d + c`` is derivative
    type r of monomial m evaluated at unit-cube corner c (A.py:110-125, 729-757).

    Corner order and monomial order are x fastest (A.py:110-111, 729-730); the
    derivative of ``x^e`` is written ``e * x^|e-1|`` as at A.py:119.
    """
    ncorner, nmono = 2 ** d, 4 ** d
    corners = np.array([[(c >> a) & 1 for a in range(d)] for c in range(ncorner)], dtype=np.float64)
    expo = np.array([[(m >> (2 * a)) & 3 for a in range(d)] for m in range(nmono)], dtype=np.int64)
    B = np.zeros((nmono, nmono), dtype=np.float64)
    for r, subset in enumerate(derivative_subsets(d)):
        for c in range(ncorner):
            row = np.ones(nmono, dtype=np.float64)
            for a in range(d):
                xa = corners[c, a]
                e = expo[:, a]
                if a in subset:
                    row = row * (e * xa ** np.abs(e - 1))
                else:
                    row = row * xa ** e
            B[r * ncorner + c] = row
    return B


def difference_matrix(d: int, reference_quirk: bool = True) -> np.ndarray:
    """Finite-difference matrix ``D`` (4^d x 4^d) mapping the 4^d neighbourhood
    values (flat index i + 4j + 16k [+ 64l], offsets -1..+2) to the b-vector in
    unit-cell coordinates (A.py:129-173, 762-876).

    ``reference_quirk`` reproduces A.py:860: the quadruple-mixed rows are filled by
    ``enumerate(range(241, 256))`` so row 240 stays zero and row ``241 + i`` uses the
    stencil centre of corner ``i`` (not ``i + 1``).  Parity with the reference requires
    it; ``False`` gives the mathematically intended matrix.
    """
    ncorner, nmono = 2 ** d, 4 ** d
    strides = [4 ** a for a in range(d)]
    base = sum(strides)  # 21 (3D) / 85 (4D): offset (0,0,..) of the cell's lower corner
    centres = [base + sum(((c >> a) & 1) * strides[a] for a in range(d)) for c in range(ncorner)]
    D = np.zeros((nmono, nmono), dtype=np.float64)
    for r, subset in enumerate(derivative_subsets(d)):
        weight = 0.5 ** len(subset)
        for c in range(ncorner):
            row = r * ncorner + c
            centre = centres[c]
            if d == 4 and len(subset) == 4 and reference_quirk:
                if c == 0:
                    continue            # row 240 never written (A.py:860)
                centre = centres[c - 1]  # row 241+i uses C[i]
            for signs in itertools.product((-1, 1), repeat=len(subset)):
                off = sum(s * strides[a] for s, a in zip(signs, subset))
                D[row, centre + off] = weight * np.prod(signs) if subset else 1.0
    return D


def a_matrix(d: int, reference_quirk: bool = True) -> np.ndarray:
    """``A = inv(B) @ D`` (A.py:175, A.py:878)."""
    return np.matmul(np.linalg.inv(hermite_matrix(d)), difference_matrix(d, reference_quirk))


# --------------------------------------------------------------------------------------
# Grid geometry (A.py:528-568, A.py:1264-1320)
# --------------------------------------------------------------------------------------

class GridGeometry:
    """Sorted field + derived geometry, as ``getFieldParams`` computes it."""

    def __init__(self, field: np.ndarray, d: int):
        self.d = d
        f = np.asarray(field)
        # Three/four successive sorts, x first (A.py:530-532, 1266-1269).  The first is
        # numpy's default (unstable) sort, the rest are stable; on a full grid the result
        # is the unique (t,z,y,x)-lexicographic order, which lexsort gives directly.
        order = np.lexsort(tuple(f[:, a] for a in range(d)))
        self.sorted = f[order]
        s = self.sorted
        # Points per axis (A.py:535-543, 1272-1291): on a full grid the reference's
        # equality scans count the distinct coordinates per axis.
        self.npts = []
        for a in range(d):
            others = [b for b in range(d) if b != a]
            sel = np.ones(len(s), dtype=bool)
            for b in others:
                sel &= s[:, b] == s[0, b]
            self.npts.append(int(sel.sum()))
        n = self.npts
        self.ncell_axis = [k - 3 for k in n]                      # A.py:545, 1288-1291 (used as n-3)
        stride = [int(np.prod(n[:a])) for a in range(d)]          # row stride of axis a in the sorted field
        self.stride = stride
        # h = |axis[0] - axis[1]| (A.py:547-549, 1293-1296)
        self.h = [np.abs(s[0, a] - s[stride[a], a]) for a in range(d)]
        # IntMin = 2nd grid value, IntMax = 2nd-to-last (A.py:551-556, 1298-1305)
        self.int_min = [s[stride[a], a] for a in range(d)]
        self.int_max = [s[-2 * stride[a], a] for a in range(d)]
        # basePointInds: flat row of each cell's lower corner (A.py:559-565, 1308-1317)
        base = sum(stride)
        grids = np.meshgrid(*[np.arange(self.ncell_axis[a]) * stride[a] for a in reversed(range(d))], indexing="ij")
        self.base_point_inds = (base + sum(grids)).ravel().astype(np.int64)
        self.nc = len(self.base_point_inds)                       # A.py:568, 1320

    @classmethod
    def from_axes(cls, axes, with_base=True):
        """Geometry of an already sorted full grid given its per-axis coordinates: the same
        expressions as ``__init__`` without materialising and sorting the (N, d+k) row array
        (used by bench.py at 256^3, where the sort alone takes minutes in numpy).
        ``with_base=False`` skips ``basePointInds`` (only ``locate`` is wanted: config 5 has 49 M cells)."""
        self = cls.__new__(cls)
        d = self.d = len(axes)
        self.sorted = None
        self.npts = [len(a) for a in axes]
        n = self.npts
        self.ncell_axis = [k - 3 for k in n]
        self.stride = [int(np.prod(n[:a])) for a in range(d)]
        self.h = [np.abs(axes[a][0] - axes[a][1]) for a in range(d)]
        self.int_min = [axes[a][1] for a in range(d)]
        self.int_max = [axes[a][-2] for a in range(d)]
        if not with_base:
            self.base_point_inds = None
            self.nc = int(np.prod(self.ncell_axis))              # A.py:568, 1320: len(basePointInds)
            return self
        base = sum(self.stride)
        grids = np.meshgrid(*[np.arange(self.ncell_axis[a]) * self.stride[a] for a in reversed(range(d))], indexing="ij")
        self.base_point_inds = (base + sum(grids)).ravel().astype(np.int64)
        self.nc = len(self.base_point_inds)
        return self

    def neighbour_inds(self, ind0):
        """Flat row indices of the 4^d block around lower-corner row(s) ``ind0``
        (A.py:605-621, 1357-1381); x fastest, offsets -1..+2.  Vectorised over ind0."""
        ind0 = np.asarray(ind0, dtype=np.int64)
        start = ind0 - sum(self.stride)
        offs = np.zeros(1, dtype=np.int64)
        for a in range(self.d):
            offs = (offs[None, :] + (np.arange(4, dtype=np.int64) * self.stride[a])[:, None]).ravel()
        return start[..., None] + offs


# --------------------------------------------------------------------------------------
# Interpolator (A.py:10-104, 344-521, 570-603; 629-723, 1064-1258, 1322-1355)
# --------------------------------------------------------------------------------------

class OracleInterp:
    """numpy restatement of ``tricubic`` (d=3) / ``quadcubic`` (d=4).

    ``mode``: 'vector' | 'norm' | 'both' for (d+3)-column fields; (d+1)-column fields are
    scalar and behave like 'norm' (A.py:24-32, 643-651).
    """

    def __init__(self, field, d: int = 3, mode: str = "vector", reference_quirk: bool = True,
                 dense: bool = True):
        self.d = d
        self.geo = GridGeometry(field, d)
        s = self.geo.sorted
        ncol = s.shape[1]
        if ncol == d + 1:
            self.mode = "norm"
            self.scalar_input = True
            self.values = {"n": s[:, d]}                                    # A.py:32, 651
        elif ncol == d + 3:
            self.mode = mode if mode in ("vector", "norm", "both") else "vector"  # A.py:75-102
            self.scalar_input = False
            self.values = {}
            if self.mode in ("vector", "both"):
                self.values.update(x=s[:, d], y=s[:, d + 1], z=s[:, d + 2])  # A.py:46-48
            if self.mode in ("norm", "both"):
                self.values["n"] = np.linalg.norm(s[:, d:], axis=1)          # A.py:58, 74
        else:
            raise SystemExit("--- Input not shaped as expected ---")        # A.py:104, 723
        self.A = a_matrix(d, reference_quirk)
        self.nmono = 4 ** d
        nc = self.geo.nc
        self.dense = dense
        if dense:
            # alpha[4^d, nc+1], last column NaN except for scalar input (A.py:31, 45, 57, 70)
            self.alpha = {k: np.zeros((self.nmono, nc + 1)) for k in self.values}
            if not self.scalar_input:
                for k in self.alpha:
                    self.alpha[k][:, -1] = np.nan
            self.alphamask = np.zeros(nc + 1, dtype=bool)
            self.alphamask[-1] = True                                        # A.py:21-22
        else:
            # compact store for grids whose dense table does not fit host RAM:
            # sorted cell ids + coefficient columns; the arithmetic is unchanged.
            self._cells = np.zeros(0, dtype=np.int64)
            self._cols = {k: np.zeros((self.nmono, 0)) for k in self.values}
            self._have = np.zeros(nc + 1, dtype=bool)
            self._have[-1] = True

    @classmethod
    def from_planes(cls, axes, values, mode, scalar_input=False, reference_quirk=True, dense=False):
        """Oracle over a pre-sorted dense grid: ``values`` maps component key ('x','y','z','n') to the
        flattened value plane (x fastest).  Same arithmetic as the row-array constructor."""
        self = cls.__new__(cls)
        self.d = len(axes)
        self.geo = GridGeometry.from_axes(axes)
        self.mode, self.scalar_input = mode, scalar_input
        self.values = {k: np.asarray(v).reshape(-1) for k, v in values.items()}
        self.A = a_matrix(self.d, reference_quirk)
        self.nmono = 4 ** self.d
        self.dense = dense
        nc = self.geo.nc
        if dense:
            self.alpha = {k: np.zeros((self.nmono, nc + 1)) for k in self.values}
            if not scalar_input:
                for k in self.alpha:
                    self.alpha[k][:, -1] = np.nan
            self.alphamask = np.zeros(nc + 1, dtype=bool)
            self.alphamask[-1] = True
        else:
            self._cells = np.zeros(0, dtype=np.int64)
            self._cols = {k: np.zeros((self.nmono, 0)) for k in self.values}
            self._have = np.zeros(nc + 1, dtype=bool)
            self._have[-1] = True
        return self

    @classmethod
    def locator(cls, axes):
        """An oracle that can only ``locate`` (bounds mask, global cell index, fractions; A.py:350-373, 1069-1092) on
        the geometry of the given axes -- for full-size index checks where no host can hold the coefficients."""
        self = cls.__new__(cls)
        self.d = len(axes)
        self.geo = GridGeometry.from_axes([np.asarray(a, dtype=np.float64) for a in axes], with_base=False)
        return self

    # ---- coefficients ------------------------------------------------------------
    def _coeff_columns(self, cells, exact_gemv=False):
        """alpha[:, cell] = A . values[neighbourhood] (A.py:573-579, 1325-1331)."""
        inds = self.geo.neighbour_inds(self.geo.base_point_inds[cells])      # (n, 4^d)
        out = {}
        for k, v in self.values.items():
            if exact_gemv:   # one dgemv per cell, the reference's exact BLAS call pattern
                col = np.empty((self.nmono, len(cells)))
                for j in range(len(cells)):
                    col[:, j] = np.dot(self.A, v[inds[j]])
                out[k] = col
            else:
                out[k] = self.A @ v[inds].T
        return out

    def calc_coefficients(self, cells, exact_gemv=False, chunk=8192):
        cells = np.asarray(cells, dtype=np.int64)
        # the reference's cheap "already there?" test first (A.py:376: alphamask[queryInds] == 0)
        have = self.alphamask if self.dense else self._have
        cells = cells[~have[cells]]
        if len(cells) == 0:
            return
        cells = np.unique(cells)
        cells = cells[cells < self.geo.nc]
        if self.dense:
            for lo in range(0, len(cells), chunk):
                c = cells[lo:lo + chunk]
                cols = self._coeff_columns(c, exact_gemv)
                for k in cols:
                    self.alpha[k][:, c] = cols[k]
                self.alphamask[c] = True
        else:
            new = np.setdiff1d(cells, self._cells, assume_unique=True)
            if len(new):
                parts = {k: [self._cols[k]] for k in self.values}
                for lo in range(0, len(new), chunk):
                    cols = self._coeff_columns(new[lo:lo + chunk], exact_gemv)
                    for k in cols:
                        parts[k].append(cols[k])
                self._have[new] = True
                allc = np.concatenate([self._cells, new])
                order = np.argsort(allc, kind="stable")
                self._cells = allc[order]
                for k in self.values:
                    self._cols[k] = np.concatenate(parts[k], axis=1)[:, order]

    def all_coeffs(self, exact_gemv=False):
        """A.py:523-525 / 1260-1262."""
        self.calc_coefficients(np.arange(self.geo.nc), exact_gemv)

    def _gather(self, key, inds):
        """``alpha[:, queryInds].T`` (A.py:385-390) -- same (N, 4^d) F-ordered view."""
        if self.dense:
            return self.alpha[key][:, inds].T
        nc = self.geo.nc
        slot = np.searchsorted(self._cells, np.minimum(inds, nc - 1))
        cols = self._cols[key][:, np.minimum(slot, max(len(self._cells) - 1, 0))]
        if (inds == nc).any():
            cols[:, inds == nc] = 0.0 if self.scalar_input else np.nan
        return cols.T

    # ---- range query ---------------------------------------------------------------
    def locate(self, query):
        """Bounds mask (in place), cell index and cell-fraction coordinates
        (A.py:350-373, 1069-1092)."""
        g, d = self.geo, self.d
        for a in range(d):
            query[np.where(query[:, a] < g.int_min[a])[0]] = np.nan
            query[np.where(query[:, a] > g.int_max[a])[0]] = np.nan
        iu = [(query[:, a] - g.int_min[a]) / g.h[a] for a in range(d)]
        ii = [np.floor(u) for u in iu]
        inds = ii[0]
        mult = 1
        for a in range(1, d):
            mult = mult * g.ncell_axis[a - 1]
            inds = inds + ii[a] * mult
        with np.errstate(invalid="ignore"):
            inds[np.where(np.isnan(inds))] = g.nc
            inds = inds.astype(int)
        frac = np.stack([iu[a] - ii[a] for a in range(d)], axis=1)
        return inds, frac

    def _monomials(self, frac):
        """Per-axis power vectors [1,u,u^2,u^3] and derivative vectors [0,1,2u,3u^2],
        expanded to (N, 4^d) in coefficient order i + 4j + 16k [+ 64l]
        (A.py:380-382, 440-442, 1097-1115)."""
        N, d = len(frac), self.d
        val, der = [], []
        for a in range(d):
            u = frac[:, a]
            vec = np.transpose(np.array([np.ones(N), u, u ** 2, u ** 3]))
            dvec = np.transpose(np.array([np.zeros(N), np.ones(N), 2 * u, 3 * u ** 2]))
            reps_inner, reps_outer = 4 ** a, 4 ** (d - 1 - a)

            def expand(m):      # np.tile(np.repeat(m, inner, axis=1), outer) without the no-op copies
                if reps_inner > 1:
                    m = np.repeat(m, reps_inner, axis=1)
                return np.tile(m, reps_outer) if reps_outer > 1 else m

            val.append(expand(vec))
            der.append(expand(dvec))
        return val, der

    @staticmethod
    def _prod(arrs):
        out = arrs[0]
        for a in arrs[1:]:
            out = out * a
        return out

    def query(self, query, exact_gemv=False):
        """Range query (A.py:344-521, 1064-1258).  ``query`` (N, >=d) float64 is
        NaN-masked in place like the reference.  Returns per mode:
        'vector' -> comps (N,3); 'norm' -> (norms (N,1), grads (N,d));
        'both' -> (comps, norms, grads).  Sets ``self.query_inds``."""
        N = len(query)
        inds, frac = self.locate(query)
        self.query_inds = inds
        self.calc_coefficients(inds, exact_gemv)
        val, der = self._monomials(frac)
        g = self.geo
        with np.errstate(invalid="ignore"):
            basis = self._prod(val)
            out = []
            if self.mode in ("vector", "both"):
                comps = [np.reshape((self._gather(k, inds) * basis).sum(axis=1), (N, 1)) for k in "xyz"]
                out.append(np.hstack(comps))
            if self.mode in ("norm", "both"):
                tn = self._gather("n", inds)
                norms = np.reshape((tn * basis).sum(axis=1), (N, 1))
                grads = []
                for a in range(self.d):
                    terms = [der[b] if b == a else val[b] for b in range(self.d)]
                    grads.append(((tn * self._prod(terms)) / g.h[a]).sum(axis=1))   # A.py:452, 1187
                out.extend([norms, np.transpose(np.array(grads))])
        return out[0] if len(out) == 1 else tuple(out)
