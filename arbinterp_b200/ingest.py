"""Grid ingest in PyTorch: unordered (x,y,z[,t],values...) rows -> dense value planes + geometry.

Host-side equivalent of the reference's ``getFieldParams`` (A.py:528-568 tricubic,
A.py:1264-1320 quadcubic).  The reference sorts the rows three/four times and scans for
equal coordinates; here each axis is reduced to its sorted distinct values
(``torch.unique``), every row gets its integer grid index and the values are scattered into a
dense ``[C][nt][nz][ny][nx]`` array (x fastest -- the reference's sorted row order).  It runs
on whatever device the tensor lives on (CUDA in production, CPU in the host-logic tests).

Also adds the checks the reference lacks (README "a regular field ... must be supplied"):
the rows must form a full tensor-product grid and each axis must be evenly spaced.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass, field as dc_field
from typing import List, Optional

import torch


@dataclass
class Geometry:
    """What ``getFieldParams`` derives (names follow the reference attributes)."""
    d: int
    npts: List[int]            # grid points per axis (4-D reference: nPosx.. ; A.py:1288-1291)
    ncell: List[int]           # interpolatable cells per axis, n-3 (3-D reference: nPos; A.py:545)
    h: List[float]             # hx.. = |axis[0]-axis[1]| (A.py:547-549)
    int_min: List[float]       # xIntMin.. = 2nd grid coordinate (A.py:551-553)
    int_max: List[float]       # xIntMax.. = 2nd-to-last grid coordinate (A.py:554-556)
    axes: List[torch.Tensor] = dc_field(default_factory=list, repr=False)
    row_index: Optional[torch.Tensor] = dc_field(default=None, repr=False)   # grid index of every input row

    @property
    def nc(self) -> int:       # A.py:568, 1320
        out = 1
        for c in self.ncell:
            out *= c
        return out


class FieldError(ValueError):
    pass


@dataclass
class IngestedField:
    """Result of :func:`ingest_field`, accepted by the ``tricubic`` / ``quadcubic`` constructors in place of
    the raw rows (multi-GPU construction ingests once and broadcasts this, see ``sharding``)."""
    planes: torch.Tensor       # [ncols-d][n_{d-1}]..[n_0] raw value planes
    geo: Geometry

    @property
    def ncols(self) -> int:
        return self.geo.d + int(self.planes.shape[0])


def geometry_from_axes(axes) -> Geometry:
    """Geometry from the sorted distinct coordinates per axis -- the same expressions as in
    :func:`ingest_field` (A.py:541-556), so every rank derives bit-identical numbers from broadcast axes."""
    d = len(axes)
    npts = [int(a.numel()) for a in axes]
    h = [float(torch.abs(a[0] - a[1])) for a in axes]
    return Geometry(d=d, npts=npts, ncell=[n - 3 for n in npts], h=h, int_min=[float(a[1]) for a in axes],
                    int_max=[float(a[-2]) for a in axes], axes=list(axes))


def ingest_field(field, d: int, device=None, spacing_rtol: float = 1e-6):
    """Return ``(planes, geometry)``.

    ``planes``: float64 tensor ``[ncols-d][n_{d-1}]..[n_1][n_0]`` (x fastest) on ``device``.
    ``field``: (N, d+1) or (N, d+3) array-like / tensor, rows in any order (README: "It is not
    necessary to order the coordinates").  The caller's array is not modified (A.py:530 copies).
    """
    t = torch.as_tensor(field)
    if t.dim() != 2 or t.shape[1] <= d:
        raise FieldError(f"field must be 2-D with more than {d} columns, got shape {tuple(t.shape)}")
    t = t.to(device=device, dtype=torch.float64)
    nrow = t.shape[0]
    if torch.isnan(t[:, :d]).any():
        raise FieldError("field coordinates contain NaN")
    axes, index, npts = [], [], []
    for a in range(d):
        ax, inv = torch.unique(t[:, a], sorted=True, return_inverse=True)
        axes.append(ax)
        index.append(inv)
        npts.append(int(ax.numel()))
    total = 1
    for n in npts:
        total *= n
    if total != nrow:
        raise FieldError(f"rows do not form a full grid: {npts} distinct coordinates per axis "
                         f"({total} points) but {nrow} rows")
    if min(npts) < 4:
        raise FieldError(f"need at least 4 grid points per axis, got {npts}")
    lin = index[d - 1]
    for a in range(d - 2, -1, -1):
        lin = lin * npts[a] + index[a]
    hits = torch.bincount(lin, minlength=total)
    if hits.numel() != total or not bool((hits == 1).all()):
        raise FieldError("rows do not form a full grid: some grid points are missing or duplicated")
    ncomp = t.shape[1] - d
    planes = torch.empty((ncomp, total), dtype=torch.float64, device=t.device)
    planes[:, lin] = t[:, d:].T
    planes = planes.reshape([ncomp] + npts[::-1])

    h, lo, hi = [], [], []
    for a in range(d):
        ax = axes[a]
        step = torch.abs(ax[0] - ax[1])                        # A.py:547-549
        diffs = ax[1:] - ax[:-1]
        dev = float(torch.max(torch.abs(diffs - step)) / step)
        if dev > spacing_rtol:
            warnings.warn(f"axis {a} is not evenly spaced (max deviation {dev:.3g} of the first step); "
                          "ARBInterp assumes a regular grid and uses the first step everywhere",
                          RuntimeWarning, stacklevel=3)
        h.append(float(step))
        lo.append(float(ax[1]))
        hi.append(float(ax[-2]))
    geo = Geometry(d=d, npts=npts, ncell=[n - 3 for n in npts], h=h, int_min=lo, int_max=hi, axes=axes,
                   row_index=lin.to(torch.int32) if total < 2 ** 31 else lin)   # kept for update_values(order='rows')
    return planes, geo


def norm_plane(planes: torch.Tensor) -> torch.Tensor:
    """``np.linalg.norm(field[:, d:], axis=1)`` (A.py:58, 74, 677, 693) on the dense planes.

    numpy evaluates sqrt((x*x + y*y) + z*z) with separately rounded products; separate torch
    kernels do the same (no FMA contraction across ops), so the plane is bit-identical."""
    sq = planes * planes
    acc = sq[0]
    for c in range(1, planes.shape[0]):
        acc = acc + sq[c]
    return torch.sqrt(acc)


def sorted_field(planes: torch.Tensor, geo: Geometry) -> torch.Tensor:
    """The reference's ``self.inputfield`` after sorting (rows x fastest), rebuilt on demand."""
    d = geo.d
    grids = torch.meshgrid(*[geo.axes[a] for a in reversed(range(d))], indexing="ij")
    cols = [g.reshape(-1) for g in reversed(grids)]
    vals = planes.reshape(planes.shape[0], -1)
    return torch.stack(cols + [vals[c] for c in range(vals.shape[0])], dim=1)
