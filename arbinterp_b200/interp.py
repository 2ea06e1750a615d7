"""Drop-in ``tricubic`` / ``quadcubic`` classes backed by the sm_100a CUDA path.

Mirrors the public surface of the reference module ``ARBInterp.py`` (A.py): constructors
``tricubic(field, *args, **kwargs)`` (A.py:10) / ``quadcubic(...)`` (A.py:629) with the positional
``'quiet'`` switch and the ``mode=`` keyword, ``Query`` / ``sQuery`` / ``rQuery`` /
``allCoeffs`` / ``calcCoefficients``, the geometry attributes, the return shapes and the NaN
semantics.  All arithmetic happens in the CUDA library (arbinterp_b200/csrc) through the C-ABI
of include/arbinterp_b200.h; there is no CPU fallback.

Differences a user can observe, all documented in DESIGN.md:
  * the full coefficient table is built eagerly on the GPU at construction (the reference
    fills it lazily per queried cell, A.py:376-377) -- ``allCoeffs`` / ``calcCoefficients`` are
    kept as no-ops and ``alphamask`` is all ones;
  * a coordinate exactly on the upper edge whose per-axis index rounds to n-3 returns NaN
    (the reference wraps into a neighbouring cell or raises);
  * ``quiet=True`` is accepted as a keyword as well as the positional string ``'quiet'``;
  * torch CUDA tensors are accepted as queries and then returned as CUDA tensors;
  * ``tricubic(..., table=False)`` / ``quadcubic(..., table=False)`` keep no coefficient table at all and
    evaluate every query from its 4^d grid neighbourhood (for fields that change often; CHANGELOG.md:9 of the
    reference); the 4-D form adds the rank-16 term that reproduces A.py:860 unless ``fixed_d4=True``;
  * ``tricubic(..., table='nodes')`` / ``quadcubic(..., table='nodes')`` keep a node (Hermite) table instead: the
    central-difference values of every grid point, 4x / 16x smaller than the cell table (csrc/arb_nodes.cuh);
  * ``save(path)`` / ``load(path)`` persist the coefficient table;
  * ``tricubic(field, devices=[0, 1, ...])`` keeps one replica of the table per listed GPU in this process and
    fans numpy range queries out over them (sharding.ReplicatedInterp / SlabShardedInterp are the
    one-process-per-GPU forms).
"""
from __future__ import annotations

import ctypes
import os
import sys

import numpy as np
import torch

from . import _lib
from .ingest import IngestedField, ingest_field, norm_plane, sorted_field

__version__ = "1.8"   # API level of the reference this mirrors (A.py:6)

_AXES = "xyzt"


def _cuda_device(device):
    if device is None:
        if not torch.cuda.is_available():
            raise _lib.ArbError("arbinterp_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.ArbError(f"arbinterp_b200 needs a CUDA device, got {device}")
    return device


class _CubicInterpolator:
    """Shared implementation; ``_d`` = 3 (tricubic) or 4 (quadcubic)."""
    __version__ = "1.8"
    _d = 3

    # ------------------------------------------------------------------ construction
    def __init__(self, field, *args, **kwargs):
        d = self._d
        self.eps = 10 * np.finfo(float).eps                      # A.py:11 (never used there either)
        quiet = ("quiet" in args) or bool(kwargs.get("quiet", False))
        pre = field if isinstance(field, IngestedField) else None
        shape = (0, pre.ncols) if pre is not None else getattr(field, "shape", None)
        if shape is None or len(shape) != 2 or shape[1] not in (d + 1, d + 3) or (pre is not None and pre.geo.d != d):
            sys.exit(f"--- Input not shaped as expected - should be N x {d + 1} or N x {d + 3} ---")  # A.py:104, 723
        devices = kwargs.get("devices")                          # single-process multi-GPU: one replica per device
        if devices is not None:
            devices = [_cuda_device(x) for x in devices]
            if not devices or len(set(devices)) != len(devices):
                raise ValueError("devices must be a non-empty list of distinct CUDA devices")
            if kwargs.get("device") is not None and _cuda_device(kwargs["device"]) != devices[0]:
                raise ValueError("device= and devices[0] disagree")
            if kwargs.get("slab") is not None:
                raise ValueError("devices=[...] replicates the table; slab-sharded tables use sharding.SlabShardedInterp")
        self._device = devices[0] if devices else _cuda_device(kwargs.get("device"))
        self._lib = _lib.load()
        slab = kwargs.get("slab")                                # (lo, hi) cell layers of the slowest axis
        self._reference_quirk = not bool(kwargs.get("fixed_d4", False))

        if pre is not None:
            planes, geo = pre.planes.to(self._device), pre.geo
        else:
            planes, geo = ingest_field(field, d, device=self._device)
        self._geo = geo
        scalar = shape[1] == d + 1
        banner = None
        if scalar:                                               # A.py:24-32, 643-651
            banner = "--- Scalar field, ignoring switches, interpolating for magnitude and gradient --- \n"
            mode = "norm"
        else:                                                    # A.py:33-102, 652-721
            mode = kwargs.get("mode", None)
            if mode == "vector":
                banner = "--- Vector field, interpolating for vector components --- \n"
            elif mode == "norm":
                banner = "--- Vector field, interpolating for magnitude and gradient --- \n"
            elif mode == "both":
                banner = "--- Vector field, interpolating vector components plus magnitude and gradient --- \n"
            elif "mode" in kwargs:
                banner = "--- Vector field, invalid option, defaulting to interpolating for vector components --- \n"
                mode = "vector"
            else:
                banner = "--- Vector field, no option selected, defaulting to interpolating for vector components --- \n"
                mode = "vector"
        self._explicit_vector = (not scalar) and kwargs.get("mode", None) == "vector"
        if not quiet:
            print(banner)
        self._scalar_input = scalar
        self._mode = mode
        self._mode_code = {"vector": _lib.MODE_VECTOR, "norm": _lib.MODE_NORM, "both": _lib.MODE_BOTH}[mode]

        self._planes = self._component_planes(planes)
        # the reference keeps the sorted field in every mode (A.py:14, 530-532); in 'norm' mode the table is built
        # from |B| only, so the three raw planes are kept beside it (not for one slab of a sharded table)
        self._raw = planes[0:3] if (mode == "norm" and not scalar and slab is None) else None

        self._set_geometry_attributes()

        nslow = geo.ncell[d - 1]
        lo, hi = (0, nslow) if slab is None else (int(slab[0]), int(slab[1]))
        if not (0 <= lo < hi <= nslow):
            raise ValueError(f"slab {slab} outside the {nslow} cell layers of the slowest axis")
        self._slab = (lo, hi)
        self._table_free = kwargs.get("table", True) is False
        self._nodes = None
        self._packed = None
        if isinstance(kwargs.get("table"), str):
            if kwargs["table"] != "nodes":
                raise ValueError("table must be True (cell coefficients), False (no table) or 'nodes' (Hermite node table)")
            if (lo, hi) != (0, nslow):
                raise ValueError("table='nodes' is available for unsharded interpolators only (it is small enough to replicate)")
            self._table = None
            self._build_nodes()
        elif self._table_free:
            if (lo, hi) != (0, nslow):
                raise ValueError("table=False is available for unsharded interpolators only")
            self._table = None
            nx = geo.npts[0]
            self._pitch = nx + (nx & 1)                           # TMA needs 16-byte row strides
            if self._pitch != nx:
                self._planes = torch.nn.functional.pad(self._planes, (0, 1)).contiguous()
            # component-interleaved grid (one 128-byte piece per neighbourhood row serves every component): the default
            # for 'vector' / 'both' in 3-D and 4-D (4-D 'vector' gathers the unused |B| slot too and is still faster
            # than the per-component TMA boxes: 0.71 against 0.63 of the cell table on 48^3x32, 0.54 against 0.29 on
            # 64^3x48, profiles/r02_tablefree_bench_4d.log); interleave=False keeps the per-component planes
            self._use_packed = (not scalar) and mode in ("vector", "both") and bool(kwargs.get("interleave", True))
            self._pack_planes()
            self._make_cgeom()
            # the planes were written by torch ops on the current stream; numpy queries run on the library's own
            # non-blocking streams, which do not order against it -- finish the writes here (as _build_table does)
            torch.cuda.current_stream(self._device).synchronize()
        else:
            self._build_table()
        self._last_cells = None
        self.queryInd = None
        self._replicas = [self]
        if devices and len(devices) > 1:
            sub = {k: v for k, v in kwargs.items() if k not in ("devices", "device", "quiet")}
            for dev in devices[1:]:
                self._replicas.append(type(self)(IngestedField(planes=planes.to(dev), geo=geo), "quiet", device=dev, **sub))
        del planes

        self._bind_mode()

    def _component_planes(self, planes):
        """Value planes the table is built from, in table component order (A.py:32, 58, 74 and the 4-D twins)."""
        if self._scalar_input:
            comp = planes[0:1]
        elif self._mode == "vector":
            comp = planes[0:3]
        elif self._mode == "norm":
            comp = norm_plane(planes[0:3]).unsqueeze(0)
        else:
            comp = torch.cat([planes[0:3], norm_plane(planes[0:3]).unsqueeze(0)], dim=0)
        return comp.contiguous()

    def update_values(self, values, order="rows"):
        """Replace the field values on the same grid and rebuild in place -- no re-ingest, no reallocation.

        The reference has no such call: a changed field means a new object and, with it, a new lazy coefficient
        fill ("not a good idea ... where you are frequently updating the field", CHANGELOG.md:9).  Here the
        rebuild of a 256^3 table is one 1.2 ms kernel, so time-varying fields can keep their table.
        ``values``: (N, 1) / (N,) for scalar input or (N, 3) for vector input, array-like or tensor (CUDA tensors
        are not copied through the host).  ``order='rows'``: in the row order of the ``field`` array the object was
        constructed from; ``order='grid'``: in sorted grid order (x fastest -- the row order of ``inputfield``).
        Results afterwards are bit-identical to a freshly constructed interpolator of the updated field."""
        if self._planes is None:
            raise ValueError("update_values() needs the field planes; an interpolator restored with load() has none")
        geo, d = self._geo, self._d
        ncol = 1 if self._scalar_input else 3
        v = torch.as_tensor(values)
        if v.dim() == 1:
            v = v.unsqueeze(1)
        total = 1
        for n in geo.npts:
            total *= n
        if v.dim() != 2 or tuple(v.shape) != (total, ncol):
            raise ValueError(f"values must have shape ({total}, {ncol}), got {tuple(v.shape)}")
        v = v.to(device=self._device, dtype=torch.float64)
        if order == "rows":
            if geo.row_index is None:
                raise ValueError("order='rows' needs the row map of the constructor's field; this interpolator was "
                                 "built from pre-ingested planes -- pass order='grid'")
            planes = torch.empty((ncol, total), dtype=torch.float64, device=self._device)
            planes[:, geo.row_index.to(self._device).long()] = v.T
        elif order == "grid":
            planes = v.T.contiguous()
        else:
            raise ValueError("order must be 'rows' or 'grid'")
        dense = planes.reshape([ncol] + list(geo.npts[::-1]))
        self._planes = self._component_planes(dense)
        if getattr(self, "_raw", None) is not None:
            self._raw = dense[0:3]
        for rep in getattr(self, "_replicas", [self])[1:]:
            rep.update_values(planes.T.to(rep._device), order="grid")
        del planes, dense
        if self._nodes is not None:
            self._build_nodes()
        elif self._table is None:
            if self._pitch != geo.npts[0]:
                self._planes = torch.nn.functional.pad(self._planes, (0, 1)).contiguous()
            self._pack_planes()
            torch.cuda.current_stream(self._device).synchronize()   # see __init__: lib streams do not order against torch's
        else:
            self._build_table()
        self._last_cells = None
        self.queryInd = None

    def release(self):
        """Drop the device tensors now.  The mode-specific entry points bound in the constructor (``self.Query = ...``,
        as the reference binds them, A.py:27-30) make every interpolator part of a reference cycle, so ``del obj`` frees
        its 8-50 GB only at the next garbage collection; call this to give the memory back at once."""
        for rep in getattr(self, "_replicas", [self])[1:]:
            rep.release()
        self._table = self._nodes = self._packed = self._planes = self._raw = self._last_cells = None

    def _bind_mode(self):
        # bind the mode-specific entry points like the reference does (A.py:27-30, 38-41, ...)
        mode = self._mode
        self.Query = {"vector": self.Query1, "norm": self.Query2, "both": self.Query3}[mode]
        self.sQuery = {"vector": self.sQuery1, "norm": self.sQuery2, "both": self.sQuery3}[mode]
        self.rQuery = {"vector": self.rQuery1, "norm": self.rQuery2, "both": self.rQuery3}[mode]
        self.calcCoefficients = self._calc_coefficients_noop

    def _set_geometry_attributes(self):
        # public geometry attributes (A.py:545-568 / 1288-1320)
        d, geo = self._d, self._geo
        if d == 3:
            self.nPos = np.array(geo.ncell)
        else:
            self.nPosx, self.nPosy, self.nPosz, self.nPost = geo.npts
        for a in range(d):
            setattr(self, "h" + _AXES[a], np.float64(geo.h[a]))
            setattr(self, _AXES[a] + "IntMin", np.float64(geo.int_min[a]))
            setattr(self, _AXES[a] + "IntMax", np.float64(geo.int_max[a]))
        self.nc = geo.nc
        self.alphamask = np.ones((self.nc + 1, 1))              # every cell is built (A.py:21-22)

    def _make_cgeom(self):
        d, geo = self._d, self._geo
        g = _lib.ArbGeom()
        g.d, g.ncomp = d, {"vector": 3, "norm": 1, "both": 4}[self._mode]
        for a in range(4):
            g.ncell[a] = geo.ncell[a] if a < d else 1
            g.int_min[a] = geo.int_min[a] if a < d else 0.0
            g.int_max[a] = geo.int_max[a] if a < d else 0.0
            g.h[a] = geo.h[a] if a < d else 1.0
        g.slab_lo, g.slab_hi = self._slab
        g.flags = 0 if self._reference_quirk else _lib.GEOM_FIXED_D4
        self._cgeom = g

    # ------------------------------------------------------------------ persistence (CHANGELOG.md:9 "save these coefficients to a file")
    _MAGIC = b"ARBTAB01"

    def save(self, path, chunk_bytes=256 << 20):
        """Write the coefficient table (or, for ``table='nodes'``, the node table) and everything needed to query it (geometry, mode, slab) to
        ``path``: 8-byte magic, uint64 header length, JSON header, zero padding to 4096, then the
        raw little-endian float64 table ``[ncell_local+1][C][4^d]``.  The field itself is not stored."""
        import json
        store = self._nodes if self._nodes is not None else self._table
        if store is None:
            raise ValueError("a table=False interpolator has no coefficient table to save")
        geo = self._geo
        header = {"format": 1, "kind": "nodes" if self._nodes is not None else "cells",
                  "d": self._d, "mode": self._mode, "scalar_input": self._scalar_input,
                  "explicit_vector": self._explicit_vector, "reference_quirk": self._reference_quirk,
                  "npts": list(geo.npts), "h": [float(v).hex() for v in geo.h],
                  "int_min": [float(v).hex() for v in geo.int_min], "int_max": [float(v).hex() for v in geo.int_max],
                  "slab": list(self._slab), "table_shape": list(store.shape), "dtype": "<f8"}
        blob = json.dumps(header).encode("utf-8")
        flat = store.reshape(-1)
        step = max(1, chunk_bytes // 8)
        with open(path, "wb") as f:
            f.write(self._MAGIC)
            f.write(np.uint64(len(blob)).tobytes())
            f.write(blob)
            f.write(b"\0" * (-(16 + len(blob)) % 4096))
            for lo in range(0, flat.numel(), step):
                flat[lo:lo + step].cpu().numpy().tofile(f)

    @classmethod
    def load(cls, path, device=None, chunk_bytes=256 << 20):
        """Rebuild an interpolator from a file written by :meth:`save` -- no ingest, no coefficient build."""
        import json
        from .ingest import Geometry
        with open(path, "rb") as f:
            if f.read(8) != cls._MAGIC:
                raise ValueError(f"{path}: not an arbinterp_b200 coefficient file")
            hlen = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            if hlen > (1 << 20):
                raise ValueError(f"{path}: implausible header length {hlen}")
            header = json.loads(f.read(hlen).decode("utf-8"))
            if header["d"] != cls._d:
                raise ValueError(f"{path} holds a {header['d']}-D table, {cls.__name__} is {cls._d}-D")
            f.seek((16 + hlen + 4095) // 4096 * 4096)
            self = cls.__new__(cls)
            self.eps = 10 * np.finfo(float).eps
            self._device = _cuda_device(device)
            self._lib = _lib.load()
            self._mode = header["mode"]
            self._mode_code = {"vector": _lib.MODE_VECTOR, "norm": _lib.MODE_NORM, "both": _lib.MODE_BOTH}[self._mode]
            self._scalar_input = header["scalar_input"]
            self._explicit_vector = header["explicit_vector"]
            self._reference_quirk = header["reference_quirk"]
            npts = header["npts"]
            unhex = lambda xs: [float.fromhex(v) for v in xs]
            self._geo = Geometry(d=cls._d, npts=npts, ncell=[n - 3 for n in npts], h=unhex(header["h"]),
                                 int_min=unhex(header["int_min"]), int_max=unhex(header["int_max"]))
            self._slab = tuple(int(v) for v in header["slab"])
            self._planes = None
            self._raw = None
            self._replicas = [self]
            self._table_free = False
            self._packed = None
            shape = tuple(int(v) for v in header["table_shape"])
            # the header is untrusted input: the table shape must be the one this geometry, slab and mode imply,
            # and the file must actually hold that many bytes, before anything is allocated
            d = cls._d
            lo, hi = self._slab if len(self._slab) == 2 else (-1, -1)
            if len(npts) != d or min(npts) < 4 or not (0 <= lo < hi <= npts[d - 1] - 3):
                raise ValueError(f"{path}: inconsistent geometry in header (npts {npts}, slab {self._slab})")
            layer = 1
            for a in range(d - 1):
                layer *= npts[a] - 3
            ncomp = {"vector": 3, "norm": 1, "both": 4}[self._mode]
            kind = header.get("kind", "cells")
            if kind == "nodes":                                  # layouts of _build_nodes
                if (lo, hi) != (0, npts[d - 1] - 3):
                    raise ValueError(f"{path}: node tables are not slab-sharded")
                if d == 3 and ncomp >= 3:
                    want = (npts[2] - 2, npts[1] - 2, npts[0] - 2, 4, 8)
                elif d == 3:
                    want = (ncomp, npts[2] - 2, npts[1] - 2, npts[0] - 3, 2, 8)
                else:
                    want = (ncomp,) + tuple(npts[a] - 2 for a in reversed(range(d))) + (16,)
            elif kind == "cells":
                want = (layer * (hi - lo) + 1, ncomp, 4 ** d)
            else:
                raise ValueError(f"{path}: unknown table kind {kind!r}")
            if shape != want:
                raise ValueError(f"{path}: table shape {shape} does not match geometry/mode (expected {want})")
            left = os.fstat(f.fileno()).st_size - f.tell()
            if left < 8 * int(np.prod(want)):
                raise ValueError(f"{path}: truncated table")
            store = torch.empty(shape, dtype=torch.float64, device=self._device)
            self._table, self._nodes = (None, store) if kind == "nodes" else (store, None)
            flat = store.reshape(-1)
            step = max(1, chunk_bytes // 8)
            for lo in range(0, flat.numel(), step):
                n = min(step, flat.numel() - lo)
                buf = np.fromfile(f, dtype="<f8", count=n)
                if buf.size != n:
                    raise ValueError(f"{path}: truncated table")
                flat[lo:lo + n].copy_(torch.from_numpy(buf))
        self._set_geometry_attributes()
        self._make_cgeom()
        self._last_cells = None
        self.queryInd = None
        self._bind_mode()
        return self

    def _pack_planes(self):
        """Table-free 'vector' / 'both': the component-interleaved grid ``[nt][nz][ny][nx][4]`` (Bx, By, Bz, |B| or 0)
        the query kernel gathers from -- one 128-byte piece per neighbourhood row serves every component."""
        if not getattr(self, "_use_packed", False):
            self._packed = None
            return
        nx = self._geo.npts[0]
        comps = [self._planes[c, ..., :nx] for c in range(self._planes.shape[0])]
        if len(comps) == 3:
            comps.append(torch.zeros_like(comps[0]))
        self._packed = torch.stack(comps, dim=-1).contiguous()

    def _build_nodes(self):
        """Node (Hermite) table (csrc/arb_nodes.cuh): the central-difference values f, fx, fy, fxy, ... of every
        interior grid point -- the rows of the reference's D matrix (A.py:129-173 / 762-876).  4-D:
        ``[C][nt-2][nz-2][ny-2][nx-2][16]``, 16x smaller than the cell table; 3-D one component:
        ``[1][nz-2][ny-2][nx-3][2][8]`` (x-adjacent nodes stored as 128-byte-aligned pairs), 4x smaller; 3-D 'vector' /
        'both': ``[nz-2][ny-2][nx-2][4][8]`` (the components of a node together), 6x / 8x smaller.  Same answers to
        round-off."""
        d, geo = self._d, self._geo
        ncomp = self._planes.shape[0]
        if d == 3 and ncomp >= 3:       # components of a node together: [nz-2][ny-2][nx-2][4][8]
            shape = [geo.npts[2] - 2, geo.npts[1] - 2, geo.npts[0] - 2, 4, 8]
        elif d == 3:                    # aligned x-pairs: [1][nz-2][ny-2][nx-3][2][8]
            shape = [ncomp, geo.npts[2] - 2, geo.npts[1] - 2, geo.npts[0] - 3, 2, 8]
        else:
            shape = [ncomp] + [geo.npts[a] - 2 for a in reversed(range(d))] + [16]
        if self._nodes is None or list(self._nodes.shape) != shape:
            self._nodes = torch.empty(shape, dtype=torch.float64, device=self._device)
        n = (ctypes.c_int64 * 4)(*([geo.npts[a] for a in range(d)] + [1] * (4 - d)))
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(self._lib.arb_build_nodes(d, self._planes.data_ptr(), ncomp, ctypes.byref(n), geo.npts[0],
                                                  self._nodes.data_ptr(), stream), "arb_build_nodes")
            torch.cuda.current_stream(self._device).synchronize()      # see _build_table
        self._make_cgeom()

    @property
    def nodes(self) -> torch.Tensor:
        """Device node table of a ``table='nodes'`` interpolator."""
        if self._nodes is None:
            raise AttributeError("this interpolator was not built with table='nodes'")
        return self._nodes

    def _build_table(self):
        d, geo = self._d, self._geo
        lo, hi = self._slab
        ncomp = self._planes.shape[0]
        nm = 4 ** d
        layer = 1
        for a in range(d - 1):
            layer *= geo.ncell[a]
        ncell_local = layer * (hi - lo)
        sub = self._planes[:, lo:hi + 3].contiguous()          # cell layer k needs grid planes k..k+3
        if getattr(self, "_table", None) is None or tuple(self._table.shape) != (ncell_local + 1, ncomp, nm):
            self._table = torch.empty((ncell_local + 1, ncomp, nm), dtype=torch.float64, device=self._device)
        n = (ctypes.c_int64 * 4)(*([geo.npts[a] for a in range(d - 1)] + [hi - lo + 3] + [1] * (4 - d)))
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            _lib.check(self._lib.arb_build_coeffs(d, sub.data_ptr(), ncomp, ctypes.byref(n), self._table.data_ptr(),
                                                   int(self._reference_quirk), stream), "arb_build_coeffs")
            # construction is one-off: finish it here so that queries issued from any stream
            # (arb_query_host uses its own copy/compute streams) see a complete table
            torch.cuda.current_stream(self._device).synchronize()
        self._make_cgeom()

    # ------------------------------------------------------------------ lazily materialised reference attributes
    @property
    def table(self) -> torch.Tensor:
        """Device coefficient table ``[ncell_local+1][C][4^d]`` (cell-major; last row NaN)."""
        if self._table is None:
            raise AttributeError("this interpolator was built with table=False and keeps no coefficient table")
        return self._table

    @property
    def A(self):
        """The reference's combined matrix ``inv(B) @ D`` (A.py:175, 878), generated exactly."""
        return _lib.get_matrix(self._d, "A", self._reference_quirk)

    @property
    def inputfield(self):
        """The sorted field (A.py:530-532 / 1266-1269): rows x fastest, columns as in the constructor's array."""
        planes = self._planes
        if planes is None:
            raise AttributeError("inputfield is unavailable after load(): a coefficient file does not store the field")
        if self._scalar_input:
            planes = planes[0:1]
        elif self._mode == "norm":
            planes = getattr(self, "_raw", None)
            if planes is None:
                raise AttributeError("inputfield is not kept for one slab of a sharded 'norm' table")
        else:
            planes = planes[0:3]                                  # 'vector' / 'both': Bx, By, Bz (+ |B| behind them)
        planes = planes[..., :self._geo.npts[0]]                  # drop the table-free TMA pitch padding, if any
        return sorted_field(planes, self._geo).cpu().numpy()

    @property
    def basePointInds(self):
        """Flat sorted-row index of every cell's lower corner (A.py:559-565, 1308-1317)."""
        geo, d = self._geo, self._d
        stride = [int(np.prod(geo.npts[:a])) for a in range(d)]
        grids = np.meshgrid(*[np.arange(geo.ncell[a]) * stride[a] for a in reversed(range(d))], indexing="ij")
        return (sum(stride) + sum(grids)).ravel()

    def _component_names(self):
        return {"vector": "xyz", "norm": "n", "both": "xyzn"}[self._mode]

    def _alpha(self, name):
        names = self._component_names()
        if name not in names:
            raise AttributeError(f"alpha{name} does not exist in mode '{self._mode}'")
        if self._table is None:
            raise AttributeError("alpha* does not exist for a table=False interpolator")
        if self._slab != (0, self._geo.ncell[self._d - 1]):
            raise AttributeError("alpha* views are only available for an unsharded table")
        col = self._table[:, names.index(name), :].T.contiguous().cpu().numpy()   # (4^d, nc+1), like A.py:31
        if self._scalar_input:
            col[:, -1] = 0.0                                   # scalar mode never NaNs the sentinel (A.py:31)
        if self._explicit_vector and self._d in (3, 4):        # stray extra column (A.py:42-45, 661-664)
            col = np.concatenate([col, np.full((col.shape[0], 1), np.nan)], axis=1)
            col[:, -2] = 0.0
        return col

    alphax = property(lambda self: self._alpha("x"))
    alphay = property(lambda self: self._alpha("y"))
    alphaz = property(lambda self: self._alpha("z"))
    alphan = property(lambda self: self._alpha("n"))

    def _plane(self, name):
        if self._planes is None:
            raise AttributeError("the field planes are not stored in a coefficient file; B* is unavailable after load()")
        names = self._component_names()
        if name not in names:
            raise AttributeError(f"B{name} does not exist in mode '{self._mode}'")
        plane = self._planes[names.index(name)]
        if getattr(self, "_table_free", False):
            plane = plane[..., :self._geo.npts[0]]
        return plane.reshape(-1).cpu().numpy()

    Bx = property(lambda self: self._plane("x"))
    By = property(lambda self: self._plane("y"))
    Bz = property(lambda self: self._plane("z"))
    Bn = property(lambda self: self._plane("n"))

    @property
    def queryInds(self):
        """Cell index of every row of the last range query; ``nc`` for NaN rows (A.py:368-370)."""
        c = self._last_cells
        if c is None:
            raise AttributeError("queryInds is set by the first range query")
        if isinstance(c, list):                                  # one share per device replica, in row order
            return np.concatenate([t.cpu().numpy() for t in c])
        return c.cpu().numpy() if isinstance(c, torch.Tensor) else c

    def allCoeffs(self):
        """A.py:523-525 / 1260-1262 -- the table is always complete here."""
        return None

    def _calc_coefficients_noop(self, alphaindex):
        return None

    # ------------------------------------------------------------------ range queries
    def _outputs(self, n, pinned):
        d, mode = self._d, self._mode
        kw = dict(dtype=torch.float64)
        mk = (lambda *s: torch.empty(*s, pin_memory=True, **kw)) if pinned else \
             (lambda *s: torch.empty(*s, device=self._device, **kw))
        comps = mk(n, 3) if mode in ("vector", "both") else None
        norm = mk(n, 1) if mode in ("norm", "both") else None
        grad = mk(n, d) if mode in ("norm", "both") else None
        return comps, norm, grad

    @staticmethod
    def _ptr(t):
        return None if t is None else t.data_ptr()

    def _range_device(self, q: torch.Tensor):
        """Queries already in HBM (torch CUDA tensor): kernel only, outputs stay on the device."""
        d = self._d
        if q.dim() != 2 or q.shape[1] < d:
            raise IndexError(f"query must be (N, >={d})")
        if q.dtype != torch.float64 or not q.is_contiguous() or q.device != self._device:
            work = q.to(device=self._device, dtype=torch.float64).contiguous()
        else:
            work = q
        n = work.shape[0]
        comps, norm, grad = self._outputs(n, pinned=False)
        cells = torch.empty(n, dtype=torch.int64, device=self._device)
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            if self._nodes is not None:
                _lib.check(self._lib.arb_query_nodes(ctypes.byref(self._cgeom), self._nodes.data_ptr(), self._mode_code,
                                                     work.data_ptr(), n, work.shape[1], self._ptr(comps), self._ptr(norm),
                                                     self._ptr(grad), cells.data_ptr(), None, None, stream), "arb_query_nodes")
            elif self._packed is not None:
                _lib.check(self._lib.arb_query_gridil(ctypes.byref(self._cgeom), self._packed.data_ptr(), self._mode_code,
                                                      work.data_ptr(), n, work.shape[1], self._ptr(comps), self._ptr(norm),
                                                      self._ptr(grad), cells.data_ptr(), None, None, stream), "arb_query_gridil")
            elif self._table is None:
                _lib.check(self._lib.arb_query_grid(ctypes.byref(self._cgeom), self._planes.data_ptr(), self._pitch,
                                                    self._mode_code, work.data_ptr(), n, work.shape[1],
                                                    self._ptr(comps), self._ptr(norm), self._ptr(grad),
                                                    cells.data_ptr(), None, None, stream), "arb_query_grid")
            else:
                _lib.check(self._lib.arb_query(ctypes.byref(self._cgeom), self._table.data_ptr(), self._mode_code,
                                               work.data_ptr(), n, work.shape[1], self._ptr(comps), self._ptr(norm),
                                               self._ptr(grad), cells.data_ptr(), None, None, stream), "arb_query")
        if work is not q:                                        # mirror the in-place NaN rows (A.py:350-355)
            bad = torch.isnan(work[:, :d]).any(dim=1) & ~torch.isnan(q[:, :d].to(self._device)).any(dim=1)
            if bool(bad.any()):
                q[bad.to(q.device)] = float("nan")
        self._last_cells = cells
        return comps, norm, grad

    def _host_outputs(self, n):
        """Host result arrays of a numpy range query, as CPU torch tensors.  Page-locked ones let the D2H copies land
        directly; beyond a few GB pinning itself becomes the cost (and can fail), so very large results are ordinary
        arrays staged through the library's pinned ring."""
        d, mode = self._d, self._mode
        if n * 8 * (3 + 1 + d) <= self._PIN_LIMIT_BYTES:
            return self._outputs(n, pinned=True)
        comps = torch.from_numpy(np.empty((n, 3))) if mode in ("vector", "both") else None
        norm = torch.from_numpy(np.empty((n, 1))) if mode in ("norm", "both") else None
        grad = torch.from_numpy(np.empty((n, d))) if mode in ("norm", "both") else None
        return comps, norm, grad

    def _host_rows(self, work, lo, hi, comps, norm, grad):
        """Rows [lo, hi) of the host batch ``work`` through this object's device: pipelined H2D / kernel / D2H inside
        the library (arb_query_host).  Returns the device tensor of cell indices (read back lazily)."""
        n = hi - lo
        dev = self._device
        row = lambda t: None if t is None else t[lo:hi].data_ptr()
        with torch.cuda.device(dev):
            # the library works on its own non-blocking streams: anything torch still has queued for the blocks
            # involved (a freed block re-used for `cells`, planes being rewritten) must be finished first
            torch.cuda.current_stream(dev).synchronize()
            cells = torch.empty(n, dtype=torch.int64, device=dev)        # stays in HBM; read back lazily
            chunk = int(os.environ.get("ARB_HOST_CHUNK_ROWS", "0"))
            qptr = work.ctypes.data + lo * work.strides[0]
            if self._nodes is not None:
                _lib.check(self._lib.arb_query_nodes_host(ctypes.byref(self._cgeom), self._nodes.data_ptr(), self._mode_code,
                                                          qptr, n, work.shape[1], row(comps), row(norm), row(grad),
                                                          cells.data_ptr(), chunk), "arb_query_nodes_host")
            elif self._packed is not None:
                _lib.check(self._lib.arb_query_gridil_host(ctypes.byref(self._cgeom), self._packed.data_ptr(), self._mode_code,
                                                           qptr, n, work.shape[1], row(comps), row(norm), row(grad),
                                                           cells.data_ptr(), chunk), "arb_query_gridil_host")
            elif self._table is None:
                _lib.check(self._lib.arb_query_grid_host(ctypes.byref(self._cgeom), self._planes.data_ptr(), self._pitch,
                                                         self._mode_code, qptr, n, work.shape[1],
                                                         row(comps), row(norm), row(grad),
                                                         cells.data_ptr(), chunk), "arb_query_grid_host")
            else:
                _lib.check(self._lib.arb_query_host(ctypes.byref(self._cgeom), self._table.data_ptr(), self._mode_code,
                                                    qptr, n, work.shape[1], row(comps), row(norm), row(grad),
                                                    cells.data_ptr(), chunk), "arb_query_host")
        return cells

    def _range_host(self, query: np.ndarray):
        """numpy queries (the reference's call, A.py:177-211).  One device: arb_query_host.  ``devices=[...]``: the
        batch is cut into one contiguous share per replica and the shares run concurrently, one host thread per
        device (the library call releases the GIL); results land in one set of output arrays."""
        d = self._d
        if query.ndim != 2 or query.shape[1] < d:
            raise IndexError(f"query must be (N, >={d})")
        direct = query.dtype == np.float64 and query.flags.c_contiguous and query.flags.writeable
        work = query if direct else np.ascontiguousarray(query, dtype=np.float64).copy()
        n = work.shape[0]
        if n <= self._SMALL_ROWS:
            return self._range_host_small(query, work, direct)
        comps, norm, grad = self._host_outputs(n)
        reps = getattr(self, "_replicas", None) or [self]
        nrep = min(len(reps), max(1, n // self._MULTI_MIN_ROWS))
        if nrep == 1:
            cells = self._host_rows(work, 0, n, comps, norm, grad)
        else:
            bounds = [n * i // nrep for i in range(nrep + 1)]
            futs = [self._pool().submit(reps[i]._host_rows, work, bounds[i], bounds[i + 1], comps, norm, grad)
                    for i in range(nrep)]
            cells = [f.result() for f in futs]
        if not direct:
            bad = np.isnan(work[:, :d]).any(axis=1) & ~np.isnan(np.asarray(query[:, :d], dtype=np.float64)).any(axis=1)
            if bad.any():
                query[np.where(bad)[0]] = np.nan                 # A.py:350-355 (raises for int arrays, as there)
        self._last_cells = cells
        return tuple(None if t is None else t.numpy() for t in (comps, norm, grad))

    _MULTI_MIN_ROWS = 1 << 17   # rows per device below which fanning a batch out costs more than it saves

    def _pool(self):
        pool = getattr(self, "_thread_pool", None)
        if pool is None:
            from concurrent.futures import ThreadPoolExecutor
            pool = self._thread_pool = ThreadPoolExecutor(max_workers=len(self._replicas),
                                                          thread_name_prefix="arb-replica")
        return pool

    _SMALL_ROWS = 8192        # same threshold as SMALL_ROWS in csrc/arb_host.cu
    _PIN_LIMIT_BYTES = 4 << 30

    def _range_host_small(self, query, work, direct):
        """Latency path for short batches and single points: plain numpy outputs, one H2D, one kernel, one
        D2H inside the library -- no pinned allocations, no stream ring."""
        d, mode, n = self._d, self._mode, work.shape[0]
        comps = np.empty((n, 3)) if mode != "norm" else None
        norm = np.empty((n, 1)) if mode != "vector" else None
        grad = np.empty((n, d)) if mode != "vector" else None
        cells = np.empty(n, dtype=np.int64)
        ptr = lambda a: None if a is None else a.ctypes.data
        prev = torch.cuda.current_device()
        if prev != self._device.index:                           # latency path: no context manager unless needed
            torch.cuda.set_device(self._device)
        try:
            args = (ctypes.byref(self._cgeom), self._nodes.data_ptr() if self._nodes is not None else
                    self._packed.data_ptr() if self._packed is not None else
                    self._planes.data_ptr() if self._table is None else self._table.data_ptr())
            if self._nodes is not None:
                rc = self._lib.arb_query_nodes_host(*args, self._mode_code, work.ctypes.data, n, work.shape[1],
                                                    ptr(comps), ptr(norm), ptr(grad), cells.ctypes.data, 0)
            elif self._packed is not None:
                rc = self._lib.arb_query_gridil_host(*args, self._mode_code, work.ctypes.data, n, work.shape[1],
                                                     ptr(comps), ptr(norm), ptr(grad), cells.ctypes.data, 0)
            elif self._table is None:
                rc = self._lib.arb_query_grid_host(*args, self._pitch, self._mode_code, work.ctypes.data, n, work.shape[1],
                                                   ptr(comps), ptr(norm), ptr(grad), cells.ctypes.data, 0)
            else:
                rc = self._lib.arb_query_host(*args, self._mode_code, work.ctypes.data, n, work.shape[1],
                                              ptr(comps), ptr(norm), ptr(grad), cells.ctypes.data, 0)
        finally:
            if prev != self._device.index:
                torch.cuda.set_device(prev)                      # the caller's current device is not ours to change
        _lib.check(rc, "arb_query_host")
        if not direct:
            bad = np.isnan(work[:, :d]).any(axis=1) & ~np.isnan(np.asarray(query[:, :d], dtype=np.float64)).any(axis=1)
            if bad.any():
                query[np.where(bad)[0]] = np.nan
        self._last_cells = cells
        return comps, norm, grad

    def _range(self, query):
        if isinstance(query, torch.Tensor):
            if query.is_cuda:
                return self._range_device(query)
            res = self._range_host(query.numpy())
            return tuple(None if r is None else torch.from_numpy(r) for r in res)
        return self._range_host(query)

    def rQuery1(self, query):
        """Vector components (A.py:344-397 / 1064-1127): returns (N,3)."""
        comps, _, _ = self._range(query)
        return comps

    def rQuery2(self, query):
        """Magnitude + gradient (A.py:399-454 / 1129-1188): returns ((N,1), (N,d))."""
        _, norm, grad = self._range(query)
        return norm, grad

    def rQuery3(self, query):
        """Components, magnitude, gradient (A.py:457-521 / 1190-1258)."""
        return self._range(query)

    # ------------------------------------------------------------------ fused query + push
    def push(self, pos, vel, dt, nsteps, kappa, gravity=None):
        """Advance particles by ``nsteps`` velocity-Verlet steps of ``dv/dt = kappa * grad(value)(x) + gravity``
        inside one kernel, where ``grad(value)`` is the gradient ``Query`` returns in 'norm'/'both'/scalar mode
        (for a magnetic trap: value = |B|, kappa = -mu/m).  ``pos``: (N,d), ``vel``: (N,3) float64 torch CUDA
        tensors, updated in place (numpy arrays are copied to the GPU and back); for a quadcubic
        (time-dependent) field the 4th column of ``pos`` is each particle's own time and advances by ``dt``
        per step.  Particles that leave the interpolation volume get NaN position and velocity (in 4-D the time
        column keeps the time at which they left).  Returns the number of particles lost."""
        if self._slab != (0, self._geo.ncell[self._d - 1]):
            raise ValueError("this interpolator holds one slab of a sharded table: use SlabShardedInterp.push(), which "
                             "moves particles between the slab owners")
        host = not isinstance(pos, torch.Tensor)
        p = torch.as_tensor(pos, dtype=torch.float64).to(self._device).contiguous() if host else pos
        v = torch.as_tensor(vel, dtype=torch.float64).to(self._device).contiguous() if host else vel
        nlost = self._push_local(p, v, None, dt, nsteps, kappa, gravity)
        if host:
            np.copyto(pos, p.cpu().numpy())
            np.copyto(vel, v.cpu().numpy())
        return nlost

    def _push_local(self, p, v, steps, dt, nsteps, kappa, gravity=None):
        """One launch of the push kernel on this table.  ``steps``: optional int64 CUDA tensor (N,) with every
        particle's next step index (arb_push_steps; a particle inside the volume but outside this table's slab
        is parked unchanged); None = plain arb_push.  Returns the number of particles lost in this launch."""
        if self._mode == "vector" or (self._table is None and self._nodes is None):
            raise ValueError("push() needs a coefficient or node table in 'norm', 'both' or scalar mode")
        if self._nodes is not None and steps is not None:
            raise ValueError("the resumable push (slab-sharded tables) needs a cell table; a node table is replicated")
        for t, w in ((p, self._d), (v, 3)):
            if t.dtype != torch.float64 or not t.is_contiguous() or t.dim() != 2 or t.shape[1] != w or t.device != self._device:
                raise ValueError(f"pos must be (N,{self._d}) and vel (N,3): contiguous float64 on the interpolator's device")
        if p.shape[0] != v.shape[0]:
            raise ValueError("pos and vel must have the same number of rows")
        if steps is not None and (steps.dtype != torch.int64 or not steps.is_contiguous() or steps.shape != (p.shape[0],)
                                  or steps.device != self._device):
            raise ValueError("steps must be a contiguous int64 (N,) tensor on the interpolator's device")
        lost = torch.zeros(1, dtype=torch.int64, device=self._device)
        grav = (ctypes.c_double * 3)(*([0.0, 0.0, 0.0] if gravity is None else [float(x) for x in gravity]))
        with torch.cuda.device(self._device):
            stream = torch.cuda.current_stream(self._device).cuda_stream
            if self._nodes is not None:
                _lib.check(self._lib.arb_push_nodes(ctypes.byref(self._cgeom), self._nodes.data_ptr(), self._mode_code,
                                                    p.data_ptr(), v.data_ptr(), p.shape[0], float(dt), int(nsteps),
                                                    float(kappa), ctypes.byref(grav), lost.data_ptr(), stream),
                           "arb_push_nodes")
            else:
                _lib.check(self._lib.arb_push_steps(ctypes.byref(self._cgeom), self._table.data_ptr(), self._mode_code,
                                                    p.data_ptr(), v.data_ptr(), None if steps is None else steps.data_ptr(),
                                                    p.shape[0], float(dt), int(nsteps), float(kappa), ctypes.byref(grav),
                                                    lost.data_ptr(), stream), "arb_push")
        return int(lost.item())

    # ------------------------------------------------------------------ single-point queries
    def _single(self, query):
        """Shared part of sQuery1/2/3 (A.py:213-342 / 916-1062): NaN outside the volume, else one
        kernel evaluation.  The reference sums with np.inner and divides by h after summing, so the
        last bits differ from its own range query; both are within the parity tolerance."""
        d, geo = self._d, self._geo
        q = np.asarray(query.detach().cpu() if isinstance(query, torch.Tensor) else query, dtype=np.float64).ravel()
        if len(q) < d:
            raise IndexError(f"single query needs {d} coordinates")
        for a in range(d):
            if q[a] < geo.int_min[a] or q[a] > geo.int_max[a]:   # A.py:215, 918
                return None
        res = self._range_host(q[:d].reshape(1, d).copy())
        self.queryInd = int(self._last_cells[0])                # A.py:231-232
        return res

    def sQuery1(self, query):
        res = self._single(query)
        if res is None:
            return np.nan                                        # A.py:216
        return res[0][0]

    def sQuery2(self, query):
        res = self._single(query)
        if res is None:
            return np.nan                                        # A.py:254
        return res[1][0, 0], res[2][0]

    def sQuery3(self, query):
        res = self._single(query)
        if res is None:
            return np.nan                                        # A.py:299
        return res[0][0], res[1][0, 0], res[2][0]

    # ------------------------------------------------------------------ dispatch (A.py:177-211 / 880-914)
    def Query1(self, query):
        try:
            if query.shape[1] > 1:
                return self.rQuery1(query)
            return self.sQuery1(query)
        except IndexError:
            return self.sQuery1(query)

    def Query2(self, query):
        try:
            if query.shape[1] > 1:
                norms, grads = self.rQuery2(query)
                return norms, grads
            norm, grad = self.sQuery2(query)
            return norm, grad
        except IndexError:
            norm, grad = self.sQuery2(query)                     # TypeError outside the volume, as A.py:198
            return norm, grad

    def Query3(self, query):
        try:
            if query.shape[1] > 1:
                comps, norms, grads = self.rQuery3(query)
                return comps, norms, grads
            comps, norm, grad = self.sQuery3(query)
            return comps, norm, grad
        except IndexError:
            comps, norm, grad = self.sQuery3(query)
            return comps, norm, grad


class tricubic(_CubicInterpolator):
    """Tricubic interpolator of a 3-D gridded field (reference class ``tricubic``, A.py:8-621)."""
    _d = 3


class quadcubic(_CubicInterpolator):
    """Quadcubic interpolator of a 4-D gridded field (reference class ``quadcubic``, A.py:627-1381)."""
    _d = 4
