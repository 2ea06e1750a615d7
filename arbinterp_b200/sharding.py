"""Multi-GPU plumbing (one process per GPU, torch.distributed): slab planning, construction-time
broadcast and query routing.  Nothing here is on the per-query compute path -- a rank's queries are
evaluated by its local table with no collective; the only exchanges are

  * construction: the source rank ingests the rows once; the dense value planes (+ axes) are broadcast
    over NCCL/NVLink and every rank builds its replica or its slab locally (cheaper than moving a
    25-400 GB table), and
  * slab-sharded tables only: one all-to-all that moves each query row to the rank owning its
    slowest-axis cell layer, and one that returns the result rows.

The reference has no counterpart (it is single-process numpy); the arithmetic that decides the
owner is the reference's cell location, ``floor((t - tIntMin) / ht)`` (A.py:1081-1086).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch


def plan_slabs(n_layers: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced ``[lo, hi)`` ranges of the slowest-axis cell layers, one per rank.
    The first ``n_layers % world`` ranks get one layer more (61 layers / 8 ranks -> 8,8,8,8,8,7,7,7).
    Ranks beyond ``n_layers`` get an empty slab ``(n_layers, n_layers)``."""
    if n_layers < 1 or world < 1:
        raise ValueError("need n_layers >= 1 and world >= 1")
    base, extra = divmod(n_layers, world)
    out, lo = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((lo, lo + size))
        lo += size
    return out


def slab_planes(slab: Tuple[int, int]) -> Tuple[int, int]:
    """Grid planes ``[first, last)`` of the slowest axis a slab needs: cell layer k uses grid points
    k .. k+3 (offsets -1..+2 around its lower corner k+1), i.e. one halo plane below, two above."""
    lo, hi = slab
    return lo, hi + 3


def owner_ranks(coord_slow: torch.Tensor, int_min: float, int_max: float, h: float,
                slabs: Sequence[Tuple[int, int]]) -> torch.Tensor:
    """Owning rank of every query from its slowest-axis coordinate.  Rows that are outside the
    volume or NaN have no owner and are sent to rank 0, whose kernel NaN-masks them like any other
    out-of-volume row.  Same expression as the kernel's locate (subtract, divide, floor)."""
    # a true IEEE division like the kernel's locate: torch turns `tensor / python_scalar` into a multiplication by the
    # reciprocal on CUDA, which is one ulp off for rows exactly on a layer boundary and would send them to the wrong rank
    layer = torch.floor((coord_slow - int_min) / torch.full((1,), h, dtype=coord_slow.dtype, device=coord_slow.device))
    valid = (coord_slow >= int_min) & (coord_slow <= int_max)
    n_layers = slabs[-1][1]
    layer = torch.where(valid, layer, torch.zeros_like(layer)).clamp_(0, n_layers - 1).to(torch.int64)
    bounds = torch.tensor([hi for _, hi in slabs], dtype=torch.int64, device=coord_slow.device)
    owner = torch.searchsorted(bounds, layer, right=True)
    return torch.where(valid, owner, torch.zeros_like(owner))


def _no_mark(name: str) -> None:
    pass


def route_rows(rows: torch.Tensor, owner: torch.Tensor, group=None, mark: Callable[[str], None] = _no_mark,
               with_home_rows: bool = False):
    """Send every row of ``rows`` to rank ``owner[row]`` (one all-to-all of the counts, one of the rows).
    Returns ``(received, order, send_split, recv_split)``: ``order`` sorts the caller's rows by owner (what was
    sent is ``rows[order]``), the splits are what :func:`return_rows` needs to send answers back.
    ``with_home_rows``: a fifth value, the row number every received row has in its sender's batch (``order`` sent
    along in a second, 8-byte-per-row all-to-all) -- what the fused return leg addresses its peer stores with.
    ``mark(name)`` is called at the end of each phase (sort / counts / permute / alltoall) for profiling."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    # owners are tiny integers: a 16-bit stable radix sort is two passes instead of eight, the per-owner
    # counts fall out of the sorted keys, and the library's row-permutation kernel moves the 24-64 byte rows
    # ~30x faster than q[order] / index_select / gather (torch's row-gather kernel: ~60 GB/s on such rows)
    skey, order = torch.sort(owner.to(torch.int16), stable=True)
    bounds = torch.searchsorted(skey, torch.arange(world + 1, dtype=torch.int16, device=owner.device))
    send_counts = (bounds[1:] - bounds[:-1]).to(torch.int64)
    mark("sort")
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send_split = send_counts.tolist()
    recv_split = recv_counts.tolist()
    mark("counts")
    from ._lib import permute_rows
    sendbuf = permute_rows(rows, order)
    mark("permute")
    recvbuf = rows.new_empty((sum(recv_split), rows.shape[1]))
    dist.all_to_all_single(recvbuf, sendbuf, recv_split, send_split, group=group)
    if with_home_rows:
        home_rows = order.new_empty(sum(recv_split))
        dist.all_to_all_single(home_rows, order.contiguous(), recv_split, send_split, group=group)
        mark("alltoall")
        return recvbuf, order, send_split, recv_split, home_rows
    mark("alltoall")
    return recvbuf, order, send_split, recv_split


def return_rows(res: torch.Tensor, order: torch.Tensor, send_split, recv_split, group=None,
                mark: Callable[[str], None] = _no_mark) -> torch.Tensor:
    """Inverse of :func:`route_rows` for per-row results: ``res`` (one row per received row) travels back and is
    put into the caller's original row order."""
    import torch.distributed as dist
    from ._lib import permute_rows
    back = res.new_empty((sum(send_split), res.shape[1]))
    dist.all_to_all_single(back, res.contiguous(), send_split, recv_split, group=group)
    mark("alltoall_back")
    out = permute_rows(back, order, scatter=True)
    mark("scatter")
    return out


def exchange_and_query(q: torch.Tensor, owner: torch.Tensor, evaluate: Callable[[torch.Tensor], torch.Tensor],
                       out_cols: int, group=None, mark: Callable[[str], None] = _no_mark) -> torch.Tensor:
    """Route rows of ``q`` to their owners, evaluate there, route the results back.

    ``evaluate(rows) -> (len(rows), out_cols)`` runs on the receiving rank (the local slab's query
    kernel).  Returns ``(len(q), out_cols)`` in the caller's row order.  Two all-to-alls of rows (plus one of
    the per-rank counts), none of them inside the query kernel."""
    recvbuf, order, send_split, recv_split = route_rows(q, owner, group, mark)
    res = evaluate(recvbuf)
    mark("kernel")
    if res.shape != (recvbuf.shape[0], out_cols):
        raise ValueError(f"evaluate returned {tuple(res.shape)}, expected {(recvbuf.shape[0], out_cols)}")
    return return_rows(res, order, send_split, recv_split, group, mark)


def push_sharded(pos: torch.Tensor, vel: torch.Tensor, nsteps: int, owner_of: Callable[[torch.Tensor], torch.Tensor],
                 advance: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], None], group=None) -> int:
    """Particle push over a slab-sharded table: particles migrate between the slab owners.

    ``pos`` (N, d), ``vel`` (N, 3): this rank's particles, updated in place.  ``owner_of(pos) -> rank`` names the rank
    whose slab holds a position; ``advance(pos, vel, step)`` runs the local push kernel in its resumable form
    (arb_push_steps): it advances every row until it is finished (``step > nsteps``; lost particles are NaN and
    finished) or leaves the local slab (parked unchanged, ``step`` = the next force evaluation).  Rounds of
    {route parked rows to their owner, advance} repeat until no rank has a parked row; every round moves each
    parked particle by at least one step, so the loop ends.  Finished rows go home in one last exchange.
    All exchanges are all-to-alls of (d + 6)-double rows; none is inside the kernel.  Returns the number of this
    rank's particles that were lost."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    n, d = pos.shape
    dev = pos.device
    # row = [pos | vel | next step | home rank | home index]; the three integers are exact in float64
    rows = torch.empty((n, d + 6), dtype=torch.float64, device=dev)
    rows[:, :d] = pos
    rows[:, d:d + 3] = vel
    rows[:, d + 3] = 0.0
    rows[:, d + 4] = float(rank)
    rows[:, d + 5] = torch.arange(n, dtype=torch.float64, device=dev)
    finished = []
    while True:
        recv, _, _, _ = route_rows(rows, owner_of(rows[:, :d]), group)
        p, v = recv[:, :d].contiguous(), recv[:, d:d + 3].contiguous()
        step = recv[:, d + 3].to(torch.int64).contiguous()
        if recv.shape[0]:
            advance(p, v, step)
        recv[:, :d], recv[:, d:d + 3], recv[:, d + 3] = p, v, step.to(torch.float64)
        done = step > nsteps
        finished.append(recv[done])
        rows = recv[~done].contiguous()
        left = torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev)
        dist.all_reduce(left, group=group)
        if int(left.item()) == 0:
            break
    fin = torch.cat(finished, dim=0)
    home, _, _, _ = route_rows(fin, fin[:, d + 4].to(torch.int64), group)
    if home.shape[0] != n:
        raise RuntimeError(f"push_sharded: {home.shape[0]} particles came home, {n} left")
    idx = home[:, d + 5].to(torch.int64)
    pos[idx] = home[:, :d]
    vel[idx] = home[:, d:d + 3]
    return int(torch.isnan(home[:, 0]).sum().item())


def broadcast_field(field, src: int = 0, device=None, group=None) -> torch.Tensor:
    """Construction-time replicate: rank ``src`` holds the (N, cols) field, everyone gets a copy."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    if rank == src:
        t = torch.as_tensor(field).to(device=device, dtype=torch.float64).contiguous()
        shape = torch.tensor(list(t.shape), dtype=torch.int64, device=t.device)
    else:
        shape = torch.empty(2, dtype=torch.int64, device=device)
    dist.broadcast(shape, src=src, group=group)
    if rank != src:
        t = torch.empty(tuple(shape.tolist()), dtype=torch.float64, device=device)
    dist.broadcast(t, src=src, group=group)
    return t


def broadcast_ingested(field, d: int, src: int = 0, device=None, group=None):
    """Construction-time replicate, done right: rank ``src`` ingests the raw rows once (sort-free grid
    indexing + scatter into dense planes), then the per-axis coordinates and the dense value planes are
    broadcast (NCCL over NVLink) -- 1/(d+C) .. 1/2 of the raw rows' bytes, and no rank repeats the ingest.
    Returns an :class:`~arbinterp_b200.ingest.IngestedField` on every rank; geometry is derived from the
    broadcast axes with the same expressions everywhere, so it is bit-identical across ranks."""
    import torch.distributed as dist
    from .ingest import IngestedField, geometry_from_axes, ingest_field
    rank = dist.get_rank(group)
    meta = torch.zeros(2 + d, dtype=torch.int64, device=device)
    if rank == src:
        planes, geo = ingest_field(field, d, device=device)
        meta[0], meta[1] = 1, planes.shape[0]
        meta[2:] = torch.tensor(geo.npts, dtype=torch.int64)
        axes = torch.cat(geo.axes).contiguous()
    dist.broadcast(meta, src=src, group=group)
    ncomp, npts = int(meta[1]), [int(v) for v in meta[2:]]
    if rank != src:
        axes = torch.empty(sum(npts), dtype=torch.float64, device=device)
        planes = torch.empty([ncomp] + npts[::-1], dtype=torch.float64, device=device)
    dist.broadcast(axes, src=src, group=group)
    dist.broadcast(planes, src=src, group=group)
    per_axis, off = [], 0
    for n in npts:
        per_axis.append(axes[off:off + n].clone())
        off += n
    out = IngestedField(planes=planes, geo=geometry_from_axes(per_axis))
    if rank == src:
        out.geo.row_index = geo.row_index          # row -> grid point map of the raw rows (update_values(order='rows'))
    return out


class PeerResults:
    """One result buffer per rank, ``(rows, ld)`` float64, mapped into every rank of the group, so that the routed form
    of the query kernel (``arb_query_routed``) can store a row's outputs straight into its home rank's buffer over
    NVLink / NVSwitch.  Mapping: ``torch.distributed._symmetric_memory`` when it is available and works for the group,
    else CUDA IPC handles exchanged with ``all_gather_object`` (``torch.multiprocessing.reductions``) plus
    ``cudaDeviceEnablePeerAccess``.  Raises when neither works (the caller falls back to the all-to-all return)."""

    def __init__(self, lib, device, group, rows: int, ld: int):
        import torch.distributed as dist
        self.rows, self.ld = int(rows), int(ld)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.how = None
        self._keep = []
        errors = []
        try:
            import torch.distributed._symmetric_memory as symm_mem
            buf = symm_mem.empty((self.rows, self.ld), dtype=torch.float64, device=device)
            name = (group if group is not None else dist.group.WORLD).group_name
            hdl = symm_mem.rendezvous(buf, name)
            self.buf, self.ptrs, self.how = buf, [int(x) for x in hdl.buffer_ptrs], "symmetric_memory"
            self._keep.append(hdl)
        except Exception as e:                                   # noqa: BLE001 -- try the IPC route
            errors.append(f"symmetric_memory: {type(e).__name__}: {e}")
        # every rank must take the same route
        ok = torch.tensor([int(self.how is not None)], dtype=torch.int64, device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if not bool(ok.item()):
            self.how = None
            try:
                from torch.multiprocessing.reductions import reduce_tensor
                self.buf = torch.empty((self.rows, self.ld), dtype=torch.float64, device=device)
                fn, args = reduce_tensor(self.buf)
                gathered = [None] * self.world
                dist.all_gather_object(gathered, args, group=group)
                self.ptrs = []
                with torch.cuda.device(device):
                    for r in range(self.world):
                        if r == self.rank:
                            self.ptrs.append(self.buf.data_ptr())
                            continue
                        t = fn(*gathered[r])
                        self._keep.append(t)
                        if t.device != self.buf.device:
                            from . import _lib
                            _lib.check(lib.arb_enable_peer_access(t.device.index), "arb_enable_peer_access")
                        self.ptrs.append(t.data_ptr())
                self.how = "cuda_ipc"
            except Exception as e:                               # noqa: BLE001
                errors.append(f"cuda_ipc: {type(e).__name__}: {e}")
            ok = torch.tensor([int(self.how is not None)], dtype=torch.int64, device=device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if not bool(ok.item()):
                raise RuntimeError("no peer mapping of the result buffers: " + "; ".join(errors))
        import ctypes
        self.cptrs = (ctypes.c_void_p * self.world)(*self.ptrs)


class ReplicatedInterp:
    """One process per GPU, the coefficient table replicated: rank ``src`` holds the field, it is ingested once, the
    dense planes are broadcast (NCCL over NVLink) and every rank builds its own table (cheaper than moving 8-33 GB of
    coefficients).  ``Query`` takes THIS rank's rows -- numpy or CUDA -- and is exactly the single-GPU call: the
    reference's return shapes, in-place NaN rows and ``queryInds`` (A.py:177-211, 350-355, 368-370), no collective.
    Every other attribute is the local interpolator's."""

    def __init__(self, cls, field, *args, group=None, src: int = 0, **kwargs):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        device = kwargs.get("device")
        if device is None and torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())
        kwargs["device"] = device
        full = broadcast_ingested(field, cls._d, src=src, device=device, group=group)
        self.local = cls(full, *args, **kwargs)

    def Query(self, q):
        return self.local.Query(q)

    def __getattr__(self, name):
        if name == "local":
            raise AttributeError(name)
        return getattr(self.local, name)


class SlabShardedInterp:
    """A tricubic/quadcubic whose coefficient table is sharded over the ranks by slowest-axis
    slabs (config 5: 96^3 x 64 'both' = 402 GB over 8 GPUs).  ``Query`` takes this rank's rows and
    returns this rank's results; rows travel to the owning rank and back."""

    def __init__(self, cls, field, *args, group=None, src: int = 0, **kwargs):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        kwargs_fused = kwargs.pop("fused", "auto")
        device = kwargs.get("device") or torch.device("cuda", torch.cuda.current_device())
        d = cls._d
        full = broadcast_ingested(field, d, src=src, device=device, group=group)
        n_slow = full.geo.ncell[d - 1]
        self.slabs = plan_slabs(n_slow, self.world)
        lo, hi = self.slabs[self.rank]
        if hi == lo:
            raise ValueError(f"rank {self.rank} has an empty slab: {n_slow} layers over {self.world} ranks")
        self.local = cls(full, *args, slab=(lo, hi), **kwargs)
        self._src = src
        self._ncol = int(full.planes.shape[0])      # 1 (scalar input) or 3 (vector input)
        del full
        self.d = d
        g = self.local._geo
        self._slow = (g.int_min[d - 1], g.int_max[d - 1], g.h[d - 1])
        # return leg fused into the query kernel (results stored into the home rank's buffer over NVLink); "auto" tries
        # it on CUDA + NCCL and falls back to the all-to-all return when the peer mapping is not available
        self.fused = kwargs_fused
        self._peer = None
        self._inbox = None
        self._peer_error = None

    def Query(self, q, timing: Optional[dict] = None):
        """Drop-in range query over the sharded table (rQuery1/2/3, A.py:1064-1258 / 344-521): ``q`` holds THIS rank's
        rows, (N, >=d), as a numpy array or a CUDA tensor; returns what the unsharded class returns for them (numpy in ->
        numpy out, tensor in -> CUDA tensors), in the caller's row order.  Side effects as in the reference: rows with a
        coordinate outside the volume are overwritten with NaN in the caller's array across all columns
        (A.py:1069-1076), and ``queryInds`` holds the GLOBAL cell index of every row, ``nc`` for NaN rows
        (A.py:1088-1090).  Collective: every rank of the group must call it (an empty batch is fine).
        ``timing``: optional dict that accumulates the milliseconds of each phase -- owner / sort / counts / permute /
        alltoall / kernel / alltoall_back / scatter / unpack (CUDA events on the current stream)."""
        d, mode = self.d, self.local._mode
        dev = self.local._device
        host = isinstance(q, np.ndarray)
        if host:
            if q.ndim != 2 or q.shape[1] < d:
                raise IndexError(f"query must be (N, >={d})")
            coords = torch.from_numpy(np.ascontiguousarray(q[:, :d], dtype=np.float64)).to(dev)
        else:
            if q.dim() != 2 or q.shape[1] < d:
                raise IndexError(f"query must be (N, >={d})")
            coords = q[:, :d].to(device=dev, dtype=torch.float64).contiguous()
        ev = []

        def mark(name):
            if timing is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream(dev))
                ev.append((name, e))

        mark("start")
        g = self.local._geo
        widths = {"vector": (3,), "norm": (1, d), "both": (3, 1, d)}[mode]
        flat = outside = None
        if self.fused in ("both", "auto") and coords.is_cuda and self.world <= 16:
            got = self._query_both_legs(coords, sum(widths) + 1, mark)      # None: peer mapping unavailable
            if got is not None:
                flat, outside = got
        if flat is not None:
            pass
        elif coords.is_cuda and self.world <= 16:
            # one kernel for both (arb_owner_keys): owner rank of every row and the out-of-volume mask
            import ctypes
            from . import _lib
            owner = torch.empty(coords.shape[0], dtype=torch.int16, device=dev)
            outside = torch.empty(coords.shape[0], dtype=torch.bool, device=dev)
            his = (ctypes.c_int64 * self.world)(*[s_[1] for s_ in self.slabs])
            with torch.cuda.device(dev):
                _lib.check(self.local._lib.arb_owner_keys(ctypes.byref(self.local._cgeom), coords.data_ptr(), coords.shape[0],
                                                          coords.shape[1], his, self.world, owner.data_ptr(),
                                                          outside.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                           "arb_owner_keys")
        else:
            lo = torch.tensor(g.int_min, dtype=torch.float64, device=dev)
            hi = torch.tensor(g.int_max, dtype=torch.float64, device=dev)
            outside = ((coords < lo) | (coords > hi)).any(dim=1)                # A.py:1069-1076
            owner = owner_ranks(coords[:, d - 1], *self._slow, self.slabs)
        if flat is None:
            mark("owner")

        def evaluate(rows):
            res = self.local.Query(rows)
            res = res if isinstance(res, tuple) else (res,)
            cells = self.local._last_cells.view(torch.float64).unsqueeze(1)     # int64 bits ride along exactly
            return torch.cat(list(res) + [cells], dim=1)

        if flat is None and self.fused and self.world <= 16:
            flat = self._query_fused(coords, owner, sum(widths) + 1, mark)
        if flat is None:
            flat = exchange_and_query(coords, owner, evaluate, sum(widths) + 1, self.group, mark)
        outs, col = [], 0
        for w in widths:
            outs.append(flat[:, col:col + w].contiguous())
            col += w
        self._last_cells = flat[:, col].contiguous().view(torch.int64)
        if host:
            bad = outside.cpu().numpy()
            if bad.any():
                q[np.where(bad)[0]] = np.nan                                    # raises for int arrays, as A.py:1069 does
            outs = [o.cpu().numpy() for o in outs]
        elif bool(outside.any()):
            q[outside.to(q.device)] = float("nan")
        mark("unpack")
        if timing is not None:
            torch.cuda.synchronize(dev)
            for (_, e0), (name, e1) in zip(ev[:-1], ev[1:]):
                timing[name + "_ms"] = timing.get(name + "_ms", 0.0) + e0.elapsed_time(e1)
            timing["total_ms"] = timing.get("total_ms", 0.0) + ev[0][1].elapsed_time(ev[-1][1])
        return outs[0] if len(outs) == 1 else tuple(outs)

    def _ensure_peer_buffers(self, n, ld, inbox: bool):
        """(Re)map, collectively, the buffers the fused legs store into: every rank's result buffer (its own batch, ``ld``
        doubles per row) and -- for the fused forward leg -- its inbox (``world`` segments of ``cap`` rows) and the
        per-sender row counts.  Returns False when no peer mapping works (remembered: the group falls back together)."""
        import torch.distributed as dist
        loc = self.local
        dev = loc._device
        if self._peer_error is not None or dev.type != "cuda" or dist.get_backend(self.group) != "nccl":
            return False
        need = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(need, op=dist.ReduceOp.MAX, group=self.group)
        need = int(need.item())
        try:
            if self._peer is None or self._peer.rows < need or self._peer.ld != ld:
                self._peer = PeerResults(loc._lib, dev, self.group, max(need, 1) * 5 // 4 + 1024, ld)
                self._inbox = None
            if inbox and getattr(self, "_inbox", None) is None:
                cap = self._peer.rows
                ld_in = (self.d + 2) // 2 * 2
                self._inbox = PeerResults(loc._lib, dev, self.group, self.world * cap, ld_in)
                self._counts = PeerResults(loc._lib, dev, self.group, 1, 16)          # int64 [world] behind a float64 row
                self._cursor = torch.zeros(self.world + 2, dtype=torch.int64, device=dev)   # [world] cursors + the ticket
                self._inbox_cap = cap
        except Exception as e:                                   # noqa: BLE001 -- remembered: the group falls back together
            self._peer, self._inbox, self._peer_error = None, None, f"{type(e).__name__}: {e}"
            if self.fused in (True, "both"):
                raise
            return False
        return True

    def _query_both_legs(self, coords, ld, mark):
        """Both legs of a routed query inside kernels, no NCCL all-to-all and no host round trip: ``arb_route_rows`` stores
        every row, with its row number here, into the inbox of the rank that owns its slab (peer memory over NVLink; the
        owner and the out-of-volume mask come out of the same kernel), a barrier, ``arb_query_inbox`` on the owner evaluates
        its inbox -- the per-sender counts are read on the device -- and stores every row's outputs into the home rank's
        result buffer at its home row, a barrier.  Returns ``(result rows (n, ld), outside mask)`` or None when the peer
        mapping is unavailable."""
        import ctypes
        import torch.distributed as dist
        from . import _lib
        loc = self.local
        dev = loc._device
        d, n, world = self.d, coords.shape[0], self.world
        ld = (ld + 1) // 2 * 2
        if not self._ensure_peer_buffers(n, ld, inbox=True):
            return None
        peer, inbox, counts = self._peer, self._inbox, self._counts
        outside = torch.empty(n, dtype=torch.bool, device=dev)
        his = (ctypes.c_int64 * world)(*[s_[1] for s_ in self.slabs])
        self._cursor.zero_()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(loc._lib.arb_route_rows(ctypes.byref(loc._cgeom), coords.data_ptr(), n, coords.shape[1], his, world,
                                               self.rank, inbox.cptrs, counts.cptrs, self._inbox_cap,
                                               self._cursor.data_ptr(), self._cursor[world + 1:].data_ptr(),
                                               outside.data_ptr(), stream), "arb_route_rows")
            mark("owner")
            dist.barrier(group=self.group)      # every rank's rows and counts have landed in the inboxes
            mark("alltoall")
            _lib.check(loc._lib.arb_query_inbox(ctypes.byref(loc._cgeom), loc.table.data_ptr(), loc._mode_code,
                                                inbox.buf.data_ptr(), counts.buf.data_ptr(), self._inbox_cap, peer.cptrs,
                                                world, peer.ld, stream), "arb_query_inbox")
            mark("kernel")
            dist.barrier(group=self.group)      # every rank's kernel has finished: all rows of this batch have landed
            mark("alltoall_back")
        return peer.buf[:n], outside

    def _query_fused(self, coords, owner, ld, mark):
        """Rows travel to their owner with their home row number; the owner's kernel (``arb_query_routed``) stores every
        row's outputs into the home rank's result buffer at that row -- peer memory over NVLink, each result row as
        16-byte pieces from adjacent lanes -- so the results arrive in the caller's order with no return all-to-all and no
        re-ordering pass.  Returns the (n, ld) result rows of
        this rank, or None when the peer mapping is unavailable (the caller then takes the all-to-all return)."""
        import ctypes
        import torch.distributed as dist
        from . import _lib
        loc = self.local
        dev = loc._device
        d, n = self.d, coords.shape[0]
        ld = (ld + 1) // 2 * 2                  # result rows leave the kernel as 16-byte pieces
        if not self._ensure_peer_buffers(n, ld, inbox=False):
            return None
        peer = self._peer
        recv, _, _, recv_split, home_rows = route_rows(coords, owner, self.group, mark, with_home_rows=True)
        m = recv.shape[0]
        world = self.world
        seg_start = (ctypes.c_int64 * (world + 1))()
        for h in range(world):
            seg_start[h + 1] = seg_start[h] + recv_split[h]
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(loc._lib.arb_query_routed(ctypes.byref(loc._cgeom), loc.table.data_ptr(), loc._mode_code,
                                                 recv.data_ptr(), m, d, seg_start, home_rows.data_ptr(), peer.cptrs, world,
                                                 peer.ld, stream), "arb_query_routed")
        mark("kernel")
        dist.barrier(group=self.group)          # every rank's kernel has finished: all rows of this batch have landed
        mark("alltoall_back")
        return peer.buf[:n]

    @property
    def queryInds(self):
        """Global cell index of every row of the last ``Query`` on this rank; ``nc`` for NaN rows (A.py:1088-1090)."""
        c = getattr(self, "_last_cells", None)
        if c is None:
            raise AttributeError("queryInds is set by the first range query")
        return c.cpu().numpy()

    @property
    def nc(self):
        return self.local.nc

    def push(self, pos: torch.Tensor, vel: torch.Tensor, dt, nsteps, kappa, gravity=None) -> int:
        """Fused query + push (``tricubic.push`` / ``quadcubic.push``) over the sharded table: this rank's particles
        (``pos`` (N, d), ``vel`` (N, 3), CUDA tensors, updated in place) are advanced by ``nsteps`` velocity-Verlet steps;
        a particle that crosses into another rank's slab is parked by the kernel and resumed there with the
        arithmetic of an unsharded run, so the result is bit-identical to ``push`` on an unsharded table.
        Particles whose position is outside the volume (or NaN) at the start are sent to rank 0, which marks them
        lost.  Returns the number of this rank's particles lost."""
        d = self.d

        def owner_of(p):
            return owner_ranks(p[:, d - 1], *self._slow, self.slabs)

        def advance(p, v, step):
            self.local._push_local(p, v, step, dt, nsteps, kappa, gravity)

        return push_sharded(pos, vel, int(nsteps), owner_of, advance, self.group)


    def update_values(self, values=None, order: str = "rows") -> None:
        """New field values on the same grid (``tricubic.update_values``) for the sharded table: the source rank
        passes ``values`` ((N, 1|3), in the row order of the field it constructed from, or ``order='grid'``), every
        other rank passes None; the dense planes are broadcast once and every rank rebuilds its slab in place."""
        import torch.distributed as dist
        loc = self.local
        geo = loc._geo
        total = 1
        for n in geo.npts:
            total *= n
        dev = loc._device
        planes = torch.empty((self._ncol, total), dtype=torch.float64, device=dev)
        if self.rank == self._src:
            v = torch.as_tensor(values)
            if v.dim() == 1:
                v = v.unsqueeze(1)
            if tuple(v.shape) != (total, self._ncol):
                raise ValueError(f"values must have shape ({total}, {self._ncol}), got {tuple(v.shape)}")
            v = v.to(device=dev, dtype=torch.float64)
            if order == "rows":
                planes[:, geo.row_index.to(dev).long()] = v.T
            elif order == "grid":
                planes.copy_(v.T)
            else:
                raise ValueError("order must be 'rows' or 'grid'")
        dist.broadcast(planes, src=self._src, group=self.group)
        loc.update_values(planes.T, order="grid")
