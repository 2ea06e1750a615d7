"""GPU tier, needs >= 2 GPUs (skipped on a 1-GPU box): replicated tables give bit-identical results on
every rank, and a t-slab-sharded quadcubic answers routed queries exactly like the unsharded table."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, field, q_all, out_dir):
    import torch.distributed as dist
    from arbinterp_b200 import quadcubic
    from arbinterp_b200.sharding import SlabShardedInterp, broadcast_field
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    # (a) replicate: rank 0 owns the field, everyone builds its own table
    rows = broadcast_field(field if rank == 0 else None, src=0, device=dev)
    whole = quadcubic(rows, "quiet", mode="both")
    res = whole.Query(q_all.copy())
    np.save(os.path.join(out_dir, f"whole{rank}.npy"), np.hstack(res))
    # (b) slab-sharded table, this rank's share of the rows routed to their owners and back
    sharded = SlabShardedInterp(quadcubic, field if rank == 0 else None, "quiet", mode="both")
    mine = torch.from_numpy(q_all[rank::world].copy()).to(dev)
    out = sharded.Query(mine)
    np.save(os.path.join(out_dir, f"slab{rank}.npy"), torch.cat(out, dim=1).cpu().numpy())
    np.save(os.path.join(out_dir, f"table_rows{rank}.npy"), np.array([sharded.local.table.shape[0], whole.table.shape[0]]))
    # (c) new field values on the same grid: rank 0 supplies them, every rank rebuilds its slab in place
    new_vals = np.cos(2.0 * field[:, 4:]) + field[:, [0]]
    sharded.update_values(new_vals if rank == 0 else None)
    fresh = quadcubic(np.concatenate([field[:, :4], new_vals], axis=1), "quiet", mode="both")
    out2 = sharded.Query(mine)
    ref2 = np.hstack(fresh.Query(q_all.copy()))[rank::world]
    np.save(os.path.join(out_dir, f"upd{rank}.npy"), np.array([np.array_equal(torch.cat(out2, dim=1).cpu().numpy(), ref2, equal_nan=True)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_replicated_and_slab_sharded_world2(tmp_path):
    import torch.multiprocessing as mp
    g = load_golden("quad_8x7x7x6")
    field, q_all = g["field"], g["both_q_in"][:, :4].copy()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), field, q_all, str(tmp_path)), nprocs=world, join=True)
    w0, w1 = np.load(tmp_path / "whole0.npy"), np.load(tmp_path / "whole1.npy")
    assert np.array_equal(w0, w1, equal_nan=True)            # same inputs -> bit-identical on every GPU
    for rank in range(world):
        got = np.load(tmp_path / f"slab{rank}.npy")
        assert np.array_equal(got, w0[rank::world], equal_nan=True)
        local_rows, whole_rows = np.load(tmp_path / f"table_rows{rank}.npy")
        assert local_rows < whole_rows                      # each rank really holds only its slab
        assert bool(np.load(tmp_path / f"upd{rank}.npy")[0]), "update_values on the sharded table"


def _push_field():
    ax = [np.linspace(-1, 1, 12), np.linspace(0, 1, 11), np.linspace(-1, 0, 10), np.linspace(0, 2, 19)]
    T, Z, Y, X = [a.ravel() for a in np.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")]
    return np.stack([X, Y, Z, T, 2 + np.sin(2 * X) * np.cos(3 * Y) * np.exp(Z) * np.cos(T) + 0.2 * X * Y * Z * T], axis=1)


def _push_particles(obj, n, rng):
    lo = np.array([obj.xIntMin, obj.yIntMin, obj.zIntMin, obj.tIntMin])
    hi = np.array([obj.xIntMax, obj.yIntMax, obj.zIntMax, obj.tIntMax])
    pos = lo + rng.uniform(0.05, 0.95, (n, 4)) * (hi - lo)
    pos[:, 3] = lo[3] + rng.uniform(0.0, 0.6, n) * (hi[3] - lo[3])       # time advances: every particle crosses slabs
    return pos, rng.normal(0, 0.2, (n, 3))


def _push_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from arbinterp_b200 import quadcubic
    from arbinterp_b200.sharding import SlabShardedInterp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    field = _push_field()
    whole = quadcubic(field.copy(), "quiet")
    pos, vel = _push_particles(whole, 4000, np.random.default_rng(9))
    dt, nsteps, kappa = 0.02, 40, -0.5
    pr, vr = torch.from_numpy(pos.copy()).to(dev), torch.from_numpy(vel.copy()).to(dev)
    lost_ref = whole.push(pr, vr, dt, nsteps, kappa, gravity=(0.0, 0.0, -0.1))
    sharded = SlabShardedInterp(quadcubic, field if rank == 0 else None, "quiet")
    p = torch.from_numpy(pos[rank::world].copy()).to(dev)
    v = torch.from_numpy(vel[rank::world].copy()).to(dev)
    lost = sharded.push(p, v, dt, nsteps, kappa, gravity=(0.0, 0.0, -0.1))
    np.savez(os.path.join(out_dir, f"push{rank}.npz"), p=p.cpu().numpy(), v=v.cpu().numpy(), lost=lost,
             pr=pr.cpu().numpy()[rank::world], vr=vr.cpu().numpy()[rank::world], lost_ref=lost_ref)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_push_equals_unsharded_world2(tmp_path):
    """SlabShardedInterp.push over NCCL: particles migrate between the two t-slab owners and end bit-identical to
    the fused push on an unsharded table (time-dependent field, gravity, some particles lost)."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_push_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    lost = 0
    for rank in range(world):
        z = np.load(tmp_path / f"push{rank}.npz")
        assert np.array_equal(z["p"], z["pr"], equal_nan=True) and np.array_equal(z["v"], z["vr"], equal_nan=True)
        assert int(z["lost"]) == int(np.isnan(z["pr"][:, 0]).sum())
        lost += int(z["lost"])
    assert lost == int(z["lost_ref"]) and lost > 0


def _fused_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from arbinterp_b200 import quadcubic, tricubic
    from arbinterp_b200.sharding import SlabShardedInterp
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    rng = np.random.default_rng(100 + rank)
    ok, ok_both, how = True, True, ""
    for cls, gname, d in ((quadcubic, "quad_8x7x7x6", 4), (tricubic, "tri_12x10x9", 3)):
        field = load_golden(gname)["field"]
        for mode in ("vector", "norm", "both"):
            whole = cls(field.copy(), "quiet", mode=mode)
            fused = SlabShardedInterp(cls, field if rank == 0 else None, "quiet", mode=mode, fused=True)
            plain = SlabShardedInterp(cls, field if rank == 0 else None, "quiet", mode=mode, fused=False)
            both = SlabShardedInterp(cls, field if rank == 0 else None, "quiet", mode=mode, fused="both")
            lo = np.array(whole._geo.int_min); hi = np.array(whole._geo.int_max)
            for n in (5000 + 37 * rank, 11, 0, 20000):            # growing, shrinking and empty batches re-use / re-map the buffers
                q = lo + rng.uniform(-0.05, 1.05, (n, d + 1))[:, :d] * (hi - lo)
                if n > 100:
                    q[::97, 1] = np.nan
                    k = np.arange(len(q[5::50]))                     # rows exactly on cell-layer (and slab) boundaries
                    q[5::50, d - 1] = lo[d - 1] + (k % whole._geo.ncell[d - 1]) * whole._geo.h[d - 1]
                qa, qb, qc = q.copy(), q.copy(), q.copy()
                ra, rb, rc = whole.Query(qa) if n else None, fused.Query(qb), plain.Query(qc)
                rb = rb if isinstance(rb, tuple) else (rb,)
                rc = rc if isinstance(rc, tuple) else (rc,)
                if n:
                    ra = ra if isinstance(ra, tuple) else (ra,)
                    ok &= all(np.array_equal(x, y, equal_nan=True) for x, y in zip(rb, ra))
                    ok &= bool(np.array_equal(fused.queryInds, whole.queryInds)) and bool(np.array_equal(qb, qa, equal_nan=True))
                ok &= all(np.array_equal(x, y, equal_nan=True) for x, y in zip(rb, rc))
                ok &= bool(np.array_equal(fused.queryInds, plain.queryInds))
                # both legs as kernels (arb_route_rows -> arb_query_inbox): no all-to-all at all
                qd = q.copy()
                rd = both.Query(qd)
                rd = rd if isinstance(rd, tuple) else (rd,)
                ok_both &= all(np.array_equal(x, y, equal_nan=True) for x, y in zip(rd, rb))
                ok_both &= bool(np.array_equal(both.queryInds, fused.queryInds)) and bool(np.array_equal(qd, qb, equal_nan=True))
                ok_both &= both._inbox is not None
            how = fused._peer.how if fused._peer is not None else "none"
    np.save(os.path.join(out_dir, f"fused{rank}.npy"), np.array([ok, ok_both]))
    with open(os.path.join(out_dir, f"how{rank}.txt"), "w") as f:
        f.write(how)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_peer_store_return_world2(tmp_path):
    """SlabShardedInterp.Query with the return leg fused into the kernel (arb_query_routed: results stored into the home
    rank's buffer over NVLink) is bit-identical to the all-to-all return and to the unsharded table -- outputs, global
    queryInds and in-place NaN rows; 3-D and 4-D, every mode, batches that grow, shrink and are empty.  The same for
    fused='both': the forward leg as a kernel too (arb_route_rows stores the rows into the owners' inboxes, arb_query_inbox
    evaluates them), no all-to-all on the query path."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_fused_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for rank in range(world):
        assert bool(np.load(tmp_path / f"fused{rank}.npy")[0])
        assert bool(np.load(tmp_path / f"fused{rank}.npy")[1]), "both legs fused (route kernel + inbox query) differ"
        assert open(tmp_path / f"how{rank}.txt").read() in ("symmetric_memory", "cuda_ipc")
