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
