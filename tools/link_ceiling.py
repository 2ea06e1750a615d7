#!/usr/bin/env python
"""Host<->device link ceiling of the box: N concurrent processes (one per GPU), each running pinned cudaMemcpyAsync
H2D and D2H copies at once, aggregate GB/s over the slowest rank -- the denominator of `e2e.link_frac` in bench.py
(which embeds the same measurement).  Separates "the box cannot move more" from "our host path is slow".

    python tools/link_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/link_ceiling.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_gpus": world}
    for bound in (False, True):
        if bound and world > 1:
            out["binding"] = bench.bind_to_gpu_cores(local, local, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        elif bound:
            continue
        for mb in (64, 256, 1024):
            r = bench.link_ceiling(torch, dist, dev, world, mb=mb, reps=4)
            out[f"{'bound' if bound else 'unbound'}_{mb}MiB"] = {k: round(v, 1) for k, v in r.items()}
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
