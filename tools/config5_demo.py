#!/usr/bin/env python
"""BASELINE config 5 at scale: 4-D vector field, quadcubic mode='both', coefficient table t-slab-sharded
over the ranks (halo planes included), trajectory-like queries routed to the owning rank.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config5_demo.py --grid 96,96,96,64

Parity at full size without a CPU table (the reference cannot allocate 402 GB): the field is
Bx = P(x,y,z,t) > 0, By = Bz = 0 with P a per-axis quadratic without an x*y*z*t monomial, so comps,
norm (= sqrt(Bx^2) = Bx exactly) and gradient are reproduced to round-off by the quadcubic interpolant
 and every routed result can be checked analytically."""
import argparse
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import quadcubic  # noqa: E402
from arbinterp_b200.sharding import SlabShardedInterp  # noqa: E402


def P(x, y, z, t):
    return 2 + 0.3 * x + 0.2 * y * z - 0.4 * z * t + 0.1 * x * x * t + 0.5 * t * t + 0.25 * y * y


def gradP(x, y, z, t):
    return torch.stack([0.3 + 0.2 * x * t, 0.2 * z + 0.5 * y, 0.2 * y - 0.4 * t, -0.4 * z + 0.1 * x * x + t], dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", default="96,96,96,64")
    ap.add_argument("--particles", type=int, default=1 << 22, help="particles per rank")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--push-steps", type=int, default=0, help="also run the sharded fused push for this many steps")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    shape = [int(v) for v in a.grid.split(",")]
    rows = None
    warm = torch.zeros(1, device=dev)
    dist.all_reduce(warm)                                   # NCCL communicator set-up is not construction time
    dist.broadcast(warm, src=0)
    torch.cuda.synchronize()
    if rank == 0:
        ax = [torch.linspace(-1, 1, shape[0], dtype=torch.float64, device=dev),
              torch.linspace(-1, 1, shape[1], dtype=torch.float64, device=dev),
              torch.linspace(-1, 1, shape[2], dtype=torch.float64, device=dev),
              torch.linspace(0, 1, shape[3], dtype=torch.float64, device=dev)]
        T, Z, Y, X = [g.reshape(-1) for g in torch.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")]
        zero = torch.zeros_like(X)
        rows = torch.stack([X, Y, Z, T, P(X, Y, Z, T), zero, zero], dim=1)
        del T, Z, Y, X, zero
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    obj = SlabShardedInterp(quadcubic, rows, "quiet", mode="both")
    del rows
    torch.cuda.synchronize(); dist.barrier()
    t_build = time.perf_counter() - t0
    loc = obj.local
    table_gb = loc.table.numel() * 8 / 1e9
    g = loc._geo
    lo = torch.tensor(g.int_min, dtype=torch.float64, device=dev)
    hi = torch.tensor(g.int_max, dtype=torch.float64, device=dev)
    gen = torch.Generator(device=dev); gen.manual_seed(777 + rank)
    n = a.particles
    pos = lo + torch.rand(n, 4, generator=gen, dtype=torch.float64, device=dev) * (hi - lo) * (1 - 1e-9)
    vel = (torch.rand(n, 4, generator=gen, dtype=torch.float64, device=dev) - 0.5) * (hi - lo) * 0.02
    vel[:, 3] = (hi[3] - lo[3]) * 0.01                      # time advances for every particle
    worst = 0.0
    times = []
    for step in range(a.steps + 2):
        # smooth random walk, reflected into the volume; time wraps around
        pos = pos + vel + 0.002 * (hi - lo) * torch.randn(n, 4, generator=gen, dtype=torch.float64, device=dev) * \
            torch.tensor([1, 1, 1, 0], dtype=torch.float64, device=dev)
        span = (hi - lo) * (1 - 1e-9)
        rel = torch.remainder(pos - lo, 2 * span)
        rel = torch.where(rel > span, 2 * span - rel, rel)
        rel[:, 3] = torch.remainder(pos[:, 3] - lo[3], span[3])
        q = (lo + rel).contiguous()
        torch.cuda.synchronize(); dist.barrier()
        t1 = time.perf_counter()
        comps, norm, grad = obj.Query(q)
        torch.cuda.synchronize(); dist.barrier()
        if step >= 2:
            times.append(time.perf_counter() - t1)
        x, y, z, t = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
        want = P(x, y, z, t)
        err = max(float((comps[:, 0] - want).abs().max()), float(comps[:, 1:].abs().max()),
                  float((norm[:, 0] - want).abs().max()),
                  float(((grad - gradP(x, y, z, t)).abs() * torch.tensor(g.h, device=dev)).max()))
        worst = max(worst, err)
    if os.environ.get("ARB_PROFILE_ROUTING") and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            obj.Query(q); torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60), flush=True)
    elif os.environ.get("ARB_PROFILE_ROUTING"):
        obj.Query(q); torch.cuda.synchronize()
    w = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    tb = torch.tensor([table_gb], dtype=torch.float64, device=dev)
    dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    if rank == 0:
        dt = sum(times) / len(times)
        print(f"[config5] grid={shape} world={world} slabs={obj.slabs} table total {float(tb):.1f} GB "
              f"({table_gb:.1f} GB on rank 0) construction (ingest on rank 0, NCCL broadcast of planes, slab build) {t_build:.2f} s", flush=True)
        print(f"[config5] {world * n} routed queries/step in {dt * 1e3:.2f} ms -> {world * n / dt:.3e} q/s "
              f"(route + query + route back); max |error| vs analytic (h-scaled for gradients) {float(w):.3e}", flush=True)
        assert float(w) < 1e-11
    if a.push_steps > 0:
        # trajectories through the sharded table: every particle's time runs through ~40 % of the t range, so
        # most particles cross at least one slab boundary and migrate to the next owner
        gen.manual_seed(4242 + rank)
        p4 = lo + torch.rand(n, 4, generator=gen, dtype=torch.float64, device=dev) * (hi - lo) * \
            torch.tensor([1, 1, 1, 0.55], dtype=torch.float64, device=dev)
        v3 = (torch.rand(n, 3, generator=gen, dtype=torch.float64, device=dev) - 0.5) * 0.05
        dtp = float(hi[3] - lo[3]) * 0.4 / a.push_steps
        t_first = p4[:, 3].clone()
        torch.cuda.synchronize(); dist.barrier()
        t1 = time.perf_counter()
        lost = obj.push(p4, v3, dtp, a.push_steps, -0.05)
        torch.cuda.synchronize(); dist.barrier()
        tp = time.perf_counter() - t1
        ok = ~torch.isnan(p4[:, 0])
        t_err = float((p4[ok, 3] - (t_first[ok] + a.push_steps * dtp)).abs().max()) if bool(ok.any()) else 0.0
        tot = torch.tensor([float(lost), float(n)], dtype=torch.float64, device=dev)
        dist.all_reduce(tot)
        if rank == 0:
            print(f"[config5] sharded fused push: {int(tot[1])} particles x {a.push_steps} steps in {tp * 1e3:.1f} ms -> "
                  f"{float(tot[1]) * a.push_steps / tp:.3e} particle-steps/s incl. migration between slab owners; "
                  f"lost {int(tot[0])}; max |t - (t0 + n dt)| of survivors {t_err:.1e}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
