#!/usr/bin/env python
"""bench.py -- throughput of the ARBInterp interpolation hot path on B200.

Workload (BASELINE.json north_star / configs[2], value+gradient variant): synthetic analytic
256^3 vector field, tricubic, uniformly random in-volume queries; each rank holds a replica of the
coefficient table and evaluates its own query batch (weak scaling, no collective on the query path).
A "step" is one pass of the query kernel over one batch of Q queries.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode norm|vector|both] [--impl reference]

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline`, `cpu_baseline`, `e2e`,
`clocks` and `gpu_launches` are described in DESIGN.md "Measurement".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES = {  # coords in + one coefficient block per component + outputs, no reuse
    (3, "norm"): 568, (3, "vector"): 1584, (3, "both"): 2128,
    (4, "norm"): 2120, (4, "vector"): 6200, (4, "both"): 8288,
}
METRIC = "interp queries/sec (value+grad)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--mode", default="norm", choices=["norm", "vector", "both"])
    ap.add_argument("--queries", type=int, default=1 << 26, help="queries per step and rank (device-resident leg)")
    ap.add_argument("--e2e-queries", type=int, default=1 << 24, help="queries per step for the host-buffer leg")
    ap.add_argument("--variant", type=int, default=None, help="query-kernel variant (arb_set_query_variant)")
    ap.add_argument("--sweep", default="", help="comma list of variants to time (prints a table to stderr)")
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--other-modes", action="store_true", help="(default at N=1) also time vector/both, device-resident")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the vector/both lines")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------
# synthetic analytic field
# ----------------------------------------------------------------------------------------
def analytic_planes(torch, n, device):
    ax = torch.linspace(-1.0, 1.0, n, dtype=torch.float64, device=device)
    Z, Y, X = torch.meshgrid(ax, ax, ax, indexing="ij")
    bx = torch.sin(2 * np.pi * X) * torch.cos(np.pi * Y) * torch.exp(-Z)
    by = X * X * Y + Z
    bz = torch.cos(X + Y + Z)
    return ax, (X, Y, Z), (bx, by, bz)


def analytic_field_rows(torch, n, device):
    ax, (X, Y, Z), (bx, by, bz) = analytic_planes(torch, n, device)
    return ax, torch.stack([t.reshape(-1) for t in (X, Y, Z, bx, by, bz)], dim=1)


class ClockSampler(threading.Thread):
    """SM clock / throttle-reason samples DURING the timed region (B200_PROFILING.md).  NVML is polled every
    5 ms from a thread (the timed region of the default run is ~60 ms, too short for `nvidia-smi -lms`);
    nvidia-smi is the fallback when pynvml is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.samples = []          # (sm_mhz, set of reasons)
        self.max_mhz = None
        self.proc = None
        self._halt = threading.Event()
        self.source = None

    def _run_nvml(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        names = (("hw_slowdown", "nvmlClocksThrottleReasonHwSlowdown"),
                 ("hw_thermal_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown"),
                 ("sw_thermal_slowdown", "nvmlClocksThrottleReasonSwThermalSlowdown"),
                 ("sw_power_cap", "nvmlClocksThrottleReasonSwPowerCap"))
        bits = [(n, getattr(pynvml, a)) for n, a in names if hasattr(pynvml, a)]
        self.source = "nvml"
        while not self._halt.is_set():
            mhz = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            self.samples.append((mhz, {n for n, b in bits if mask & b}))
            time.sleep(0.005)

    def _run_smi(self):
        self.source = "nvidia-smi"
        self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu), "-lms", "100"], stdout=subprocess.PIPE, text=True)
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                reasons = {n for n, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                            ("sw_power_cap", 8)) if r[col].lower().startswith("active")}
                self.samples.append((float(r[1]), reasons))
                self.max_mhz = float(r[2])
            except (ValueError, IndexError):
                continue

    def run(self):
        try:
            self._run_nvml()
        except Exception:
            try:
                self._run_smi()
            except Exception:
                pass

    def wait_first_sample(self, timeout=5.0):
        t0 = time.perf_counter()
        while not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.01)

    def stop(self):
        self._halt.set()
        if self.proc:
            self.proc.terminate()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "source": self.source}
        sm = sorted(s[0] for s in self.samples)
        reasons = set().union(*[s[1] for s in self.samples])
        busy = sm[len(sm) // 2:]                      # the upper half: samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons),
                "samples": len(sm), "source": self.source}


def bind_to_gpu_numa_node(gpu_index):
    """Pin this rank to the CPU cores NVML reports as local to its GPU, so the pinned host buffers of the
    end-to-end leg are allocated on (and copied from) the NUMA node the GPU's PCIe root hangs off."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cores [{allowed[0]}..{allowed[-1]}]"
    except Exception:
        return None
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(mode):
    """DRAM bytes per query from the committed ncu --set full capture (profiles/roofline_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(mode)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (port of the reference) on the host cores
# ----------------------------------------------------------------------------------------
def make_cpu_oracle(n, mode):
    from oracle.arb_oracle import OracleInterp
    ax = np.linspace(-1.0, 1.0, n)
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")
    bx = np.sin(2 * np.pi * X) * np.cos(np.pi * Y) * np.exp(-Z)
    by = X * X * Y + Z
    bz = np.cos(X + Y + Z)
    vals = {}
    if mode in ("vector", "both"):
        vals.update(x=bx.ravel(), y=by.ravel(), z=bz.ravel())
    if mode in ("norm", "both"):
        vals["n"] = np.sqrt(bx * bx + by * by + bz * bz).ravel()
    return OracleInterp.from_planes([ax, ax, ax], vals, mode)


def cpu_queries(ora, count, seed):
    rng = np.random.default_rng(seed)
    g = ora.geo
    q = np.empty((count, 3))
    for a in range(3):
        q[:, a] = g.int_min[a] + rng.uniform(0, 1, count) * (g.int_max[a] - g.int_min[a]) * (1 - 1e-12)
    return q


def cpu_warm_rate(ora, q, chunk=100_000):
    """Warm rate: coefficients of the touched cells are filled first (the reference's lazy fill,
    A.py:376-377), then the timed pass is the reference's rQuery arithmetic only."""
    t0 = time.perf_counter()
    for lo in range(0, len(q), chunk):
        ora.query(q[lo:lo + chunk].copy())
    cpu_warm_rate.cold = len(q) / (time.perf_counter() - t0)      # first pass: includes the lazy coefficient fill
    t0 = time.perf_counter()
    for lo in range(0, len(q), chunk):
        ora.query(q[lo:lo + chunk].copy())
    return len(q) / (time.perf_counter() - t0)


_PARENT_ORACLE = None


def _worker(args):
    count, seed, steps, warmup = args
    ora = _PARENT_ORACLE                               # forked: planes shared copy-on-write
    q = cpu_queries(ora, count, seed)
    for lo in range(0, count, 100_000):
        ora.query(q[lo:lo + 100_000].copy())          # fill
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for lo in range(0, count, 100_000):
            ora.query(q[lo:lo + 100_000].copy())
        times.append(time.perf_counter() - t0)
    return times[warmup:]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (numpy port in oracle/, the
    reference itself being pure Python that cannot travel to the GPU box) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    global _PARENT_ORACLE
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    cores = len(os.sched_getaffinity(0))
    procs = max(1, min(cores, 256))
    per = max(20_000, min(200_000, args.cpu_sample // 5))
    n = args.grid
    t0 = time.perf_counter()
    _PARENT_ORACLE = make_cpu_oracle(n, args.mode)
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_worker, [(per, 100 + i, args.steps, args.warmup) for i in range(procs)])
    wall = time.perf_counter() - t0
    per_step = np.max(np.array(res), axis=0)                     # slowest worker per step
    total_t = float(per_step.sum())
    value = procs * per * args.steps / total_t
    sample = f"{procs} forked workers x {per} warm queries/step on a {n}^3 grid (steps={args.steps}); wall {wall:.1f}s"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"tricubic {args.grid}^3 analytic vector field, mode={args.mode}, uniform random in-volume queries",
                   "cpu_grid": n},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": procs, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------
def time_device(torch, lib, obj, mode, d, q, outs, cells, steps, warmup, dist=None):
    """K launches of the query kernel on device-resident queries; returns (elapsed_s, per-launch ms)."""
    from arbinterp_b200 import _lib
    code = {"vector": _lib.MODE_VECTOR, "norm": _lib.MODE_NORM, "both": _lib.MODE_BOTH}[mode]
    stream = torch.cuda.current_stream()
    n, ldq = q.shape

    def launch():
        _lib.check(lib.arb_query(ctypes.byref(obj._cgeom), obj.table.data_ptr(), code, q.data_ptr(), n, ldq,
                                 outs[0].data_ptr() if outs[0] is not None else None,
                                 outs[1].data_ptr() if outs[1] is not None else None,
                                 outs[2].data_ptr() if outs[2] is not None else None,
                                 cells.data_ptr(), None, None, stream.cuda_stream), "arb_query")

    for _ in range(warmup):
        launch()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ev[0].record(stream)
    for i in range(steps):
        launch()
        ev[i + 1].record(stream)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    return ev[0].elapsed_time(ev[steps]) / 1e3, per


def alloc_outputs(torch, mode, d, n, device):
    kw = dict(dtype=torch.float64, device=device)
    comps = torch.empty(n, 3, **kw) if mode in ("vector", "both") else None
    norm = torch.empty(n, 1, **kw) if mode in ("norm", "both") else None
    grad = torch.empty(n, d, **kw) if mode in ("norm", "both") else None
    return comps, norm, grad


def uniform_queries(torch, obj, n, seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    q = torch.rand(n, 3, generator=g, dtype=torch.float64, device=device)
    lo = torch.tensor([obj.xIntMin, obj.yIntMin, obj.zIntMin], dtype=torch.float64, device=device)
    hi = torch.tensor([obj.xIntMax, obj.yIntMax, obj.zIntMax], dtype=torch.float64, device=device)
    return lo + q * (hi - lo) * (1 - 1e-12)


def run_b200(args):
    import torch
    from arbinterp_b200 import tricubic, _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    affinity = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    lib = _lib.load()
    if args.variant is not None:
        lib.arb_set_query_variant(args.variant)

    # ---- construction (untimed): rank 0 makes the field, NCCL broadcast, every rank builds its replica
    n = args.grid
    t0 = time.perf_counter()
    rows = analytic_field_rows(torch, n, device)[1] if rank == 0 else None
    if dist is not None:
        from arbinterp_b200.sharding import broadcast_ingested
        rows = broadcast_ingested(rows, 3, src=0, device=device)     # ingest once, NCCL-broadcast the planes
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    obj = tricubic(rows, "quiet", mode=args.mode)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t1
    del rows
    torch.cuda.empty_cache()

    d = 3
    Q = args.queries
    q = uniform_queries(torch, obj, Q, 1234 + rank, device)
    outs = alloc_outputs(torch, args.mode, d, Q, device)
    cells = torch.empty(Q, dtype=torch.int64, device=device)

    if args.sweep:
        for v in [int(s) for s in args.sweep.split(",")]:
            lib.arb_set_query_variant(v)
            el, per = time_device(torch, lib, obj, args.mode, d, q, outs, cells, args.steps, args.warmup)
            if rank == 0:
                print(f"[sweep] mode={args.mode} variant={v}: {Q * args.steps / el:.4e} q/s  "
                      f"({ALG_BYTES[(d, args.mode)] * Q * args.steps / el / 1e9:.0f} GB/s algorithmic)  "
                      f"min launch {min(per):.3f} ms", file=sys.stderr, flush=True)
        lib.arb_set_query_variant(args.variant if args.variant is not None else 0)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first_sample()
        sampler.samples.clear()                       # keep only what is sampled from the warm-up on
    elapsed, per_launch = time_device(torch, lib, obj, args.mode, d, q, outs, cells, args.steps, args.warmup, dist)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([elapsed], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = world * Q * args.steps / elapsed

    # sanity inside the bench: outputs are finite for in-volume queries
    assert bool(torch.isfinite(outs[1] if outs[1] is not None else outs[0]).all()), "non-finite outputs in bench"

    # ---- the coefficient-build kernel on its own (rebuilds the table in place; HBM-write roofline)
    build_info = None
    if rank == 0 and world == 1:
        import ctypes
        lo, hi = obj._slab
        sub = obj._planes[:, lo:hi + 3].contiguous()
        npts = (ctypes.c_int64 * 4)(*([obj._geo.npts[a] for a in range(d - 1)] + [hi - lo + 3] + [1] * (4 - d)))
        stream = torch.cuda.current_stream()
        times = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            _lib.check(lib.arb_build_coeffs(d, sub.data_ptr(), sub.shape[0], ctypes.byref(npts), obj.table.data_ptr(), 1,
                                            stream.cuda_stream), "build")
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        del sub
        ms = float(np.mean(times[1:]))
        gb = obj.table.numel() * 8 / 1e9
        build_info = {"kernel": "coefficient build (arb_build_coeffs, separable FP64-pipe kernel + NaN sentinel row)",
                      "ms": ms, "table_gb": gb, "achieved_gbs": gb / ms * 1e3, "bound": "hbm (table write)",
                      "frac_of_measured_hbm": gb / ms * 1e3 / measured_peak()[0],
                      "cell_components_per_s": obj.nc * obj.table.shape[1] / ms * 1e3, "launches": 3}

    # ---- other modes (device-resident), optional
    others = {}
    if not args.no_other_modes and rank == 0 and world == 1:
        del obj, outs
        torch.cuda.empty_cache()
        for m in ("vector", "both"):
            if m == args.mode:
                continue
            _, rows = analytic_field_rows(torch, n, device)
            o2 = tricubic(rows, "quiet", mode=m)
            del rows
            torch.cuda.empty_cache()
            Q2 = Q // 2
            outs2 = alloc_outputs(torch, m, d, Q2, device)
            el, per = time_device(torch, lib, o2, m, d, q[:Q2], outs2, cells[:Q2], args.steps, args.warmup)
            rate = Q2 * args.steps / el
            others[m] = {"value": rate, "unit": "queries/s", "alg_bytes_per_query": ALG_BYTES[(d, m)],
                         "achieved_gbs": ALG_BYTES[(d, m)] * rate / 1e9,
                         "frac_of_measured_hbm": ALG_BYTES[(d, m)] * rate / 1e9 / measured_peak()[0],
                         "queries_per_step": Q2, "table_gb": o2.table.numel() * 8 / 1e9}
            del o2, outs2
            torch.cuda.empty_cache()
        # quadcubic (time-dependent field, 256x256 Lekien-Marsden matrix with the A.py:860 quirk), value + 4-gradient
        from arbinterp_b200 import quadcubic
        shape4 = (48, 48, 48, 32)
        axes = [torch.linspace(-1.0, 1.0, m, dtype=torch.float64, device=device) for m in shape4[:3]]
        axes.append(torch.linspace(0.0, 1.0, shape4[3], dtype=torch.float64, device=device))
        T, Z, Y, X = torch.meshgrid(*reversed(axes), indexing="ij")
        u = torch.sin(2 * np.pi * X) * torch.cos(np.pi * Y) * torch.exp(-Z) * torch.cos(2 * T) + X * X * Y + Z * (1 + T)
        t1 = time.perf_counter()
        o4 = quadcubic(torch.stack([t.reshape(-1) for t in (X, Y, Z, T, u)], dim=1), "quiet")
        torch.cuda.synchronize()
        t_ctor4 = time.perf_counter() - t1
        del T, Z, Y, X, u
        Q4 = Q // 4
        g4 = torch.Generator(device=device)
        g4.manual_seed(4321)
        lo4 = torch.tensor(o4._geo.int_min, dtype=torch.float64, device=device)
        hi4 = torch.tensor(o4._geo.int_max, dtype=torch.float64, device=device)
        q4 = lo4 + torch.rand(Q4, 4, generator=g4, dtype=torch.float64, device=device) * (hi4 - lo4) * (1 - 1e-12)
        outs4 = alloc_outputs(torch, "norm", 4, Q4, device)
        el, per = time_device(torch, lib, o4, "norm", 4, q4, outs4, cells[:Q4], args.steps, args.warmup)
        rate = Q4 * args.steps / el
        others["quadcubic_norm"] = {"value": rate, "unit": "queries/s", "alg_bytes_per_query": ALG_BYTES[(4, "norm")],
                                    "achieved_gbs": ALG_BYTES[(4, "norm")] * rate / 1e9,
                                    "frac_of_measured_hbm": ALG_BYTES[(4, "norm")] * rate / 1e9 / measured_peak()[0],
                                    "queries_per_step": Q4, "table_gb": o4.table.numel() * 8 / 1e9,
                                    "workload": "quadcubic scalar field %dx%dx%dx%d (config 4 stand-in), uniform random "
                                                "(x,y,z,t) queries" % shape4, "constructor_s": t_ctor4}
        del o4, outs4, q4
        torch.cuda.empty_cache()
        _, rows = analytic_field_rows(torch, n, device)
        obj = tricubic(rows, "quiet", mode=args.mode)
        del rows

    # ---- end-to-end leg: public API with host (pinned) numpy buffers, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        QE = args.e2e_queries
        qh = torch.empty(QE, 3, dtype=torch.float64, pin_memory=True)
        qh.copy_(q[:QE])
        qnp = qh.numpy()
        for _ in range(max(1, min(args.warmup, 2))):
            res = obj.Query(qnp)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = obj.Query(qnp)
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([te], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            te = float(t.item())
        res = res if isinstance(res, tuple) else (res,)
        d2h = sum(r.nbytes for r in res if r is not None)
        e2e = {"value": world * QE * args.steps / te, "unit": "queries/s", "h2d_bytes_per_step": int(qnp.nbytes),
               "d2h_bytes_per_step": int(d2h), "queries_per_step": QE,
               "api": "tricubic.Query(numpy float64 [N,3] in pinned host memory) -> numpy outputs"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        ncpu = n
        ora = make_cpu_oracle(ncpu, args.mode)
        qc = cpu_queries(ora, args.cpu_sample, 7)
        rate = cpu_warm_rate(ora, qc)
        cpu = {"value": rate, "unit": "queries/s", "cores": 1, "kind": "port",
               "sample": f"{args.cpu_sample} warm queries (second pass, coefficients cached) through the numpy oracle "
                         f"on a {ncpu}^3 grid of the same analytic field, single process",
               "cold_value": cpu_warm_rate.cold,
               "cold_note": "first pass over the same sample incl. the lazy coefficient fill (A.py:376-377); the port "
                            "fills touched cells with one batched dgemm, the reference's per-cell Python loop is slower"}
        # full-size parity on the same sample (oracle as the checker): indices exact, values within 1e-12 scaled
        ref = ora.query(qc.copy())
        ref = ref if isinstance(ref, tuple) else (ref,)
        got = obj.Query(qc.copy())
        got = got if isinstance(got, tuple) else (got,)
        scale = max(float(np.abs(v).max()) for v in ora.values.values())
        worst = 0.0
        for a, b in zip(got, ref):
            sc = scale / np.array(ora.geo.h)[None, :] if b.shape[1] == d and args.mode != "vector" and a is got[-1] else scale
            worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), sc))))
        parity = {"n": int(len(qc)), "max_scaled_err": worst, "tolerance": 1e-12,
                  "indices_equal": bool(np.array_equal(obj.queryInds, ora.query_inds)),
                  "checker": "oracle/arb_oracle.py (numpy restatement, bit-equal to the live reference on tests/golden)"}
        assert parity["indices_equal"] and worst <= 1e-12, f"parity failure in bench: {parity}"

    peak, peak_src = measured_peak()
    alg = ALG_BYTES[(d, args.mode)]
    avg_ms = float(np.mean(per_launch))
    achieved = alg * Q / (avg_ms * 1e-3) / 1e9
    traffic_per_query = ncu_traffic(args.mode)
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"tricubic {n}^3 analytic vector field, mode={args.mode} (value+gradient), "
                               f"uniform random in-volume queries",
                   "queries_per_step_per_gpu": Q, "table_gb": obj.table.numel() * 8 / 1e9,
                   "cache": "inputs larger than L2 (query batch %.1f GB, table %.1f GB vs 126 MB L2)" %
                            (q.numel() * 8 / 1e9, obj.table.numel() * 8 / 1e9),
                   "parallelism": f"replicated table, queries sharded x{world}",
                   "variant": args.variant if args.variant is not None else 0, "build_s": t_build,
                   "cpu_affinity": affinity},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (traffic_per_query * Q) if traffic_per_query else None,
                     "kernel": "query kernel (arb_query)", "alg_bytes_per_query": alg,
                     "avg_launch_ms": avg_ms, "peak_source": peak_src},
        "cpu_baseline": cpu,
        "parity": parity,
        "e2e": e2e,
        "gpu_launches": args.steps * world,
        "clocks": clocks,
    }
    if build_info:
        line["build"] = build_info
    if others:
        line["other_modes"] = others
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
