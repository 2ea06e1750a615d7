#!/usr/bin/env python
"""bench.py -- throughput of the ARBInterp interpolation hot path on B200.

Workload (BASELINE.json north_star / configs[2], value+gradient variant): synthetic analytic
256^3 vector field, tricubic, uniformly random in-volume queries; each rank holds a replica of the
coefficient table and evaluates its own query batch (weak scaling, no collective on the query path).
A "step" is one pass of the query kernel over one batch of Q queries.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode norm|vector|both] [--impl reference]

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline`, `cpu_baseline`, `e2e`,
`clocks`, `gpu_launches`, `other_modes` and (N > 1) `sharded` are described in DESIGN.md "Measurement".
"""
import argparse
import ctypes
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time
_T_PROCESS_START = time.perf_counter()
import warnings

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
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="rows of the in-bench oracle parity check")
    ap.add_argument("--ref-sample", type=int, default=200_000, help="rows of the unmodified-reference cpu_baseline")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other-modes", action="store_true", help="skip the vector/both/quadcubic lines")
    ap.add_argument("--no-sharded", action="store_true", help="(N > 1) skip the slab-sharded config-5 object")
    ap.add_argument("--sharded-particles", type=int, default=1 << 22, help="trajectory queries per rank and step")
    ap.add_argument("--force-port", action="store_true", help="CPU arm: use the numpy port even if the reference is staged")
    return ap.parse_args()


def workload_string(grid, mode):
    """The same string in both arms (the driver compares `config.workload` to decide `same_config`)."""
    return f"tricubic {grid}^3 analytic vector field, mode={mode} (value+gradient), uniform random in-volume queries"


# ----------------------------------------------------------------------------------------
# synthetic analytic fields
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


def analytic_field_rows_numpy(n):
    ax = np.linspace(-1.0, 1.0, n)
    Z, Y, X = np.meshgrid(ax, ax, ax, indexing="ij")
    bx = np.sin(2 * np.pi * X) * np.cos(np.pi * Y) * np.exp(-Z)
    by = X * X * Y + Z
    bz = np.cos(X + Y + Z)
    return np.stack([t.ravel() for t in (X, Y, Z, bx, by, bz)], axis=1)


def field4_rows(torch, shape, device):
    """SURVEY 4 / 8d config 5 seed field: every component non-trivial in all four variables; By carries an explicit
    x*y*z*t monomial, so the A.py:860 rank-16 term is excited."""
    axes = [torch.linspace(-1.0, 1.0, m, dtype=torch.float64, device=device) for m in shape[:3]]
    axes.append(torch.linspace(0.0, 1.0, shape[3], dtype=torch.float64, device=device))
    T, Z, Y, X = [g.reshape(-1) for g in torch.meshgrid(*reversed(axes), indexing="ij")]
    bx = torch.sin(2 * X) * torch.cos(3 * Y) * torch.exp(-Z) * torch.cos(2 * T)
    by = X * X * Y + Z * T + 0.3 * X * Y * Z * T
    bz = torch.cos(X + Y + Z + T)
    return axes, torch.stack([X, Y, Z, T, bx, by, bz], dim=1)


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


def bind_to_gpu_cores(gpu_index, local_rank, local_world):
    """Give this rank its OWN cores: the CPUs NVML reports as local to its GPU (the NUMA node its PCIe root hangs
    off), cut into equal shares among the ranks of the box that report the same set (on a single-node VM every GPU
    reports all cores -- eight ranks bound to the same 32 cores is no binding at all).  The staging threads of the
    host-buffer path and the pinned allocations then stay on this rank's cores."""
    try:
        import pynvml
        pynvml.nvmlInit()
        allowed_now = sorted(os.sched_getaffinity(0))

        def local_cpus(idx):
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
            return tuple(sorted(cpus & set(allowed_now))) or tuple(allowed_now)

        mine = local_cpus(gpu_index)
        sharers = [r for r in range(local_world) if local_cpus(r) == mine]      # local rank r drives GPU r
        k, n = sharers.index(local_rank), len(sharers)
        share = list(mine[len(mine) * k // n: len(mine) * (k + 1) // n]) or list(mine)
        os.sched_setaffinity(0, share)
        return f"{len(share)} cores [{share[0]}..{share[-1]}] of {len(mine)} GPU-local, {n} ranks share them"
    except Exception as e:                                  # noqa: BLE001 -- binding is best effort
        return f"unbound ({type(e).__name__})"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key):
    """DRAM bytes per query from the committed ncu --set full captures (profiles/roofline_traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------
# CPU arm.  Preferred: the UNMODIFIED reference module, staged by oracle/stage_reference.py under oracle/_ref
# (or found in baseline/_ref / $ARB_REFERENCE); fallback: the numpy port in oracle/arb_oracle.py.
# ----------------------------------------------------------------------------------------
def load_reference_module():
    """The reference's own ARBInterp.py as a module object, or None.  Loaded by file path under a private name:
    the repository root holds drop-in packages of the same names (ARBInterp/, ARBTools/) that must not shadow it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        from stage_reference import module_path
    finally:
        sys.path.pop(0)
    roots = [os.environ.get("ARB_REFERENCE"), os.path.join(ROOT, "oracle", "_ref"), os.path.join(ROOT, "baseline", "_ref")]
    for root in roots:
        path = root and module_path(root)
        if path:
            spec = importlib.util.spec_from_file_location("arb_reference_unmodified", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod, path
    return None, None


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


def uniform_in_volume(lo, hi, count, seed):
    rng = np.random.default_rng(seed)
    q = np.empty((count, len(lo)))
    for a in range(len(lo)):
        q[:, a] = lo[a] + rng.uniform(0, 1, count) * (hi[a] - lo[a]) * (1 - 1e-12)
    return q


def cpu_queries(ora, count, seed):
    return uniform_in_volume(ora.geo.int_min, ora.geo.int_max, count, seed)


class CpuArm:
    """One CPU implementation of the path behind a common face: `query(q)` (range query, NaN-masks q in place)."""

    def __init__(self, n, mode, force_port=False):
        self.n, self.mode = n, mode
        mod, path = (None, None) if force_port else load_reference_module()
        if mod is not None:
            self.kind = "reference"
            self.what = "unmodified reference module " + os.path.relpath(path, ROOT)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                self.obj = mod.tricubic(analytic_field_rows_numpy(n), "quiet", mode=mode)
            self.lo = [self.obj.xIntMin, self.obj.yIntMin, self.obj.zIntMin]
            self.hi = [self.obj.xIntMax, self.obj.yIntMax, self.obj.zIntMax]
        else:
            self.kind = "port"
            self.what = "numpy port oracle/arb_oracle.py (bit-equal to the reference on tests/golden)"
            self.obj = make_cpu_oracle(n, mode)
            self.lo, self.hi = self.obj.geo.int_min, self.obj.geo.int_max

    def query(self, q):
        if self.kind == "reference":
            with warnings.catch_warnings():                   # NaN -> int cast warning at A.py:370
                warnings.simplefilter("ignore")
                return self.obj.Query(q)
        return self.obj.query(q)

    def passes(self, q, chunk):
        t0 = time.perf_counter()
        for lo in range(0, len(q), chunk):
            self.query(q[lo:lo + chunk].copy())
        return len(q) / (time.perf_counter() - t0)


_PARENT_ARM = None


def _worker(args):
    idx, per, seed, steps, warmup, chunk = args
    arm = _PARENT_ARM                                  # forked: field and filled coefficients shared copy-on-write
    q = uniform_in_volume(arm.lo, arm.hi, per, seed)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for lo in range(0, per, chunk):
            arm.query(q[lo:lo + chunk].copy())
        times.append(time.perf_counter() - t0)
    return times[warmup:]


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host cores -- the unmodified module
    when it is staged (kind "reference"), else the numpy port (kind "port").  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    global _PARENT_ARM
    os.environ["OMP_NUM_THREADS"] = os.environ["OPENBLAS_NUM_THREADS"] = "1"
    cores = len(os.sched_getaffinity(0))
    procs = max(1, min(cores, 256))
    n = args.grid
    t0 = time.perf_counter()
    arm = CpuArm(n, args.mode, args.force_port)
    t_ctor = time.perf_counter() - t0
    if arm.kind == "reference":
        # the reference fills coefficients lazily, one Python call per cell (~2e4 cells/s, A.py:376-377): the parent
        # fills the cells of every worker's sample once (bounded to ~20 s), the workers then time warm passes
        per = max(5_000, min(50_000, 400_000 // procs))
        chunk = 50_000
    else:
        per = max(20_000, min(200_000, args.cpu_sample // 5))
        chunk = 100_000
    seeds = [100 + i for i in range(procs)]
    t1 = time.perf_counter()
    for sd in seeds:                                    # lazy fill in the parent (cold pass, untimed by the metric)
        arm.query(uniform_in_volume(arm.lo, arm.hi, per, sd))
    cold = procs * per / (time.perf_counter() - t1)
    _PARENT_ARM = arm
    with mp.get_context("fork").Pool(procs) as pool:
        res = pool.map(_worker, [(i, per, seeds[i], args.steps, args.warmup, chunk) for i in range(procs)])
    wall = time.perf_counter() - t0
    per_step = np.max(np.array(res), axis=0)                     # slowest worker per step
    total_t = float(per_step.sum())
    value = procs * per * args.steps / total_t
    sample = (f"{arm.what}: {procs} forked workers x {per} warm queries/step on a {n}^3 grid (steps={args.steps}); "
              f"constructor {t_ctor:.1f}s, single-process cold pass incl. lazy coefficient fill {cold:.3g} q/s; wall {wall:.1f}s")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(args.grid, args.mode), "cpu_grid": n},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": procs, "kind": arm.kind, "sample": sample,
                         "cold_value": cold},
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


def time_public_device(torch, obj, q, steps, warmup):
    """K calls of obj.Query on a device-resident batch (the public call with a CUDA tensor: kernel + output allocation
    from torch's caching allocator); queries/s from CUDA events around the K calls."""
    # these legs follow seconds of CPU-side checking with the GPU idle: warm up for >= 100 ms so that the clocks are
    # back up before the timed calls (a 2 ms kernel timed right after an idle period ran at half its rate)
    # (the warm-up keeps the previous call's outputs alive while the next call allocates, exactly like the timed loop:
    # otherwise the first timed call needs a second set of output blocks that the caching allocator does not hold yet,
    # and one cudaMalloc of ~0.5 GB inside ten 2 ms kernels read as "0.3-0.5 of the usual rate" in earlier sessions)
    t0 = time.perf_counter()
    done = 0
    res = None
    while done < max(warmup, 2) or time.perf_counter() - t0 < 0.1:
        res = obj.Query(q)
        torch.cuda.synchronize()
        done += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        res = obj.Query(q)
    e1.record()
    torch.cuda.synchronize()
    res = res if isinstance(res, tuple) else (res,)
    return q.shape[0] * steps / (e0.elapsed_time(e1) / 1e3), all(bool(torch.isfinite(r).all()) for r in res)


def alloc_outputs(torch, mode, d, n, device):
    kw = dict(dtype=torch.float64, device=device)
    comps = torch.empty(n, 3, **kw) if mode in ("vector", "both") else None
    norm = torch.empty(n, 1, **kw) if mode in ("norm", "both") else None
    grad = torch.empty(n, d, **kw) if mode in ("norm", "both") else None
    return comps, norm, grad


def uniform_queries(torch, obj, n, seed, device):
    d = obj._d
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    q = torch.rand(n, d, generator=g, dtype=torch.float64, device=device)
    lo = torch.tensor(obj._geo.int_min, dtype=torch.float64, device=device)
    hi = torch.tensor(obj._geo.int_max, dtype=torch.float64, device=device)
    return lo + q * (hi - lo) * (1 - 1e-12)


def scaled_error(got, ref, d, h, scale):
    """max |got - ref| / max(|ref|, S) over every output (S/h for gradient columns); NaN masks must agree."""
    got = got if isinstance(got, tuple) else (got,)
    ref = ref if isinstance(ref, tuple) else (ref,)
    worst = 0.0
    for a, b in zip(got, ref):
        a, b = np.asarray(a), np.asarray(b)
        if a.shape != b.shape or not np.array_equal(np.isnan(a), np.isnan(b)):
            return float("inf")
        is_grad = b.shape[1] == d and a is got[-1] and len(got) > 1
        sc = scale / np.asarray(h, dtype=np.float64)[None, :] if is_grad else scale
        sc = np.broadcast_to(sc, b.shape)
        ok = ~np.isnan(b)
        if ok.any():
            worst = max(worst, float(np.max(np.abs(a[ok] - b[ok]) / np.maximum(np.abs(b[ok]), sc[ok]))))
    return worst


def oracle_sample_check(obj, axes, values, mode, d, nrows, seed, scalar=False, chunk=25_000):
    """In-run guard: `nrows` uniformly random in-volume queries through the GPU object and through the numpy oracle
    (compact coefficient store) built on the SAME planes; indices exact, values within 1e-12 scaled."""
    from oracle.arb_oracle import OracleInterp
    ora = OracleInterp.from_planes(axes, values, mode, scalar_input=scalar)
    q = cpu_queries(ora, nrows, seed)
    scale = max(float(np.abs(v).max()) for v in values.values())
    worst, idx_ok, finite = 0.0, True, True
    for lo in range(0, nrows, chunk):
        qc = q[lo:lo + chunk]
        ref = ora.query(qc.copy())
        got = obj.Query(qc.copy())
        worst = max(worst, scaled_error(got, ref, d, ora.geo.h, scale))
        idx_ok &= bool(np.array_equal(obj.queryInds, ora.query_inds))
        finite &= all(bool(np.isfinite(np.asarray(g)).all()) for g in (got if isinstance(got, tuple) else (got,)))
    return {"n": int(nrows), "max_scaled_err": worst, "tolerance": 1e-12, "indices_equal": idx_ok, "finite": finite,
            "ok": bool(idx_ok and finite and worst <= 1e-12)}


def link_ceiling(torch, dist, device, world, mb=256, reps=6):
    """What the host<->device links of this box can move when all `world` ranks copy at once: every rank runs
    `reps` pinned H2D copies and `reps` pinned D2H copies of `mb` MiB concurrently on two streams (cudaMemcpyAsync),
    between barriers; GB/s = bytes of all ranks / slowest rank's time.  Also each direction alone."""
    n = mb << 20
    hin = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    hout = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    din = torch.empty(n, dtype=torch.uint8, device=device)
    dout = torch.empty(n, dtype=torch.uint8, device=device)
    s1, s2 = torch.cuda.Stream(device), torch.cuda.Stream(device)
    out = {}
    for name, do_in, do_out in (("h2d", True, False), ("d2h", False, True), ("bidir", True, True)):
        for timed in (False, True):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps if timed else 1):
                if do_in:
                    with torch.cuda.stream(s1):
                        din.copy_(hin, non_blocking=True)
                if do_out:
                    with torch.cuda.stream(s2):
                        hout.copy_(dout, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        out[name + "_gbs"] = world * reps * n * (int(do_in) + int(do_out)) / dt / 1e9
    return out


def time_e2e(torch, dist, device, obj, qnp, steps, warmup, world):
    """`steps` calls of the public API on a host batch; wall clock between barriers + syncs, max over ranks."""
    for _ in range(max(1, min(warmup, 2))):
        res = obj.Query(qnp)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        res = obj.Query(qnp)
    torch.cuda.synchronize()
    te = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([te], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = float(t.item())
    res = res if isinstance(res, tuple) else (res,)
    d2h = sum(r.nbytes for r in res if r is not None)
    return world * len(qnp) * steps / te, int(qnp.nbytes), int(d2h)


# ----------------------------------------------------------------------------------------
# N > 1: BASELINE configs[4] -- quadcubic 'both', coefficient table t-slab-sharded with halos, trajectory-like
# queries routed to the owning rank -- with in-run parity
# ----------------------------------------------------------------------------------------
def run_sharded(torch, dist, args, rank, world, device):
    from arbinterp_b200 import quadcubic
    from arbinterp_b200.sharding import SlabShardedInterp
    from oracle.arb_oracle import OracleInterp
    out = {}
    rng = np.random.default_rng(5150)

    # ---- (i) small 4-D VECTOR field with an xyzt monomial in every component: routed == unsharded, bit for bit
    small = (14, 12, 11, 3 * world + 5)
    ax = [np.linspace(-1, 1, small[0]), np.linspace(-.5, .5, small[1]), np.linspace(0, 2, small[2]),
          np.linspace(0, 1, small[3])]
    T, Z, Y, X = [a.ravel() for a in np.meshgrid(ax[3], ax[2], ax[1], ax[0], indexing="ij")]
    xyzt = X * Y * Z * T
    fs = np.stack([X, Y, Z, T, np.sin(2 * X) * np.cos(3 * Y) * np.exp(-Z) * np.cos(2 * T) + 0.7 * xyzt,
                   X * X * Y + Z * T - 0.4 * xyzt, np.cos(X + Y + Z + T) + 1.1 * xyzt], axis=1)
    whole = quadcubic(fs.copy(), "quiet", mode="both", device=device)
    sh = SlabShardedInterp(quadcubic, fs if rank == 0 else None, "quiet", mode="both", device=device)
    lo_s = np.array([a[1] for a in ax]); hi_s = np.array([a[-2] for a in ax])
    qs = lo_s + rng.uniform(-0.03, 1.03, (20000, 4)) * (hi_s - lo_s)          # ~11 % of the rows leave the volume
    qs[::97, 1] = np.nan
    mine = np.ascontiguousarray(qs[rank::world])
    q_ref, q_got = mine.copy(), mine.copy()
    ref = whole.Query(q_ref)
    got = sh.Query(q_got)
    same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(got, ref))
    same &= bool(np.array_equal(sh.queryInds, whole.queryInds)) and bool(np.array_equal(q_got, q_ref, equal_nan=True))
    flag = torch.tensor([int(same)], dtype=torch.int64, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["routed_equals_unsharded"] = bool(flag.item())
    out["small_check"] = {"grid": list(small), "mode": "both", "rows_per_rank": int(len(mine)),
                          "what": "outputs, queryInds and the in-place NaN rows of SlabShardedInterp.Query(numpy) are "
                                  "bit-identical to the unsharded quadcubic on the same rank; every component has an "
                                  "x*y*z*t term (A.py:860 rank-16 term excited)"}
    whole.release()
    sh.local.release()
    del whole, sh
    torch.cuda.empty_cache()

    # ---- config 5 scaled to the world: 96^3 x nt 'both', 8 cell layers per rank (61 layers = the full 96^3 x 64 at 8)
    layers = min(61, 8 * world)
    shape = (96, 96, 96, layers + 3)
    rows = None
    crop = None
    if rank == 0:
        axes_t, rows = field4_rows(torch, shape, device)
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    obj = SlabShardedInterp(quadcubic, rows, "quiet", mode="both", device=device)
    torch.cuda.synchronize()
    dist.barrier()
    t_build = time.perf_counter() - t0
    loc = obj.local
    g = loc._geo
    if rank == 0:
        # cropped sub-field straddling the rank 0 / rank 1 slab boundary (coefficients are local to the 4^4 neighbourhood)
        hi0 = obj.slabs[0][1]
        win = [(40, 52), (38, 50), (44, 56), (hi0 - 3, hi0 + 5)]               # grid-point windows x, y, z, t
        r5 = rows.reshape(shape[3], shape[2], shape[1], shape[0], 7)
        crop = r5[win[3][0]:win[3][1], win[2][0]:win[2][1], win[1][0]:win[1][1], win[0][0]:win[0][1]].reshape(-1, 7).cpu().numpy()
        full_scale = float(rows[:, 4:].abs().max())
    del rows
    torch.cuda.empty_cache()
    table_gb = torch.tensor([loc.table.numel() * 8 / 1e9], dtype=torch.float64, device=device)
    tb_max = table_gb.clone()
    dist.all_reduce(table_gb)
    dist.all_reduce(tb_max, op=dist.ReduceOp.MAX)

    # ---- (ii) global cell indices against the oracle's locate() on 10^4 rows of the FULL geometry (rank 0's rows)
    axes_np = [a.cpu().numpy() for a in g.axes]
    lo = np.array(g.int_min); hi = np.array(g.int_max)
    if rank == 0:
        qi = lo + rng.uniform(-0.02, 1.02, (10000, 4)) * (hi - lo)
        qi[::113, 2] = np.nan
        qi[5::211, 0] = np.inf
    else:
        qi = np.empty((0, 4))
    qi_in = qi.copy()
    obj.Query(qi_in)
    idx_ok = True
    if rank == 0:
        loc_ora = OracleInterp.locator(axes_np)
        q_or = qi.copy()
        with np.errstate(invalid="ignore"):
            want, _ = loc_ora.locate(q_or)
        idx_ok = bool(np.array_equal(obj.queryInds, want)) and bool(np.array_equal(qi_in, q_or, equal_nan=True))
        out["index_check"] = {"n": 10000, "nan_rows": int((want == loc_ora.geo.nc).sum()), "nc": int(loc_ora.geo.nc)}
    # ---- (iii) values against the oracle on the cropped sub-field
    if rank == 0:
        ora = OracleInterp(crop, 4, mode="both", dense=False)
        qc = uniform_in_volume(ora.geo.int_min, ora.geo.int_max, 10000, 99)
    else:
        qc = np.empty((0, 4))
    got = obj.Query(qc.copy())
    crop_err = 0.0
    if rank == 0:
        ref = ora.query(qc.copy())
        crop_err = scaled_error(got, ref, 4, ora.geo.h, full_scale)
        out["crop_check"] = {"n": 10000, "window_points": [w[1] - w[0] for w in win], "max_scaled_err": crop_err,
                             "tolerance": 1e-12, "t_layers_spanned": [int(hi0 - 3), int(hi0 + 2)],
                             "owners_hit": sorted(set(int(v) for v in np.searchsorted(
                                 [s[1] for s in obj.slabs], np.unique(obj.queryInds // (93 ** 3)), side="right")))}
    flags = torch.tensor([int(idx_ok), int(crop_err <= 1e-12)], dtype=torch.int64, device=device)
    dist.broadcast(flags, src=0)
    out["global_indices_equal_oracle"] = bool(flags[0].item())
    out["cropped_values_within_1e-12"] = bool(flags[1].item())

    # ---- the same field as a REPLICATED node (Hermite) table (csrc/arb_nodes.cuh): 16x smaller than the cell table,
    # so every rank holds all of it and answers its own rows with no exchange at all
    from arbinterp_b200.sharding import ReplicatedInterp
    rows = field4_rows(torch, shape, device)[1] if rank == 0 else None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rep = ReplicatedInterp(quadcubic, rows, "quiet", mode="both", table="nodes", device=device)
    torch.cuda.synchronize()
    dist.barrier()
    t_build_nodes = time.perf_counter() - t0
    del rows
    torch.cuda.empty_cache()

    # ---- throughput: trajectory-like queries (P particles per rank, smooth random walk reflected into the volume,
    # time advancing and wrapping), routed to the slab owners and back
    n = args.sharded_particles
    lo_t = torch.tensor(g.int_min, dtype=torch.float64, device=device)
    hi_t = torch.tensor(g.int_max, dtype=torch.float64, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(777 + rank)
    pos = lo_t + torch.rand(n, 4, generator=gen, dtype=torch.float64, device=device) * (hi_t - lo_t) * (1 - 1e-9)
    vel = (torch.rand(n, 4, generator=gen, dtype=torch.float64, device=device) - 0.5) * (hi_t - lo_t) * 0.02
    vel[:, 3] = (hi_t[3] - lo_t[3]) * 0.01
    only_space = torch.tensor([1, 1, 1, 0], dtype=torch.float64, device=device)
    span = (hi_t - lo_t) * (1 - 1e-9)
    timing, timing_plain, timing_ret = {}, {}, {}
    step_ms, plain_ms, ret_ms = [], [], []
    node_ms = []
    finite, same_fused = True, True
    for step in range(args.warmup + args.steps):
        pos = pos + vel + 0.002 * (hi_t - lo_t) * torch.randn(n, 4, generator=gen, dtype=torch.float64, device=device) * only_space
        rel = torch.remainder(pos - lo_t, 2 * span)
        rel = torch.where(rel > span, 2 * span - rel, rel)
        rel[:, 3] = torch.remainder(pos[:, 3] - lo_t[3], span[3])
        q = (lo_t + rel).contiguous()
        timed = step >= args.warmup
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        res = obj.Query(q, timing=timing if timed else None)
        e1.record()
        torch.cuda.synchronize()
        cells_plain = obj._last_cells
        # the same rows with the return leg as an NCCL all-to-all + re-ordering pass (fused peer-store return switched off)
        was, obj.fused = obj.fused, False
        dist.barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        res_plain = obj.Query(q, timing=timing_plain if timed else None)
        e5.record()
        torch.cuda.synchronize()
        same_fused &= all(bool(torch.equal(a.view(torch.int64), b.view(torch.int64))) for a, b in zip(res, res_plain)) and \
            bool(torch.equal(cells_plain, obj._last_cells))
        # ... and with only the return leg fused (round 2's first form: rows travel by NCCL all-to-all after a sort)
        obj.fused = True
        dist.barrier()
        e6, e7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e6.record()
        res_ret = obj.Query(q, timing=timing_ret if timed else None)
        e7.record()
        torch.cuda.synchronize()
        obj.fused = was
        same_fused &= all(bool(torch.equal(a.view(torch.int64), b.view(torch.int64))) for a, b in zip(res, res_ret)) and \
            bool(torch.equal(cells_plain, obj._last_cells))
        if timed:
            plain_ms.append(e4.elapsed_time(e5))
            ret_ms.append(e6.elapsed_time(e7))
        dist.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        res_n = rep.Query(q)
        e3.record()
        torch.cuda.synchronize()
        if timed:
            step_ms.append(e0.elapsed_time(e1))
            node_ms.append(e2.elapsed_time(e3))
            finite &= all(bool(torch.isfinite(r).all()) for r in res)
    # the two answers to the last step's rows agree (cell table, routed / node table, local)
    m = min(n, 200_000)
    node_err = scaled_error(tuple(r[:m].cpu().numpy() for r in res_n), tuple(r[:m].cpu().numpy() for r in res), 4,
                            g.h, float(torch.stack([r.abs().max() for r in res[:2]]).max()))
    same_cells = bool(torch.equal(rep.local._last_cells, obj._last_cells))
    tn = torch.tensor([sum(node_ms), node_err, float(not same_cells), sum(plain_ms), float(not same_fused), sum(ret_ms)],
                      dtype=torch.float64, device=device)
    dist.all_reduce(tn, op=dist.ReduceOp.MAX)
    ret_s = float(tn[5]) / 1e3
    phr = torch.tensor([timing_ret.get(k + "_ms", 0.0) for k in
                        ("owner", "sort", "counts", "permute", "alltoall", "kernel", "alltoall_back", "scatter", "unpack")],
                       dtype=torch.float64, device=device)
    dist.all_reduce(phr, op=dist.ReduceOp.MAX)
    phr = (phr / args.steps).tolist()
    both_legs = getattr(obj, "_inbox", None) is not None
    node_s, node_err, cells_differ = float(tn[0]) / 1e3, float(tn[1]), bool(tn[2] > 0)
    plain_s, fused_differs = float(tn[3]) / 1e3, bool(tn[4] > 0)
    php = torch.tensor([timing_plain.get(k + "_ms", 0.0) for k in
                        ("owner", "sort", "counts", "permute", "alltoall", "kernel", "alltoall_back", "scatter", "unpack")],
                       dtype=torch.float64, device=device)
    dist.all_reduce(php, op=dist.ReduceOp.MAX)
    php = (php / args.steps).tolist()
    peer_how = obj._peer.how if getattr(obj, "_peer", None) is not None else None
    node_gb = rep.local.nodes.numel() * 8 / 1e9
    t = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_s = float(t.item()) / 1e3
    ph = torch.tensor([timing.get(k + "_ms", 0.0) for k in
                       ("owner", "sort", "counts", "permute", "alltoall", "kernel", "alltoall_back", "scatter", "unpack")],
                      dtype=torch.float64, device=device)
    dist.all_reduce(ph, op=dist.ReduceOp.MAX)
    ph = (ph / args.steps).tolist()
    fl = torch.tensor([int(finite)], dtype=torch.int64, device=device)
    dist.all_reduce(fl, op=dist.ReduceOp.MIN)
    rate = world * n * args.steps / total_s
    out.update({
        "workload": f"quadcubic {shape[0]}x{shape[1]}x{shape[2]}x{shape[3]} analytic vector field (xyzt term), mode=both, "
                    f"coefficient table t-slab-sharded over {world} ranks with halo planes, trajectory-like queries "
                    f"routed to the owning rank and back (BASELINE configs[4]; {layers} of its 61 t layers)",
        "value": rate, "unit": "queries/s", "queries_per_step_per_gpu": n, "steps": args.steps,
        "ms_per_step": 1e3 * total_s / args.steps, "slabs": [list(s) for s in obj.slabs],
        "table_gb_total": float(table_gb.item()), "table_gb_max_per_rank": float(tb_max.item()),
        "construction_s": t_build, "outputs_finite": bool(fl.item()),
        "alg_bytes_per_query": ALG_BYTES[(4, "both")],
        "phase_ms_per_step_max_over_ranks": {
            "kernel": ph[5], "all_to_all (counts + rows + results)": ph[2] + ph[4] + ph[6],
            "sort + permute + scatter": ph[1] + ph[3] + ph[7], "owner + unpack": ph[0] + ph[8],
            "each": dict(zip(("owner", "sort", "counts", "permute", "alltoall", "kernel", "alltoall_back", "scatter", "unpack"), ph))},
        "kernel_frac_of_measured_hbm": (ALG_BYTES[(4, "both")] * n / (ph[5] * 1e-3) / 1e9 / measured_peak()[0]) if ph[5] > 0 else None,
        "forward_leg": ("fused into a kernel (the library default): arb_route_rows stores every row, with its home row number, "
                        "into the inbox of the rank that owns its slab over NVLink -- no sort, no counts exchange, no all-to-all, "
                        "no host round trip; the owner's kernel (arb_query_inbox) reads the per-sender counts on the device; "
                        "in the phase split 'owner' is that kernel and 'alltoall' the barrier after it") if both_legs else
                       "NCCL all-to-all of the rows after a 16-bit sort by owner (route_rows)",
        "return_leg": ("fused into the query kernel (the library default): every row's outputs are stored into its home "
                       f"rank's result buffer over NVLink (peer mapping: {peer_how})") if peer_how else
                      ("NCCL all-to-all of the result rows + a re-ordering pass (peer mapping of the result buffers "
                       "unavailable: " + str(getattr(obj, "_peer_error", None)) + ")"),
        "all_to_all_return": {
            "what": "the same rows with the return leg as an NCCL all-to-all of the result rows + a re-ordering pass (fused=False)",
            "value": world * n * args.steps / plain_s, "unit": "queries/s", "ms_per_step": 1e3 * plain_s / args.steps,
            "fused_speedup": plain_s / total_s, "bit_identical_to_fused": not fused_differs,
            "phase_ms_per_step_max_over_ranks": dict(zip(("owner", "sort", "counts", "permute", "alltoall", "kernel",
                                                          "alltoall_back", "scatter", "unpack"), php))},
        "return_leg_only_fused": {
            "what": "the same rows with only the return leg fused (fused=True): rows sorted by owner, exchanged with two NCCL "
                    "all-to-alls, results stored into the home rank's buffer by the query kernel",
            "value": world * n * args.steps / ret_s, "unit": "queries/s", "ms_per_step": 1e3 * ret_s / args.steps,
            "both_legs_speedup": ret_s / total_s,
            "phase_ms_per_step_max_over_ranks": dict(zip(("owner", "sort", "counts", "permute", "alltoall", "kernel",
                                                          "alltoall_back", "scatter", "unpack"), phr))},
        "replicated_node_table": {
            "what": "the same field as a node (Hermite) table replicated on every rank (quadcubic(table='nodes') via "
                    "sharding.ReplicatedInterp): each rank answers its own rows, no exchange; same queries as above",
            "value": world * n * args.steps / node_s, "unit": "queries/s", "ms_per_step": 1e3 * node_s / args.steps,
            "node_table_gb_per_rank": node_gb, "construction_s": t_build_nodes,
            "speedup_over_slab_sharded": total_s / node_s,
            "frac_of_measured_hbm": ALG_BYTES[(4, "both")] * n * args.steps / node_s / 1e9 / measured_peak()[0],
            "max_scaled_diff_vs_slab_sharded": node_err, "tolerance": 1e-12, "cell_indices_equal": not cells_differ,
            "ok": bool(node_err <= 1e-12 and not cells_differ)},
    })
    obj.local.release()
    rep.local.release()
    del obj, loc, rep
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import gc
    import torch
    from arbinterp_b200 import tricubic, _lib

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    dist = None
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    affinity = bind_to_gpu_cores(local, local, local_world) if (world > 1 and not os.environ.get("ARB_NO_BIND")) else None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=device)
        dist = dist_mod
    lib = _lib.load()
    if args.variant is not None:
        lib.arb_set_query_variant(args.variant)

    # ---- construction (untimed): rank 0 makes the field, NCCL broadcast, every rank builds its replica
    n = args.grid
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

    # ---- end-to-end leg: public API with HOST numpy buffers, copies inside the timed region.  Headline: ordinary
    # (pageable) numpy input, what a user of the reference has; `pinned`: the same call on page-locked input.
    e2e = None
    if not args.no_e2e:
        QE = args.e2e_queries
        q_page = q[:QE].cpu().numpy()                                     # ordinary numpy array
        qh = torch.empty(QE, 3, dtype=torch.float64, pin_memory=True)
        qh.copy_(q[:QE])
        v_page, h2d, d2h = time_e2e(torch, dist, device, obj, q_page, args.steps, args.warmup, world)
        v_pin, _, _ = time_e2e(torch, dist, device, obj, qh.numpy(), args.steps, args.warmup, world)
        link = link_ceiling(torch, dist, device, world)
        bytes_per_q = (h2d + d2h) / QE
        e2e = {"value": v_page, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "queries_per_step": QE,
               "api": "tricubic.Query(ordinary pageable numpy float64 [N,3]) -> numpy outputs",
               "pinned": {"value": v_pin, "api": "the same call with the caller's array page-locked"},
               "link_ceiling_gbs": link["bidir_gbs"], "link_h2d_only_gbs": link["h2d_gbs"],
               "link_d2h_only_gbs": link["d2h_gbs"],
               "link_frac": v_pin * bytes_per_q / 1e9 / link["bidir_gbs"],
               "link_frac_pageable": v_page * bytes_per_q / 1e9 / link["bidir_gbs"],
               "link_note": f"ceiling = {world} rank(s) x concurrent pinned cudaMemcpyAsync H2D + D2H, 256 MiB each way, "
                            "aggregate GB/s over the slowest rank (tools/link_ceiling.py is the stand-alone form); "
                            "link_frac = e2e bytes moved per second (pinned input) / that ceiling"}
        del qh, q_page

    # ---- N > 1: single-process multi-GPU drop-in (rank 0 drives every GPU of the job; the other ranks wait on a
    # CPU-side (gloo) barrier -- an NCCL barrier would keep a spinning kernel of THEIR context on the GPUs rank 0 is
    # about to use, and two contexts time-slice a GPU)
    if world > 1 and not args.no_e2e:
        torch.cuda.synchronize()
        cpu_group = dist.new_group(backend="gloo")
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                # the other ranks are idle now: this process may use every core of the box (threads inherit the mask)
                all_cores = set(range(os.cpu_count()))
                for tid in os.listdir("/proc/self/task"):
                    try:
                        os.sched_setaffinity(int(tid), all_cores)
                    except OSError:
                        pass
                from arbinterp_b200.ingest import IngestedField
                QS = args.e2e_queries
                planes = analytic_planes(torch, n, device)[2]
                from arbinterp_b200.ingest import geometry_from_axes
                ax = torch.linspace(-1.0, 1.0, n, dtype=torch.float64, device=device)
                multi = tricubic(IngestedField(planes=torch.stack(planes), geo=geometry_from_axes([ax, ax, ax])),
                                 "quiet", mode=args.mode, devices=list(range(world)))
                del planes
                qm = uniform_queries(torch, obj, QS, 4242, device).cpu().numpy()
                one = obj.Query(qm.copy())
                many = multi.Query(qm.copy())
                one = one if isinstance(one, tuple) else (one,)
                many = many if isinstance(many, tuple) else (many,)
                same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(one, many)) and \
                    bool(np.array_equal(obj.queryInds, multi.queryInds))
                v_multi, _, _ = time_e2e(torch, None, device, multi, qm, args.steps, args.warmup, 1)
                qpin = torch.empty(QS, 3, dtype=torch.float64, pin_memory=True)
                qpin.copy_(torch.from_numpy(qm))
                v_multi_pin, _, _ = time_e2e(torch, None, device, multi, qpin.numpy(), args.steps, args.warmup, 1)
                e2e["single_process_multi_gpu"] = {
                    "value": v_multi, "unit": "queries/s", "devices": world, "queries_per_step": QS,
                    "pinned": {"value": v_multi_pin, "api": "the same call with the caller's array page-locked"},
                    "bit_identical_to_one_gpu": bool(same),
                    "api": f"tricubic(field, devices=[0..{world - 1}]).Query(pageable numpy) in ONE process, other ranks idle"}
                multi.release()
                del multi
            except Exception as e:                                   # noqa: BLE001 -- report, do not lose the line
                e2e["single_process_multi_gpu"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        dist.barrier(group=cpu_group)

    # ---- other modes and the 4-D path (device-resident), each guarded by an oracle sample (N = 1)
    others = {}
    if not args.no_other_modes and rank == 0 and world == 1:
        obj.release()
        del obj, outs
        torch.cuda.empty_cache()
        ax3, _, raw = analytic_planes(torch, n, device)
        raw_np = [p.cpu().numpy().ravel() for p in raw]
        ax_np = ax3.cpu().numpy()
        del raw
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
            vals = dict(x=raw_np[0], y=raw_np[1], z=raw_np[2])
            if m == "both":
                vals["n"] = np.linalg.norm(np.stack(raw_np, axis=1), axis=1)          # A.py:74
            finite = all(bool(torch.isfinite(t_).all()) for t_ in outs2 if t_ is not None)
            guard = oracle_sample_check(o2, [ax_np] * 3, vals, m, 3, 100_000, 11)
            guard["timed_outputs_finite"] = finite
            others[m] = {"value": rate, "unit": "queries/s", "alg_bytes_per_query": ALG_BYTES[(d, m)],
                         "achieved_gbs": ALG_BYTES[(d, m)] * rate / 1e9,
                         "frac_of_measured_hbm": ALG_BYTES[(d, m)] * rate / 1e9 / measured_peak()[0],
                         "traffic_bytes_per_query_ncu": ncu_traffic(m),
                         "queries_per_step": Q2, "table_gb": o2.table.numel() * 8 / 1e9, "parity": guard}
            assert guard["ok"] and finite, f"parity failure in bench (mode {m}): {guard}"
            o2.release()
            del o2, outs2
            gc.collect()
            torch.cuda.empty_cache()
            # the memory-light forms of the same mode: node (Hermite) table and table-free (no table at all)
            for form, kw in (("node_table", {"table": "nodes"}), ("table_free", {"table": False})):
                _, rows = analytic_field_rows(torch, n, device)
                o3 = tricubic(rows, "quiet", mode=m, **kw)
                del rows
                torch.cuda.empty_cache()
                r3, fin3 = time_public_device(torch, o3, q[:Q2 // 2], max(3, args.steps // 2), args.warmup)
                g3 = oracle_sample_check(o3, [ax_np] * 3, vals, m, 3, 50_000, 13)
                g3["timed_outputs_finite"] = fin3
                store = o3._nodes if o3._nodes is not None else o3._packed if o3._packed is not None else o3._planes
                others[m][form] = {"value": r3, "unit": "queries/s", "of_cell_table": r3 / rate,
                                   "memory_gb": store.numel() * 8 / 1e9, "parity": g3}
                assert g3["ok"] and fin3, f"parity failure in bench ({form}, mode {m}): {g3}"
                o3.release()
                del o3, store
                gc.collect()
                torch.cuda.empty_cache()
            del vals
        # quadcubic (time-dependent field, 256x256 Lekien-Marsden matrix with the A.py:860 quirk), value + 4-gradient
        from arbinterp_b200 import quadcubic
        shape4 = (48, 48, 48, 32)
        axes = [torch.linspace(-1.0, 1.0, m, dtype=torch.float64, device=device) for m in shape4[:3]]
        axes.append(torch.linspace(0.0, 1.0, shape4[3], dtype=torch.float64, device=device))
        T, Z, Y, X = torch.meshgrid(*reversed(axes), indexing="ij")
        u = torch.sin(2 * np.pi * X) * torch.cos(np.pi * Y) * torch.exp(-Z) * torch.cos(2 * T) + X * X * Y + Z * (1 + T) \
            + 0.25 * X * Y * Z * T
        t1 = time.perf_counter()
        o4 = quadcubic(torch.stack([t.reshape(-1) for t in (X, Y, Z, T, u)], dim=1), "quiet")
        torch.cuda.synchronize()
        t_ctor4 = time.perf_counter() - t1
        u_np = u.cpu().numpy().ravel()
        del T, Z, Y, X, u
        Q4 = Q // 4
        q4 = uniform_queries(torch, o4, Q4, 4321, device)
        outs4 = alloc_outputs(torch, "norm", 4, Q4, device)
        el, per = time_device(torch, lib, o4, "norm", 4, q4, outs4, cells[:Q4], args.steps, args.warmup)
        rate = Q4 * args.steps / el
        finite = all(bool(torch.isfinite(t_).all()) for t_ in outs4 if t_ is not None)
        guard = oracle_sample_check(o4, [a.cpu().numpy() for a in axes], {"n": u_np}, "norm", 4, 100_000, 12, scalar=True)
        guard["timed_outputs_finite"] = finite
        others["quadcubic_norm"] = {"value": rate, "unit": "queries/s", "alg_bytes_per_query": ALG_BYTES[(4, "norm")],
                                    "achieved_gbs": ALG_BYTES[(4, "norm")] * rate / 1e9,
                                    "frac_of_measured_hbm": ALG_BYTES[(4, "norm")] * rate / 1e9 / measured_peak()[0],
                                    "traffic_bytes_per_query_ncu": ncu_traffic("4d_norm"),
                                    "queries_per_step": Q4, "table_gb": o4.table.numel() * 8 / 1e9,
                                    "workload": "quadcubic scalar field %dx%dx%dx%d with an xyzt term (config 4 stand-in), "
                                                "uniform random (x,y,z,t) queries" % shape4, "constructor_s": t_ctor4,
                                    "parity": guard}
        assert guard["ok"] and finite, f"parity failure in bench (quadcubic): {guard}"
        rate4 = rate
        o4.release()
        del o4, outs4
        gc.collect()
        torch.cuda.empty_cache()
        o4n = quadcubic(torch.stack([t.reshape(-1) for t in torch.meshgrid(*reversed(axes), indexing="ij")][::-1] +
                                    [torch.from_numpy(u_np).to(device)], dim=1), "quiet", table="nodes")
        r4n, fin4n = time_public_device(torch, o4n, q4, max(3, args.steps // 2), args.warmup)
        g4n = oracle_sample_check(o4n, [a.cpu().numpy() for a in axes], {"n": u_np}, "norm", 4, 50_000, 14, scalar=True)
        others["quadcubic_norm"]["node_table"] = {"value": r4n, "unit": "queries/s", "of_cell_table": r4n / rate4,
                                                  "memory_gb": o4n.nodes.numel() * 8 / 1e9, "parity": g4n,
                                                  "timed_outputs_finite": fin4n}
        assert g4n["ok"] and fin4n, f"parity failure in bench (quadcubic node table): {g4n}"
        o4n.release()
        del o4n, q4
        gc.collect()
        torch.cuda.empty_cache()
        # the headline mode on the node table and table-free
        vals_main = {}
        if args.mode in ("vector", "both"):
            vals_main.update(x=raw_np[0], y=raw_np[1], z=raw_np[2])
        if args.mode in ("norm", "both"):
            vals_main["n"] = np.linalg.norm(np.stack(raw_np, axis=1), axis=1)
        others[args.mode + "_memory_light"] = {}
        for form, kw in (("node_table", {"table": "nodes"}), ("table_free", {"table": False})):
            _, rows = analytic_field_rows(torch, n, device)
            o3 = tricubic(rows, "quiet", mode=args.mode, **kw)
            del rows
            torch.cuda.empty_cache()
            # measured twice (before and after the CPU-side parity check) with the SM clock sampled: these latency-bound
            # kernels ran at 0.3-0.5 of their usual rate in two of five sessions, at the end of a long run on a warm GPU
            smp = ClockSampler(local)
            smp.start()
            r3a, fin3 = time_public_device(torch, o3, q[:Q // 4], max(3, args.steps // 2), args.warmup)
            clk3 = smp.stop()
            g3 = oracle_sample_check(o3, [ax_np] * 3, vals_main, args.mode, 3, 50_000, 15)
            r3b, fin3b = time_public_device(torch, o3, q[:Q // 4], max(3, args.steps // 2), args.warmup)
            r3, fin3 = max(r3a, r3b), fin3 and fin3b
            print(f"[bench] {args.mode} {form}: {r3a:.3e} q/s, again after the parity check {r3b:.3e} q/s; "
                  f"allocated {torch.cuda.memory_allocated() / 1e9:.1f} GB, reserved {torch.cuda.memory_reserved() / 1e9:.1f} GB",
                  file=sys.stderr, flush=True)
            store = o3._nodes if o3._nodes is not None else o3._packed if o3._packed is not None else o3._planes
            others[args.mode + "_memory_light"][form] = {"value": r3, "unit": "queries/s", "of_cell_table": r3 / value,
                                                         "measurements": [r3a, r3b], "clocks_first_measurement": clk3,
                                                         "memory_gb": store.numel() * 8 / 1e9, "parity": g3,
                                                         "timed_outputs_finite": fin3}
            assert g3["ok"] and fin3, f"parity failure in bench ({form}, mode {args.mode}): {g3}"
            o3.release()
            del o3, store
            gc.collect()
            torch.cuda.empty_cache()
        _, rows = analytic_field_rows(torch, n, device)
        obj = tricubic(rows, "quiet", mode=args.mode)
        del rows

    # ---- N > 1: the slab-sharded config-5 path with in-run parity
    sharded = None
    if world > 1 and not args.no_sharded:
        del outs, cells, q
        table_gb_main = obj.table.numel() * 8 / 1e9
        qbytes = Q * 3 * 8
        obj.release()
        del obj
        obj = None
        torch.cuda.empty_cache()
        sharded = run_sharded(torch, dist, args, rank, world, device)
    else:
        table_gb_main = obj.table.numel() * 8 / 1e9
        qbytes = Q * 3 * 8

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (rank 0, N=1 only): the unmodified reference when staged, and the in-bench parity
    # of 10^6 rows against the numpy oracle (the checker)
    cpu = None
    parity = None
    if world == 1 and not args.no_cpu:
        ora = make_cpu_oracle(n, args.mode)
        qc = cpu_queries(ora, args.cpu_sample, 7)
        t0 = time.perf_counter()
        ref = ora.query(qc.copy())
        port_cold = len(qc) / (time.perf_counter() - t0)
        ref_inds = ora.query_inds.copy()                      # the chunked warm pass below overwrites query_inds
        t0 = time.perf_counter()
        for lo_ in range(0, len(qc), 100_000):
            ora.query(qc[lo_:lo_ + 100_000].copy())
        port_warm = len(qc) / (time.perf_counter() - t0)
        got = obj.Query(qc.copy())
        scale = max(float(np.abs(v).max()) for v in ora.values.values())
        worst = scaled_error(got, ref, d, ora.geo.h, scale)
        parity = {"n": int(len(qc)), "max_scaled_err": worst, "tolerance": 1e-12,
                  "indices_equal": bool(np.array_equal(obj.queryInds, ref_inds)),
                  "checker": "oracle/arb_oracle.py (numpy restatement, bit-equal to the live reference on tests/golden)"}
        assert parity["indices_equal"] and worst <= 1e-12, f"parity failure in bench: {parity}"
        del ora, ref
        arm = CpuArm(n, args.mode, args.force_port)
        if arm.kind == "reference":
            qs = uniform_in_volume(arm.lo, arm.hi, args.ref_sample, 8)
            cold = arm.passes(qs, 50_000)                     # first pass: lazy per-cell coefficient fill (A.py:376-377)
            warm = arm.passes(qs, 50_000)
            # the unmodified reference against the GPU path on its own sample
            r_ref = arm.query(qs.copy())
            r_got = obj.Query(qs.copy())
            ref_err = scaled_error(r_got, r_ref, d, [arm.obj.hx, arm.obj.hy, arm.obj.hz], scale)
            ref_idx = bool(np.array_equal(obj.queryInds, arm.obj.queryInds))
            assert ref_idx and ref_err <= 1e-12, f"parity failure against the unmodified reference: {ref_err} {ref_idx}"
            cpu = {"value": warm, "unit": "queries/s", "cores": 1, "kind": "reference",
                   "sample": f"{args.ref_sample} warm queries (second pass, coefficients cached) through the {arm.what} "
                             f"on a {n}^3 grid of the same analytic field, single process (the reference is single-threaded)",
                   "cold_value": cold,
                   "cold_note": "first pass over the same sample incl. the reference's lazy per-cell coefficient fill (A.py:376-377)",
                   "parity_vs_gpu": {"n": int(len(qs)), "max_scaled_err": ref_err, "indices_equal": ref_idx},
                   "port_value": port_warm, "port_cold_value": port_cold,
                   "port_note": f"the numpy port (oracle/arb_oracle.py) on {args.cpu_sample} rows, warm / cold"}
        else:
            cpu = {"value": port_warm, "unit": "queries/s", "cores": 1, "kind": "port",
                   "sample": f"{args.cpu_sample} warm queries (second pass, coefficients cached) through the numpy oracle "
                             f"on a {n}^3 grid of the same analytic field, single process (reference not staged under oracle/_ref)",
                   "cold_value": port_cold}
        del arm

    peak, peak_src = measured_peak()
    alg = ALG_BYTES[(d, args.mode)]
    avg_ms = float(np.mean(per_launch))
    achieved = alg * Q / (avg_ms * 1e-3) / 1e9
    traffic_per_query = ncu_traffic(args.mode)
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(n, args.mode),
                   "queries_per_step_per_gpu": Q, "table_gb": table_gb_main,
                   "cache": "inputs larger than L2 (query batch %.1f GB, table %.1f GB vs 126 MB L2)" %
                            (qbytes / 1e9, table_gb_main),
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
        # per step and rank: the sortedness probe + the two gated query-kernel launches (one of them returns at once)
        "gpu_launches": 3 * args.steps * world,
        "clocks": clocks,
    }
    if build_info:
        line["build"] = build_info
    if others:
        line["other_modes"] = others
    if sharded:
        line["sharded"] = sharded
    line["wall_s"] = time.perf_counter() - _T_PROCESS_START      # the whole run, imports and every extra leg included
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
