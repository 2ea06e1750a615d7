#!/usr/bin/env python
"""e2e (host numpy -> Query -> host numpy) throughput vs chunk size of the stream ring."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
from tools.perf_sweep import field_rows
dev = torch.device("cuda", 0)
obj = tricubic(field_rows((128,) * 3, dev), "quiet", mode="norm")
n = 1 << 24
qh = torch.empty(n, 3, dtype=torch.float64, pin_memory=True)
lo = np.array(obj._geo.int_min); hi = np.array(obj._geo.int_max)
qh.copy_(torch.from_numpy(lo + np.random.default_rng(0).uniform(0, 1, (n, 3)) * (hi - lo) * (1 - 1e-12)))
qnp = qh.numpy()
qpage = qnp.copy()
chunks = [int(c) for c in os.environ.get("ARB_E2E_CHUNKS", "0,262144,524288,1048576,2097152,4194304").split(",")]
print(f"[e2e] ARB_COPY_THREADS={os.environ.get('ARB_COPY_THREADS', 'default')} cores={len(os.sched_getaffinity(0))} (chunk 0 = the library's own choice)", flush=True)
for chunk in chunks:
    os.environ["ARB_HOST_CHUNK_ROWS"] = str(chunk)
    for name, arr in (("pinned", qnp), ("pageable", qpage)):
        obj.Query(arr)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            obj.Query(arr)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
        print(f"[e2e] chunk={chunk:8d} {name:8s}: {n / dt:.3e} q/s  (H2D {24 * n / dt / 1e9:.1f} GB/s, D2H {32 * n / dt / 1e9:.1f} GB/s)", flush=True)
