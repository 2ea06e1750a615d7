#!/usr/bin/env python
"""Where does construction time go (ingest vs build)?  Wall-clock sections with device syncs."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from arbinterp_b200 import tricubic
from arbinterp_b200.ingest import ingest_field, norm_plane
from tools.perf_sweep import field_rows
dev = torch.device("cuda", 0)
rows = field_rows((256,) * 3, dev)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); planes, geo = ingest_field(rows, 3, device=dev); torch.cuda.synchronize(); t1 = time.perf_counter()
    nrm = norm_plane(planes[0:3]); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"[ctor] rep{rep} ingest_field {1e3 * (t1 - t0):.1f} ms, norm_plane {1e3 * (t2 - t1):.1f} ms", flush=True)
    del planes, nrm
for mode in ("norm", "both"):
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        obj = tricubic(rows, "quiet", mode=mode); torch.cuda.synchronize(); t1 = time.perf_counter()
        print(f"[ctor] tricubic(256^3, mode={mode}) rep{rep}: {1e3 * (t1 - t0):.1f} ms total", flush=True)
        del obj
# section timing inside ingest
import torch.autograd.profiler as prof
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as p:
    planes, geo = ingest_field(rows, 3, device=dev); torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cpu_time_total", row_limit=14, max_name_column_width=50))
