import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from arbinterp_b200 import tricubic
from tools.nodes_bench import rate as lib_rate
dev = torch.device("cuda", 0)
n = 256
_, rows = bench.analytic_field_rows(torch, n, dev)
o = tricubic(rows, "quiet", mode="norm", table="nodes")
del rows
Q = 1 << 26
q = bench.uniform_queries(torch, o, Q, 1234, dev)
for nq in (Q // 4, Q // 8, Q):
    qq = q[:nq]
    r1, _ = bench.time_public_device(torch, o, qq, 10, 5)
    r2 = lib_rate(o, qq)[0]
    qc = qq.clone()
    r3, _ = bench.time_public_device(torch, o, qc, 10, 5)
    print(f"nq={nq}: Query(slice) {r1:.3e}  lib direct {r2:.3e}  Query(clone) {r3:.3e}", flush=True)
from tools.perf_sweep import field_rows
o2 = tricubic(field_rows((256,) * 3, dev), "quiet", mode="norm", table="nodes")
print("perf_sweep field:", bench.time_public_device(torch, o2, q[:Q // 4], 10, 5)[0], lib_rate(o2, q[:Q // 4])[0])
print("h", o.hx, o2.hx, o.xIntMin, o2.xIntMin)
