#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the shipped library uses
(B200_PROFILING.md): UBLKCP (cp.async.bulk, TMA engine), UTMALDG (cp.async.bulk.tensor, TMA tiles), LDGSTS (cp.async),
DMMA (FP64 tensor-core MMA), DFMA (FP64 pipe), SYNCS (mbarrier), MATCH (warp de-duplication), and the cubin targets.
Usage: python tools/sass_summary.py [arbinterp_b200/libarbinterp_b200.so] > profiles/r02_sass_summary.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "arbinterp_b200/libarbinterp_b200.so"
OPS = ["UBLKCP", "UTMALDG", "LDGSTS", "DMMA", "DFMA", "SYNCS", "MATCH", "SHFL", "LDS", "UTCHMMA", "UTCQMMA", "LDTM", "STTM"]
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
print(f"# {lib}")
print("# cubins: " + ", ".join(sorted(set(re.findall(r"sm_\d+a?", elf)))))
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*\)$", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[cur][o] += 1
print("# " + " ".join(f"{o:>8s}" for o in OPS) + "  kernel")
tot = collections.Counter()
for k, c in counts.items():
    tot.update(c)
    print("  " + " ".join(f"{c[o]:8d}" for o in OPS) + "  " + k)
print("# " + " ".join(f"{tot[o]:8d}" for o in OPS) + "  TOTAL over %d kernels" % len(counts))
print("# no UTC*MMA / LDTM / STTM: tcgen05.mma has no f64 kind (SURVEY 7.2); the dense-contraction variants use DMMA")
