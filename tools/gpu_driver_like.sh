#!/bin/bash
# what the driver runs at round end, in its order
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
echo "pytest wall ${SECONDS}s"; SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "smoke wall ${SECONDS}s"; SECONDS=0
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/driver_ref.json 2> gpurun_out/driver_ref.err
echo "reference arm wall ${SECONDS}s"; SECONDS=0
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/driver_bench.json 2> gpurun_out/driver_bench.err
echo "bench wall ${SECONDS}s"
python - <<PY
import json
b=json.loads([l for l in open("gpurun_out/driver_bench.json") if l.startswith("{")][-1])
r=json.loads([l for l in open("gpurun_out/driver_ref.json") if l.startswith("{")][-1])
print("value %.4e frac %.3f e2e %.4e cpu1 %.3e parity %s build_s %.3f" % (b["value"], b["roofline"]["frac"], b["e2e"]["value"], b["cpu_baseline"]["value"], b["parity"]["max_scaled_err"], b["config"]["build_s"]))
print("reference arm %.4e q/s on %d cores -> e2e ratio %.0f" % (r["value"], r["cpu_baseline"]["cores"], b["e2e"]["value"]/r["value"]))
PY
tail -n 2 gpurun_out/driver_bench.err gpurun_out/driver_ref.err
