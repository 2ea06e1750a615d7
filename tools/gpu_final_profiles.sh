#!/bin/bash
# End-of-round evidence: what the driver runs (GPU tests, smoke, reference arm, bench) + the ncu launch list of the
# bench command.  Run under gpurun from the repository root; outputs in gpurun_out/, summaries go to profiles/.
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 > gpurun_out/pytest.log
echo "pytest wall ${SECONDS}s" >> gpurun_out/pytest.log; cat gpurun_out/pytest.log; SECONDS=0
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "smoke wall ${SECONDS}s"; SECONDS=0
timeout 600 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > gpurun_out/driver_ref.json 2> gpurun_out/driver_ref.err
echo "reference arm wall ${SECONDS}s"; SECONDS=0
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/driver_bench.json 2> gpurun_out/driver_bench.err
echo "bench wall ${SECONDS}s"; SECONDS=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-other-modes --queries 16777216 --e2e-queries 2097152 > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list wall ${SECONDS}s"
python - <<PY
import json
b=json.loads([l for l in open("gpurun_out/driver_bench.json") if l.startswith("{")][-1])
r=json.loads([l for l in open("gpurun_out/driver_ref.json") if l.startswith("{")][-1])
print("value %.4e frac %.3f e2e %.4e cpu1 %.3e (cold %.3e) parity %s build %.3f ms (%.3f of peak)" % (b["value"], b["roofline"]["frac"], b["e2e"]["value"], b["cpu_baseline"]["value"], b["cpu_baseline"]["cold_value"], b["parity"]["max_scaled_err"], b["build"]["ms"], b["build"]["frac_of_measured_hbm"]))
print({k:(float("%.3e" % v["value"]),round(v["frac_of_measured_hbm"],3)) for k,v in b.get("other_modes",{}).items()})
print("reference arm %.4e q/s on %d cores -> e2e ratio %.0f" % (r["value"], r["cpu_baseline"]["cores"], b["e2e"]["value"]/r["value"]))
PY
tail -n 2 gpurun_out/driver_bench.err gpurun_out/driver_ref.err
