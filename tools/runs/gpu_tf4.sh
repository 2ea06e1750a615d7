#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q -x -k "table_free" 2>&1 | tail -6 > gpurun_out/pytest_tf4.log
echo "pytest wall ${SECONDS}s" >> gpurun_out/pytest_tf4.log
cat gpurun_out/pytest_tf4.log
timeout 600 python tools/stress_parity.py --cases 120 --seed 5 2>&1 | tail -3 > gpurun_out/stress_tf4.log; cat gpurun_out/stress_tf4.log
timeout 600 python tools/perf_sweep.py --d 4 --variants 0 --build-variants 0 --table-free 2>&1 | grep "\[query\]" > gpurun_out/sweep_tf4.log
timeout 300 python tools/perf_sweep.py --d 4 --grid4 32,32,32,16 --modes norm --variants 0 --build-variants 0 --table-free 2>&1 | grep "\[query\]" >> gpurun_out/sweep_tf4.log
cat gpurun_out/sweep_tf4.log
