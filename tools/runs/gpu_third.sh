#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest.log
timeout 900 python tools/perf_sweep.py --variants 0 --modes norm,both > gpurun_out/sweep_build.log 2>&1
timeout 600 python tools/e2e_sweep.py > gpurun_out/e2e_sweep.log 2>&1
cat gpurun_out/pytest.log | tail -5; grep build gpurun_out/sweep_build.log; cat gpurun_out/e2e_sweep.log
