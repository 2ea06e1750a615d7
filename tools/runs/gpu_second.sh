#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest.log
timeout 1500 python tools/perf_sweep.py > gpurun_out/sweep.log 2>&1
echo "sweep exit: $?" >> gpurun_out/sweep.log
tail -8 gpurun_out/pytest.log; cat gpurun_out/sweep.log
