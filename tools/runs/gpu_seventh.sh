#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest.log
echo "pytest wall ${SECONDS}s" >> gpurun_out/pytest.log
timeout 300 python tools/latency_bench.py > gpurun_out/latency.log 2>&1
tail -n 6 gpurun_out/pytest.log; cat gpurun_out/latency.log | tail -n 8
