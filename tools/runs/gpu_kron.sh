#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "build_tile or coefficient_table or quirk" 2>&1 | tail -12 > gpurun_out/pytest_kron.log
timeout 600 python tools/perf_sweep.py --variants 0 --modes norm --build-variants 0,3,4 2>&1 | grep "\[build\]" > gpurun_out/sweep_kron.log
cat gpurun_out/pytest_kron.log | tail -8; cat gpurun_out/sweep_kron.log
