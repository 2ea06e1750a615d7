#!/bin/bash
# separable build kernels: parity tests, variant sweep, ncu captures
mkdir -p gpurun_out
SECONDS=0
timeout 300 python -m pytest tests -m gpu -q -x -k "build_tile or coefficient_table" 2>&1 | tail -4 > gpurun_out/pytest_sep.log
echo "pytest wall ${SECONDS}s" >> gpurun_out/pytest_sep.log
cat gpurun_out/pytest_sep.log
timeout 300 python tools/build_sweep.py 2>&1 | grep -v Warning > gpurun_out/build_sweep.log
cat gpurun_out/build_sweep.log | cut -c1-230
timeout 300 ncu --set full --clock-control none --import-source on -k regex:build_sep3 -c 1 -o gpurun_out/prof_build3d_sep \
    python tools/build_sweep.py --profile 5 --d 3 --modes norm > gpurun_out/prof_sep3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:build_sep4 -c 1 -o gpurun_out/prof_build4d_sep \
    python tools/build_sweep.py --profile 5 --d 4 --modes norm > gpurun_out/prof_sep4.log 2>&1
tail -n 2 gpurun_out/prof_sep3.log gpurun_out/prof_sep4.log
