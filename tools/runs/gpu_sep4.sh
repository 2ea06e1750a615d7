#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q -x -k "build_tile or coefficient_table or update_values or plain_c or quirk or slab or multi_tile" 2>&1 | tail -5 > gpurun_out/pytest_sep4.log
echo "pytest wall ${SECONDS}s" >> gpurun_out/pytest_sep4.log
cat gpurun_out/pytest_sep4.log
timeout 300 python tools/build_sweep.py --d 4 --variants 0,5,7,9 2>&1 | grep -v Warning > gpurun_out/build_sweep3.log
cat gpurun_out/build_sweep3.log | cut -c1-150
timeout 300 ncu --set full --clock-control none --import-source on -k regex:build_sep4 -c 1 -o gpurun_out/prof_build4d_sep3 \
    python tools/build_sweep.py --profile 0 --d 4 --modes norm > gpurun_out/prof_sep4.log 2>&1
tail -n 1 gpurun_out/prof_sep4.log
