#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest.log
echo "pytest wall ${SECONDS}s" >> gpurun_out/pytest.log
timeout 600 python tools/perf_sweep.py --variants 0 --modes norm,both --build-variants 0,1,4 2>&1 | grep "\[build\]" > gpurun_out/sweep_kron.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_kron -c 1 -o gpurun_out/prof_build3d_kron \
    python tools/profile_target.py --mode norm --grid 160 --launches 1 --queries 1048576 > gpurun_out/prof_kron3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:build_kron -c 1 -o gpurun_out/prof_build4d_kron \
    python tools/profile_target.py --d 4 --mode norm --launches 1 --queries 1048576 > gpurun_out/prof_kron4.log 2>&1
cat gpurun_out/pytest.log; cat gpurun_out/sweep_kron.log
