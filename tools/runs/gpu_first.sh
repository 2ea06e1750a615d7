#!/bin/bash
# first GPU trip: parity tests, smoke, variant sweep + bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest.log
echo "pytest exit: ${PIPESTATUS[0]}" >> gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke exit: $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 --sweep 0,1,2,10,11,12 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit: $?" >> gpurun_out/bench.err
tail -5 gpurun_out/pytest.log; tail -3 gpurun_out/smoke.log; cat gpurun_out/bench.err | tail -20; cat gpurun_out/bench.json
