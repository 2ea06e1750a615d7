#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest.log
timeout 900 python tools/perf_sweep.py --d 3 --variants 0 --build-variants 0 --table-free > gpurun_out/sweep_tf.log 2>&1
timeout 900 python tools/perf_sweep.py --d 3 --grid3 128 --variants 0 --build-variants 0 --table-free --modes norm >> gpurun_out/sweep_tf.log 2>&1
tail -n 12 gpurun_out/pytest.log; grep -E "query|Error|error" gpurun_out/sweep_tf.log
