#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/pytest.log
echo "pytest wall ${SECONDS}s exit ${PIPESTATUS[0]}" >> gpurun_out/pytest.log
cat gpurun_out/pytest.log
