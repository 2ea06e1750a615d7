#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  SECONDS=0
  timeout 420 compute-sanitizer --tool $tool --kernel-regex kns=arb python tools/sanitize_target.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $? (${SECONDS}s)"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize target done|Error|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -8
done
