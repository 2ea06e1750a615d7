#!/bin/bash
# new default build: full GPU suite, smoke, bench, stress parity, ncu captures of the default build kernels
mkdir -p gpurun_out
SECONDS=0
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest.log
echo "pytest wall ${SECONDS}s" >> gpurun_out/pytest.log
cat gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
SECONDS=0
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_sep.json 2> gpurun_out/bench_sep.err
echo "bench wall ${SECONDS}s"; tail -n 3 gpurun_out/bench_sep.err
python - <<PY
import json
b=json.loads([l for l in open("gpurun_out/bench_sep.json") if l.startswith("{")][-1])
print("value %.4e frac %.3f e2e %.4e build %s" % (b["value"], b["roofline"]["frac"], b["e2e"]["value"], b.get("build")))
print({k:(v["value"],round(v["frac_of_measured_hbm"],3)) for k,v in b.get("other_modes",{}).items()})
PY
timeout 600 python tools/stress_parity.py 2>&1 | tail -3 > gpurun_out/stress_sep.log; cat gpurun_out/stress_sep.log
timeout 300 python tools/ctor_profile.py > gpurun_out/ctor_profile2.log 2>&1; tail -n 12 gpurun_out/ctor_profile2.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:build_sep3 -c 1 -o gpurun_out/prof_build3d_sep_default \
    python tools/build_sweep.py --profile 0 --d 3 --modes norm > gpurun_out/prof_sep3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:build_sep3 -c 1 -o gpurun_out/prof_build3d_sep_both \
    python tools/build_sweep.py --profile 0 --d 3 --modes both > gpurun_out/prof_sep3b.log 2>&1
tail -n 1 gpurun_out/prof_sep3.log gpurun_out/prof_sep3b.log
