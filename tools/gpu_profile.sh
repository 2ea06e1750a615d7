#!/bin/bash
mkdir -p gpurun_out
python tools/fp64_peak.py > gpurun_out/fp64_peak.txt 2>&1
# (a) launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-other-modes --queries 16777216 --e2e-queries 2097152 > gpurun_out/bench_under_ncu.log 2>&1
# (b) full capture of the query kernel (norm), (c) of the build kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_block -s 1 -c 1 -o gpurun_out/prof_query_norm \
    python tools/profile_target.py --mode norm > gpurun_out/prof_query.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_kernel -c 1 -o gpurun_out/prof_build3d \
    python tools/profile_target.py --mode norm --launches 1 > gpurun_out/prof_build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_kernel -c 1 -o gpurun_out/prof_build4d \
    python tools/profile_target.py --d 4 --mode norm --launches 1 --queries 1048576 > gpurun_out/prof_build4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:query_block -s 1 -c 1 -o gpurun_out/prof_query_4dboth \
    python tools/profile_target.py --d 4 --mode both --queries 4194304 > gpurun_out/prof_query4.log 2>&1
# real bench lines (not under a profiler)
timeout 900 python bench.py --steps 10 --warmup 3 --other-modes > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/fp64_peak.txt; tail -2 gpurun_out/prof_query.log gpurun_out/prof_build.log gpurun_out/prof_build4.log gpurun_out/prof_query4.log; cat gpurun_out/bench.json gpurun_out/bench_ref.json; tail -3 gpurun_out/bench.err gpurun_out/bench_ref.err; ls -la gpurun_out
