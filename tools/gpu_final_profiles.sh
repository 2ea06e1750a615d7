#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-other-modes --queries 16777216 --e2e-queries 2097152 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:query_grid_kernel -s 1 -c 1 -o gpurun_out/prof_query_grid \
    python tools/profile_extra.py > gpurun_out/prof_extra1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_kernel -c 1 -o gpurun_out/prof_push \
    python tools/profile_extra.py > gpurun_out/prof_extra2.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --other-modes > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err
tail -n 2 gpurun_out/prof_extra1.log gpurun_out/prof_extra2.log; cut -c1-300 gpurun_out/bench_final.json; ls -la gpurun_out | tail -8
