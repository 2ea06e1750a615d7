#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python tools/ctor_profile.py > gpurun_out/ctor_profile.log 2>&1
timeout 600 python tools/e2e_sweep.py > gpurun_out/e2e_sweep2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_kernel -c 1 -o gpurun_out/prof_build3d_v2 \
    python tools/profile_target.py --mode norm --launches 1 > gpurun_out/prof_build_v2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:build_kernel -c 1 -o gpurun_out/prof_build4d_v2 \
    python tools/profile_target.py --d 4 --mode norm --launches 1 --queries 1048576 > gpurun_out/prof_build4_v2.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/config5_demo.py --grid 64,64,64,40 > gpurun_out/config5_n$N.log 2>&1
grep ctor gpurun_out/ctor_profile.log; tail -n 25 gpurun_out/ctor_profile.log; cat gpurun_out/e2e_sweep2.log; grep -v "^\*\|OMP" gpurun_out/config5_n$N.log | tail -n 8
