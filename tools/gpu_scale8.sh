#!/bin/bash
# 8-GPU box: weak-scaling bench at N = 1, 2, 4, 8 and config 5 (96^3 x 64 'both', t-slab sharded) on 8 ranks
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus8.txt
for n in 1 2 4 8; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
        bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    tools/config5_demo.py --grid 96,96,96,64 --particles 4194304 > gpurun_out/config5_n8.log 2>&1
for n in 1 2 4 8; do python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
    print("N=$n value %.4e q/s  e2e %.4e  ms/step %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]))
except Exception as e:
    print("N=$n failed", e)
PY
done
grep config5 gpurun_out/config5_n8.log; tail -n 3 gpurun_out/config5_n8.log
