#!/bin/bash
# multi-GPU trip (gpurun --gpus N): slab-sharding test + bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_multi.log
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
done
cat gpurun_out/pytest_multi.log; cat gpurun_out/bench_n*.json; tail -n 5 gpurun_out/bench_n$N.err
