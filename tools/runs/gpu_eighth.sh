#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
python - <<PY
import json
for n in (1,2):
    try:
        d=json.loads([l for l in open(f"gpurun_out/bench_n{n}.json") if l.startswith("{")][-1])
        print(f"N={n} value {d['value']:.4e} e2e {d['e2e']['value']:.4e} parity {d.get('parity')} affinity {d['config'].get('cpu_affinity')}")
    except Exception as e:
        print(n, "failed", e)
PY
head -12 gpurun_out/topo.txt; tail -n 3 gpurun_out/bench_n1.err
