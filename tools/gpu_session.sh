#!/bin/bash
# One parameterised GPU session (replaces round 1's tools/runs/gpu_*.sh).  Run under gpurun from the repo root:
#     gpurun --timeout 1500 -- 'bash tools/gpu_session.sh tests smoke ref bench'
#     gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_session.sh multi bench:2'
# Stages (any order, any subset); everything lands in gpurun_out/<tag>_*.{log,json} (tag = $ARB_TAG, default "s"):
#   tests        pytest -m gpu (whole suite)            tests:<expr>  pytest -m gpu -k <expr>
#   multi        tests/test_multi_gpu.py                smoke         __graft_entry__.smoke()
#   ref          bench.py --impl reference              bench         bench.py at N=1 (driver flags)
#   bench:<N>    bench.py under torchrun at N ranks     link:<N>      tools/link_ceiling.py at N ranks
#   launches     ncu launch list of the bench command   ncu:<name>    ncu --set full of tools/$ARB_TARGET $ARB_ARGS (kernel regex $ARB_KERNEL) -> <tag>_<name>.ncu-rep
#   py:<script>  python tools/<script> (extra args via $ARB_ARGS)
#   sanitize     compute-sanitizer memcheck + racecheck over tools/sanitize_target.py
TAG=${ARB_TAG:-s}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_gpus.txt 2>&1
STEPS=${ARB_STEPS:-20}; WARM=${ARB_WARMUP:-5}
torchrun_n() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 "$@"; }
for stage in "$@"; do
  SECONDS=0
  case $stage in
    tests)      timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_pytest.log ;;
    tests:*)    timeout 1500 python -m pytest tests -m gpu -x -q -k "${stage#tests:}" 2>&1 | tail -40 > $OUT/${TAG}_pytest_k.log; cat $OUT/${TAG}_pytest_k.log ;;
    multi)      timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_round2.py -m gpu -q -k "world2 or devices or current_device" 2>&1 | tail -25 > $OUT/${TAG}_pytest_multi.log; cat $OUT/${TAG}_pytest_multi.log ;;
    smoke)      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log ;;
    ref)        timeout 900 python bench.py --impl reference --gpus 1 --steps $STEPS --warmup $WARM > $OUT/${TAG}_ref.json 2> $OUT/${TAG}_ref.err; tail -c 1500 $OUT/${TAG}_ref.json; tail -3 $OUT/${TAG}_ref.err ;;
    bench)      timeout 900 python bench.py --gpus 1 --steps $STEPS --warmup $WARM $ARB_BENCH_ARGS > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; cat $OUT/${TAG}_bench_n1.json; tail -5 $OUT/${TAG}_bench_n1.err ;;
    bench:*)    n=${stage#bench:}; timeout 1200 bash -c "$(declare -f torchrun_n); torchrun_n $n bench.py --gpus $n --steps $STEPS --warmup $WARM $ARB_BENCH_ARGS" > $OUT/${TAG}_bench_n$n.json 2> $OUT/${TAG}_bench_n$n.err; grep '^{' $OUT/${TAG}_bench_n$n.json; tail -8 $OUT/${TAG}_bench_n$n.err ;;
    link:*)     n=${stage#link:}; if [ "$n" = "1" ]; then timeout 300 python tools/link_ceiling.py; else timeout 300 bash -c "$(declare -f torchrun_n); torchrun_n $n tools/link_ceiling.py"; fi 2>&1 | grep '^{' | tee $OUT/${TAG}_link_n$n.json ;;
    launches)   timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-other-modes > $OUT/${TAG}_launches_bench.log 2>&1; python tools/launch_list.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.txt 2>&1; tail -20 $OUT/${TAG}_launches.txt ;;
    ncu:*)      t=${stage#ncu:}; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${ARB_KERNEL:-.}" -c ${ARB_NCU_COUNT:-1} -s ${ARB_NCU_SKIP:-0} -f -o $OUT/${TAG}_$t python tools/${ARB_TARGET:-profile_target.py} $ARB_ARGS > $OUT/${TAG}_ncu_$t.log 2>&1; tail -3 $OUT/${TAG}_ncu_$t.log; python tools/ncu_summary.py $OUT/${TAG}_$t.ncu-rep > $OUT/${TAG}_${t}_ncu.txt 2>&1; head -60 $OUT/${TAG}_${t}_ncu.txt ;;
    py:*)       s=${stage#py:}; timeout 1200 python tools/$s $ARB_ARGS > $OUT/${TAG}_${s%.py}.log 2>&1; tail -60 $OUT/${TAG}_${s%.py}.log ;;
    sanitize)   bash tools/gpu_sanitize.sh ;;
    *)          echo "unknown stage $stage" ;;
  esac
  echo "== stage $stage: ${SECONDS}s"
done
