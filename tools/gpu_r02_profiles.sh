#!/bin/bash
# Round-2 ncu captures (one GPU; run under gpurun from the repo root): `ncu --set full` of the LAST launch of the
# query kernel for each configuration below -> gpurun_out/r02_<name>.ncu-rep + gpurun_out/r02_<name>_ncu.txt
# (tools/ncu_summary.py; copy the .txt files to profiles/).  Kernel regex picks the launch that actually runs work:
# the probed default issues probe + two gated launches, so captures name a forced variant (25 = 32 slots, 43 = 16-slot
# rings + prefetch) to profile exactly one kernel.
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name units_per_launch args...   (SKIP = query-kernel launches before the captured one)
  name=$1; units=$2; shift 2
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX:-query_block_kernel}" -s ${SKIP:-2} -c 1 -f \
      -o $OUT/r02_$name python tools/profile_target.py --launches 3 "$@" > $OUT/r02_ncu_$name.log 2>&1
  tail -1 $OUT/r02_ncu_$name.log
  python tools/ncu_summary.py $OUT/r02_$name.ncu-rep $units > $OUT/r02_${name}_ncu.txt 2>&1
  grep -E "derived|gpu__time_duration|dram_throughput|lts__t_sector_hit" $OUT/r02_${name}_ncu.txt
  case " $KEEP_REPS " in *" $name "*) ;; *) rm -f $OUT/r02_$name.ncu-rep ;; esac    # reports are 10-30 MB each
}
KEEP_REPS=${KEEP_REPS:-"query3d_norm_sorted_16slots nodes3d_norm"}
Q=16777216
for c in "$@"; do
  case $c in
    vector)       cap query3d_vector $Q --mode vector --variant 25 ;;
    both)         cap query3d_both $Q --mode both --variant 25 ;;
    norm4)        cap query4d_norm 4194304 --d 4 --mode norm --queries 4194304 --variant 25 ;;
    both4)        cap query4d_both 4194304 --d 4 --mode both --queries 4194304 --variant 25 ;;
    sorted32)     SKIP=3 cap query3d_norm_sorted_32slots $Q --mode norm --order sorted --variant 25 ;;
    sorted16)     SKIP=3 cap query3d_norm_sorted_16slots $Q --mode norm --order sorted --variant 43 ;;
    nodes)        cap nodes3d_norm $Q --mode norm --table nodes ;;
    nodes_sorted) SKIP=3 cap nodes3d_norm_sorted $Q --mode norm --table nodes --order sorted ;;
    nodes4)       cap nodes4d_both 4194304 --d 4 --mode both --queries 4194304 --table nodes ;;
    gridil4)      KREGEX=query_gridil4_kernel SKIP=1 cap gridil4_both 4194304 --d 4 --mode both --queries 4194304 --table free ;;
    nodes_vec)    cap nodes3d_vector $Q --mode vector --table nodes ;;
    *) echo "unknown capture $c" ;;
  esac
done
