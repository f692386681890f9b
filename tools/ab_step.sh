#!/bin/bash
# serialised per-kernel times of one bench block for a few A/B settings (ncu launch list)
o=gpurun_out; mkdir -p $o
B="python bench.py --steps 32 --warmup 16 --no-also --no-cpu-baseline"
i=0
for env in "$@"; do
  i=$((i+1))
  env $env $B > $o/ab_plain.log 2>&1 && env $env timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/ab_$i.csv $B > $o/ab_ncu.log 2>&1
  echo "== $env"; python tools/launch_summary.py $o/ab_$i.csv | grep -E "k_modal|k_tail_near|k_tail_far_mma|dgemm_tma"
done
