#!/bin/bash
# ncu --set full of the ensemble-persistent kernel on the config-2 shape, exported to text on the box
tag=${1:-r02_ens}
o=gpurun_out; mkdir -p $o
B="python bench.py --workload c2 --steps 256 --warmup 16 --no-also --no-cpu-baseline"
$B > $o/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_md_ens -s 1 -c 1 -f -o $o/${tag} $B > $o/${tag}_ncu.log 2>&1
r=$o/${tag}.ncu-rep
if [ -f $r ]; then
  ncu -i $r --page details > $o/${tag}_ncu_details.txt 2>&1
  ncu -i $r --page raw --csv > $o/${tag}_ncu_raw.csv 2>&1
  python tools/ncu_hot.py $r 60 > $o/${tag}_ncu_hot_lines.txt 2>&1
  rm -f $r
fi
tail -2 $o/${tag}_ncu.log
