#!/bin/bash
# ncu --set full of one kernel of the default bench step, exported to text on the box.  usage: bash tools/ncu_one.sh <tag> <kernel regex> [skip] [count] [extra bench args]
tag=$1; re=$2; skip=${3:-2}; cnt=${4:-1}; shift 4
o=gpurun_out; mkdir -p $o
B="python bench.py --steps 32 --warmup 16 --no-also --no-cpu-baseline $@"
$B > $o/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$re" -s $skip -c $cnt -f -o $o/${tag} $B > $o/${tag}_ncu.log 2>&1
r=$o/${tag}.ncu-rep
if [ -f $r ]; then
  ncu -i $r --page details > $o/${tag}_ncu_details.txt 2>&1
  ncu -i $r --page raw --csv > $o/${tag}_ncu_raw.csv 2>&1
  python tools/ncu_hot.py $r 60 > $o/${tag}_ncu_hot_lines.txt 2>&1
  rm -f $r
fi
tail -3 $o/${tag}_ncu.log
