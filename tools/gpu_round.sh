#!/bin/bash
# One gpurun call: GPU tests, the default bench, the serialised launch list of the same bench command and ncu --set full captures
# of the kernels of the headline step and of the noise generator, exported to text on the box (the .ncu-rep files are too large
# for the 64 MiB return path).  usage: gpurun -- 'bash tools/gpu_round.sh <tag> [pytest args]'
tag=${1:-r02}
shift
o=gpurun_out
mkdir -p $o
(timeout 900 python -m pytest ${@:-tests -m gpu} -x -q --durations=5) > $o/${tag}_pytest.log 2>&1; tail -9 $o/${tag}_pytest.log
(timeout 600 python bench.py) > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err; tail -3 $o/${tag}_bench_n1.err
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1) > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_ref.err
export_rep() {   # details page, raw csv, hottest source lines; then drop the report
    local r=$o/$1.ncu-rep
    [ -f $r ] || return
    ncu -i $r --page details > $o/$1_ncu_details.txt 2>&1
    ncu -i $r --page raw --csv > $o/$1_ncu_raw.csv 2>&1
    python tools/ncu_hot.py $r 40 > $o/$1_ncu_hot_lines.txt 2>&1
    rm -f $r
}
B="python bench.py --steps 32 --warmup 16 --no-also --no-cpu-baseline"
$B > $o/${tag}_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $o/${tag}_launches_bench_c5diag.csv $B > $o/${tag}_ncu_launch.log 2>&1
$B > $o/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dgemm_tma_kernel|k_modal|k_tail_near' -s 40 -c 6 -f -o $o/${tag}_step $B > $o/${tag}_ncu_step.log 2>&1
export_rep ${tag}_step
$B > $o/${tag}_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_tail_far_mma" -s 70 -c 2 -f -o $o/${tag}_far $B > $o/${tag}_ncu_far.log 2>&1
export_rep ${tag}_far
N="python tools/probe_noise.py 256 300 8192"
$N > $o/${tag}_plain_noise.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_fft|k_fill_xi|k_chol|k_factor|dgemm_tma_kernel' -c 10 -f -o $o/${tag}_noise $N > $o/${tag}_ncu_noise.log 2>&1
export_rep ${tag}_noise
du -sh $o; ls -la $o | tail -30
