#!/bin/bash
# round-2 profiling visit (one GPU): ncu launch lists of the bench commands and --set full captures of the top kernels.
# Raw / source / details pages are exported to CSV on the box (the .ncu-rep files would exceed gpurun_out's 64 MiB).
TAG=${1:-r2e}
OUT=gpurun_out; mkdir -p $OUT
B1="python bench.py --steps 2 --warmup 3 --no-cpu --no-configs --min-seconds 0"
B2="python bench.py --workload 2d3v --steps 2 --warmup 3 --no-cpu --no-configs --min-seconds 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fused.csv $B1 > $OUT/${TAG}_ncu_l1.log 2>&1; echo "launch list 1d rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_2d3v.csv $B2 > $OUT/${TAG}_ncu_l2.log 2>&1; echo "launch list 2d3v rc=$?"
export_rep () {  # <rep basename>
    ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
    ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_source.csv 2>/dev/null
    rm -f $OUT/$1.ncu-rep
}
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:OpStrangFused -s 4 -c 1 -f -o $OUT/${TAG}_prof_fused $B1 > $OUT/${TAG}_ncu_f1.log 2>&1; echo "full fused rc=$?"
export_rep ${TAG}_prof_fused
timeout 1500 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:OpLoopTail -s 3 -c 1 -f -o $OUT/${TAG}_prof_looptail $B1 > $OUT/${TAG}_ncu_f2.log 2>&1; echo "full loop tail rc=$?"
export_rep ${TAG}_prof_looptail
# 2d3v: after set-up (sort, charge) and 3 warm-up steps, one whole step of particle kernels
timeout 2400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k2_sorted|k2_pass|k2_hist" -s 14 -c 5 -f -o $OUT/${TAG}_prof_2d3v $B2 > $OUT/${TAG}_ncu_f3.log 2>&1; echo "full 2d3v rc=$?"
export_rep ${TAG}_prof_2d3v
du -sh $OUT; ls -la $OUT | grep ${TAG}
