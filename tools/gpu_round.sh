#!/bin/bash
# One GPU-box visit: parity tests, bench (fused + unfused), ncu launch list, ncu --set full of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag>
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_fused.json 2> $OUT/${TAG}_bench_fused.err; echo "bench fused rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 --fuse 0 --no-cpu > $OUT/${TAG}_bench_unfused.json 2> $OUT/${TAG}_bench_unfused.err; echo "bench unfused rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "bench ref rc=$?"
cat $OUT/${TAG}_bench_fused.json | cut -c1-1500
# launch list (every launch with its device time) of the same bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_fused.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_launch_fused.log 2>&1; echo "ncu launches fused rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_unfused.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --fuse 0 > $OUT/${TAG}_ncu_launch_unfused.log 2>&1; echo "ncu launches unfused rc=$?"
# full capture of the particle passes; raw + source pages are exported to CSV on the box (the .ncu-rep of
# more than one or two launches would push gpurun_out/ over its 64 MiB limit)
export_rep () {  # <rep basename>
    ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
    ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_source.csv 2>/dev/null
    ncu -i $OUT/$1.ncu-rep --page details --csv > $OUT/$1_details.csv 2>/dev/null
}
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:OpStrangFused -s 3 -c 2 -f -o $OUT/${TAG}_prof_fused \
    python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_full_fused.log 2>&1; echo "ncu full fused rc=$?"
export_rep ${TAG}_prof_fused
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 4 -c 5 -f -o $OUT/${TAG}_prof_unfused \
    python bench.py --steps 2 --warmup 3 --no-cpu --fuse 0 > $OUT/${TAG}_ncu_full_unfused.log 2>&1; echo "ncu full unfused rc=$?"
export_rep ${TAG}_prof_unfused
rm -f $OUT/${TAG}_prof_unfused.ncu-rep
du -sh $OUT; ls -la $OUT
