#!/bin/bash
# 2d3v visit: bench at the config-5 per-GPU share, launch list, ncu --set full of the four passes.
TAG=${1:-d2}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python bench.py --workload 2d3v --steps 5 --warmup 3 "$@" > $OUT/${TAG}_bench_2d3v.json 2> $OUT/${TAG}_bench_2d3v.err; echo "bench rc=$?"; tail -c 1500 $OUT/${TAG}_bench_2d3v.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_2d3v.json"))
    print("value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "cpu", d.get("cpu_baseline"))
    for k, v in d["roofline"]["all_passes"].items(): print("  ", k, v)
except Exception as e:
    print("bench parse failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_2d3v.csv \
    python bench.py --workload 2d3v --steps 2 --warmup 3 --no-cpu --particles 50000000 > $OUT/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k2_pass -s 12 -c 5 -f \
    -o $OUT/${TAG}_prof2d python bench.py --workload 2d3v --steps 2 --warmup 3 --no-cpu --particles 50000000 > $OUT/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i $OUT/${TAG}_prof2d.ncu-rep --page raw --csv > $OUT/${TAG}_prof2d_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof2d.ncu-rep --page source --csv > $OUT/${TAG}_prof2d_source.csv 2>/dev/null
rm -f $OUT/${TAG}_prof2d.ncu-rep
ls -la $OUT | tail -8
