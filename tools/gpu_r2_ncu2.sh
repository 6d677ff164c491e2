TAG=r2l; OUT=gpurun_out; mkdir -p $OUT
B2="python bench.py --workload 2d3v --steps 2 --warmup 3 --no-cpu --no-configs --min-seconds 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_2d3v.csv $B2 > $OUT/${TAG}_l2.log 2>&1; echo "launch list rc=$?"
timeout 2400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k2_sorted|k2_pass" -s 11 -c 4 -f -o $OUT/${TAG}_prof_2d3v $B2 > $OUT/${TAG}_f3.log 2>&1; echo "full rc=$?"
ncu -i $OUT/${TAG}_prof_2d3v.ncu-rep --page raw --csv > $OUT/${TAG}_prof_2d3v_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof_2d3v.ncu-rep --page source --csv > $OUT/${TAG}_prof_2d3v_source.csv 2>/dev/null
rm -f $OUT/${TAG}_prof_2d3v.ncu-rep; ls -la $OUT | grep $TAG
