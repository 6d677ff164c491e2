#!/bin/bash
# one visit: the whole GPU test suite and one bench line per workload (no ncu).  Usage: bash tools/gpu_check.sh <tag>
TAG=${1:-c}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
for wl in weibel landau boris 2d3v; do
    ST=50; [ $wl = 2d3v ] && ST=10
    timeout 900 python bench.py --steps $ST --warmup 3 --no-cpu --workload $wl > $OUT/${TAG}_bench_${wl}.json 2> $OUT/${TAG}_bench_${wl}.err
    python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_${wl}.json"))
    print("$wl: %.4g p-steps/s, ms/step %.4f, e2e %.4g, frac %.3f, clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"]))
    print("   ", {k.replace("operator",""): round(v["avg_ms"], 4) for k, v in d["roofline"]["all_passes"].items()})
except Exception as e:
    print("$wl bench parse failed", e); print(open("$OUT/${TAG}_bench_${wl}.err").read()[-1500:])
PY
done
