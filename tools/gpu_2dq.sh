#!/bin/bash
# quick 2d3v visit: 2D parity tests + the 2d3v bench
TAG=${1:-q2}; shift
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_2d3v.py tests/test_gpu_maxwell2d.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --workload 2d3v --steps 10 --warmup 3 --no-cpu "$@" > $OUT/${TAG}_bench_2d3v.json 2> $OUT/${TAG}_bench_2d3v.err; echo "bench rc=$?"; tail -c 1500 $OUT/${TAG}_bench_2d3v.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench_2d3v.json"))
    print("value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k, v in d["roofline"]["all_passes"].items(): print("  ", k, v)
except Exception as e:
    print("bench parse failed", e)
PY
