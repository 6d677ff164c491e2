#!/bin/bash
# round-2 single-GPU visit: parity tests, smoke, default bench (with configs block).  Usage: bash tools/gpu_r2.sh <tag>
TAG=${1:-r2}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.4g ms/step %.4f sustained %.4g e2e %.4g loop %s frac %.3f" % (d["value"], d["ms_per_step"], d["sustained"]["value"], d["e2e"]["value"], d["e2e_loop"] and "%.4g" % d["e2e_loop"]["value"], d["roofline"]["frac"]), d["clocks"])
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.4g ms/step %.4f frac %s" % (v["value"], v["ms_per_step"], v["frac"]), {p: round(x["frac"], 3) for p, x in (v["passes"] or {}).items()})
except Exception as e:
    print("bench parse failed", e)
PY
