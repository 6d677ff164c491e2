#!/bin/bash
# tuning visit: register / prefetch variants of the sorted 2d3v passes (k2_sorted), one short bench each
OUT=gpurun_out; mkdir -p $OUT
for V in "2p 3n" "2t 3t" "2t 2t" "2p 4t"; do
  set -- $V
  GEMPIC_K2_HEAD=$1 GEMPIC_K2_TAIL=$2 timeout 300 python bench.py --workload 2d3v --steps 6 --no-cpu --no-configs --min-seconds 0.2 > $OUT/var.json 2> $OUT/var.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/var.json").read().strip().splitlines()[-1])
    p = d["roofline"]["all_passes"]
    g = lambda k: next((round(v["avg_ms"], 3) for kk, v in p.items() if kk.startswith(k)), None)
    print("head", sys.argv[1], g("fused[HE,HE,Hp3,Hp2]"), " tail", sys.argv[2], g("operatorHp3{2,3} sorted"), " step", round(d["ms_per_step"], 3))
except Exception as e:
    print("failed", sys.argv[1:], e, open("gpurun_out/var.err").read()[-800:])
PY
done
