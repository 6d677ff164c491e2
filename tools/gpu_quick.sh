#!/bin/bash
# Short GPU-box visit while iterating on a kernel: parity tests, one bench line, one ncu --set full capture
# of the fused Strang pass (raw/source pages exported to CSV on the box, .ncu-rep dropped).
# Usage: bash tools/gpu_quick.sh <tag> [extra bench args]
TAG=${1:-q}; shift
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_bench.json"))
    print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
    for k, v in d["roofline"]["all_passes"].items(): print("  ", k, v)
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/${TAG}_bench.err").read()[-3000:])
PY
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:${KREGEX:-OpStrangFused} -s ${KSKIP:-4} -c 1 -f \
    -o $OUT/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-cpu "$@" > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/${TAG}_prof.ncu-rep --page raw --csv > $OUT/${TAG}_prof_raw.csv 2>/dev/null
ncu -i $OUT/${TAG}_prof.ncu-rep --page source --csv > $OUT/${TAG}_prof_source.csv 2>/dev/null
rm -f $OUT/${TAG}_prof.ncu-rep
