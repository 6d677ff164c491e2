#!/bin/bash
# Kernel-variant sweep on one GPU box: for every variants/lib*.so swap it in as the product library, run the fused-pass
# parity tests and one bench line per workload.  Usage: bash tools/gpu_variants.sh <tag> [bench args]
TAG=${1:-v}; shift
OUT=gpurun_out
mkdir -p $OUT
cp gempic.jl_b200/libgempic_b200.so /tmp/lib_orig.so
for lib in variants/lib*.so; do
    name=$(basename $lib .so)
    cp $lib gempic.jl_b200/libgempic_b200.so
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or deferred or operators_1d2v" > $OUT/${TAG}_${name}_pytest.log 2>&1
    echo "$name pytest rc=$? $(tail -1 $OUT/${TAG}_${name}_pytest.log)"
    for wl in weibel landau; do
        timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload $wl "$@" > $OUT/${TAG}_${name}_${wl}.json 2> $OUT/${TAG}_${name}_${wl}.err
        python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_${name}_${wl}.json"))
    print("  $name $wl value %.4g ms/step %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]),
          {k: round(v["avg_ms"], 4) for k, v in d["roofline"]["all_passes"].items()})
except Exception as e:
    print("  $name $wl bench parse failed", e); print(open("$OUT/${TAG}_${name}_${wl}.err").read()[-2000:])
PY
    done
done
cp /tmp/lib_orig.so gempic.jl_b200/libgempic_b200.so
