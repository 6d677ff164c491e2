#!/bin/bash
# one visit: 2D + Boris parity tests, Boris bench, 2d3v bench (+ optional ncu via tools/gpu_2d.sh)
TAG=${1:-r}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_2d3v.py tests/test_gpu_parity.py -m gpu -x -q -k "2d3v or boris or operators or strang_resident or invariants or displacements" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload boris > $OUT/${TAG}_bench_boris.json 2> $OUT/${TAG}_bench_boris.err
python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_boris.json')); print('boris value %.4g ms/step %.4f' % (d['value'], d['ms_per_step']), {k: round(v['avg_ms'],4) for k,v in d['roofline']['all_passes'].items()})"
bash tools/gpu_2d.sh $TAG --no-cpu
