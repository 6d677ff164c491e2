#!/bin/bash
# round-2 multi-GPU visit: all GPU tests (the NCCL and one-process multi-device tests run up to the box's GPU count),
# then the default bench at N = all GPUs.  Usage: bash tools/gpu_r2_multi.sh <tag> <N>
TAG=${1:-r2m}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/${TAG}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N > $OUT/${TAG}_bench_n${N}.json 2> $OUT/${TAG}_bench_n${N}.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench_n${N}.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n${N}.json").read().strip().splitlines()[-1])
    print("N=$N value %.4g ms/step %.4f sustained %.4g e2e %.4g frac %.3f" % (d["value"], d["ms_per_step"], d["sustained"]["value"], d["e2e"]["value"], d["roofline"]["frac"]), d["clocks"])
    print("parity", d["sharded_parity"])
    for k, v in (d.get("configs") or {}).items():
        print(k, "value %.4g ms/step %.4f frac %s" % (v["value"], v["ms_per_step"], v["frac"]))
except Exception as e:
    print("bench parse failed", e)
PY
