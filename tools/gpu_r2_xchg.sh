#!/bin/bash
# multi-GPU visit for the peer-memory exchange: NCCL / multi-device tests, then bench with and without it
TAG=${1:-r2x}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $OUT/${TAG}_pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/${TAG}_pytest_multi.log
GEMPIC_NO_XCHG=1 timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "sharded and 2" > $OUT/${TAG}_pytest_multi_nccl.log 2>&1; echo "pytest (NCCL only) rc=$?"; tail -2 $OUT/${TAG}_pytest_multi_nccl.log
run () {  # <name> <env> <args...>
  local NAME=$1 E=$2; shift 2
  env $E timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N "$@" > $OUT/${TAG}_${NAME}_n${N}.json 2> $OUT/${TAG}_${NAME}_n${N}.err; echo "$NAME rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_${NAME}_n${N}.json").read().strip().splitlines()[-1])
    print("$NAME N=$N value %.4g ms/step %.4f sustained %.4g launches %s" % (d["value"], d["ms_per_step"], d["sustained"]["value"], d["gpu_launches"]), d["clocks"]["sm_mhz"], d["roofline"]["avg_launch_ms"])
    for k, v in (d.get("configs") or {}).items():
        print("   ", k, "value %.4g ms/step %.4f" % (v["value"], v["ms_per_step"]))
except Exception as e:
    print("$NAME parse failed", e); print(open("$OUT/${TAG}_${NAME}_n${N}.err").read()[-1500:])
PY
}
run xchg "A=1"
run nccl "GEMPIC_NO_XCHG=1"
