#!/bin/bash
# 8-GPU visit, trimmed: N = 1 and N = NMAX for the Weibel workload, N = NMAX for 2d3v / boris / landau, NCCL parity test.
TAG=${1:-s8}; NMAX=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
run () {  # <N> <workload> <extra>
  local N=$1 W=$2; shift 2
  if [ "$N" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --warmup 3 --workload $W --no-cpu "$@" > $OUT/${TAG}_${W}_n${N}.json 2> $OUT/${TAG}_${W}_n${N}.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $N --warmup 3 --workload $W "$@" > $OUT/${TAG}_${W}_n${N}.json 2> $OUT/${TAG}_${W}_n${N}.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${W}_n${N}.json") if l.startswith("{")][-1])
    print("$W N=$N value %.4g ms/step %.4f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$W N=$N FAILED", e); print(open("$OUT/${TAG}_${W}_n${N}.err").read()[-1500:])
PY
}
run 1 weibel --steps 50
run $NMAX weibel --steps 50
run $NMAX 2d3v --steps 10
run 1 2d3v --steps 10
run $NMAX boris --steps 50
run $NMAX landau --steps 50
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
