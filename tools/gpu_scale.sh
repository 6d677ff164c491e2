#!/bin/bash
# Scaling visit on an N-GPU box: bench.py at 1,2,4,..,N GPUs for the default (Weibel 1d2v) workload and for the
# 2d3v workload, launched exactly like the driver does.   Usage: bash tools/gpu_scale.sh <tag> <Nmax>
TAG=${1:-s}; NMAX=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | wc -l
run () {  # <N> <workload> <extra>
  local N=$1 W=$2; shift 2
  if [ "$N" = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --workload $W --no-cpu "$@" > $OUT/${TAG}_${W}_n${N}.json 2> $OUT/${TAG}_${W}_n${N}.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $N --steps 10 --warmup 3 --workload $W "$@" > $OUT/${TAG}_${W}_n${N}.json 2> $OUT/${TAG}_${W}_n${N}.err
  fi
  python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_${W}_n${N}.json") if l.startswith("{")][-1])
    print("$W N=$N value %.4g ms/step %.3f e2e %.4g" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$W N=$N FAILED", e); print(open("$OUT/${TAG}_${W}_n${N}.err").read()[-1500:])
PY
}
N=1
while [ $N -le $NMAX ]; do
  run $N weibel
  run $N 2d3v --steps 5
  N=$((N*2))
done
run $NMAX boris
run $NMAX landau
run 1 2d3v --steps 6 --sort-interval 2
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
