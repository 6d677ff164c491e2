#!/bin/bash
# N-GPU visit: the NCCL parity test and the bench at N GPUs (torchrun, one rank per GPU), plus the
# single-GPU bench of the other BASELINE workloads.   Usage: bash tools/gpu_multi.sh <tag> <N>
TAG=${1:-m}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
for W in weibel landau boris; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 --workload $W > $OUT/${TAG}_bench_${W}_n${N}.json 2> $OUT/${TAG}_bench_${W}_n${N}.err
  echo "bench $W N=$N rc=$?"; tail -c 600 $OUT/${TAG}_bench_${W}_n${N}.err
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_${W}_n${N}.json')); print('$W N=$N value %.4g ms/step %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline']['kernel'], '%.0f GB/s' % d['roofline']['achieved'])"
done
for W in landau boris; do
  timeout 600 python bench.py --steps 10 --warmup 3 --workload $W --no-cpu > $OUT/${TAG}_bench_${W}_n1.json 2> $OUT/${TAG}_bench_${W}_n1.err
  python -c "
import json; d=json.load(open('$OUT/${TAG}_bench_${W}_n1.json')); print('$W N=1 value %.4g ms/step %.3f e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d['roofline']['kernel'], '%.0f GB/s' % d['roofline']['achieved'])"
done
