#!/bin/bash
# Usage: bash tools/gpu_scale.sh <tag> <workload...>   (run under gpurun --gpus 8): 1/2/4/8-GPU bench lines
set -u
TAG=${1:-sc}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
NG=$(nvidia-smi -L | wc -l)
for W in "$@"; do
  for n in 1 2 4 8; do
    [ $n -gt $NG ] && continue
    if [ "$n" = "1" ]; then
      timeout 600 python bench.py --gpus 1 --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${W}_n1.json 2> $OUT/bench_${W}_n1.err
    else
      NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
         bench.py --gpus $n --workload $W --steps 5 --warmup 3 > $OUT/bench_${W}_n$n.json 2> $OUT/bench_${W}_n$n.err
    fi
    echo "bench $W n=$n rc=$?" | tee -a $OUT/summary.txt
    tail -1 $OUT/bench_${W}_n$n.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'n', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'E', d['e_xc'])"
    tail -2 $OUT/bench_${W}_n$n.err
  done
done
