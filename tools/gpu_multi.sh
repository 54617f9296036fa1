#!/bin/bash
# Usage: bash tools/gpu_multi.sh <tag> <ngpus> <workload...>   (run under gpurun --gpus N)
set -u
TAG=${1:-m}; N=${2:-2}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.csv 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for W in "$@"; do
  for n in 1 $N; do
    if [ "$n" = "1" ]; then
      timeout 600 python bench.py --gpus 1 --workload $W --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench_${W}_n1.json 2> $OUT/bench_${W}_n1.err
    else
      NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
         bench.py --gpus $n --workload $W --steps 5 --warmup 3 > $OUT/bench_${W}_n$n.json 2> $OUT/bench_${W}_n$n.err
    fi
    echo "bench $W n=$n rc=$?" | tee -a $OUT/summary.txt
    tail -1 $OUT/bench_${W}_n$n.json | cut -c1-400
    tail -3 $OUT/bench_${W}_n$n.err
  done
done
