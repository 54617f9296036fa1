#!/bin/bash
# Usage: bash tools/gpu_multi3.sh <tag> <ngpus> [workload...]  (under gpurun --gpus N): one N-rank bench line per workload
# (each with its parity record: reference CUDA, 1-GPU recomputation, rank agreement), bounded by its own timeout.
set -u
TAG=${1:-m}; N=${2:-2}; shift; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.csv 2>&1
timeout 120 python tools/canary.py > $OUT/canary.log 2>&1 || { echo "canary failed" | tee -a $OUT/summary.txt; exit 1; }
for W in "$@"; do
  NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
     bench.py --gpus $N --workload $W --steps 10 --warmup 3 > $OUT/bench_${W}_n$N.json 2> $OUT/bench_${W}_n$N.err
  echo "bench $W n=$N rc=$?" | tee -a $OUT/summary.txt
  tail -1 $OUT/bench_${W}_n$N.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'n', d['n_gpus'], 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value'],2), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'dens', d['roofline'].get('density_ms'), 'vxc', d['roofline'].get('vxc_ms'), 'parity ok', d['parity']['ok'])"
  tail -2 $OUT/bench_${W}_n$N.err
done
