#!/bin/bash
mkdir -p gpurun_out/m7
for K in 10 50 200; do
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload C4 --steps $K --warmup 3 --no-parity > gpurun_out/m7/bench_C4_n2_k$K.json 2> gpurun_out/m7/err_$K.txt
tail -1 gpurun_out/m7/bench_C4_n2_k$K.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('K', d['steps'], 'ms', round(d['ms_per_step'],4), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'dens', round(d['roofline']['density_ms'],4), 'vxc', round(d['roofline']['vxc_ms'],4))"
done
