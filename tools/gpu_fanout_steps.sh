#!/bin/bash
# Single-process fan-out measurements: tools/gpu_fanout_steps.sh <tag> <devices> [workloads...]
# One bench line per (workload, fan_threads on/off) under gpurun_out/.
tag=$1; ndev=$2; shift 2
for W in "$@"; do
  for T in ${FAN_THREADS:-1 0}; do
    out=gpurun_out/${tag}_bench_${W}_devices${ndev}_threads${T}
    timeout 120 python bench.py --workload $W --devices $ndev --steps 10 --no-cpu-baseline --opt fan_threads=$T > $out.json 2> $out.err
    python - "$out.json" <<'PY'
import json, sys
l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], l["n_gpus"], "step %.3f ms" % l["ms_per_step"], "e2e %.3f ms" % l["e2e"]["ms_per_step"],
      "parity", (l["parity"] or {}).get("ok"), "dens %.3f vxc %.3f" % (l["roofline"]["density_ms"], l["roofline"]["vxc_ms"]))
PY
    tail -2 $out.err
  done
done
