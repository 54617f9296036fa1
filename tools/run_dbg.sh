#!/bin/bash
# debugging session on the GPU box: a few random cases through every TMA variant
mkdir -p gpurun_out/dbg
for a in "GGA 3000 36" "LDA 3000 36" "GGA 20000 152" "B3LYP 20000 377" "LDA 20000 377" "GGA 20001 64" "GGA 10000 255" "B3LYP 5000 7" "GGA 30000 300"; do
  echo "== $a"; timeout 120 python tools/debug_case.py $a 2>&1 | tail -10
done 2>&1 | tee gpurun_out/dbg/dbg.log
