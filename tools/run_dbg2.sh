#!/bin/bash
mkdir -p gpurun_out/dbg2
for fn in LDA B3LYP; do for shape in "4000 377" "4000 200" "4000 152" "3000 36" "20000 377"; do for prod in 2 1; do for dyn in 1 0; do
  timeout 60 python tools/debug_density_producers.py $fn $shape $prod $dyn 2>&1 | grep -v "^\[dft_b200\] mbarrier" | tail -2
  echo "rc=$? ($fn $shape prod=$prod dyn=$dyn)"
done; done; done; done > gpurun_out/dbg2/out.txt 2>&1
