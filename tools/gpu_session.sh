#!/bin/bash
# Usage (on the GPU box, through gpurun): bash tools/gpu_session.sh <tag> <what...>
# Runs a bounded sequence of checks; every step has its own timeout and log under gpurun_out/<tag>/.
set -u
TAG=${1:-s}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.csv 2>&1
# every session starts with a bounded canary: a new kernel that deadlocks (e.g. a setmaxnreg.inc that can never be
# satisfied has no timeout) must cost one minute, not the whole call -- and must not be followed by more launches
timeout 120 python tools/canary.py > $OUT/canary.log 2>&1
rc=$?
echo "canary rc=$rc" | tee -a $OUT/summary.txt
tail -5 $OUT/canary.log
if [ $rc -ne 0 ]; then echo "canary failed: session aborted" | tee -a $OUT/summary.txt; exit 1; fi
for what in "$@"; do
  case $what in
    tests)   timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/tests.log 2>&1; echo "tests rc=$?" | tee -a $OUT/summary.txt; tail -15 $OUT/tests.log ;;
    pyt:*)   K=${what#pyt:}; timeout 1800 python -m pytest tests -m gpu -x -q -s -k "$K" > $OUT/pyt.log 2>&1; echo "pytest -k $K rc=$?" | tee -a $OUT/summary.txt; tail -25 $OUT/pyt.log ;;
    golden)  timeout 300 python tools/make_reference_golden.py $OUT/golden > $OUT/golden.log 2>&1; echo "golden rc=$?" | tee -a $OUT/summary.txt; tail -20 $OUT/golden.log ;;
    smoke)   timeout 300 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt; tail -5 $OUT/smoke.log ;;
    micro)   timeout 120 python -c "
from quantum_compute_dft_b200.solver import load_library
l=load_library()
print('DMMA TFLOP/s', l.DFT_MicrobenchDMMA(8192)); print('DFMA TFLOP/s', l.DFT_MicrobenchDFMA(8192))" > $OUT/micro.log 2>&1; echo "micro rc=$?" | tee -a $OUT/summary.txt; cat $OUT/micro.log ;;
    bench:*) W=${what#bench:}; timeout 600 python bench.py --workload $W --steps 5 --warmup 3 > $OUT/bench_$W.json 2> $OUT/bench_$W.err; echo "bench $W rc=$?" | tee -a $OUT/summary.txt; cat $OUT/bench_$W.json; tail -3 $OUT/bench_$W.err ;;
    benchx:*) A=${what#benchx:}; W=${A%%:*}; X=${A#*:}; timeout 600 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu-baseline $X > $OUT/benchx_${W}_$(echo $X | tr -d ' -').json 2> $OUT/benchx_$W.err; echo "benchx $W [$X] rc=$?" | tee -a $OUT/summary.txt; tail -1 $OUT/benchx_${W}_$(echo $X | tr -d ' -').json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'ms', d['ms_per_step'], 'dens', d['roofline']['density_ms'], 'vxc', d['roofline']['vxc_ms'])" ;;
    ref:*)   W=${what#ref:}; timeout 900 python bench.py --impl reference --workload $W --steps 3 --warmup 1 > $OUT/ref_$W.json 2> $OUT/ref_$W.err; echo "ref $W rc=$?" | tee -a $OUT/summary.txt; cat $OUT/ref_$W.json; tail -3 $OUT/ref_$W.err ;;
    ncu:*)   W=${what#ncu:}; CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline"
             timeout 600 $CMD > $OUT/ncu_plain_$W.log 2>&1 && \
             timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$W.csv $CMD > $OUT/ncu_list_$W.log 2>&1 && \
             timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'density_tma|vxc_tma|vxc_staged|xc_point|eval_kernel|xc_small' -c 7 -o $OUT/prof_$W $CMD > $OUT/ncu_full_$W.log 2>&1
             echo "ncu $W rc=$?" | tee -a $OUT/summary.txt; tail -3 $OUT/ncu_full_$W.log ;;
    sweep8)  SWEEP_RANKS=8 timeout 600 python tools/vxc_sweep.py C5 "dyn_sched=1" "dyn_sched=0" "dyn_sched=1,density_unit=1" "dyn_sched=0,density_unit=1" "dyn_sched=1" "dyn_sched=0" > $OUT/sweep8.txt 2>&1; echo "sweep8 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweep8.txt ;;
    sweepv)  timeout 600 python tools/vxc_sweep.py C5 "vxc_skip_mode=4" "vxc_skip_mode=1" "vxc_skip=0" "vxc_skip_mode=4,vxc_scatter=0" > $OUT/sweepv.txt 2>&1; echo "sweepv rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepv.txt ;;
    sweepv2) timeout 600 python tools/vxc_sweep.py C5 "vxc_skip_mode=1" "vxc_skip_mode=2" "vxc_skip_mode=5" "vxc_skip_mode=6" "vxc_skip_mode=5,vxc_producers=2" "vxc_skip_mode=6,vxc_producers=2" "vxc_skip_mode=5,vxc_scatter=0" "vxc_skip_mode=1" > $OUT/sweepv2.txt 2>&1; echo "sweepv2 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepv2.txt ;;
    sweepv3) timeout 600 python tools/vxc_sweep.py C5 "vxc_skip_mode=2" "vxc_skip_mode=3" "vxc_skip_mode=2,vxc_vk=16" "vxc_skip_mode=2,vxc_producers=2" "vxc_skip_mode=3,vxc_producers=2" "vxc_skip_mode=2,vxc_scatter=0" "vxc_skip_mode=1" > $OUT/sweepv3.txt 2>&1; echo "sweepv3 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepv3.txt ;;
    aotime:*) W=${what#aotime:}; timeout 600 python tools/ao_time.py $W > $OUT/aotime_$W.txt 2>&1; echo "aotime $W rc=$?" | tee -a $OUT/summary.txt; cat $OUT/aotime_$W.txt ;;
    sweepdef) timeout 600 python tools/vxc_sweep.py C5 "" "vxc_producers=1" "l2_prefetch=1" "stagger_min=100000" > $OUT/sweepdef.txt 2>&1; timeout 600 python tools/vxc_sweep.py C4 "" "vxc_producers=1" "l2_prefetch=1" >> $OUT/sweepdef.txt 2>&1; echo "sweepdef rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepdef.txt ;;
    sweepv4) timeout 600 python tools/vxc_sweep.py C5 "vxc_skip_mode=2" "vxc_skip_mode=7" "vxc_skip_mode=7,vxc_producers=1" "vxc_skip_mode=7,vxc_scatter=0" "vxc_skip_mode=2" > $OUT/sweepv4.txt 2>&1; echo "sweepv4 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepv4.txt ;;
    sweepv5) timeout 600 python tools/vxc_sweep.py C5 "" "vxc_skip_mode=1" "vxc_skip_mode=3" "vxc_skip_mode=7" "vxc_skip=0" "zero_skip=0" > $OUT/sweepv5.txt 2>&1; timeout 600 python tools/vxc_sweep.py C4 "" "vxc_skip=1" >> $OUT/sweepv5.txt 2>&1; echo "sweepv5 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepv5.txt ;;
    phase:*) A=${what#phase:}; timeout 600 python tools/phase_timing.py C5 $(echo $A | tr ',' ' ') > $OUT/phase_$(echo $A | tr -d '=,').txt 2>&1; echo "phase $A rc=$?" | tee -a $OUT/summary.txt; cat $OUT/phase_$(echo $A | tr -d '=,').txt ;;
    sweepv6) SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "vxc_rebalance=0" "vxc_skip_mode=1" "vxc_skip_mode=1,vxc_rebalance=0" "" > $OUT/sweepv6.txt 2>&1; echo "sweepv6 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepv6.txt ;;
    sweepv7) SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "vxc_prefetch=4" "vxc_prefetch=8" "vxc_prefetch=16" "vxc_prefetch=8,vxc_scatter=0" "vxc_scatter=0" > $OUT/sweepv7.txt 2>&1; echo "sweepv7 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepv7.txt ;;
    sweep8b) SWEEP_RANKS=8 SWEEP_STEPS=10 timeout 600 python tools/vxc_sweep.py C5 "" "stagger_min=100000" "stagger_min=100000,density_unit=1" "vxc_rebalance=0" "" > $OUT/sweep8b.txt 2>&1; SWEEP_RANKS=8 SWEEP_STEPS=10 timeout 600 python tools/vxc_sweep.py C4 "" "stagger_min=100000" "vxc_shape=128" "vxc_shape=64" >> $OUT/sweep8b.txt 2>&1; echo "sweep8b rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweep8b.txt ;;
    sweepd2) SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "density_producers=1" "density_producers=2,dyn_sched=0" "density_producers=2,density_unit=1" "" > $OUT/sweepd2.txt 2>&1; SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C4 "" "density_producers=1" >> $OUT/sweepd2.txt 2>&1;  SWEEP_RANKS=8 SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "density_producers=1" >> $OUT/sweepd2.txt 2>&1; echo "sweepd2 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepd2.txt ;;
    sweepd3) SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "density_scatter=0" "density_scatter=1,dyn_sched=0" "density_scatter=1,density_unit=1" "" > $OUT/sweepd3.txt 2>&1; SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C4 "" "density_scatter=0" >> $OUT/sweepd3.txt 2>&1;  SWEEP_RANKS=8 SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "density_scatter=0" >> $OUT/sweepd3.txt 2>&1; echo "sweepd3 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepd3.txt ;;
    sweepd4) SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "density_wide=1" "density_wide=1,density_producers=2" "density_wide=1,dyn_sched=0" "density_wide=1,zero_skip=0" "zero_skip=0" > $OUT/sweepd4.txt 2>&1; SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C4 "" "density_wide=1" >> $OUT/sweepd4.txt 2>&1;  SWEEP_RANKS=8 SWEEP_STEPS=8 timeout 600 python tools/vxc_sweep.py C5 "" "density_wide=1" >> $OUT/sweepd4.txt 2>&1; echo "sweepd4 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepd4.txt ;;
    sweepd)  SWEEP_LIB=diag timeout 600 python tools/vxc_sweep.py C5 "vxc_skip_mode=4" "vxc_skip_mode=4,debug_nodmma=1" "vxc_skip_mode=4,debug_nodmma=1,vxc_scatter=0" > $OUT/sweepd.txt 2>&1; echo "sweepd rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweepd.txt ;;
    sweep1)  timeout 600 python tools/vxc_sweep.py C5 "dyn_sched=1" "dyn_sched=0" > $OUT/sweep1.txt 2>&1; echo "sweep1 rc=$?" | tee -a $OUT/summary.txt; cat $OUT/sweep1.txt ;;
    *) echo "unknown step $what" ;;
  esac
done
