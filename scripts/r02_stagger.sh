#!/bin/bash
# Experiment: start the second warp of every scheduler late (PB200_STAGGER_CYCLES), 30-D global optimisation.
# The macro (a clock64 spin at the top of anneal_kernel for warps with %warpid bit 2 set) was removed after the measurement;
# libraries were built with -DPB200_STAGGER_CYCLES=... into prospect_public_b200/_lib/ab/ (PB200_LIB picks one).
mkdir -p gpurun_out
run() { name=$1; reps=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-subrecords --jobs-per-step 4 --reps-per-gpu $reps 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$name chains $reps', '%.4g' % j['value'], 'kernel_ms %.4f' % j['roofline']['kernel_ms'], flush=True)"; }
AB=$PWD/prospect_public_b200/_lib/ab
for reps in ${SIZES:-4096}; do
  run shipped $reps
  for v in ${VARIANTS:-s450 s900 s1400 w4s900}; do run $v $reps PB200_LIB=$AB/libpb200_$v.so; done
done 2>&1 | tee gpurun_out/stagger_ab.log
