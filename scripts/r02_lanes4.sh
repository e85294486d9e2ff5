#!/bin/bash
# 4 lanes per chain (8 coordinates per lane, 8 chains per warp) against the shipped choice, 30-D global optimisation
mkdir -p gpurun_out
: # parity tests of the 4-lane geometry: tests/test_gpu_parity.py (lanes = 4 cases)
run() { name=$1; reps=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-subrecords --jobs-per-step 4 --reps-per-gpu $reps 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$name chains $reps', '%.4g' % j['value'], 'kernel_ms %.4f' % j['roofline']['kernel_ms'], flush=True)"; }
for reps in ${SIZES:-3072 4096 8192 16384 65536}; do
  run shipped $reps
  run lanes4 $reps PB200_LANES_PER_CHAIN=4
done 2>&1 | tee gpurun_out/lanes4_ab.log
