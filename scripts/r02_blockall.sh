#!/bin/bash
# A/B: four-steps-per-reduction (block_step) for the 16- and 32-lane geometries vs the shipped choice
run() { name=$1; reps=$2; shift 2
  env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-subrecords --jobs-per-step 4 --reps-per-gpu $reps 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$name chains $reps', '%.4g' % j['value'], 'kernel_ms %.4f' % j['roofline']['kernel_ms'], flush=True)"; }
AB=$PWD/prospect_public_b200/_lib/ab/libpb200_blockall.so
for reps in 1024 2048 4096 8192; do
  run shipped $reps
  for l in 16 32; do
    run one_step_l$l $reps PB200_LANES_PER_CHAIN=$l
    run block_l$l $reps PB200_LANES_PER_CHAIN=$l PB200_LIB=$AB
  done
done
