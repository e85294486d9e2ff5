#!/bin/bash
# shipped lanes choice per batch size (chain-steps/s of the 30-D global-optimisation job)
run() { name=$1; reps=$2; shift 2
  env "$@" python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-subrecords --jobs-per-step 4 --reps-per-gpu $reps 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$name chains $reps', '%.4g' % j['value'], 'kernel_ms %.4f' % j['roofline']['kernel_ms'], flush=True)"; }
for reps in 512 1024 1536 2048 2560 3072 4096; do
  run shipped $reps
  for l in 8 16 32; do run lanes$l $reps PB200_LANES_PER_CHAIN=$l; done
done
