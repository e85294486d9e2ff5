#!/bin/bash
run() { name=$1; shift
  env "$@" python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-subrecords --jobs-per-step 4 2>/dev/null | tail -1 | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('$name', '%.4g' % j['value'], 'kernel_ms %.4f' % j['roofline']['kernel_ms'])"; }
run default
for v in pipe4 pipe2 base2; do run $v PB200_LIB=$PWD/prospect_public_b200/_lib/ab/libpb200_$v.so; done
