#!/bin/bash
# A/B of library builds on the K4 bench: PB200_LIB picks the .so (built here with different -D flags).
mkdir -p gpurun_out
run() { # name env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload profile256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_$name.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/ab_$name.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('$name', '%.4g' % j['value'], '%.3f ms' % j['ms_per_step'])
else: print('$name FAILED', open('gpurun_out/ab_$name.log').read()[-600:])
PY
}
run ws0 PB200_WIDE_WS=0
run ws1_u8 PB200_WIDE_WS=1
for u in 2 4; do run ws1_u$u PB200_WIDE_WS=1 PB200_LIB=$PWD/prospect_public_b200/_lib/ab/libpb200_u$u.so; done
