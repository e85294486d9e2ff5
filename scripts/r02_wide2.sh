#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or k4" 2>&1 | tail -15 > gpurun_out/r02_wide_tests.log
for ws in 0 1; do
  PB200_WIDE_WS=$ws timeout 300 python bench.py --workload profile256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k4_ws$ws.log 2>&1
done
cat gpurun_out/r02_wide_tests.log
for ws in 0 1; do python - <<PY
import json
l=[x for x in open('gpurun_out/r02_k4_ws$ws.log') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('ws$ws', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'])
else: print(open('gpurun_out/r02_k4_ws$ws.log').read()[-800:])
PY
done
python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 140 --csv --log-file gpurun_out/r02_k4_launches2.csv \
    python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
