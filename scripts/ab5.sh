python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for st in 0 1; do for cfg in "4096 8" "4096 16" "4096 32" "16384 8" "65536 8"; do set -- $cfg
  PB200_STREAMED=$st PB200_LANES_PER_CHAIN=$2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --reps-per-gpu $1 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('streamed $st reps $1 lanes $2 ms_per_step', round(d['ms_per_step'],3), 'value %.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])
"; done; done
