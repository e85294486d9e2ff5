for f in ab_5_4_7 ab_4_3_6 ab_5_3_7 ab_4_4_5; do for cfg in "4096 32" "4096 16" "4096 8" "65536 16" "65536 8"; do set -- $cfg
  PB200_LIB=$PWD/prospect_public_b200/_lib/$f.so PB200_LANES_PER_CHAIN=$2 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --reps-per-gpu $1 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$f reps $1 lanes $2 kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"; done; done
PB200_LIB=$PWD/prospect_public_b200/_lib/ab_5_4_7.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
