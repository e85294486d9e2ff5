for f in ab_w1 ab_w2 ab_w4; do for cfg in "4096 16" "4096 32" "8192 8" "65536 8"; do set -- $cfg
  PB200_LIB=$PWD/prospect_public_b200/_lib/$f.so PB200_LANES_PER_CHAIN=$2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --reps-per-gpu $1 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$f reps $1 lanes $2 kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"; done; done
