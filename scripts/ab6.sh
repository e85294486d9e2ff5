for f in ab_m8_4 ab_m8_3 ab_m8_2; do for cfg in "4096 8" "8192 8" "65536 8"; do set -- $cfg
  PB200_STREAMED=0 PB200_LIB=$PWD/prospect_public_b200/_lib/$f.so PB200_LANES_PER_CHAIN=$2 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --reps-per-gpu $1 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('$f reps $1 lanes $2 kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"; done; done
