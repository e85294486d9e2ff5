for reps in 4096 16384 65536; do for l in 32 16 8; do PB200_LANES_PER_CHAIN=$l python bench.py --steps 6 --warmup 3 --no-cpu-baseline --reps-per-gpu $reps 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); print('reps $reps lanes $l kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'], 'frac %.2f'%d['roofline']['frac'])
"; done; done
