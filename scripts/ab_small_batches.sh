# small batches: lanes per chain x producer warps
run() {  # lanes ws reps
  PB200_LANES_PER_CHAIN=$1 PB200_WS=$2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --reps-per-gpu $3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lanes=$1 ws=$2 chains=$3', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"
}
for reps in 512 1024 2048 3072; do for l in 32 16 8; do run $l 0 $reps; run $l 1 $reps; done; done
