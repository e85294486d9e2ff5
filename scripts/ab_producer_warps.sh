# parity, then kernel timing at several batch sizes with and without producer warps
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() {  # ws reps
  PB200_WS=$1 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --reps-per-gpu $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('ws=$1 chains=$2', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"
}
for reps in 4096 8192 65536; do run 0 $reps; run 1 $reps; done
