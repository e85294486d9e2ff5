# parity, then kernel time of every geometry that matters: 30-D at 4096 / 65536 chains (8, 16, 32 lanes), small batch
# with producer warps, 256-D profile workload
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
run() {  # lanes reps [extra bench args]
  PB200_LANES_PER_CHAIN=$1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --reps-per-gpu $2 $3 $4 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('lanes=$1 chains=$2 $3 $4', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"
}
run 0 4096; run 0 65536; run 0 1024; run 16 4096; run 32 4096; run 16 65536; run 32 65536
run 0 0 --workload profile256
run 0 0 --workload profile30
