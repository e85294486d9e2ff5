# producer-warp kernel: parity first, then timing of CTA shapes and batch sizes
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
run() {  # lib ws reps
  PB200_WS=$2 PB200_LIB=$PWD/prospect_public_b200/_lib/$1.so timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --reps-per-gpu $3 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 ws=$2 chains=$3', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"
}
for reps in 4096 8192 16384 32768; do run libpb200 0 $reps; run libpb200 1 $reps; done
for f in ws_p1 ws_p2_i0 ws_p4 ws_p4_i0; do run $f 1 4096; run $f 1 16384; done
PB200_LANES_PER_CHAIN=16 run libpb200 1 4096
PB200_LANES_PER_CHAIN=32 run libpb200 1 4096
PB200_LANES_PER_CHAIN=32 run libpb200 1 1024
PB200_LANES_PER_CHAIN=32 run libpb200 0 1024
