for f in ab_lt1_c1 ab_lt1_c4 ab_lt1_c5 ab_lt0_c6 ab_lt0_c7 ab_lt0_c8; do
  PB200_LIB=$PWD/prospect_public_b200/_lib/$f.so python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$f', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'ms_per_step', round(d['ms_per_step'],3), 'value %.3e'%d['value'])
"
done
PB200_LIB=$PWD/prospect_public_b200/_lib/ab_lt0_c6.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
