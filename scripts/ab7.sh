# timing split of the fused kernel: full vs stand-in RNG vs no chain update (experiment libraries, results meaningless)
for f in libpb200 exp_FAKE_RNG exp_FAKE_UPDATE; do
 for reps in 4096 65536; do
  PB200_LIB=$PWD/prospect_public_b200/_lib/$f.so python bench.py --steps 6 --warmup 3 --no-cpu-baseline --reps-per-gpu $reps 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$f $reps', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"
 done
done
