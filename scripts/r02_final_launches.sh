#!/bin/bash
# Launch list (ncu, device time per kernel) of the bench headline command on the FINAL build of round 2.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-subrecords --no-cpu-baseline --jobs-per-step 2"
timeout 120 $CMD > gpurun_out/final_plain.log 2>&1 &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_global30_launches.csv \
    $CMD > gpurun_out/final_ncu.log 2>&1
echo rc=$?
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/r02_final_global30_launches.csv') if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
k, v = hdr.index('Kernel Name'), hdr.index('Metric Value')
tot = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[k].split('(')[0][:70]
    tot[name][0] += 1; tot[name][1] += float(r[v].replace(',', ''))
s = sum(t for _, t in tot.values())
for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f'{t / 1e3:10.1f} us {100 * t / s:5.1f} %  x{n:<4d} {name}')
PY
