#!/bin/bash
# Launch list (device time per kernel) of the K4 bench step on the wide path.
mkdir -p gpurun_out
python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 140 --csv --log-file gpurun_out/r02_k4_launches.csv \
    python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
tail -3 gpurun_out/ncu_launches.log | cut -c1-300
