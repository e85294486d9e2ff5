#!/bin/bash
# Round-2 "before" numbers: K4 (256-D profile) with / without producer warps, K5 probe, and one ncu capture of the K4 kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_tests_before.log
python bench.py --workload profile256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k4_before.log 2>&1
PB200_WS=1 python bench.py --workload profile256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k4_before_ws.log 2>&1
python scripts/mcmc_probe.py > gpurun_out/r02_k5_before.log 2>&1
python bench.py --workload profile256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:anneal -s 2 -c 1 -o gpurun_out/r02_k4_before \
    python bench.py --workload profile256 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k4_before.log 2>&1
tail -3 gpurun_out/r02_tests_before.log gpurun_out/r02_k4_before.log gpurun_out/r02_k4_before_ws.log gpurun_out/r02_k5_before.log
