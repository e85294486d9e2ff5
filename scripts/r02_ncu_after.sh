#!/bin/bash
# "after" captures: the three kernels of the wide path (K4) and the MH kernel (K5), --set full, one launch each.
mkdir -p gpurun_out
python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wide_steps|wide_materialise|wide_anchor" -s 31 -c 3 -o gpurun_out/r02_k4_after \
    python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k4_after.log 2>&1
python scripts/mcmc_probe.py > gpurun_out/plain_mh.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 1 -c 1 -o gpurun_out/r02_k5_mh \
    python scripts/mcmc_probe.py > gpurun_out/ncu_k5.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
