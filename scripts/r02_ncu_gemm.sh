#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wide_anchor|wide_materialise" -s 21 -c 2 -o gpurun_out/r02_k4_gemm \
    python bench.py --workload profile256 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log | cut -c1-200
