#!/bin/bash
# Wide-path bring-up: parity tests first (bounded), then the K4 bench in the three modes.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or k4" 2>&1 | tail -25 > gpurun_out/r02_wide_tests.log
for mode in simt mma fused; do
  case $mode in simt) export PB200_WIDE=1 PB200_WIDE_MMA=0;; mma) export PB200_WIDE=1 PB200_WIDE_MMA=1;; fused) export PB200_WIDE=0;; esac
  timeout 300 python bench.py --workload profile256 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_k4_$mode.log 2>&1
done
cat gpurun_out/r02_wide_tests.log
for mode in simt mma fused; do tail -c 600 gpurun_out/r02_k4_$mode.log; echo; done
