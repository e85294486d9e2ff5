#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or k4 or bit_exact" 2>&1 | tail -15
for ws in 0 1; do echo "producer warps $ws"; PB200_WIDE_WS=$ws python scripts/k4_probe.py 8x128 8x16 8x512 2>&1 | awk '/wide timing/{l=$0} /chains=/{print l; print $0}'; done
echo fused; PB200_WIDE=0 python scripts/k4_probe.py 8x128 2>&1 | tail -1
