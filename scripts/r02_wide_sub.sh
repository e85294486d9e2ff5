#!/bin/bash
# K4: sub-batches of profile points on parallel streams (PB200_WIDE_SUB), 8 x 128 and neighbours
for sub in 1 2 4 8; do
  echo "== PB200_WIDE_SUB=$sub"
  PB200_WIDE_TIMING=0 PB200_WIDE_SUB=$sub python scripts/k4_probe.py 8x32 8x128 8x512 16x64 2>&1 | grep "chains="
done
