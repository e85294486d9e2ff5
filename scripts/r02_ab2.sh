#!/bin/bash
for v in 0 200 2000; do
  echo "suspend hint $v ns, producer warps on:"; PB200_WIDE_WS=1 PB200_LIB=$PWD/prospect_public_b200/_lib/ab/libpb200_s$v.so python scripts/k4_probe.py 8x128 2>&1 | tail -2
done
echo "default build (20000 ns), producer warps on:"; PB200_WIDE_WS=1 python scripts/k4_probe.py 8x128 8x16 2>&1 | tail -4
