#!/bin/bash
for v in NO_RNG NO_STEPS; do echo $v; PB200_WIDE_WS=1 PB200_LIB=$PWD/prospect_public_b200/_lib/ab/libpb200_$v.so python scripts/k4_probe.py 8x16 8x128 2>&1 | awk '/wide timing/{l=$0} /chains=/{print l; print $0}'; done
