#!/bin/bash
# Full GPU check: parity + e2e tests, smoke, bench line (stdout of each kept under gpurun_out/).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 2>&1 | tee gpurun_out/bench_latest.json | tail -2
