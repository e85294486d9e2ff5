#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "wide or k4" 2>&1 | tail -5
python scripts/k4_probe.py 8x16 8x128 8x512 2>&1 | awk '/wide timing/{l=$0} /chains=/{print l; print $0}'
python bench.py --workload profile256 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('bench', j['value'], j['ms_per_step'], 'e2e', j['e2e']['value'])"
