# register budget of the 8-lane kernel at issue-bound batch sizes.  Variant libraries first, e.g.
#   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -Iinclude \
#        -DPB200_MIN_CTAS_8=5 prospect_public_b200/csrc/pb200_kernels.cu -o prospect_public_b200/_lib/r8_5.so
run() {  # lib reps
  PB200_LIB=$PWD/prospect_public_b200/_lib/$1.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --reps-per-gpu $2 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$1 chains=$2', 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'value %.3e'%d['value'])
"
}
for f in libpb200 r8_5 r8_6; do run $f 8192; run $f 65536; done
