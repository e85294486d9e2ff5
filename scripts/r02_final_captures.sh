#!/bin/bash
# Final round-2 captures for profiles/: the bench line, launch list + full capture of the bench headline kernel, and the
# K5 kernel (in-block box test with the TF32 mma.sync filter).  One ncu at a time, each after its plain command exited 0.
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err
python bench.py --steps 2 --warmup 3 --no-subrecords --no-cpu-baseline --jobs-per-step 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_global30_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-subrecords --no-cpu-baseline --jobs-per-step 2 > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 2 --warmup 3 --no-subrecords --no-cpu-baseline --jobs-per-step 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:anneal_kernel -s 6 -c 1 -o gpurun_out/r02_global30 \
    python bench.py --steps 2 --warmup 3 --no-subrecords --no-cpu-baseline --jobs-per-step 2 > gpurun_out/ncu_f.log 2>&1
python scripts/k5_probe2.py > gpurun_out/plain_k5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mh_kernel -s 1 -c 1 -o gpurun_out/r02_k5_final \
    python scripts/k5_probe2.py > gpurun_out/ncu_k5b.log 2>&1
mkdir -p profiles
python scripts/ncu_summary.py gpurun_out/r02_global30_launches.csv gpurun_out/r02_global30.ncu-rep r02_global30_lanes8 20480000 > /dev/null
cp profiles/r02_global30_lanes8.json gpurun_out/; cp gpurun_out/r02_global30_launches.csv gpurun_out/r02_global30_launches_copy.csv
python scripts/ncu_kernels.py gpurun_out/r02_k5_final.ncu-rep gpurun_out/r02_final_k5_mh_kernel.json > /dev/null
rm -f gpurun_out/r02_k5_final.ncu-rep gpurun_out/r02_global30.ncu-rep     # (the merge-back limit is 64 MiB; the summaries are what is kept)
tail -c 300 gpurun_out/r02_bench_final.err; cat gpurun_out/plain_k5.log | tail -2; ls -la gpurun_out | tail -8
