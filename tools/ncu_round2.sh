#!/bin/bash
# round-2 profile set: (1) launch list of one bench step, (2) ncu --set full of every C2 kernel at batch 256, (3) C3 / C4 kernels
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_batch256.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'_kernel' -c 40 -f -o gpurun_out/r2_full_c2_batch256 python tools/prof_c2.py 256 2 > gpurun_out/prof_c2.log 2>&1
tail -2 gpurun_out/prof_c2.log
ncu --set full --clock-control none --import-source on -k regex:'_kernel' -c 24 -f -o gpurun_out/r2_full_c34 python tools/prof_c34.py > gpurun_out/prof_c34.log 2>&1
tail -2 gpurun_out/prof_c34.log
ls -la gpurun_out/*.ncu-rep
