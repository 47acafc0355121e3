#!/bin/bash
# ncu --set full capture of every kernel of one C2 iteration (batch $1, default 64)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'_kernel' -s 11 -c 11 -f -o gpurun_out/prof_full python tools/prof_c2.py ${1:-64} 2 > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/prof_full.log
