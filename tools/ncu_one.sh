#!/bin/bash
# ncu --set full of the kernels matching $1 (regex) in the 2nd C2 iteration; batch $2 (default 64); output name $3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${4:-1} -c ${5:-1} -f -o gpurun_out/${3:-prof_one} python tools/prof_c2.py ${2:-64} 2 > gpurun_out/prof_one.log 2>&1
tail -2 gpurun_out/prof_one.log
