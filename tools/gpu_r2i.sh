#!/bin/bash
mkdir -p gpurun_out
for a in "1 720 1280 0" "1 720 1280 none" "2 360 640 0" "5 360 640 none" "1 96 448 0"; do
  timeout 120 python tools/dbg_warp_tma.py $a 2>&1 | tail -2
done
timeout 300 compute-sanitizer --tool memcheck python tools/dbg_warp_tma.py 2 360 640 0 2>&1 | grep -v "^$" | head -60 | tee gpurun_out/sanitizer.log
