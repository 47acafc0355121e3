#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
for k in resample_stream color_fast huff_sync_intra idct_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/r2_src_$k python tools/prof_c2.py 256 2 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
