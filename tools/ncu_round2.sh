#!/bin/bash
# round-2 profile set: (1) launch list of one bench step, (2) ncu --set full of every C2 kernel at batch 256, (3) C3 / C4 kernels.
# The .ncu-rep files are summarised ON THE BOX (gpurun_out/ is capped at 64 MiB): per-kernel JSON + text + gzipped raw CSV stay,
# one small report with source correlation is kept for the dominant kernel.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_batch256.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
ncu --set full --clock-control none -k regex:'_kernel' -c 40 -f -o /tmp/r2_full_c2_batch256 python tools/prof_c2.py 256 2 > gpurun_out/prof_c2.log 2>&1
tail -2 gpurun_out/prof_c2.log
python tools/ncu_traffic.py /tmp/r2_full_c2_batch256.ncu-rep 256 gpurun_out/r2_ncu_dram_traffic_batch256.json gpurun_out/r2_ncu_full_c2_batch256.txt > /dev/null
ncu -i /tmp/r2_full_c2_batch256.ncu-rep --page raw --csv | gzip -9 > gpurun_out/r2_ncu_full_c2_batch256.raw.csv.gz
ncu --set full --clock-control none -k regex:'_kernel' -c 24 -f -o /tmp/r2_full_c34 python tools/prof_c34.py > gpurun_out/prof_c34.log 2>&1
tail -2 gpurun_out/prof_c34.log
python tools/ncu_traffic.py /tmp/r2_full_c34.ncu-rep 0 gpurun_out/r2_ncu_c34.json gpurun_out/r2_ncu_full_c34.txt > /dev/null
ncu -i /tmp/r2_full_c34.ncu-rep --page raw --csv | gzip -9 > gpurun_out/r2_ncu_full_c34.raw.csv.gz
ncu --set full --clock-control none --import-source on -k regex:'huff_write' -s 1 -c 1 -f -o gpurun_out/r2_huff_write_batch256 python tools/prof_c2.py 256 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'resample_stream' -s 1 -c 1 -f -o gpurun_out/r2_resample_stream_batch256 python tools/prof_c2.py 256 2 > /dev/null 2>&1
ls -la gpurun_out/ | head -40
du -sh gpurun_out
