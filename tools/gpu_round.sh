#!/bin/bash
# One GPU call: parity tests, bench, ncu launch list, ncu full capture of the C2 kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_c2.py 256 2 > gpurun_out/prof_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'_kernel' -s 12 -c 12 -f -o gpurun_out/prof_full python tools/prof_c2.py 64 2 > gpurun_out/prof_full.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/bench.json | head -c 600
