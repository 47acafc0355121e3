#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_pipeline.py tests/test_gpu_ops_f4.py tests/test_gpu_jpeg.py tests/test_plugin_cpu.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'resample_planar' -s 1 -c 1 -f -o gpurun_out/prof_planar python tools/prof_c2.py 256 2 > gpurun_out/prof_planar.log 2>&1
tail -2 gpurun_out/prof_planar.log
