#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jpeg.py -m gpu -x -q 2>&1 | tail -4
for cfg in "1 3" "0 3" "1 2" "1 4"; do timeout 300 python tools/e2e_diag.py $cfg 2>&1 | tail -1; done
DALIB200_NO_MEMCPY_BATCH=1 timeout 300 python tools/e2e_diag.py 1 3 2>&1 | tail -1
DALIB200_HOST_TIMING=1 timeout 300 python tools/e2e_diag.py 1 3 2>&1 | grep "host timing" | tail -6
