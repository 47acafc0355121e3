#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
bash tools/ncu_round2.sh
