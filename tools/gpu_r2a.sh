#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/cpu_scaling.py > gpurun_out/cpu_scaling.log 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k jpeg 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 3))
print({k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'huff_write|unstuff|huff_sync' -s 8 -c 8 -f -o gpurun_out/prof_r2a python tools/prof_c2.py 256 2 > gpurun_out/prof_r2a.log 2>&1
tail -2 gpurun_out/prof_r2a.log
cat gpurun_out/cpu_scaling.log
