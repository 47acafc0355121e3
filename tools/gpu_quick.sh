#!/bin/bash
# quick GPU check: parity tests then the bench line ($1 = pytest -k filter, optional)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q ${1:+-k "$1"} 2>&1 | tail -15
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 3), 'parity_mismatch', d['cpu_baseline']['parity_mismatching_elements'])
print({k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
PY
