#!/bin/bash
# full gpu test suite + bench with secondary workloads
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 3), 'parity_mismatch', d['cpu_baseline']['parity_mismatching_elements'])
print({k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
s = d.get('secondary') or {}
for k, v in s.items():
    if isinstance(v, dict): print(k, round(v['value']), v['unit'], round(v['ms_per_step'], 3), {a: round(b, 3) for a, b in v['kernels_ms'].items()}, v.get('parity_mismatching_elements'), v.get('stft_max_abs_err_over_max'))
    else: print(k, v)
PY
