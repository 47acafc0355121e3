#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_ops_f4.py tests/test_gpu_jpeg.py tests/test_plugin_cpu.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'], 3), 'parity_mismatch', (d.get('cpu_baseline') or {}).get('parity_mismatching_elements'))
print({k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
print('cpu', {k: v for k, v in (d.get('cpu_baseline') or {}).items() if k in ('value', 'cores', 'workers', 'single_core_value')})
PY
