#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest.log
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
tail -c 1500 gpurun_out/r2_bench_reference_arm.json; echo
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "median", round(d["ms_per_step_median"], 3), "e2e", round(d["e2e"]["value"]), d["e2e"].get("equals_device_resident_path"))
print("clocks", d["clocks"]); print("roofline", d["roofline"]); print("cpu", {k: d["cpu_baseline"][k] for k in ("value", "cores", "kind", "parity_mismatching_elements") if k in d["cpu_baseline"]})
print({k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
s = d["secondary"]
print("C3", round(s["c3_video"]["value"]), {k: round(v, 3) for k, v in s["c3_video"]["kernels_ms"].items()}, s["c3_video"]["parity_mismatching_elements"])
print("C4", round(s["c4_audio"]["value"]), s["c4_audio"]["kernels_ms"], s["c4_audio"]["fused_equals_two_kernel_chain_bitwise"], s["c4_audio"]["two_kernel_chain"])
PY
