#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest.log
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/c3_band.log
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tag, env in (("band+tma", {}), ("band", {"DALIB200_WARP_NO_TMA": "1"}), ("generic", {"DALIB200_WARP_GENERIC": "1"})):
    for k in ("DALIB200_WARP_NO_TMA", "DALIB200_WARP_GENERIC"): os.environ.pop(k, None)
    os.environ.update(env)
    r = bench.secondary_workloads(7000.0, flush, 10, 3)
    c3 = r["c3_video"]
    print(tag, "C3", round(c3["value"]), "frames/s", round(c3["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in c3["kernels_ms"].items()},
          "mismatch", c3["parity_mismatching_elements"])
PY
