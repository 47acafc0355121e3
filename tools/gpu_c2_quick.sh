#!/bin/bash
# tests + per-kernel timing of the C2 chain at batch 256 (10 L2-flushed launches)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest.log
timeout 600 python - <<'PY'
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from dali_b200 import capi
from dali_b200.hotpath import ImagePipelineC2
streams = bench.make_batch(256, 0, 16)
mirror = np.random.default_rng(0).integers(0, 2, 256)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
pipe = ImagePipelineC2(256)
pipe.setup(streams, mirror); pipe.upload()
for _ in range(3): out = pipe.launch()
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
capi.profiling(True); capi.profiling_collect()
for a, b in ev:
    flush.fill_(1); a.record(); out = pipe.launch(); b.record()
torch.cuda.synchronize()
prof = capi.profiling_collect(); capi.profiling(False)
agg = {}
for k, ms in prof: agg[k] = agg.get(k, 0) + ms / 10
t = float(np.mean([a.elapsed_time(b) for a, b in ev]))
print("step", round(t, 3), "ms", round(256 / t * 1e3), "img/s", {k: round(v, 3) for k, v in agg.items()}, "chk", float(out.float().sum().item()))
PY
