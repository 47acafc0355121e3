#!/bin/bash
mkdir -p gpurun_out
timeout 120 tools/probe/tma_probe 2>&1 | tee gpurun_out/tma_probe.log
export DALIB200_WARP_NO_TMA=1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest.log
timeout 600 python - <<'PY' 2>&1 | tail -60 | tee gpurun_out/e2e_host.log
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from dali_b200 import fn, types, pipeline_def
from dali_b200.hotpath import IMAGENET_MEAN, IMAGENET_STD
batch = 256
streams = bench.make_batch(batch, 0, 16)
mirror = [np.array(m, np.int32) for m in np.random.default_rng(0).integers(0, 2, batch)]
for depth in (2, 3, 4):
    @pipeline_def(batch_size=batch, num_threads=8, device_id=0, prefetch_queue_depth=depth)
    def c2():
        jpegs = fn.external_source(source=lambda i: streams, name="jpegs")
        mir = fn.external_source(source=lambda i: mirror, name="mirror")
        img = fn.decoders.image(jpegs, device="mixed", output_type=types.RGB)
        img = fn.resize(img, resize_x=224, resize_y=224)
        return fn.crop_mirror_normalize(img, dtype=types.FLOAT16, output_layout="CHW", crop=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD, mirror=mir)
    p = c2(); p.build()
    def step():
        (out,) = p.run()
        t = torch.as_tensor(out.as_tensor(), device="cuda")
        return float(t[:, 0, 0, 0].float().sum().item())
    for _ in range(3): step()
    t0 = time.perf_counter()
    for _ in range(20): step()
    dt = (time.perf_counter() - t0) / 20
    print(f"depth {depth}: {dt*1e3:.3f} ms/step  {batch/dt:.0f} img/s", flush=True)
    if depth == 3:
        os.environ["DALIB200_HOST_TIMING"] = "1"
        # static flag: only effective in a fresh process; run a child below instead
    del p
PY
DALIB200_HOST_TIMING=1 timeout 300 python - <<'PY' 2>&1 | tail -40 | tee -a gpurun_out/e2e_host.log
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from dali_b200 import fn, types, pipeline_def
from dali_b200.hotpath import IMAGENET_MEAN, IMAGENET_STD
batch = 256
streams = bench.make_batch(batch, 0, 16)
mirror = [np.array(m, np.int32) for m in np.random.default_rng(0).integers(0, 2, batch)]
@pipeline_def(batch_size=batch, num_threads=8, device_id=0, prefetch_queue_depth=3)
def c2():
    jpegs = fn.external_source(source=lambda i: streams, name="jpegs")
    mir = fn.external_source(source=lambda i: mirror, name="mirror")
    img = fn.decoders.image(jpegs, device="mixed", output_type=types.RGB)
    img = fn.resize(img, resize_x=224, resize_y=224)
    return fn.crop_mirror_normalize(img, dtype=types.FLOAT16, output_layout="CHW", crop=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD, mirror=mir)
p = c2(); p.build()
for i in range(6):
    t0 = time.perf_counter()
    (out,) = p.run()
    t1 = time.perf_counter()
    t = torch.as_tensor(out.as_tensor(), device="cuda"); v = float(t[:, 0, 0, 0].float().sum().item())
    t2 = time.perf_counter()
    print(f"[py] run {1e3*(t1-t0):.3f} ms  readback {1e3*(t2-t1):.3f} ms", file=sys.stderr, flush=True)
PY
