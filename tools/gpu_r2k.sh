#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest.log
timeout 900 python - <<'PY' 2>&1 | tee gpurun_out/c3_tma.log
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for tag, env in (("tma", None), ("generic", "1")):
    if env: os.environ["DALIB200_WARP_NO_TMA"] = env
    else: os.environ.pop("DALIB200_WARP_NO_TMA", None)
    r = bench.secondary_workloads(7000.0, flush, 10, 3)
    c3 = r["c3_video"]
    print(tag, "C3", round(c3["value"]), "frames/s", round(c3["ms_per_step"], 3), "ms", {k: round(v, 3) for k, v in c3["kernels_ms"].items()},
          "mismatch", c3["parity_mismatching_elements"])
    if tag == "tma":
        c4 = r["c4_audio"]
        print("C4", round(c4["value"]), c4["unit"], round(c4["ms_per_step"], 3), {k: round(v, 3) for k, v in c4["kernels_ms"].items()})
PY
timeout 600 python - <<'PY' 2>&1 | tail -20 | tee gpurun_out/e2e_host.log
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from dali_b200 import fn, types, pipeline_def, capi
from dali_b200.hotpath import IMAGENET_MEAN, IMAGENET_STD
batch = 256
raw = bench.make_batch(batch, 0, 16)
arena = capi.pinned_empty(sum((s.size + 63) & ~63 for s in raw))
off, pinned = 0, []
for s in raw:
    v = arena[off:off + s.size]; v[:] = s; pinned.append(v); off += (s.size + 63) & ~63
mirror = [np.array(m, np.int32) for m in np.random.default_rng(0).integers(0, 2, batch)]
for streams, nc, tag in ((raw, False, "pageable"), (pinned, True, "pinned no_copy")):
  for depth in (2, 3, 4):
    @pipeline_def(batch_size=batch, num_threads=8, device_id=0, prefetch_queue_depth=depth)
    def c2():
        jpegs = fn.external_source(source=lambda i: streams, name="jpegs", no_copy=nc)
        mir = fn.external_source(source=lambda i: mirror, name="mirror")
        img = fn.decoders.image(jpegs, device="mixed", output_type=types.RGB)
        img = fn.resize(img, resize_x=224, resize_y=224)
        return fn.crop_mirror_normalize(img, dtype=types.FLOAT16, output_layout="CHW", crop=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD, mirror=mir)
    p = c2(); p.build()
    def step():
        (out,) = p.run()
        t = torch.as_tensor(out.as_tensor(), device="cuda")
        return float(t[:, 0, 0, 0].float().sum().item())
    for _ in range(3): chk = step()
    t0 = time.perf_counter()
    for _ in range(30): chk = step()
    dt = (time.perf_counter() - t0) / 30
    print(f"{tag} depth {depth}: {dt*1e3:.3f} ms/step  {batch/dt:.0f} img/s  chk {chk:.3f}", flush=True)
    del p
PY
