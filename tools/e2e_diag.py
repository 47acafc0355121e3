"""e2e diagnosis: python tools/e2e_diag.py [readback 0|1] [depth]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dali_b200 import fn, types, pipeline_def, capi
from dali_b200.hotpath import IMAGENET_MEAN, IMAGENET_STD
readback, depth = int(sys.argv[1]), int(sys.argv[2])
batch = 256
raw = bench.make_batch(batch, 0, 16)
arena = capi.pinned_empty(sum((s.size + 63) & ~63 for s in raw))
off, streams = 0, []
for s in raw:
    v = arena[off:off + s.size]; v[:] = s; streams.append(v); off += (s.size + 63) & ~63
mirror = [np.array(m, np.int32) for m in np.random.default_rng(0).integers(0, 2, batch)]
@pipeline_def(batch_size=batch, num_threads=8, device_id=0, prefetch_queue_depth=depth)
def c2():
    jpegs = fn.external_source(source=lambda i: streams, name="jpegs", no_copy=True)
    mir = fn.external_source(source=lambda i: mirror, name="mirror")
    img = fn.decoders.image(jpegs, device="mixed", output_type=types.RGB)
    img = fn.resize(img, resize_x=224, resize_y=224)
    return fn.crop_mirror_normalize(img, dtype=types.FLOAT16, output_layout="CHW", crop=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD, mirror=mir)
p = c2(); p.build()
tr, tb = [], []
def step():
    t0 = time.perf_counter()
    (out,) = p.run()
    t1 = time.perf_counter()
    if readback:
        t = torch.as_tensor(out.as_tensor(), device="cuda")
        v = float(t[:, 0, 0, 0].float().sum().item())
    t2 = time.perf_counter()
    tr.append(t1 - t0); tb.append(t2 - t1)
for _ in range(4): step()
tr.clear(); tb.clear()
torch.cuda.synchronize()
T0 = time.perf_counter()
for _ in range(40): step()
torch.cuda.synchronize()
dt = (time.perf_counter() - T0) / 40
print(f"readback={readback} depth={depth} batchcopy={'off' if os.environ.get('DALIB200_NO_MEMCPY_BATCH') else 'on'}: {dt*1e3:.3f} ms/step {batch/dt:.0f} img/s   "
      f"run() {1e3*np.mean(tr):.3f} ms  readback {1e3*np.mean(tb):.3f} ms", flush=True)
