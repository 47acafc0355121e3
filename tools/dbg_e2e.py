import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from dali_b200 import fn, types, pipeline_def
from dali_b200.hotpath import IMAGENET_MEAN, IMAGENET_STD
N = 256
streams = bench.make_batch(N, 0, 16)
mirror = [np.array(m, np.int32) for m in np.random.default_rng(0).integers(0, 2, N)]
@pipeline_def(batch_size=N, num_threads=8, device_id=0)
def pipe():
    j = fn.external_source(source=lambda i: streams, name="jpegs")
    m = fn.external_source(source=lambda i: mirror, name="mirror")
    img = fn.decoders.image(j, device="mixed", output_type=types.RGB)
    img = fn.resize(img, resize_x=224, resize_y=224)
    return fn.crop_mirror_normalize(img, dtype=types.FLOAT16, output_layout="CHW", crop=(224, 224), mean=IMAGENET_MEAN, std=IMAGENET_STD, mirror=m)
p = pipe(); p.build()
for _ in range(3): p.run()
torch.cuda.synchronize()
ts, tw, tc = [], [], []
t00 = time.perf_counter()
for _ in range(10):
    t0 = time.perf_counter(); p.schedule_run() if len(p._sched) < p._depth else None; t1 = time.perf_counter()
    outs = p.share_outputs(); t2 = time.perf_counter()
    ts.append(t1 - t0); tw.append(t2 - t1)
torch.cuda.synchronize()
print("total/step ms", 1e3 * (time.perf_counter() - t00) / 10, "schedule ms", 1e3 * np.mean(ts), "wait ms", 1e3 * np.mean(tw))
# split the schedule: callbacks+feed vs backend run
slot = p._next_slot
t0 = time.perf_counter(); p._run_input_callbacks(slot); t1 = time.perf_counter(); p._slots[slot].run(); t2 = time.perf_counter()
p._slots[slot].wait()
print("callbacks+feed ms", 1e3 * (t1 - t0), "backend.run (setup+enqueue) ms", 1e3 * (t2 - t1))
# --- pure H2D of the same size, and depth-3 prefetch
x = torch.empty(129 << 20, dtype=torch.uint8).pin_memory(); y = torch.empty_like(x, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); print("H2D 129 MiB ms", 1e3 * (time.perf_counter() - t0) / 5)
for depth in (2, 3):
    p3 = pipe(prefetch_queue_depth=depth); p3.build()
    for _ in range(4): p3.run()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): p3.run()
    torch.cuda.synchronize(); print("depth", depth, "ms/step", 1e3 * (time.perf_counter() - t0) / 10)
