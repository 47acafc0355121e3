"""Profiling driver (used under ncu): runs the C2 chain over the C-ABI a few times on a small batch.
  python tools/prof_c2.py [batch] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from dali_b200.hotpath import ImagePipelineC2

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
streams = bench.make_batch(batch, 0, min(os.cpu_count(), 16))
mirror = np.random.default_rng(0).integers(0, 2, batch)
pipe = ImagePipelineC2(batch)
pipe.setup(streams, mirror)
pipe.upload()
for _ in range(iters):
    pipe.launch()
torch.cuda.synchronize()
assert all(s == 0 for s in pipe.status())
print("ok")
