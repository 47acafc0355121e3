import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
from dali_b200 import capi
from dali_b200.hotpath import ImagePipelineC2
N = 64
streams = bench.make_batch(N, 0, 16)
pipe = ImagePipelineC2(N)
pipe.run(streams, np.zeros(N, np.int64)); torch.cuda.synchronize()
out = (C.c_ulonglong * 40)()
capi.lib().dalib200JpegDebugHuffStats(out)
v = list(out)
print("ctas-rounds", v[32], "chains per compacted round:", v[:20])
