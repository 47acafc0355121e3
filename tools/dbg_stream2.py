import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dali_b200.hotpath import ImagePipelineC2
from oracle import pyoracle as po
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
streams = bench.make_batch(N, 0, 16)
mirror = np.zeros(N, np.int64)
pipe = ImagePipelineC2(N)
for rep in range(4):
    pipe.run(streams, mirror)
    torch.cuda.synchronize()
    bad = 0
    for i in range(0, N, 1):
        dec = pipe.decoded(i).cpu().numpy()
        got = pipe.resized(i).cpu().numpy()
        if rep == 0 or True:
            want = po.resample(dec, (224, 224))
        d = np.argwhere(got != want)
        if len(d):
            bad += 1
            if bad <= 8:
                ys, xs = np.unique(d[:, 0]), np.unique(d[:, 1])
                print("rep", rep, "img", i, "n", len(d), "rows", ys[:12], "cols", xs[:12], "diff", (got.astype(int) - want.astype(int))[tuple(d[0])], d[:4].tolist())
    print("rep", rep, "bad images", bad, flush=True)
