import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gpu_helpers as g
from oracle import pyoracle as po
rng = np.random.default_rng(78)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 56
imgs = [rng.integers(0, 256, (1080, 1920, 3)).astype(np.uint8) for _ in range(N)]
want = [po.resample(im, (224, 224)) for im in imgs]
for rep in range(4):
    got, paths = g.resample(imgs, [(224, 224)] * N, want_path=True)
    bad = 0
    print('paths', sum(paths), len(paths))
    for i, (o, w) in enumerate(zip(got, want)):
        d = np.argwhere(o != w)
        if len(d):
            bad += 1
            if bad <= 6:
                ys, xs = np.unique(d[:, 0]), np.unique(d[:, 1])
                print("rep", rep, "img", i, "n", len(d), "rows", ys[:12], "cols", xs.min(), xs.max(), len(xs), "diff", (o.astype(int) - w.astype(int))[tuple(d[0])], d[:4].tolist())
    print("rep", rep, "bad images", bad)
