import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import gpu_helpers as g
from oracle import pyoracle as po
n, H, W, fill = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), (None if sys.argv[4] == "none" else float(sys.argv[4]))
rng = np.random.default_rng(1)
imgs = [rng.integers(0, 256, (H, W, 3)).astype(np.uint8) for _ in range(n)]
mats = [np.float32([[1, 0.03, 0.4 + k], [-0.03, 1, 0.3]]) for k in range(n)]
got, path = g.warp_affine(imgs, mats, None, 1, fill, np.uint8, contiguous=True, want_path=True)
ok = all(np.array_equal(o, po.warp_affine(im, M, None, 1, fill, np.uint8)) for im, M, o in zip(imgs, mats, got))
print("n", n, H, W, fill, "path", path, "ok", ok)
