"""Decoder determinism / parity stress: decode the C2 batch several times, compare every image with cv2.imdecode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, cv2
import concurrent.futures as cf
import bench
from dali_b200.hotpath import ImagePipelineC2
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
streams = bench.make_batch(N, 0, 16)
cv2.setNumThreads(1)
with cf.ThreadPoolExecutor(16) as ex:
    want = list(ex.map(lambda s: cv2.imdecode(s, cv2.IMREAD_COLOR)[..., ::-1], streams))
mirror = np.zeros(N, np.int64)
pipe = ImagePipelineC2(N)
for rep in range(reps):
    pipe.run(streams, mirror)
    torch.cuda.synchronize()
    bad = 0
    for i in range(N):
        got = pipe.decoded(i).cpu().numpy()
        d = np.argwhere(got != want[i])
        if len(d):
            bad += 1
            if bad <= 5:
                print("rep", rep, "img", i, "n", len(d), "rows", d[:, 0].min(), d[:, 0].max(), "cols", d[:, 1].min(), d[:, 1].max(), flush=True)
    print("rep", rep, "bad images", bad, "status", sum(pipe.status()), flush=True)
