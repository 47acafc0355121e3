import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, cv2
import gpu_helpers as g
from dali_b200 import capi
from oracle import pyoracle as po
rng = np.random.default_rng(21)
imgs = [g.synth_image(300, 420, 30), g.synth_image(211, 333, 31),
        np.full((256, 384, 3), 128, np.uint8), np.zeros((64, 1024, 3), np.uint8),
        rng.integers(0, 256, (120, 200, 3)).astype(np.uint8), rng.integers(0, 256, (64, 64, 3)).astype(np.uint8),
        g.synth_image(17, 9, 32), g.synth_image(640, 640, 33)]
streams = []
for i, im in enumerate(imgs):
    params = [cv2.IMWRITE_JPEG_QUALITY, [90, 75, 90, 50, 100, 98, 90, 85][i]]
    if i % 2 == 1: params += [cv2.IMWRITE_JPEG_OPTIMIZE, 1]
    if i == 7: params += [cv2.IMWRITE_JPEG_RST_INTERVAL, 3]
    if i == 4: params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444]
    streams.append(cv2.imencode(".jpg", im, params)[1].tobytes())
def coefs_want(s):
    comps = po.jpeg_coeffs(s); info = po.jpeg_info(s)
    hs, vs, mcux, mcuy = info["hs"], info["vs"], info["mcux"], info["mcuy"]
    blocks = []
    for my in range(mcuy):
        for mx in range(mcux):
            for c in range(info["ncomp"]):
                for v in range(vs[c]):
                    for hh in range(hs[c]):
                        blocks.append(comps[c][my * vs[c] + v, mx * hs[c] + hh])
    return np.stack(blocks).reshape(-1)
for label, order in (("single", None), ("batch", list(range(8))), ("batch-rev", [7, 5, 3, 1, 6, 4, 2, 0])):
    if order is None:
        for k, s in enumerate(streams):
            plan = capi.Plan("Jpeg", 1)
            outs, status = g.jpeg_decode([s], plan=plan)
            co = coefs_want(s); got = g.jpeg_coefs(plan, 0, co.size)
            cd = np.flatnonzero(got != co)
            print(label, k, "len", len(s), "status", status, "px", int((outs[0] != po.jpeg_decode(s)).sum()), "coef", len(cd), np.unique(cd // 64)[:6], flush=True)
    else:
        plan = capi.Plan("Jpeg", 8)
        outs, status = g.jpeg_decode([streams[k] for k in order], plan=plan)
        for pos_, (k, o) in enumerate(zip(order, outs)):
            co = coefs_want(streams[k]); got = g.jpeg_coefs(plan, pos_, co.size)
            cd = np.flatnonzero(got != co)
            print(label, k, "status", status[pos_], "px", int((o != po.jpeg_decode(streams[k])).sum()), "coef", len(cd), np.unique(cd // 64)[:6], flush=True)
