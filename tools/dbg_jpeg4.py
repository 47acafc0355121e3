import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, cv2
import gpu_helpers as g
from dali_b200 import capi
from oracle import pyoracle as po
rng = np.random.default_rng(21)
for (h, w, q, ss) in [(120, 200, 100, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444), (120, 200, 100, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420), (120, 200, 95, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444), (64, 64, 98, None)]:
    im = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    params = [cv2.IMWRITE_JPEG_QUALITY, q] + ([cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss] if ss else [])
    s = cv2.imencode(".jpg", im, params)[1].tobytes()
    plan = capi.Plan("Jpeg", 1)
    outs, status = g.jpeg_decode([s], plan=plan)
    want = po.jpeg_decode(s)
    comps = po.jpeg_coeffs(s); info = po.jpeg_info(s)
    hs, vs, mcux, mcuy = info["hs"], info["vs"], info["mcux"], info["mcuy"]
    blocks = []
    for my in range(mcuy):
        for mx in range(mcux):
            for c in range(info["ncomp"]):
                for v in range(vs[c]):
                    for hh in range(hs[c]):
                        blocks.append(comps[c][my * vs[c] + v, mx * hs[c] + hh])
    co = np.stack(blocks).reshape(-1)
    got = g.jpeg_coefs(plan, 0, co.size)
    d = np.argwhere(outs[0] != want)
    cd = np.flatnonzero(got != co)
    print((h, w, q), "len", len(s), "status", status, "pixel mismatches", len(d), "coef mismatches", len(cd), "first coef idx", cd[:5], "blocks", np.unique(cd // 64)[:8], "nblocks", co.size // 64, [(int(got[i]), int(co[i])) for i in cd[:4]])
