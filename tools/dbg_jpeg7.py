import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, cv2
import gpu_helpers as g
from dali_b200 import capi
from oracle import pyoracle as po
exec(open(os.path.join(os.path.dirname(__file__), "dbg_jpeg5.py")).read().split("for label, order in")[0])
s = streams[7]
b = np.frombuffer(s, np.uint8)
sos = s.find(b"\xff\xda"); hdr = int.from_bytes(s[sos+2:sos+4], "big"); start = sos + 2 + hdr
pos = [i for i in range(start, len(s) - 1) if b[i] == 0xFF and 0xD0 <= b[i+1] <= 0xD7]
bounds = [start] + [p + 2 for p in pos]; ends = pos + [len(s) - 2]
lens = [e - st for st, e in zip(bounds, ends)]
cum = np.cumsum([0] + [(l + 31) // 32 for l in lens])
print("units", len(lens), "lens[46:54]", lens[46:54], "first_subseq[46:54]", cum[46:54].tolist())
stuffed = [sum(1 for i in range(st, e - 1) if b[i] == 0xFF and b[i+1] == 0) for st, e in zip(bounds, ends)]
print("stuffed[46:54]", stuffed[46:54])
plan = capi.Plan("Jpeg", 1)
outs, status = g.jpeg_decode([s], plan=plan)
co = coefs_want(s); got = g.jpeg_coefs(plan, 0, co.size)
cd = np.flatnonzero(got != co)
print("mismatch blocks", np.unique(cd // 64).tolist()[:40])
print([(int(i // 64), int(i % 64), int(got[i]), int(co[i])) for i in cd[:24]])
