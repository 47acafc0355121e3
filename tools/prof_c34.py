"""Profiling driver (used under ncu): one launch of the C3 (video) chain on 256 frames of 720p and of the C4 (audio) chain on 64 clips,
the latter also through the optional tensor-core mel path.   python tools/prof_c34.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dali_b200 import capi
from dali_b200.hotpath import VideoPipelineC3, AudioPipelineC4
from oracle import pyoracle as po

nseq, flen, fh, fw = 16, 16, 720, 1280
nfr = nseq * flen
rng = np.random.default_rng(3)
g = torch.Generator(device="cuda"); g.manual_seed(3)
frames = torch.randint(0, 256, (nfr, fh, fw, 3), dtype=torch.uint8, device="cuda", generator=g)
inv, hsvp, mir = [], [], []
for q in range(nseq):
    ang, sc = np.deg2rad(rng.uniform(-10, 10)), rng.uniform(0.95, 1.05)
    cx, cy = fw / 2, fh / 2
    a, b = sc * np.cos(ang), sc * np.sin(ang)
    m = po.affine_inv(np.array([[a, -b, cx - a * cx + b * cy], [b, a, cy - b * cx - a * cy]], np.float32))
    inv += [m] * flen; hsvp += [(rng.uniform(-30, 30), rng.uniform(0.7, 1.3), rng.uniform(0.8, 1.2))] * flen; mir += [int(rng.integers(0, 2))] * flen
v = VideoPipelineC3(nfr, (fh, fw))
v.setup(inv, hsvp, mir)
for _ in range(2):
    v.launch(frames)
torch.cuda.synchronize()
del v, frames
nclip, clen = 64, 160000
clips = torch.from_numpy(np.random.default_rng(4).uniform(-1, 1, (nclip, clen)).astype(np.float32)).cuda()
au = AudioPipelineC4(nclip, clen)
for _ in range(2):
    au.launch(clips)
capi.check(capi.lib().dalib200MelPlanSetTensorCores(au.mel.handle, 1))
for _ in range(2):
    au.launch(clips)
torch.cuda.synchronize()
print("ok")
