"""Generates tests/golden/resample3d_ref.npz -- run HERE (container with /root/reference), commit the output.

Source of truth: oracle/_ref/libdali_ref_cpu.so = the reference's own SeparableResampleCPU<Out, In, 3>
(dali/kernels/imgproc/resample/separable_cpu.h) compiled in place (oracle/Makefile `make ref`, shim oracle/ref_shim.cc).
Usage:  python tests/golden/make_resample3d_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402

L, T, G, CU, LZ, NN = po.F_LINEAR, po.F_TRIANGULAR, po.F_GAUSSIAN, po.F_CUBIC, po.F_LANCZOS3, po.F_NN
CASES = [   # (in DHWC, out DHW, min filters [z, y, x], mag filters, roi)
    ((12, 20, 24, 1), (6, 10, 12), [(L, 1, 0)] * 3, [(L, 0, 0)] * 3, None),                    # exact 2x: rounding ties
    ((9, 14, 33, 3), (5, 9, 17), [(L, 1, 0)] * 3, [(L, 0, 0)] * 3, None),
    ((6, 7, 8, 2), (11, 13, 40), [(L, 1, 0)] * 3, [(L, 0, 0)] * 3, None),                      # upscale
    ((10, 12, 16, 1), (4, 20, 9), [(CU, 1, 0), (T, 1, 0), (G, 1, 0)], [(LZ, 0, 0), (CU, 0, 0), (L, 0, 0)], None),
    ((8, 9, 10, 4), (8, 5, 21), [(L, 1, 0)] * 3, [(CU, 0, 0)] * 3, None),                      # depth untouched
    ((11, 13, 15, 3), (5, 6, 7), [(L, 1, 0)] * 3, [(L, 0, 0)] * 3, ((1.5, 11.0, 2.25), (9.0, 1.5, 14.5))),   # ROI, y flipped
    ((7, 8, 9, 3), (3, 12, 20), [(NN, 0, 0)] * 3, [(NN, 0, 0)] * 3, None),                     # pure nearest neighbour
    ((7, 10, 12, 1), (4, 5, 30), [(NN, 0, 0), (L, 1, 0), (L, 1, 0)], [(NN, 0, 0), (L, 0, 0), (CU, 0, 0)], None),   # NN pass among FIR passes
]


def main():
    po.build(ref=True)
    assert po.have_ref(), "needs /root/reference to build oracle/_ref"
    rng = np.random.default_rng(4321)
    g = {}
    for i, (shape, out, fmin, fmag, roi) in enumerate(CASES):
        vol = rng.integers(0, 256, shape).astype(np.uint8)
        out8, order = po.ref_resample3d(vol, out, fmin, fmag, np.uint8, roi, want_order=True)
        outf = po.ref_resample3d(vol, out, fmin, fmag, np.float32, roi)
        volf = (vol.astype(np.float32) * 1.37 - 20).astype(np.float32)
        outff = po.ref_resample3d(volf, out, fmin, fmag, np.float32, roi)
        g[f"in_{i}"] = vol
        g[f"out_u8_{i}"], g[f"out_f32_{i}"], g[f"out_f32f32_{i}"] = out8, outf, outff
        g[f"meta_{i}"] = np.array(list(out) + [v for f in fmin for v in f[:2]] + [v for f in fmag for v in f[:2]] + order, np.int32)
        g[f"roi_{i}"] = np.array(list(roi[0]) + list(roi[1]) if roi else [np.nan] * 6, np.float64)
    np.savez_compressed(os.path.join(HERE, "resample3d_ref.npz"), **g)
    print("wrote", len(CASES), "cases")


if __name__ == "__main__":
    main()
