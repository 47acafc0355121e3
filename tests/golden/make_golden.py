"""Generates tests/golden/*.npz -- run HERE (container with /root/reference), commit the output.

Sources of truth:
  * JPEG: cv2.imdecode (OpenCV 4.13 / libjpeg-turbo 3.1.2), the stand-in for the reference's
    nvimgcodec libjpeg-turbo CPU backend.
  * everything else: oracle/_ref/libdali_ref_cpu.so = the reference's own CPU kernels compiled in place
    (oracle/Makefile `make ref`).
Usage:  python tests/golden/make_golden.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pyoracle as po  # noqa: E402


def synth_image(h, w, seed):
    """SURVEY.md 8(d) C2 recipe: bicubic-upsampled low-frequency noise + sigma=5 gaussian noise."""
    r = np.random.default_rng(seed)
    lo = r.uniform(0, 255, (max(2, h // 32), max(2, w // 32), 3)).astype(np.float32)
    img = cv2.resize(lo, (w, h), interpolation=cv2.INTER_CUBIC) + r.normal(0, 5, (h, w, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    po.build(ref=True)
    assert po.have_ref(), "needs /root/reference to build oracle/_ref"
    rng = np.random.default_rng(1234)
    # ---------------- JPEG
    jp = {}
    cases = [(48, 64, "420", 90, 0), (33, 47, "420", 75, 0), (40, 40, "444", 95, 0), (31, 50, "422", 60, 0),
             (64, 48, "420", 85, 2), (17, 23, "440", 80, 0), (24, 56, "411", 70, 0), (20, 20, "gray", 90, 0)]
    ss = {"420": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, "444": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444,
          "422": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, "440": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440,
          "411": cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411}
    for i, (h, w, s, q, rst) in enumerate(cases):
        img = synth_image(h, w, 100 + i)
        params = [cv2.IMWRITE_JPEG_QUALITY, q]
        if s == "gray":
            img = img[..., 0]
        else:
            params += [cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss[s]]
        if rst:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
        ok, enc = cv2.imencode(".jpg", img, params)
        dec = cv2.imdecode(enc, cv2.IMREAD_COLOR)[..., ::-1]
        jp[f"enc_{i}"] = enc.ravel()
        jp[f"dec_{i}"] = np.ascontiguousarray(dec)
    np.savez_compressed(os.path.join(HERE, "jpeg_cv2.npz"), **jp)
    # ---------------- resample (reference CPU kernel)
    rs = {}
    rcases = [
        ((54, 96, 3), (11, 11), (po.F_LINEAR, 1, 0), (po.F_LINEAR, 1, 0), None),     # C2 scale, triangular antialias
        ((48, 64, 3), (22, 22), (po.F_LINEAR, 1, 0), (po.F_LINEAR, 1, 0), None),     # C1 scale
        ((30, 40, 3), (15, 20), (po.F_LINEAR, 1, 0), (po.F_LINEAR, 1, 0), None),     # exact 2x -> rounding ties
        ((20, 31, 3), (45, 50), (po.F_LINEAR, 1, 0), (po.F_LINEAR, 0, 0), None),     # upscale, linear
        ((33, 47, 1), (17, 60), (po.F_CUBIC, 1, 0), (po.F_LANCZOS3, 0, 0), None),
        ((40, 40, 4), (13, 29), (po.F_GAUSSIAN, 1, 0), (po.F_CUBIC, 0, 0), None),
        ((40, 52, 3), (16, 16), (po.F_LINEAR, 1, 0), (po.F_LINEAR, 0, 0), ((30.5, 4.25), (3.0, 47.75))),  # ROI + flip y
        ((25, 25, 3), (9, 31), (po.F_NN, 0, 0), (po.F_NN, 0, 0), None),
    ]
    for i, (shape, out_hw, fmin, fmag, roi) in enumerate(rcases):
        img = rng.integers(0, 256, shape).astype(np.uint8)
        out8, order = po.ref_resample(img, out_hw, fmin, fmag, np.uint8, roi, want_order=True)
        outf = po.ref_resample(img, out_hw, fmin, fmag, np.float32, roi)
        rs[f"in_{i}"] = img
        rs[f"out_u8_{i}"] = out8
        rs[f"out_f32_{i}"] = outf
        rs[f"meta_{i}"] = np.array([out_hw[0], out_hw[1], fmin[0], fmin[1], fmag[0], fmag[1], order], np.int32)
        rs[f"roi_{i}"] = np.array([roi[0][0], roi[0][1], roi[1][0], roi[1][1]] if roi else [np.nan] * 4, np.float64)
    np.savez_compressed(os.path.join(HERE, "resample_ref.npz"), **rs)
    # ---------------- CMN / half (reference CPU kernel + half_float)
    cm = {}
    mean, inv = po.cmn_norm_args([0.485 * 255, 0.456 * 255, 0.406 * 255], [0.229 * 255, 0.224 * 255, 0.225 * 255])
    cm["mean"], cm["inv_std"] = mean, inv
    ccases = [((22, 22, 3), (0, 0), (22, 22), 0, None, None), ((22, 30, 3), (0, 4), (22, 22), 1, None, None),
              ((17, 33, 3), (3, 5), (9, 16), 1, 4, [0, 0, 0, 42]), ((10, 10, 3), (-2, -3), (15, 14), 0, None, [1.0, 2.0, 3.0])]
    for i, (shape, anchor, crop, mirror, padc, fill) in enumerate(ccases):
        img = (np.arange(np.prod(shape)) % 256).astype(np.uint8).reshape(shape) if i == 0 else rng.integers(0, 256, shape).astype(np.uint8)
        cm[f"in_{i}"] = img
        cm[f"args_{i}"] = np.array([anchor[0], anchor[1], crop[0], crop[1], mirror, padc or 0], np.int32)
        cm[f"fill_{i}"] = np.array(fill if fill is not None else [], np.float32)
        for dt, nm in ((np.float32, "f32"), (np.float16, "f16")):
            for layout in ("CHW", "HWC"):
                cm[f"out_{nm}_{layout}_{i}"] = po.ref_cmn(img, anchor, crop, mirror, mean, inv, dt, layout, padc, fill)
    hx = np.concatenate([rng.normal(0, 2, 500).astype(np.float32),
                         (np.arange(0, 256, dtype=np.uint32) * 65536 + 0x3F800000 + 0x1000).view(np.float32),   # exact ties
                         np.float32([0, -0.0, 1e-8, 6e-8, 5.96e-8, 3e-8, 1e-5, 65504, 65520, 70000, -70000])])
    cm["half_in"] = hx
    cm["half_out"] = po.ref_float2half(hx).view(np.uint16)
    np.savez_compressed(os.path.join(HERE, "cmn_ref.npz"), **cm)
    # ---------------- warp / hsv / csc
    wp = {}
    img = rng.integers(0, 256, (40, 300, 3)).astype(np.uint8)     # > 256 wide: exercises the coordinate re-anchoring
    wp["in"] = img
    ang, s = 0.12, 1.03
    M = np.float32([[s * np.cos(ang), -s * np.sin(ang), 3.5], [s * np.sin(ang), s * np.cos(ang), -2.25]])
    wp["M"] = M
    wp["Minv"] = po.affine_inv(M, use_ref=True)
    for interp in (0, 1):
        for fill, fn in ((None, "clamp"), (0.0, "fill0")):
            wp[f"out_{interp}_{fn}"] = po.ref_warp_affine(img, M, None, interp, fill, np.uint8)
    hs = [(0, 1, 1), (25.0, 1.2, 0.9), (-30.0, 0.7, 1.2), (170.0, 1.9, 0.3)]
    wp["hsv_args"] = np.float32(hs)
    for i, (h, sa, v) in enumerate(hs):
        Mh, Th = po.color_twist_matrix(h, sa, v, use_ref=True)
        wp[f"hsv_M_{i}"] = Mh
        wp[f"hsv_out_{i}"] = po.linear_transform(img, Mh, Th, np.uint8, use_ref=True)
    cube = np.stack(np.meshgrid(np.arange(0, 256, 15), np.arange(0, 256, 17), np.arange(0, 256, 13), indexing="ij"), -1)
    cube = cube.reshape(-1, 1, 3).astype(np.uint8)
    wp["csc_in"] = cube
    wp["csc_rgb2ycbcr"] = po.csc(cube, po.IT_RGB, po.IT_YCBCR, use_ref=True)
    wp["csc_ycbcr2rgb"] = po.csc(cube, po.IT_YCBCR, po.IT_RGB, use_ref=True)
    wp["csc_ycbcr2gray"] = po.csc(cube, po.IT_YCBCR, po.IT_GRAY, use_ref=True)
    wp["csc_rgb2gray_cv2"] = cv2.cvtColor(cube, cv2.COLOR_RGB2GRAY)[..., None]
    np.savez_compressed(os.path.join(HERE, "warp_color_ref.npz"), **wp)
    # ---------------- audio
    au = {}
    t = np.arange(4000) / 16000.0
    sig = (0.5 * np.sin(2 * np.pi * 440 * t) + 0.3 * np.sin(2 * np.pi * 1234.5 * t) + 0.05 * rng.normal(0, 1, t.size)).astype(np.float32)
    au["sig"] = sig
    au["hann512"] = po.hann_window(512, use_ref=True)
    au["windows_512_256_reflect"] = po.extract_windows(sig, au["hann512"], 512, 256, True, True, use_ref=True)
    au["windows_512_256_zero"] = po.extract_windows(sig, au["hann512"], 512, 256, True, False, use_ref=True)
    au["windows_400_160_nocenter"] = po.extract_windows(sig, po.hann_window(400, use_ref=True), 400, 160, False, False, use_ref=True)
    wins = au["windows_512_256_reflect"].astype(np.float64)
    buf = np.zeros((wins.shape[0], 1024)); buf[:, 256:768] = wins       # window centred in nfft (fft_cpu_impl_ffts.cc:111)
    au["spec_nfft1024_power2_float64"] = (np.abs(np.fft.rfft(buf, axis=1)) ** 2).T
    spec = au["spec_nfft1024_power2_float64"].astype(np.float32)
    au["mel_128_16k_slaney_norm"] = po.mel_filter_bank(spec, 128, 16000.0, 0.0, 8000.0, "slaney", True, use_ref=True)
    au["mel_40_16k_htk_nonorm"] = po.mel_filter_bank(spec, 40, 16000.0, 20.0, 7600.0, "htk", False, use_ref=True)
    np.savez_compressed(os.path.join(HERE, "audio_ref.npz"), **au)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
