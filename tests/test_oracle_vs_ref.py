"""CPU tests: the plain-C restatement against the reference's own CPU kernels (oracle/_ref), on random
inputs.  Skipped when oracle/_ref/libdali_ref_cpu.so is absent (it can only be built where
/root/reference exists; the prebuilt .so travels to the GPU box with the snapshot)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built (no /root/reference)")

FILTERS = [po.F_NN, po.F_LINEAR, po.F_TRIANGULAR, po.F_GAUSSIAN, po.F_CUBIC, po.F_LANCZOS3]


def test_resample_random_sweep():
    rng = np.random.default_rng(11)
    ncase = 0
    for it in range(250):
        H, W = rng.integers(1, 80, 2)
        C = int(rng.choice([1, 3, 4]))
        oh, ow = [int(v) for v in rng.integers(1, 64, 2)]
        if it % 7 == 0:
            oh, ow = max(1, H // 2), max(1, W // 2)      # exact 2x: exercises .5 rounding ties
        fm = (int(rng.choice(FILTERS)), int(rng.integers(0, 2)), 0.0)
        fg = (int(rng.choice(FILTERS)), int(rng.integers(0, 2)), 0.0)
        in_f32 = bool(rng.integers(0, 2))
        odt = np.float32 if in_f32 or rng.integers(0, 2) else np.uint8
        img = rng.integers(0, 256, (H, W, C)).astype(np.uint8)
        if in_f32:
            img = img.astype(np.float32) * 1.37 - 20
        roi = None
        if it % 3 == 0:
            y0, y1 = sorted(rng.uniform(0, H, 2))
            x0, x1 = sorted(rng.uniform(0, W, 2))
            if y1 - y0 < 0.5 or x1 - x0 < 0.5:
                continue
            if it % 6 == 0:
                y0, y1 = y1, y0                          # flipped ROI
            roi = ((float(y0), float(x0)), (float(y1), float(x1)))
            if fm[0] == po.F_NN or fg[0] == po.F_NN:
                continue   # reference NN pass reads past the cropped ROI (resampling_impl_cpu.h:534-571): excluded
        a, oa = po.resample(img, (oh, ow), fm, fg, odt, roi, want_order=True)
        b, ob = po.ref_resample(img, (oh, ow), fm, fg, odt, roi, want_order=True)
        assert oa == ob
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), (H, W, C, oh, ow, fm, fg, roi)
        ncase += 1
    assert ncase > 150


def test_resample_c1_c2_shapes():
    rng = np.random.default_rng(12)
    for (H, W) in [(480, 640), (1080, 1920)]:
        img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        assert np.array_equal(po.resample(img, (224, 224)), po.ref_resample(img, (224, 224)))


def test_cmn_random_sweep():
    rng = np.random.default_rng(13)
    for it in range(120):
        H, W = [int(v) for v in rng.integers(1, 50, 2)]
        img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
        ch, cw = int(rng.integers(1, H + 1)), int(rng.integers(1, W + 1))
        ay, ax = int(rng.integers(0, H - ch + 1)), int(rng.integers(0, W - cw + 1))
        pad = it % 5 == 0
        if pad:
            ay -= 2; ax -= 3; ch += 3; cw += 4
        mean, inv = po.cmn_norm_args([0.485 * 255, 0.456 * 255, 0.406 * 255], [0.229 * 255, 0.224 * 255, 0.225 * 255])
        padc = 4 if it % 3 == 0 else None
        fill = [0, 0, 0, 42] if padc else ([1.5, 2.5, 3.5] if pad else None)
        for dt in (np.float32, np.float16):
            for layout in ("CHW", "HWC"):
                kw = dict(anchor=(ay, ax), crop=(ch, cw), mirror=bool(it & 1), mean=mean, inv_std=inv, out_dtype=dt,
                          layout=layout, pad_channels=padc, fill=fill)
                a, b = po.cmn(img, **kw), po.ref_cmn(img, **kw)
                assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), kw


def test_float2half_ties_away():
    ties = (np.arange(0, 1024, dtype=np.uint32) * 8192 + 0x3F800000 + 0x1000).view(np.float32)
    x = np.concatenate([ties, -ties, np.float32([0, 1e-8, 5.96e-8, 6e-8, 65504, 65520, 1e6])])
    assert np.array_equal(po.float2half(x).view(np.uint16), po.ref_float2half(x).view(np.uint16))
    # and it is NOT numpy's ties-to-even on exact ties
    assert (po.float2half(ties).view(np.uint16) != ties.astype(np.float16).view(np.uint16)).any()


def test_warp_hsv_random():
    rng = np.random.default_rng(14)
    for it in range(25):
        H, W = [int(v) for v in rng.integers(2, 330, 2)]
        C = 3 if it % 4 else 1
        img = rng.integers(0, 256, (H, W, C)).astype(np.uint8)
        ang, s = rng.uniform(-0.5, 0.5), rng.uniform(0.7, 1.4)
        M = np.float32([[s * np.cos(ang), -s * np.sin(ang), rng.uniform(-20, 20)], [s * np.sin(ang), s * np.cos(ang), rng.uniform(-20, 20)]])
        assert np.array_equal(po.affine_inv(M), po.affine_inv(M, use_ref=True))
        for interp in (0, 1):
            for fill in (None, 42.0):
                assert np.array_equal(po.warp_affine(img, M, None, interp, fill), po.ref_warp_affine(img, M, None, interp, fill))
    for it in range(200):
        h, s, v = rng.uniform(-180, 180), rng.uniform(0, 2), rng.uniform(0, 2)
        Ma, Ta = po.color_twist_matrix(h, s, v)
        Mb, Tb = po.color_twist_matrix(h, s, v, use_ref=True)
        assert np.array_equal(Ma, Mb) and np.array_equal(Ta, Tb)
    img = rng.integers(0, 256, (31, 45, 3)).astype(np.uint8)
    M, T = po.color_twist_matrix(33.0, 1.3, 0.8)
    for dt in (np.uint8, np.float32):
        assert np.array_equal(po.linear_transform(img, M, T, dt), po.linear_transform(img, M, T, dt, use_ref=True))


def test_audio_windows_and_mel():
    rng = np.random.default_rng(15)
    sig = rng.normal(0, 0.3, 9000).astype(np.float32)
    for (wl, st, ce, rf) in [(512, 256, True, True), (512, 256, True, False), (400, 160, False, False), (1024, 256, True, True)]:
        w = po.hann_window(wl)
        assert np.array_equal(w, po.hann_window(wl, use_ref=True))
        assert np.array_equal(po.extract_windows(sig, w, wl, st, ce, rf), po.extract_windows(sig, w, wl, st, ce, rf, use_ref=True))
    spec = po.spectrogram(sig, nfft=1024, window_length=1024, window_step=256)
    for (nf, sr, fl, fh, formula, norm) in [(128, 16000, 0, 8000, "slaney", True), (80, 16000, 20, 7600, "htk", False),
                                            (64, 44100, 0, 0, "slaney", True), (40, 22050, 100, 9000, "htk", True)]:
        assert np.array_equal(po.mel_filter_bank(spec, nf, sr, fl, fh, formula, norm),
                              po.mel_filter_bank(spec, nf, sr, fl, fh, formula, norm, use_ref=True))
