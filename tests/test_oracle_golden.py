"""CPU tests: the oracle (plain-C restatement) against the committed golden vectors.

Golden vectors were produced by tests/golden/make_golden.py from the reference's own CPU kernels
compiled in place (oracle/_ref) and, for JPEG, from cv2.imdecode (libjpeg-turbo).  Integer/byte
outputs must be bit-exact; fp32 outputs are compared bit-for-bit too (same operation order, no FMA).
"""
import os

import numpy as np
import pytest


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_jpeg_matches_libjpeg_turbo_golden(oracle, golden_dir):
    g = _load(golden_dir, "jpeg_cv2.npz")
    n = len([k for k in g.files if k.startswith("enc_")])
    assert n >= 8
    for i in range(n):
        dec = oracle.jpeg_decode(g[f"enc_{i}"].tobytes())
        assert np.array_equal(dec, g[f"dec_{i}"]), f"jpeg case {i}"


def test_jpeg_matches_cv2_live(oracle):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(7)
    for (h, w, ss, q, rst) in [(120, 160, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 90, 0),
                               (57, 91, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, 50, 4),
                               (64, 64, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 100, 0),
                               (250, 3, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 80, 0),
                               (1, 1, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 80, 0)]:
        lo = rng.uniform(0, 255, (max(2, h // 16), max(2, w // 16), 3)).astype(np.float32)
        img = np.clip(cv2.resize(lo, (w, h), interpolation=cv2.INTER_CUBIC) + rng.normal(0, 5, (h, w, 3)), 0, 255).astype(np.uint8)
        params = [cv2.IMWRITE_JPEG_QUALITY, q, cv2.IMWRITE_JPEG_SAMPLING_FACTOR, ss]
        if rst:
            params += [cv2.IMWRITE_JPEG_RST_INTERVAL, rst]
        ok, enc = cv2.imencode(".jpg", img, params)
        ref = cv2.imdecode(enc, cv2.IMREAD_COLOR)[..., ::-1]
        assert np.array_equal(oracle.jpeg_decode(enc.tobytes()), ref)


def test_jpeg_rejects_garbage(oracle):
    with pytest.raises(ValueError):
        oracle.jpeg_info(b"not a jpeg at all")
    with pytest.raises(ValueError):
        oracle.jpeg_info(b"\xff\xd8\xff\xd9")


def _roi(arr):
    if np.isnan(arr[0]):
        return None
    return ((float(arr[0]), float(arr[1])), (float(arr[2]), float(arr[3])))


def test_resample_golden(oracle, golden_dir):
    g = _load(golden_dir, "resample_ref.npz")
    n = len([k for k in g.files if k.startswith("in_")])
    for i in range(n):
        oh, ow, tmin, amin, tmag, amag, order = [int(v) for v in g[f"meta_{i}"]]
        roi = _roi(g[f"roi_{i}"])
        out8, o = oracle.resample(g[f"in_{i}"], (oh, ow), (tmin, amin, 0.0), (tmag, amag, 0.0), np.uint8, roi, want_order=True)
        assert o == order, f"processing order case {i}"
        assert np.array_equal(out8, g[f"out_u8_{i}"]), f"resample u8 case {i}"
        outf = oracle.resample(g[f"in_{i}"], (oh, ow), (tmin, amin, 0.0), (tmag, amag, 0.0), np.float32, roi)
        assert np.array_equal(outf.view(np.uint32), g[f"out_f32_{i}"].view(np.uint32)), f"resample f32 case {i}"


def test_resample_c2_derived_constants(oracle):
    """SURVEY.md Appendix B.4 (from params.h:40-56, resampling_setup.cc:131-192)."""
    import ctypes as C
    lib = oracle.lib()
    def q(H, W, oh, ow):
        f = (oracle.FilterDesc * 2)(oracle.FilterDesc(oracle.F_LINEAR, 1, 0), oracle.FilterDesc(oracle.F_LINEAR, 1, 0))
        use, z = (C.c_int * 2)(0, 0), (C.c_float * 2)(0, 0)
        sy, sx = C.c_int(), C.c_int()
        order = lib.oracle_resample_order(H, W, oh, ow, f, f, use, z, z, C.byref(sy), C.byref(sx))
        return order, sy.value, sx.value
    assert q(1080, 1920, 224, 224) == (1, 10, 18)      # vertical first, supports 10 (y) / 18 (x)
    assert q(480, 640, 224, 224) == (0, 5, 6)
    assert q(720, 1280, 224, 224) == (0, 7, 12)


def test_triangular_support_kat(oracle):
    """dali/kernels/imgproc/resample/resampling_impl_cpu_test.cc:58-90: in_w=479 -> w=93 => support 11."""
    import ctypes as C
    f = (oracle.FilterDesc * 2)(oracle.FilterDesc(oracle.F_TRIANGULAR, 1, 0), oracle.FilterDesc(oracle.F_TRIANGULAR, 1, 0))
    use, z = (C.c_int * 2)(0, 0), (C.c_float * 2)(0, 0)
    sy, sx = C.c_int(), C.c_int()
    oracle.lib().oracle_resample_order(479, 479, 93, 93, f, f, use, z, z, C.byref(sy), C.byref(sx))
    assert sx.value == 11 and sy.value == 11


def test_cmn_and_half_golden(oracle, golden_dir):
    g = _load(golden_dir, "cmn_ref.npz")
    mean, inv = oracle.cmn_norm_args([0.485 * 255, 0.456 * 255, 0.406 * 255], [0.229 * 255, 0.224 * 255, 0.225 * 255])
    assert np.array_equal(mean, g["mean"]) and np.array_equal(inv, g["inv_std"])
    n = len([k for k in g.files if k.startswith("in_")])
    for i in range(n):
        ay, ax, ch, cw, mirror, padc = [int(v) for v in g[f"args_{i}"]]
        fill = g[f"fill_{i}"]
        for dt, nm in ((np.float32, "f32"), (np.float16, "f16")):
            for layout in ("CHW", "HWC"):
                out = oracle.cmn(g[f"in_{i}"], (ay, ax), (ch, cw), mirror, mean, inv, dt, layout, padc or None,
                                 fill if fill.size else None)
                want = g[f"out_{nm}_{layout}_{i}"]
                assert np.array_equal(out.view(np.uint16 if dt == np.float16 else np.uint32),
                                      want.view(np.uint16 if dt == np.float16 else np.uint32)), (i, nm, layout)
    assert np.array_equal(oracle.float2half(g["half_in"]).view(np.uint16), g["half_out"])


def test_crop_anchor_rounding(oracle):
    # crop_attr.cc:224-239: round(double(pos) * (in - crop)), half away from zero; "truncate" option
    assert oracle.crop_anchor(0.5, 33, 32) == 1        # 0.5 -> 1 (half away)
    assert oracle.crop_anchor(0.5, 33, 32, truncate=True) == 0
    assert oracle.crop_anchor(0.5, 1920, 224) == 848
    assert oracle.crop_anchor(0.25, 10, 3) == 2        # 1.75 -> 2


def test_warp_color_golden(oracle, golden_dir):
    g = _load(golden_dir, "warp_color_ref.npz")
    img, M = g["in"], g["M"]
    assert np.array_equal(oracle.affine_inv(M), g["Minv"])
    for interp in (0, 1):
        for fill, fn in ((None, "clamp"), (0.0, "fill0")):
            assert np.array_equal(oracle.warp_affine(img, M, None, interp, fill), g[f"out_{interp}_{fn}"]), (interp, fn)
    for i, (h, s, v) in enumerate(g["hsv_args"]):
        Mh, Th = oracle.color_twist_matrix(float(h), float(s), float(v))
        assert np.array_equal(Mh, g[f"hsv_M_{i}"])
        assert np.array_equal(oracle.linear_transform(img, Mh, Th), g[f"hsv_out_{i}"])
    cube = g["csc_in"]
    assert np.array_equal(oracle.csc(cube, oracle.IT_RGB, oracle.IT_YCBCR), g["csc_rgb2ycbcr"])
    assert np.array_equal(oracle.csc(cube, oracle.IT_YCBCR, oracle.IT_RGB), g["csc_ycbcr2rgb"])
    assert np.array_equal(oracle.csc(cube, oracle.IT_YCBCR, oracle.IT_GRAY), g["csc_ycbcr2gray"])
    assert np.array_equal(oracle.csc(cube, oracle.IT_RGB, oracle.IT_GRAY), g["csc_rgb2gray_cv2"])
    assert np.array_equal(oracle.csc(cube, oracle.IT_RGB, oracle.IT_BGR), cube[..., ::-1])


def test_hsv_identity(oracle):
    img = np.random.default_rng(0).integers(0, 256, (16, 16, 3)).astype(np.uint8)
    assert np.array_equal(oracle.hsv(img, 0.0, 1.0, 1.0), img)


def test_audio_golden(oracle, golden_dir):
    g = _load(golden_dir, "audio_ref.npz")
    sig = g["sig"]
    assert np.array_equal(oracle.hann_window(512), g["hann512"])
    # Hann KAT: w[t] = 0.5 (1 - cos(2 pi (t + 0.5) / N))   (window_functions.h:26-33)
    t = np.arange(512)
    assert np.allclose(g["hann512"], 0.5 * (1 - np.cos(2 * np.pi * (t + 0.5) / 512)), atol=1e-7)
    assert np.array_equal(oracle.extract_windows(sig, g["hann512"], 512, 256, True, True), g["windows_512_256_reflect"])
    assert np.array_equal(oracle.extract_windows(sig, g["hann512"], 512, 256, True, False), g["windows_512_256_zero"])
    assert np.array_equal(oracle.extract_windows(sig, oracle.hann_window(400), 400, 160, False, False), g["windows_400_160_nocenter"])
    assert oracle.num_windows(160000, 1024, 256, True) == 626           # C4: 626 frames / clip
    spec = oracle.spectrogram(sig, nfft=1024, window_length=512, window_step=256, power=2)
    want = g["spec_nfft1024_power2_float64"]
    assert spec.shape == want.shape == (513, 16)
    assert np.abs(spec - want).max() <= 1e-6 * want.max()
    mag = oracle.spectrogram(sig, nfft=1024, window_length=512, window_step=256, power=1)
    assert np.abs(mag - np.sqrt(want)).max() <= 1e-6 * np.sqrt(want.max())
    tf = oracle.spectrogram(sig, nfft=1024, window_length=512, window_step=256, power=2, layout="tf")
    assert np.array_equal(tf.T, spec)
    s32 = want.astype(np.float32)
    assert np.array_equal(oracle.mel_filter_bank(s32, 128, 16000.0, 0.0, 8000.0, "slaney", True), g["mel_128_16k_slaney_norm"])
    assert np.array_equal(oracle.mel_filter_bank(s32, 40, 16000.0, 20.0, 7600.0, "htk", False), g["mel_40_16k_htk_nonorm"])
    # dense-matrix formulation (what the tensor-core GEMM computes) agrees to fp32 round-off
    W = oracle.mel_weights(513, 128, 16000.0, 0.0, 8000.0, "slaney", True)
    dense = W.astype(np.float64) @ s32.astype(np.float64)
    assert np.abs(dense - g["mel_128_16k_slaney_norm"]).max() <= 2e-6 * np.abs(dense).max()


def test_non_pow2_dft(oracle):
    rng = np.random.default_rng(3)
    sig = rng.normal(0, 1, 3000).astype(np.float32)
    spec = oracle.spectrogram(sig, nfft=400, window_length=400, window_step=160, power=2)
    wins = oracle.extract_windows(sig, oracle.hann_window(400), 400, 160, True, True).astype(np.float64)
    want = (np.abs(np.fft.rfft(wins, axis=1)) ** 2).T
    assert np.abs(spec - want).max() <= 1e-6 * want.max()
