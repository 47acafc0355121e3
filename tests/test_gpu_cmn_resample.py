"""-m gpu parity tests for CropMirrorNormalize and Resize: CUDA path (through the C-ABI) vs the oracle.
Bit-exact: u8, fp16 and fp32 outputs are compared as raw bits (same operation order as the reference
CPU backend, no FMA contraction, reference rounding rules)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import pyoracle as po  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


IMAGENET = ([0.485 * 255, 0.456 * 255, 0.406 * 255], [0.229 * 255, 0.224 * 255, 0.225 * 255])


def test_cmn_reference_shape_matrix():
    """The reference's optimised-kernel matrix (dali/test/python/operator_1/test_crop_mirror_normalize.py:982-1037):
    16 shapes x dtype x pad x mirror x crop x layout, input arange % 256, mean [0.1,0.2,0.3]."""
    import gpu_helpers as g
    shapes = [(1, 1, 3), (1, 10, 3), (1, 31, 3), (1, 32, 3), (1, 33, 3), (1, 127, 3), (1, 128, 3), (1, 129, 3),
              (1, 3071, 3), (1, 3072, 3), (1, 3073, 3), (8, 3071, 3), (8, 3072, 3), (8, 3073, 3), (1024, 1024, 3), (999, 999, 3)]
    crops = [(1.0, 0.25), (0.25, 0.25), (0.25, 1.0), (0.5, 0.75), None]
    mean, inv = po.cmn_norm_args([0.1, 0.2, 0.3], [1.0])
    for dt in (np.float32, np.float16):
        for pad in (False, True):
            for mirror in (False, True):
                for layout in ("CHW", "HWC"):
                    imgs, anchors, cr = [], [], []
                    for si, sh in enumerate(shapes):
                        img = (np.arange(np.prod(sh)) % 256).astype(np.uint8).reshape(sh)
                        c = crops[si % len(crops)]
                        ch, cw = (sh[0], sh[1]) if c is None else (max(1, int(sh[0] * c[0])), max(1, int(sh[1] * c[1])))
                        ay, ax = po.crop_anchor(0.5, sh[0], ch), po.crop_anchor(0.5, sh[1], cw)
                        imgs.append(img); anchors.append((ay, ax)); cr.append((ch, cw))
                    oc = 4 if pad else 3
                    fill = [0, 0, 0, 42] if pad else None
                    outs = g.cmn(imgs, anchors, cr, [mirror] * len(imgs), mean, inv, dt, layout, oc, fill)
                    for img, a, c, o in zip(imgs, anchors, cr, outs):
                        want = po.cmn(img, a, c, mirror, mean, inv, dt, layout, oc if pad else None, fill)
                        assert np.array_equal(bits(o), bits(want)), (img.shape, a, c, dt, pad, mirror, layout)


def test_half_conversion_exhaustive():
    """The kernels convert float -> float16 with ONE hardware conversion (round to nearest even of `bits | 1`); it must equal the
    integer restatement of the reference's half_float rounding (ties away from zero, after the +-65504 clamp) for all 2^32 floats."""
    import ctypes as C
    from dali_b200 import capi
    bad = C.c_uint64(123)
    capi.check(capi.lib().dalib200DebugCheckHalfConversion(C.byref(bad)))
    assert bad.value == 0


def test_cmn_random_windows_and_padding():
    import gpu_helpers as g
    rng = np.random.default_rng(21)
    mean, inv = po.cmn_norm_args(*IMAGENET)
    for dt in (np.float16, np.float32):
        for layout in ("CHW", "HWC"):
            imgs, anchors, crops, mirrors = [], [], [], []
            for it in range(40):
                H, W = [int(v) for v in rng.integers(1, 300, 2)]
                img = rng.integers(0, 256, (H, W, 3)).astype(np.uint8)
                ch, cw = int(rng.integers(1, H + 1)), int(rng.integers(1, W + 1))
                ay, ax = int(rng.integers(0, H - ch + 1)), int(rng.integers(0, W - cw + 1))
                if it % 5 == 0:       # window leaves the image -> out_of_bounds_policy="pad"
                    ay -= int(rng.integers(0, 6)); ax -= int(rng.integers(0, 6)); ch += 5; cw += 7
                imgs.append(img); anchors.append((ay, ax)); crops.append((ch, cw)); mirrors.append(bool(rng.integers(0, 2)))
            fill = [1.5, 2.5, 3.5]
            outs = g.cmn(imgs, anchors, crops, mirrors, mean, inv, dt, layout, None, fill)
            for img, a, c, m, o in zip(imgs, anchors, crops, mirrors, outs):
                want = po.cmn(img, a, c, m, mean, inv, dt, layout, None, fill)
                assert np.array_equal(bits(o), bits(want)), (img.shape, a, c, m, dt, layout)


def test_cmn_c2_shape_all_bytes():
    """224x224 fp16 CHW (the C2 output); every byte value against every channel's mean/std."""
    import gpu_helpers as g
    mean, inv = po.cmn_norm_args(*IMAGENET)
    img = (np.arange(224 * 224 * 3) % 256).astype(np.uint8).reshape(224, 224, 3)
    for mirror in (False, True):
        out = g.cmn([img], [(0, 0)], [(224, 224)], [mirror], mean, inv, np.float16, "CHW")[0]
        assert np.array_equal(bits(out), bits(po.cmn(img, (0, 0), (224, 224), mirror, mean, inv, np.float16, "CHW")))


def test_cmn_rejects_bad_arguments():
    from dali_b200 import capi
    plan = capi.Plan("Cmn", 2)
    s = (capi.CmnSample * 1)()
    s[0].in_h, s[0].in_w, s[0].channels = 4, 4, 3
    s[0].crop_h, s[0].crop_w = 2, 2
    assert capi.lib().dalib200CmnPlanSetup(plan.handle, 1, s, capi.UINT8, capi.LAYOUT_CHW, 3) != 0     # bad dtype
    assert b"not supported" in capi.lib().dalib200GetLastError()
    assert capi.lib().dalib200CmnPlanSetup(plan.handle, 3, s, capi.FLOAT, capi.LAYOUT_CHW, 3) != 0      # batch > capacity


FILTERS = [po.F_NN, po.F_LINEAR, po.F_TRIANGULAR, po.F_GAUSSIAN, po.F_CUBIC, po.F_LANCZOS3]


def test_resample_golden(golden_dir):
    import gpu_helpers as g
    gz = np.load(os.path.join(golden_dir, "resample_ref.npz"))
    n = len([k for k in gz.files if k.startswith("in_")])
    for i in range(n):
        oh, ow, tmin, amin, tmag, amag, order = [int(v) for v in gz[f"meta_{i}"]]
        r = gz[f"roi_{i}"]
        roi = None if np.isnan(r[0]) else ((float(r[0]), float(r[1])), (float(r[2]), float(r[3])))
        (o8,), (od,) = g.resample([gz[f"in_{i}"]], [(oh, ow)], (tmin, amin, 0.0), (tmag, amag, 0.0), np.uint8, [roi], want_order=True)
        assert od == order
        assert np.array_equal(o8, gz[f"out_u8_{i}"]), f"u8 case {i}"
        (of,) = g.resample([gz[f"in_{i}"]], [(oh, ow)], (tmin, amin, 0.0), (tmag, amag, 0.0), np.float32, [roi])
        assert np.array_equal(bits(of), bits(gz[f"out_f32_{i}"])), f"f32 case {i}"


def test_resample_random_batch_vs_oracle():
    import gpu_helpers as g
    rng = np.random.default_rng(31)
    for trial in range(12):
        fm = (int(rng.choice(FILTERS)), int(rng.integers(0, 2)), 0.0)
        fg = (int(rng.choice(FILTERS)), int(rng.integers(0, 2)), 0.0)
        in_f32 = trial % 4 == 3
        odt = np.float32 if in_f32 or trial % 2 else np.uint8
        imgs, outs, rois = [], [], []
        for it in range(16):
            H, W = [int(v) for v in rng.integers(1, 120, 2)]
            C = int(rng.choice([1, 3, 4]))
            img = rng.integers(0, 256, (H, W, C)).astype(np.uint8)
            if in_f32:
                img = img.astype(np.float32) * 1.37 - 20
            oh, ow = [int(v) for v in rng.integers(1, 90, 2)]
            if it % 5 == 0:
                oh, ow = max(1, H // 2), max(1, W // 2)
            roi = None
            if it % 3 == 0 and fm[0] != po.F_NN and fg[0] != po.F_NN:
                y0, y1 = sorted(rng.uniform(0, H, 2)); x0, x1 = sorted(rng.uniform(0, W, 2))
                if y1 - y0 >= 0.5 and x1 - x0 >= 0.5:
                    if it % 6 == 0:
                        y0, y1 = y1, y0
                    roi = ((float(y0), float(x0)), (float(y1), float(x1)))
            imgs.append(img); outs.append((oh, ow)); rois.append(roi)
        got, orders = g.resample(imgs, outs, fm, fg, odt, rois, want_order=True)
        for img, hw, roi, o, od in zip(imgs, outs, rois, got, orders):
            want, wo = po.resample(img, hw, fm, fg, odt, roi, want_order=True)
            assert od == wo
            assert np.array_equal(bits(o), bits(want)), (img.shape, hw, fm, fg, roi, odt)


@pytest.mark.parametrize("shape", [(480, 640), (720, 1280), (1080, 1920)])
def test_resample_baseline_shapes(shape):
    """C1 / C3 / C2 input sizes -> 224x224, default filters (triangular antialias), batch with mixed content."""
    import gpu_helpers as g
    rng = np.random.default_rng(41)
    imgs = [rng.integers(0, 256, shape + (3,)).astype(np.uint8) for _ in range(3)]
    got = g.resample(imgs, [(224, 224)] * 3)
    for img, o in zip(imgs, got):
        assert np.array_equal(o, po.resample(img, (224, 224)))


def test_resample_large_upscale_and_odd_sizes():
    import gpu_helpers as g
    rng = np.random.default_rng(42)
    cases = [((37, 53, 3), (301, 411)), ((1, 1, 3), (17, 9)), ((5, 700, 3), (40, 33)), ((700, 5, 1), (9, 64)), ((64, 64, 3), (64, 64))]
    imgs = [rng.integers(0, 256, c[0]).astype(np.uint8) for c in cases]
    got = g.resample(imgs, [c[1] for c in cases])
    for img, c, o in zip(imgs, cases, got):
        assert np.array_equal(o, po.resample(img, c[1])), c


def test_resample_streaming_path_vs_oracle():
    """Down-scales of 16-byte aligned rows go through the streaming (TMA ring) kernel: mixed batches, every FIR filter,
    u8 and f32 outputs, ROIs, strips/segments/chunks with ragged tails."""
    import gpu_helpers as g
    rng = np.random.default_rng(77)
    streamed = 0
    for trial in range(10):
        ftype = [po.F_TRIANGULAR, po.F_LINEAR, po.F_GAUSSIAN, po.F_CUBIC, po.F_LANCZOS3][trial % 5]
        fm = (ftype, 1, 0.0)
        fg = (po.F_LINEAR, 0, 0.0)
        odt = np.float32 if trial % 3 == 2 else np.uint8
        imgs, outs, rois = [], [], []
        for it in range(10):
            C = int(rng.choice([1, 3, 4]))
            W = 16 * int(rng.integers(4, 60)) if it % 4 else int(rng.integers(40, 900))
            H = int(rng.integers(60, 700))
            img = rng.integers(0, 256, (H, W, C)).astype(np.uint8)
            oh = int(rng.integers(3, max(4, H // 2)))
            ow = int(rng.integers(3, max(4, W // 2)))
            if it == 0:
                oh, ow = 224, 224
                img = rng.integers(0, 256, (1080, 1920, 3)).astype(np.uint8)
            roi = None
            if it % 3 == 1:
                y0, y1 = sorted(rng.uniform(0, H, 2)); x0, x1 = sorted(rng.uniform(0, W, 2))
                if y1 - y0 >= 2 * oh and x1 - x0 >= 2 * ow:
                    roi = ((float(y0), float(x0)), (float(y1), float(x1)))
            imgs.append(img); outs.append((oh, ow)); rois.append(roi)
        got, paths = g.resample(imgs, outs, fm, fg, odt, rois, want_path=True)
        streamed += sum(paths)
        for img, hw, roi, o in zip(imgs, outs, rois, got):
            want = po.resample(img, hw, fm, fg, odt, roi)
            assert np.array_equal(bits(o), bits(want)), (img.shape, hw, fm, roi, odt)
    assert streamed >= 20, streamed


def test_resample_streaming_many_items_per_cta():
    """More strips than resident CTAs: every CTA walks several work items back to back (ring and accumulator hand-over
    between items), repeated to catch ordering races."""
    import gpu_helpers as g
    rng = np.random.default_rng(78)
    imgs = [rng.integers(0, 256, (1080, 1920, 3)).astype(np.uint8) for _ in range(28)]
    want = [po.resample(im, (224, 224)) for im in imgs]
    for rep in range(3):
        got, paths = g.resample(imgs, [(224, 224)] * len(imgs), want_path=True)
        assert all(paths)
        for i, (o, w) in enumerate(zip(got, want)):
            assert np.array_equal(o, w), (rep, i, int((o != w).sum()))
