"""CPU tests of the host layer: schema registry -> fn.* generation, graph construction, argument validation, the Resize
size arithmetic against the reference's own known-answer vectors, and the 'no CPU fallback' contract."""
import ctypes as C

import numpy as np
import pytest

from dali_b200 import backend, fn, types, pipeline_def, Pipeline
from oracle import pyoracle as po


def test_schema_registry_and_fn_names():
    names = set(backend.schema_names())
    assert {"decoders__Image", "Resize", "CropMirrorNormalize", "WarpAffine", "Hsv", "ColorSpaceConversion", "Spectrogram",
            "MelFilterBank"} <= names
    # ops/_names.py:24-72 naming
    for f in (fn.decoders.image, fn.resize, fn.crop_mirror_normalize, fn.warp_affine, fn.hsv, fn.color_space_conversion, fn.spectrogram,
              fn.mel_filter_bank, fn.external_source):
        assert callable(f)
    assert fn.crop_mirror_normalize.schema_name == "CropMirrorNormalize"
    # registered for GPU / mixed only: the product has no CPU implementation of the hot-path ops
    assert backend.operator_registered("decoders__Image", "mixed") and not backend.operator_registered("decoders__Image", "cpu")
    for s in ("Resize", "CropMirrorNormalize", "WarpAffine", "Hsv", "ColorSpaceConversion", "Spectrogram", "MelFilterBank"):
        assert backend.operator_registered(s, "gpu") and not backend.operator_registered(s, "cpu")
    args = backend.schema_args("Resize")
    assert args["resize_x"][0] and "antialias" in args and "minibatch_size" in args
    assert backend.schema_args("ColorSpaceConversion")["image_type"][1]       # required argument


def test_nvidia_dali_alias_imports():
    import nvidia.dali as dali
    from nvidia.dali import fn as nfn, types as ntypes, pipeline_def as npd  # noqa: F401
    from nvidia.dali.plugin.pytorch import DALIGenericIterator  # noqa: F401
    assert nfn.resize is fn.resize and ntypes.FLOAT16 == 8 and dali.Pipeline is Pipeline


def test_graph_construction_and_errors():
    @pipeline_def(batch_size=8, num_threads=2, device_id=None)
    def pipe():
        j = fn.external_source(source=lambda i: [np.zeros(4, np.uint8)] * 8, name="jpegs")
        img = fn.decoders.image(j, device="mixed")
        img = fn.resize(img, resize_x=224, resize_y=224)
        return fn.crop_mirror_normalize(img, dtype=types.FLOAT16, output_layout="CHW", mean=[1, 2, 3], std=[4, 5, 6])
    p = pipe()
    with pytest.raises(backend.BackendError, match="needs a GPU"):
        p.build()
    with pytest.raises(RuntimeError, match="inside a pipeline definition"):
        fn.resize(None)

    @pipeline_def(batch_size=2, num_threads=1, device_id=None)
    def bad_arg():
        j = fn.external_source(source=lambda i: [np.zeros((4, 4, 3), np.uint8)] * 2, device="gpu")
        return fn.resize(j, resize_q=3)
    with pytest.raises(TypeError, match="unexpected 'resize_q' argument"):
        bad_arg().build()

    @pipeline_def(batch_size=2, num_threads=1, device_id=None)
    def cpu_op():
        j = fn.external_source(source=lambda i: [np.zeros((4, 4, 3), np.uint8)] * 2)
        return fn.resize(j, resize_x=2, device="cpu")
    with pytest.raises(backend.BackendError, match="no CPU fallback"):
        cpu_op().build()


def _resize_params(mode, req, lo, hi, subpixel=True, max_size=None):
    lib = backend.lib()
    f2 = lambda v: (C.c_float * 2)(*v)
    dst, olo, ohi = (C.c_int * 2)(), (C.c_float * 2)(), (C.c_float * 2)()
    ms = f2(max_size) if max_size is not None else None
    assert lib.dalihTestResizeParams(mode, f2(req), f2(lo), f2(hi), int(subpixel), ms, dst, olo, ohi) == 0
    return list(dst), list(olo), list(ohi)


DEFAULT, STRETCH, NOT_LARGER, NOT_SMALLER = 0, 1, 2, 3


def test_resize_size_rules_reference_kats():
    """dali/operators/image/resize/resize_attr_test.cc (2-D cases) and test_resize.py:303-315 (H, W order)."""
    # resize_x=480 (:97-113): (768,1024) -> {360,480}; (320,240) -> {640,480}
    assert _resize_params(DEFAULT, (0, 480), (0, 0), (768, 1024))[0] == [360, 480]
    assert _resize_params(DEFAULT, (0, 480), (0, 0), (320, 240))[0] == [640, 480]
    # resize_y=480 (:114-129): -> {480,640}, {480,360}
    assert _resize_params(DEFAULT, (480, 0), (0, 0), (768, 1024))[0] == [480, 640]
    assert _resize_params(DEFAULT, (480, 0), (0, 0), (320, 240))[0] == [480, 360]
    # resize_shorter=600 (:425-454): (400,800) -> {600,1200}; (500,250) -> {1200,600}; with max_size [800,1000] -> {500,1000}, {800,400}
    assert _resize_params(NOT_SMALLER, (600, 600), (0, 0), (400, 800))[0] == [600, 1200]
    assert _resize_params(NOT_SMALLER, (600, 600), (0, 0), (500, 250))[0] == [1200, 600]
    assert _resize_params(NOT_SMALLER, (600, 600), (0, 0), (400, 800), max_size=(800, 1000))[0] == [500, 1000]
    assert _resize_params(NOT_SMALLER, (600, 600), (0, 0), (500, 250), max_size=(800, 1000))[0] == [800, 400]
    # resize_longer=600 (:456-475): (400,800) -> {300,600}; (500,250) -> {600,300}
    assert _resize_params(NOT_LARGER, (600, 600), (0, 0), (400, 800))[0] == [300, 600]
    assert _resize_params(NOT_LARGER, (600, 600), (0, 0), (500, 250))[0] == [600, 300]
    # python mirror (test_resize.py:303-315), (W,H) there -> (H,W) here
    assert _resize_params(NOT_SMALLER, (600, 600), (0, 0), (480, 640), max_size=(720, 720))[0] == [540, 720]
    assert _resize_params(DEFAULT, (0, 600), (0, 0), (480, 640))[0] == [450, 600]
    assert _resize_params(DEFAULT, (600, 0), (0, 0), (480, 640))[0] == [600, 800]
    # ROI + flip (:477-515): resize_x=200, resize_y=-100, roi (7,200)-(330,40) on (400,800) -> dst {100,200}, lo {330,200}, hi {7,40}
    dst, lo, hi = _resize_params(DEFAULT, (-100, 200), (7, 200), (330, 40))
    assert dst == [100, 200] and lo == [330.0, 200.0] and hi == [7.0, 40.0]
    # C2: resize_x = resize_y = 224 on 1080p
    assert _resize_params(DEFAULT, (224, 224), (0, 0), (1080, 1920)) == ([224, 224], [0.0, 0.0], [1080.0, 1920.0])


def test_resize_subpixel_roi_adjustment():
    """resize_attr_base.h:97-113: when the rounded size differs from the requested fractional size the ROI is scaled
    about its centre.  (:368-423, last axes of the not_smaller / max_size case: dst 384x288 from 320x240.)"""
    dst, lo, hi = _resize_params(NOT_SMALLER, (480, 600), (0, 0), (320, 240), max_size=(384, 400))
    assert dst == [384, 288] and lo == [0.0, 0.0] and hi == [320.0, 240.0]
    dst, lo, hi = _resize_params(DEFAULT, (0, 100.4), (0, 0), (30, 40))      # height 75.3 -> 75: ROI shrinks about the centre
    assert dst == [75, 100]
    assert lo[0] > 0 and abs((hi[0] - lo[0]) - 30 * 75 / 75.3) < 1e-3 and abs((lo[0] + hi[0]) / 2 - 15) < 1e-4


def _resize_params3(mode, req, hi, lo=(0, 0, 0), subpixel=True, max_size=None):
    lib = backend.lib()
    f3 = lambda v: (C.c_float * 3)(*v)
    dst, olo, ohi = (C.c_int * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    ms = f3(max_size) if max_size is not None else None
    assert lib.dalihTestResizeParams3D(mode, f3(req), f3(lo), f3(hi), int(subpixel), ms, dst, olo, ohi) == 0
    return list(dst), list(olo), list(ohi)


def test_resize_size_rules_for_volumes_reference_kats():
    """dali/operators/image/resize/resize_attr_test.cc, the Resize3D* cases (D, H, W order)."""
    import math
    # resize_z=480 alone (:129-143): the other extents follow the depth scale
    assert _resize_params3(DEFAULT, (480, 0, 0), (512, 768, 1024))[0] == [480, 720, 960]
    assert _resize_params3(DEFAULT, (480, 0, 0), (400, 320, 240))[0] == [480, 384, 288]
    # separate arguments (:146-172)
    assert _resize_params3(DEFAULT, (140, 130, 120), (123, 234, 345)) == ([140, 130, 120], [0.0, 0.0, 0.0], [123.0, 234.0, 345.0])
    # missing y, default mode: geometric mean of the two scales; the ROI of the rounded axis shrinks about its centre (:174-229)
    for (z, shape) in ((140, (123, 234, 345)), (150, (321, 432, 543))):
        h = shape[1]
        sub = h * math.sqrt(z / shape[0] * 120 / shape[2])
        y = int(round(sub))
        dst, lo, hi = _resize_params3(DEFAULT, (z, 0, 120), shape)
        assert dst == [z, y, 120]
        c = h * 0.5
        assert lo[0] == 0 and lo[2] == 0 and hi[0] == shape[0] and hi[2] == shape[2]
        assert abs(lo[1] - (c - c * y / sub)) < 1e-3 and abs(hi[1] - (c + c * y / sub)) < 1e-3
        dst, lo, hi = _resize_params3(DEFAULT, (z, 0, 120), shape, subpixel=False)
        assert dst == [z, y, 120] and lo == [0.0, 0.0, 0.0] and hi == [float(v) for v in shape]
        # stretch: the missing extent is left alone (:231-258)
        assert _resize_params3(STRETCH, (z, 0, 120), shape)[0] == [z, h, 120]
    # flips (:303-333): negative sizes swap lo and hi
    y0 = int(round(234 * math.sqrt(140.0 / 123 * 120 / 345)))
    dst, lo, hi = _resize_params3(DEFAULT, (-140, 0, 120), (123, 234, 345), subpixel=False)
    assert dst == [140, y0, 120] and lo == [123.0, 0.0, 0.0] and hi == [0.0, 234.0, 345.0]
    dst, lo, hi = _resize_params3(STRETCH, (150, 0, -120), (321, 432, 543), subpixel=False)
    assert dst == [150, 432, 120] and lo == [0.0, 0.0, 543.0] and hi == [321.0, 432.0, 0.0]
    # not_larger (:335-366)
    assert _resize_params3(NOT_LARGER, (0, 400, 500), (1536, 768, 1024))[0] == [750, 375, 500]
    assert _resize_params3(NOT_LARGER, (0, 400, 500), (32, 320, 240))[0] == [40, 400, 300]
    assert _resize_params3(NOT_LARGER, (600, 400, 500), (1536, 768, 1024))[0] == [600, 300, 400]
    assert _resize_params3(NOT_LARGER, (600, 400, 500), (32, 320, 240))[0] == [40, 400, 300]
    # not_smaller (:368-410), with max_size
    assert _resize_params3(NOT_SMALLER, (0, 480, 600), (1536, 768, 1024))[0] == [960, 480, 640]
    assert _resize_params3(NOT_SMALLER, (0, 480, 600), (32, 320, 240))[0] == [80, 800, 600]
    assert _resize_params3(NOT_SMALLER, (160, 480, 600), (32, 320, 240))[0] == [160, 1600, 1200]
    assert _resize_params3(NOT_SMALLER, (160, 480, 600), (1536, 768, 1024), max_size=(720, 720, 720))[0] == [720, 360, 480]
    assert _resize_params3(NOT_SMALLER, (160, 480, 600), (32, 320, 240), max_size=(720, 720, 720))[0] == [72, 720, 540]
    # per-axis max_size (:410-423): the depth of the second sample rounds 38.4 -> 38 and its ROI shrinks about the centre
    dst, lo, hi = _resize_params3(NOT_SMALLER, (160, 480, 600), (32, 320, 240), max_size=(720, 384, 400))
    assert dst == [38, 384, 288] and lo[1:] == [0.0, 0.0] and hi[1:] == [320.0, 240.0]
    assert abs(lo[0] - 0.166667) < 1e-5 and abs(hi[0] - 31.83333) < 1e-4


def test_resize_layout_parsing_reference_kats():
    """resize_attr_test.cc:22-51: layout -> (spatial_ndim, first_spatial_dim); layouts outside the schema are rejected."""
    lib = backend.lib()
    def parse(l):
        sd, fs = C.c_int(), C.c_int()
        return (sd.value, fs.value) if lib.dalihTestResizeLayout(l.encode(), C.byref(sd), C.byref(fs)) == 0 else None
    want = {"HWC": (2, 0), "CHW": (2, 1), "DHWC": (3, 0), "CDHW": (3, 1), "FHWC": (2, 1), "FCHW": (2, 2), "FDHWC": (3, 1), "FCDHW": (3, 2),
            "CFHW": (2, 2), "CFDHW": (3, 2)}
    for l, v in want.items():
        assert parse(l) == v, l
    for l in ("HCW", "FWCH", "HW", "", "DHW"):
        assert parse(l) is None, l


def test_external_source_feeding_modes():
    calls = []

    def per_sample(info):
        calls.append((info.idx_in_epoch, info.idx_in_batch, info.iteration))
        return np.full((2, 2, 3), info.idx_in_batch, np.uint8)
    with Pipeline(batch_size=4, num_threads=1, device_id=None) as p:
        a = fn.external_source(source=per_sample, batch=False)
        b = fn.external_source(source=iter([[np.zeros(3, np.float32)] * 4] * 2), name="b")
        p.set_outputs(a, b)
    p.build()
    outs = p.run()
    assert [o.shape for o in (outs[0].at(0), outs[1].at(3))] == [(2, 2, 3), (3,)]
    assert outs[0].at(2)[0, 0, 0] == 2 and calls[:4] == [(0, 0, 0), (1, 1, 0), (2, 2, 0), (3, 3, 0)]
    p.run()
    assert calls[4] == (4, 0, 1)
    with pytest.raises(StopIteration):
        p.run()


@pytest.mark.skipif(not po.have_ref(), reason="needs oracle/_ref")
def test_rotate_params_match_reference_transform_code():
    """fn.rotate: canvas size (RotatedCanvasSize + parity vote) and the 2x3 matrix equal the reference's own
    translation * rotation2D * translation (rotate_params.h:36-55,222-236,279-297), bit for bit."""
    import ctypes as C
    from dali_b200 import backend
    h = backend.lib()
    rng = np.random.default_rng(3)
    for t in range(400):
        ang = float(rng.uniform(-400, 400)) if t > 8 else [0.0, 90.0, -90.0, 180.0, 45.0, -45.0, 30.0, 270.0, 360.0][t]
        H, W = int(rng.integers(1, 900)), int(rng.integers(1, 900))
        mode = t % 3
        size = (float(rng.uniform(1, 500)), float(rng.uniform(1, 500))) if mode == 1 else None
        hw = (C.c_int * 2)()
        M = np.empty(6, np.float32)
        sz = None if size is None else (C.c_float * 2)(*size)
        h.dalihTestRotateParams(C.c_float(ang), H, W, int(mode == 2), sz, hw, M.ctypes.data_as(C.c_void_p))
        want_hw, want_M = po.ref_rotate_params(ang, H, W, keep_size=mode == 2, size=size)
        assert (hw[0], hw[1]) == want_hw, (ang, H, W, mode)
        assert np.array_equal(M.view(np.uint32), want_M.reshape(-1).view(np.uint32)), (ang, H, W, mode)


@pytest.mark.skipif(not po.have_ref(), reason="needs oracle/_ref")
def test_random_crop_generator_matches_reference():
    """decoders.image_random_crop / random_resized_crop draw the reference's windows: Philox4x32-10 with RandomCropAttr's per-sample
    states feeding the same libstdc++ distributions (random_crop_generator_util.cc, random_crop_attr.h:36-72)."""
    import ctypes as C
    from dali_b200 import backend
    h = backend.lib()
    rng = np.random.default_rng(1)
    for t in range(200):
        seed, si = int(rng.integers(0, 2 ** 31)), int(rng.integers(0, 300))
        H, W = int(rng.integers(1, 2000)), int(rng.integers(1, 2000))
        ar, area, na = sorted(rng.uniform(0.2, 3.0, 2)), sorted(rng.uniform(0.01, 1.0, 2)), int(rng.integers(1, 12))
        a = (C.c_int * 40)()
        h.dalihTestRandomCrop(C.c_int64(seed), si, H, W, C.c_float(ar[0]), C.c_float(ar[1]), C.c_float(area[0]), C.c_float(area[1]), na, 10, a)
        want = po.ref_random_crop(seed, si, H, W, ar, area, na, 10)
        assert [tuple(a[4 * k:4 * k + 4]) for k in range(10)] == want, (seed, si, H, W)


def _host():
    import ctypes as C
    from dali_b200 import backend
    return backend.lib(), C


def test_crop_window_known_answers():
    """CropAttr::CalculateAnchor (crop_attr.cc:224-239): anchor = round(pos * (in - crop)) with halves away from zero, or truncated
    with rounding="truncate"; a zero crop extent keeps the whole axis."""
    lib, C = _host()

    def win(ch, cw, py, px, H, W, trunc=False):
        out = (C.c_int64 * 4)()
        assert lib.dalihTestCropWindow(C.c_float(ch), C.c_float(cw), C.c_float(py), C.c_float(px), int(trunc), H, W, out) == 0
        return tuple(out)
    assert win(224, 224, 0.5, 0.5, 256, 341) == (16, 59, 224, 224)           # 0.5 * 117 = 58.5 -> 59 (half away from zero)
    assert win(224, 224, 0.5, 0.5, 256, 341, trunc=True) == (16, 58, 224, 224)
    assert win(0, 100, 0.3, 1.0, 50, 200) == (0, 100, 50, 100)                # crop_h = 0: the whole axis, anchor 0
    assert win(10, 10, 0.0, 0.25, 11, 13) == (0, 1, 10, 10)                   # round(0.25 * 3) = 1
    assert win(3, 5, 1.0, 0.0, 3, 5) == (0, 0, 3, 5)


def test_slice_window_known_answers():
    """slice_attr.h (NamedSliceAttr): start / end from absolute or relative arguments in `axis_names` order ("WH"), rounded with
    std::llround; rel_start + rel_shape are summed BEFORE the multiplication by the extent."""
    lib, C = _host()

    def sl(mode, a, b, H, W):
        out = (C.c_int64 * 4)()
        assert lib.dalihTestSliceWindow(mode, (C.c_float * 2)(*a), (C.c_float * 2)(*b), H, W, out) == 0
        return tuple(out)                                                     # (y0, y1, x0, x1)
    assert sl(0, (10, 20), (30, 40), 100, 200) == (20, 60, 10, 40)            # start + shape
    assert sl(1, (0.1, 0.25), (0.5, 0.5), 100, 200) == (25, 75, 20, 120)      # rel_start, rel_shape
    assert sl(2, (10, 20), (60, 70), 100, 200) == (20, 70, 10, 60)            # start, end
    # float32(0.255) * 100 = 25.4999995 -> 25 and float32(0.755) * 100 = 75.4999995 -> 75 (the products are formed in double from the
    # float arguments, as the reference does); 0.105f * 200 = 21.00000016 -> 21
    assert sl(3, (0.105, 0.255), (0.6, 0.755), 100, 200) == (25, 75, 21, 120)
    assert sl(1, (0.5, 0.5), (0.5, 0.5), 7, 9) == (4, 7, 5, 9)                # llround(3.5) = 4, llround(4.5) = 5: halves away from zero
