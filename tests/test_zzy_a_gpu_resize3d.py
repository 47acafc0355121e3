"""GPU parity tests of the 3-D (DHWC) resize: the CUDA path through the C-ABI against the oracle / the committed
golden vectors of the reference's SeparableResampleCPU<.., 3>, bit for bit.

Collected last on purpose: this kernel was written after the round's GPU budget was spent.  Its planner and per-element arithmetic are
pinned on the CPU (tests/test_resample3d_cpu.py runs the kernel body compiled for the host); what these tests add is the launch itself."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
L, T, G, CU, LZ, NN = po.F_LINEAR, po.F_TRIANGULAR, po.F_GAUSSIAN, po.F_CUBIC, po.F_LANCZOS3, po.F_NN


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def test_resample3d_golden(golden_dir):
    import gpu_helpers as gh
    from test_resample3d_cpu import _golden_cases
    for i, g, out, fmin, fmag, roi, order in _golden_cases(golden_dir):
        vol = np.ascontiguousarray(g[f"in_{i}"])
        (o8,), (o,) = gh.resample3d([vol], [out], fmin, fmag, np.uint8, [roi], want_order=True)
        assert o == order, f"pass order, case {i}"
        assert np.array_equal(o8, g[f"out_u8_{i}"]), f"u8 case {i}"
        (of,) = gh.resample3d([vol], [out], fmin, fmag, np.float32, [roi])
        assert np.array_equal(_bits(of), _bits(g[f"out_f32_{i}"])), f"u8->f32 case {i}"
        volf = (vol.astype(np.float32) * 1.37 - 20).astype(np.float32)
        (off,) = gh.resample3d([volf], [out], fmin, fmag, np.float32, [roi])
        assert np.array_equal(_bits(off), _bits(g[f"out_f32f32_{i}"])), f"f32 case {i}"


def test_resample3d_random_batches_against_oracle():
    """Ragged batches (every sample its own shape, ROI and pass order; one filter set per batch as the operator builds them)."""
    import gpu_helpers as gh
    from test_resample3d_cpu import random_case
    rng = np.random.default_rng(31)
    it = 0
    for batch in range(12):
        vols, outs, rois, fmin, fmag, odt, dt = [], [], [], None, None, None, None
        while len(vols) < 6:
            it += 1
            vol, out, a, b, o, roi = random_case(rng, it)
            if fmin is None:
                fmin, fmag, odt, dt = a, b, o, vol.dtype
            if any(x[0] == NN for x in fmin + fmag):
                roi = None
            vols.append(np.ascontiguousarray(vol.astype(dt)))
            outs.append(out)
            rois.append(roi)
        if dt == np.float32:
            odt = np.float32
        got, orders = gh.resample3d(vols, outs, fmin, fmag, odt, rois, want_order=True)
        for k in range(len(vols)):
            want, order = po.resample3d(vols[k], outs[k], fmin, fmag, odt, rois[k], want_order=True)
            assert orders[k] == order, (batch, k)
            assert np.array_equal(_bits(got[k]), _bits(want)), (batch, k, vols[k].shape, outs[k], fmin, fmag, rois[k])


def test_resample3d_medical_volume_shapes():
    """A CT-like volume down and up (128 x 160 x 160 -> 64 x 96 x 96 and -> 144 x 176 x 200), default linear + antialias, u8 and f32;
    also through a plan that is reused with a larger batch (temporaries grow)."""
    import gpu_helpers as gh
    from dali_b200 import capi
    rng = np.random.default_rng(32)
    base = rng.uniform(0, 255, (16, 20, 20, 1)).astype(np.float32)
    vol = np.clip(np.kron(base, np.ones((8, 8, 8, 1), np.float32)) + rng.normal(0, 6, (128, 160, 160, 1)), 0, 255).astype(np.uint8)
    f_min, f_mag = [(L, 1, 0.0)] * 3, [(L, 0, 0.0)] * 3
    plan = capi.Plan("Resample3D", 4)
    for out in ((64, 96, 96), (144, 176, 200)):
        for odt in (np.uint8, np.float32):
            (got,) = gh.resample3d([vol], [out], f_min, f_mag, odt, plan=plan)
            want = po.resample3d(vol, out, f_min, f_mag, odt)
            assert np.array_equal(_bits(got), _bits(want)), (out, odt)
    small = np.ascontiguousarray(vol[:40, :50, :60])
    got = gh.resample3d([vol, small, small], [(32, 40, 40), (20, 25, 30), (50, 50, 70)], f_min, f_mag, np.uint8, plan=plan)
    for g_, (v, o) in zip(got, [(vol, (32, 40, 40)), (small, (20, 25, 30)), (small, (50, 50, 70))]):
        assert np.array_equal(g_, po.resample3d(v, o, f_min, f_mag, np.uint8))


# ---- through the operator: fn.resize on DHWC / FDHWC data -----------------------------------------------------------------------
def _resize_params(in_shape, req):
    """Output size and source ROI of a volume: the operator's ResizeAttr arithmetic (resize_attr_base.{h,cc}, default mode,
    subpixel_scale), pinned on the CPU against the reference's Resize3D* vectors by tests/test_host_cpu.py."""
    import ctypes as C
    from dali_b200 import backend
    f3 = lambda v: (C.c_float * 3)(*[float(x) for x in v])
    dst, lo, hi = (C.c_int * 3)(), (C.c_float * 3)(), (C.c_float * 3)()
    assert backend.lib().dalihTestResizeParams3D(0, f3(req), f3((0, 0, 0)), f3(in_shape), 1, None, dst, lo, hi) == 0
    return tuple(dst), (list(lo), list(hi))


def _run_volumes(vols, build, layout="DHWC"):
    from dali_b200 import fn, pipeline_def

    @pipeline_def(batch_size=len(vols), num_threads=1, device_id=0)
    def pipe():
        x = fn.external_source(source=lambda i: vols, device="gpu", layout=layout)
        return build(fn, x)
    p = pipe()
    p.build()
    return [o.as_cpu() for o in p.run()]


def test_fn_resize_on_volumes():
    from dali_b200 import types
    rng = np.random.default_rng(33)
    vols = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((20, 30, 40, 1), (16, 16, 16, 1), (9, 33, 21, 1))]
    f_min, f_mag = [(T, 1, 0.0)] * 3, [(L, 0, 0.0)] * 3           # defaults: linear, antialias on -> triangular when shrinking
    a, b, c = _run_volumes(vols, lambda fn, x: (
        fn.resize(x, size=[12, 18, 25]),
        fn.resize(x, resize_z=10, resize_y=24, dtype=types.FLOAT),
        fn.resize(x, size=[24, 20, 30], interp_type=types.INTERP_CUBIC, antialias=False)))
    for i, v in enumerate(vols):
        assert a[i].shape == (12, 18, 25, 1)
        assert np.array_equal(np.asarray(a[i]), po.resample3d(v, (12, 18, 25), f_min, f_mag, np.uint8)), i
        dst, roi = _resize_params(v.shape[:3], (10, 24, 0))
        assert tuple(b[i].shape) == dst + (1,), (i, b[i].shape, dst)
        want = po.resample3d(v, dst, f_min, f_mag, np.float32, roi)
        assert np.array_equal(_bits(np.asarray(b[i])), _bits(want)), i
        cub = [(CU, 0, 0.0)] * 3
        assert np.array_equal(np.asarray(c[i]), po.resample3d(v, (24, 20, 30), cub, cub, np.uint8)), i


def test_fn_resize_on_sequences_of_volumes_with_roi():
    rng = np.random.default_rng(34)
    seqs = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((2, 12, 20, 24, 3), (3, 10, 10, 14, 3))]
    f_min, f_mag = [(T, 1, 0.0)] * 3, [(L, 0, 0.0)] * 3
    (a,) = _run_volumes(seqs, lambda fn, x: (fn.resize(x, size=[6, 9, 11], roi_start=[0.25, 0.0, 0.5], roi_end=[1.0, 0.75, 0.0], roi_relative=True),),
                        layout="FDHWC")
    for i, s in enumerate(seqs):
        got = np.asarray(a[i])
        assert got.shape == (s.shape[0], 6, 9, 11, 3)
        D, H, W = s.shape[1:4]
        roi = ([0.25 * D, 0.0, 0.5 * W], [1.0 * D, 0.75 * H, 0.0])       # x flipped (start > end)
        roi = ([float(np.float32(v)) for v in roi[0]], [float(np.float32(v)) for v in roi[1]])
        for f in range(s.shape[0]):
            want = po.resample3d(np.ascontiguousarray(s[f]), (6, 9, 11), f_min, f_mag, np.uint8, roi)
            assert np.array_equal(got[f], want), (i, f)


def test_fn_resize_on_channel_first_volumes():
    """CDHW / FCDHW: the dimensions in front of the spatial ones are frames (resize_op_impl.h:56-101) -- C volumes of one channel"""
    from dali_b200 import types
    rng = np.random.default_rng(35)
    vols = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((2, 10, 14, 18), (3, 8, 8, 8))]
    f_min, f_mag = [(T, 1, 0.0)] * 3, [(L, 0, 0.0)] * 3
    a, b = _run_volumes(vols, lambda fn, x: (fn.resize(x, size=[6, 9, 11]), fn.resize(x, resize_x=12, resize_y=7, resize_z=5, dtype=types.FLOAT)),
                        layout="CDHW")
    for i, v in enumerate(vols):
        ga, gb = np.asarray(a[i]), np.asarray(b[i])
        assert ga.shape == (v.shape[0], 6, 9, 11) and gb.shape == (v.shape[0], 5, 7, 12)
        for c in range(v.shape[0]):
            plane = np.ascontiguousarray(v[c][..., None])
            assert np.array_equal(ga[c], po.resample3d(plane, (6, 9, 11), f_min, f_mag, np.uint8)[..., 0]), (i, c)
            assert np.array_equal(_bits(gb[c]), _bits(po.resample3d(plane, (5, 7, 12), f_min, f_mag, np.float32)[..., 0])), (i, c)
    seqs = [rng.integers(0, 256, (2, 2, 6, 7, 8), dtype=np.uint8)]
    (c5,) = _run_volumes(seqs, lambda fn, x: (fn.resize(x, size=[3, 4, 5]),), layout="FCDHW")
    got = np.asarray(c5[0])
    assert got.shape == (2, 2, 3, 4, 5)
    for f in range(2):
        for c in range(2):
            want = po.resample3d(np.ascontiguousarray(seqs[0][f, c][..., None]), (3, 4, 5), f_min, f_mag, np.uint8)[..., 0]
            assert np.array_equal(got[f, c], want), (f, c)
