"""GPU parity test of fn.resize on channel-first 2-D layouts (CHW, FCHW, CFHW): the dimensions in front of H, W are frames of
one-channel images (resize_op_impl.h:56-101).  Host-side change only (the kernels are the validated 2-D ones); collected last because
it was written after the round's GPU budget was spent."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _run(batch, layout, build):
    from dali_b200 import fn, pipeline_def

    @pipeline_def(batch_size=len(batch), num_threads=1, device_id=0)
    def pipe():
        x = fn.external_source(source=lambda i: batch, device="gpu", layout=layout)
        return build(fn, x)
    p = pipe()
    p.build()
    return [o.as_cpu() for o in p.run()]


def test_fn_resize_on_channel_first_images():
    from dali_b200 import types
    rng = np.random.default_rng(41)
    f_min, f_mag = (po.F_TRIANGULAR, 1, 0.0), (po.F_LINEAR, 0, 0.0)
    imgs = [rng.integers(0, 256, s, dtype=np.uint8) for s in ((3, 97, 131), (3, 64, 64), (1, 120, 75))]
    a, b = _run(imgs, "CHW", lambda fn, x: (fn.resize(x, size=[48, 56]), fn.resize(x, resize_x=200, dtype=types.FLOAT, subpixel_scale=False)))
    for i, im in enumerate(imgs):
        ga, gb = np.asarray(a[i]), np.asarray(b[i])
        oh = int(np.floor(im.shape[1] * np.float32(200) / np.float32(im.shape[2]) + 0.5))
        assert ga.shape == (im.shape[0], 48, 56) and gb.shape == (im.shape[0], oh, 200), (ga.shape, gb.shape, oh)
        for c in range(im.shape[0]):
            plane = np.ascontiguousarray(im[c][..., None])
            assert np.array_equal(ga[c], po.resample(plane, (48, 56), f_min, f_mag)[..., 0]), (i, c)
            want = po.resample(plane, (oh, 200), f_min, f_mag, np.float32)[..., 0]
            assert np.array_equal(gb[c].view(np.uint32), want.view(np.uint32)), (i, c)
    for layout, shape in (("FCHW", (2, 3, 40, 52)), ("CFHW", (3, 2, 40, 52))):
        seq = [rng.integers(0, 256, shape, dtype=np.uint8)]
        (o,) = _run(seq, layout, lambda fn, x: (fn.resize(x, size=[20, 33]),))
        got = np.asarray(o[0])
        assert got.shape == shape[:2] + (20, 33)
        for u in range(shape[0]):
            for v in range(shape[1]):
                want = po.resample(np.ascontiguousarray(seq[0][u, v][..., None]), (20, 33), f_min, f_mag)[..., 0]
                assert np.array_equal(got[u, v], want), (layout, u, v)
