"""GPU parity tests of PROGRESSIVE (SOF2) JPEG decoding: the CUDA path through the C-ABI / decoders.image against libjpeg-turbo
(cv2.imdecode of the progressive stream), the oracle's decode of the baseline twin (same coefficients) and the twin's coefficients.

Collected last on purpose: the progressive entropy stage was written after the round's GPU budget was spent.  Its planner and scan
decoder are pinned on the CPU (tests/test_jpeg_prog_cpu.py runs the kernel bodies compiled for the host, bit for bit against the
baseline twins); what these tests add is the launch itself and the hand-over to the shared DC / IDCT / colour stages."""
import numpy as np
import pytest

from dali_b200 import capi
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _cases():
    import cv2
    from test_jpeg_prog_cpu import synth, twins
    ss = [cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422,
          cv2.IMWRITE_JPEG_SAMPLING_FACTOR_440, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_411]
    out = []
    for k, (h, w, s, q, rst) in enumerate([(48, 64, 0, 90, 0), (33, 47, 1, 75, 0), (97, 61, 2, 60, 3), (136, 200, 0, 95, 0), (17, 23, 3, 80, 0),
                                           (24, 56, 4, 70, 2), (8, 8, 0, 90, 0), (1, 1, 1, 90, 0), (250, 333, 0, 85, 5), (480, 640, 0, 90, 0)]):
        out.append(twins(synth(h, w, 300 + k), q, ss[s], rst))
    out.append(twins(synth(65, 130, 399)[..., 0], 85))          # single component
    return out


def test_progressive_streams_decode_like_libjpeg():
    import cv2
    import gpu_helpers as g
    from test_jpeg_prog_cpu import mcu_order
    cases = _cases()
    streams = [p for _, p in cases]
    outs, status, plan = g.jpeg_decode(streams, want_coefs=True)
    assert status == [0] * len(streams)
    for i, (base, prog) in enumerate(cases):
        want_c = mcu_order(base).reshape(-1)
        assert np.array_equal(g.jpeg_coefs(plan, i, want_c.size), want_c), f"coefficients of stream {i}"
        want = cv2.imdecode(np.frombuffer(prog, np.uint8), cv2.IMREAD_COLOR)[..., ::-1]
        assert np.array_equal(outs[i], want), f"pixels of stream {i} vs libjpeg-turbo"
        assert np.array_equal(outs[i], po.jpeg_decode(base)), f"pixels of stream {i} vs the oracle's decode of the baseline twin"


def test_mixed_batches_windows_and_output_types():
    """baseline and progressive samples in one batch (the progressive ones have no units in the subsequence decode), a batch of only
    progressive samples, regions of interest, GRAY / BGR / float output, box upsampling, plan reuse"""
    import gpu_helpers as g
    cases = _cases()
    plan = capi.Plan("Jpeg", 16)
    mixed = [cases[0][0], cases[3][1], cases[9][0], cases[9][1], cases[2][1], cases[4][0]]
    twin = [cases[0][0], cases[3][0], cases[9][0], cases[9][0], cases[2][0], cases[4][0]]
    for _ in range(2):
        outs, status = g.jpeg_decode(mixed, plan=plan)
        assert status == [0] * len(mixed)
        for i, o in enumerate(outs):
            assert np.array_equal(o, po.jpeg_decode(twin[i])), i
    only = [cases[3][1], cases[5][1], cases[8][1]]
    outs, status = g.jpeg_decode(only, plan=plan, fancy=False)
    for o, (b, _) in zip(outs, [cases[3], cases[5], cases[8]]):
        assert np.array_equal(o, po.jpeg_decode(b, fancy=False))
    rois = [(8, 3, 120, 77), None, (100, 50, 333, 250)]
    a, st = g.jpeg_decode_ex(only, rois=rois, plan=plan)
    b, _ = g.jpeg_decode_ex([cases[3][0], cases[5][0], cases[8][0]], rois=rois)
    assert st == [0, 0, 0] and all(np.array_equal(x, y) for x, y in zip(a, b))
    for ot, dt in ((capi.GRAY, capi.UINT8), (capi.BGR, capi.UINT8), (capi.YCbCr, capi.FLOAT)):
        a, _ = g.jpeg_decode_ex(only, output_type=ot, dtype=dt, plan=plan)
        b, _ = g.jpeg_decode_ex([cases[3][0], cases[5][0], cases[8][0]], output_type=ot, dtype=dt)
        assert all(np.array_equal(x.view(np.uint8), y.view(np.uint8)) for x, y in zip(a, b)), (ot, dt)


def test_truncated_progressive_streams_report_a_status():
    import gpu_helpers as g
    base, prog = _cases()[3]
    last = prog.rfind(b"\xff\xda")
    outs, status = g.jpeg_decode([prog, prog[: (last + len(prog)) // 2], prog[: last - 3], base])
    assert status == [0, 1, 1, 0]
    assert np.array_equal(outs[0], po.jpeg_decode(base)) and np.array_equal(outs[3], outs[0])
    assert outs[1].shape == outs[0].shape and np.abs(outs[1].astype(int) - outs[0]).mean() < 8      # all but the last bit of Y arrived


def test_decoders_image_accepts_progressive_files():
    from dali_b200 import fn, pipeline_def
    cases = _cases()
    files = [cases[9][1], cases[0][0], cases[8][1], cases[3][1]]
    twin = [cases[9][0], cases[0][0], cases[8][0], cases[3][0]]

    @pipeline_def(batch_size=len(files), num_threads=2, device_id=0)
    def pipe():
        j = fn.external_source(source=lambda i: [np.frombuffer(f, np.uint8) for f in files])
        img = fn.decoders.image(j, device="mixed")
        return img, fn.resize(img, size=[64, 80])
    p = pipe()
    p.build()
    for _ in range(2):
        img, small = [o.as_cpu() for o in p.run()]
        for i in range(len(files)):
            want = po.jpeg_decode(twin[i])
            assert np.array_equal(np.asarray(img[i]), want), i
            assert np.array_equal(np.asarray(small[i]), po.resample(want, (64, 80), (po.F_TRIANGULAR, 1, 0.0), (po.F_LINEAR, 0, 0.0))), i


def test_sequential_frames_coded_in_several_scans():
    """SOF0 streams whose components come in separate scans (re-coded from interleaved baseline streams by the test encoder of
    tests/test_jpeg_prog_cpu.py, which libjpeg-turbo accepts): one wave, a warp per component"""
    import cv2
    import gpu_helpers as g
    from test_jpeg_prog_cpu import multiscan_expected, synth, to_multiscan_baseline, twins
    streams, bases, exp = [], [], []
    for k, (h, w, s, rst) in enumerate([(48, 64, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 0), (97, 61, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_444, 5),
                                        (33, 47, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_422, 0), (240, 320, cv2.IMWRITE_JPEG_SAMPLING_FACTOR_420, 7)]):
        base, _ = twins(synth(h, w, 600 + k), 85, s)
        multi, hblk, wblk = to_multiscan_baseline(base, rst)
        streams.append(multi)
        bases.append(base)
        exp.append(multiscan_expected(base, hblk, wblk).reshape(-1))
    outs, status, plan = g.jpeg_decode(streams + [bases[0]], want_coefs=True)
    assert status == [0] * 5
    for i, multi in enumerate(streams):
        assert np.array_equal(g.jpeg_coefs(plan, i, exp[i].size), exp[i]), f"coefficients of stream {i}"
        want = cv2.imdecode(np.frombuffer(multi, np.uint8), cv2.IMREAD_COLOR)[..., ::-1]
        assert np.array_equal(outs[i], want), f"pixels of stream {i} vs libjpeg-turbo"
        assert np.array_equal(outs[i], po.jpeg_decode(bases[i])), f"pixels of stream {i} vs the oracle's decode of the interleaved twin"
    assert np.array_equal(outs[4], outs[0])
