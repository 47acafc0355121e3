"""The C2 hot path (JPEG decode -> Resize -> CropMirrorNormalize) driven directly over the C-ABI.

`ImagePipelineC2` is the minimal batched executor the operators of dali_b200.pipeline sit on: it owns one
plan per stage and the inter-stage device buffers, and enqueues the whole chain on one CUDA stream with no
host synchronisation between the stages.  bench.py uses it for the device-resident (`value`) measurement and
__graft_entry__.smoke() for the smoke run; user code goes through dali_b200.fn / pipeline_def.
"""
import ctypes as C

import numpy as np

from . import capi

IMAGENET_MEAN = [0.485 * 255, 0.456 * 255, 0.406 * 255]
IMAGENET_STD = [0.229 * 255, 0.224 * 255, 0.225 * 255]


def cmn_norm_args(mean, std, scale=1.0, shift=0.0):
    """dali/operators/image/crop/crop_mirror_normalize.h:135-141: double arithmetic, stored as float."""
    mean = np.atleast_1d(np.asarray(mean, np.float32))
    std = np.atleast_1d(np.asarray(std, np.float32))
    n = max(mean.size, std.size)
    m, s = np.empty(n, np.float32), np.empty(n, np.float32)
    for d in range(n):
        mean_val, std_val = np.float64(mean[d % mean.size]), np.float64(std[d % std.size])
        # reference: std::fma(-shift, std / scale, mean) in double (exact for shift == 0, the hot-path case)
        m[d] = np.float32(mean_val if shift == 0 else np.float64(-shift) * (std_val / np.float64(scale)) + mean_val)
        s[d] = np.float32(np.float64(scale) / std_val)
    return m, s


class ImagePipelineC2:
    """decode (mixed) -> resize(out_h, out_w) -> crop_mirror_normalize(fp16/fp32, CHW) for one batch."""

    def __init__(self, max_batch, out_hw=(224, 224), out_dtype="float16", mean=IMAGENET_MEAN, std=IMAGENET_STD, device=None):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_batch = max_batch
        self.out_hw = tuple(out_hw)
        self.out_dtype = torch.float16 if out_dtype in ("float16", torch.float16) else torch.float32
        self.jpeg = capi.Plan("Jpeg", max_batch)
        self.resample = capi.Plan("Resample", max_batch)
        self.cmn = capi.Plan("Cmn", max_batch)
        self.mean, self.inv_std = cmn_norm_args(mean, std)
        self._decoded = None
        self._resized = None
        self.output = None
        self.n = 0
        self.shapes = []
        self.staged_bytes = 0

    # ---- per batch host work: header parse + staging, shape inference, descriptor build
    def setup(self, streams, mirror=None):
        torch, lib = self.torch, capi.lib()
        n = len(streams)
        self._keep = [np.frombuffer(s, np.uint8) if not isinstance(s, np.ndarray) else s for s in streams]
        ptrs = (C.c_void_p * n)(*[b.ctypes.data for b in self._keep])
        lens = (C.c_size_t * n)(*[b.size for b in self._keep])
        capi.check(lib.dalib200JpegPlanSetup(self.jpeg.handle, n, ptrs, lens, capi.RGB, 1))
        self.staged_bytes = int(lib.dalib200JpegPlanStagedBytes(self.jpeg.handle))
        info = capi.JpegInfo()
        shapes = []
        for i in range(n):
            capi.check(lib.dalib200JpegPlanGetInfo(self.jpeg.handle, i, C.byref(info)))
            shapes.append((info.height, info.width))
        self.shapes, self.n = shapes, n
        oh, ow = self.out_hw
        rs = (capi.ResampleSample * n)()
        cm = (capi.CmnSample * n)()
        # what fn.resize passes by default (resampling_attr.cc:76-133): min = Triangular + antialias, mag = Linear
        fmin = capi.FilterDesc(capi.FILTER_TRIANGULAR, 1, 0.0)
        fmag = capi.FilterDesc(capi.FILTER_LINEAR, 0, 0.0)
        for i, (h, w) in enumerate(shapes):
            r = rs[i]
            r.in_h, r.in_w, r.channels, r.out_h, r.out_w = h, w, 3, oh, ow
            for d in range(2):
                r.min_filter[d] = fmin; r.mag_filter[d] = fmag; r.use_roi[d] = 0
            c = cm[i]
            c.in_h, c.in_w, c.channels = oh, ow, 3
            c.anchor_y, c.anchor_x, c.crop_h, c.crop_w = 0, 0, oh, ow
            c.mirror = int(mirror[i]) if mirror is not None else 0
            for k in range(3):
                c.mean[k] = float(self.mean[k]); c.inv_std[k] = float(self.inv_std[k]); c.fill[k] = 0.0
            c.mean[3] = 0.0; c.inv_std[3] = 1.0; c.fill[3] = 0.0
        capi.check(lib.dalib200ResamplePlanSetup(self.resample.handle, n, rs, capi.UINT8, capi.UINT8))
        capi.check(lib.dalib200CmnPlanSetup(self.cmn.handle, n, cm, capi.FLOAT16 if self.out_dtype == torch.float16 else capi.FLOAT,
                                            capi.LAYOUT_CHW, 3))
        # inter-stage buffers (grow only; one allocation per stage, samples packed back to back)
        dec_bytes = sum(h * w * 3 for h, w in shapes)
        if self._decoded is None or self._decoded.numel() < dec_bytes:
            self._decoded = torch.empty(dec_bytes, dtype=torch.uint8, device=self.device)
        if self._resized is None or self._resized.shape[0] < n:
            self._resized = torch.empty((max(n, self.max_batch), oh, ow, 3), dtype=torch.uint8, device=self.device)
        if self.output is None or self.output.shape[0] < n:
            self.output = torch.empty((max(n, self.max_batch), 3, oh, ow), dtype=self.out_dtype, device=self.device)
        base = self._decoded.data_ptr()
        offs, o = [], 0
        for h, w in shapes:
            offs.append(base + o)
            o += h * w * 3
        self._dec_ptrs = capi.ptr_array(offs)
        rbase, rstride = self._resized.data_ptr(), oh * ow * 3
        self._res_ptrs = capi.ptr_array([rbase + i * rstride for i in range(n)])
        obase, ostride = self.output.data_ptr(), 3 * oh * ow * self.output.element_size()
        self._out_ptrs = capi.ptr_array([obase + i * ostride for i in range(n)])

    def upload(self, stream=None):
        capi.check(capi.lib().dalib200JpegUpload(self.jpeg.handle, capi.stream_handle(stream)))

    def launch(self, stream=None):
        lib, s = capi.lib(), capi.stream_handle(stream)
        capi.check(lib.dalib200JpegLaunch(self.jpeg.handle, self._dec_ptrs, s))
        capi.check(lib.dalib200ResampleLaunch(self.resample.handle, self._dec_ptrs, self._res_ptrs, s))
        capi.check(lib.dalib200CmnLaunch(self.cmn.handle, self._res_ptrs, self._out_ptrs, s))
        return self.output[: self.n]

    def run(self, streams, mirror=None, stream=None):
        self.setup(streams, mirror)
        self.upload(stream)
        return self.launch(stream)

    def decoded(self, i):
        h, w = self.shapes[i]
        off = sum(hh * ww * 3 for hh, ww in self.shapes[:i])
        return self._decoded[off: off + h * w * 3].view(h, w, 3)

    def resized(self, i):
        return self._resized[i]

    def status(self):
        st = (C.c_int32 * self.n)()
        capi.check(capi.lib().dalib200JpegGetStatus(self.jpeg.handle, st))
        return list(st)
