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

    def __init__(self, max_batch, out_hw=(224, 224), out_dtype="float16", mean=IMAGENET_MEAN, std=IMAGENET_STD, device=None, fused=False):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.max_batch = max_batch
        self.out_hw = tuple(out_hw)
        self.out_dtype = torch.float16 if out_dtype in ("float16", torch.float16) else torch.float32
        # fused: samples that qualify (4:2:0 YCbCr streams, stream-eligible resize) are resized straight from the decoder's planes
        # (dalib200ResampleLaunchPlanar); the decoded RGB image is then never written.  The others take the two-kernel path.
        # Off by default: less DRAM traffic but ~4 % more time per batch at 1080p (both variants are issue-bound).
        self.fused = bool(fused)
        self.jpeg = capi.Plan("Jpeg", max_batch)
        self.resample = capi.Plan("Resample", max_batch)
        self.resample_planar = capi.Plan("Resample", max_batch) if self.fused else None
        self.planar = []
        self.cmn = capi.Plan("Cmn", max_batch)
        self.mean, self.inv_std = cmn_norm_args(mean, std)
        self._decoded = None
        self._resized = None
        self.output = None
        self._bound = False
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
        self.planar = [0] * n
        if self.fused:
            ok = (C.c_uint8 * n)()
            capi.check(lib.dalib200ResamplePlanSetupPlanar(self.resample_planar.handle, n, rs, ok))
            granted = (C.c_uint8 * n)()
            capi.check(lib.dalib200JpegPlanSetPlanesOnly(self.jpeg.handle, ok, granted))
            self.planar = list(granted)
            if any(o and not g for o, g in zip(ok, granted)):
                # the resampler's item list covers the samples IT found eligible: re-run its setup with the decoder's verdict
                for i in range(n):
                    if not granted[i]:
                        rs[i].channels = 1          # 1-channel samples are never planar-eligible
                capi.check(lib.dalib200ResamplePlanSetupPlanar(self.resample_planar.handle, n, rs, ok))
                for i in range(n):
                    rs[i].channels = 3
        self._rest = [i for i in range(n) if not self.planar[i]]
        if self._rest:
            rs_b = (capi.ResampleSample * len(self._rest))(*[rs[i] for i in self._rest])
            capi.check(lib.dalib200ResamplePlanSetup(self.resample.handle, len(self._rest), rs_b, capi.UINT8, capi.UINT8))
        capi.check(lib.dalib200CmnPlanSetup(self.cmn.handle, n, cm, capi.FLOAT16 if self.out_dtype == torch.float16 else capi.FLOAT,
                                            capi.LAYOUT_CHW, 3))
        # inter-stage buffers (grow only; one allocation per stage, samples packed back to back)
        dec_bytes = sum(h * w * 3 for h, w in shapes)
        if self._decoded is None or self._decoded.numel() < dec_bytes:
            self._decoded = torch.empty(dec_bytes, dtype=torch.uint8, device=self.device)
        if self._resized is None or self._resized.shape[0] < n:
            self._resized = torch.empty((max(n, self.max_batch), oh, ow, 3), dtype=torch.uint8, device=self.device)
        if self._bound:
            if tuple(self.output.shape[1:]) != (3, oh, ow) or self.output.shape[0] < n or self.output.dtype != self.out_dtype:
                raise ValueError("bind_output: the bound tensor does not fit this batch")
        elif self.output is None or self.output.shape[0] < n:
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

    def bind_output(self, tensor):
        """CropMirrorNormalize writes into `tensor` ([>= batch, 3, H, W], contiguous) from the next setup() on -- e.g. this rank's
        slice of a sharding.GatherBuffer, so that a following all-gather needs no copy."""
        if not tensor.is_contiguous():
            raise ValueError("bind_output: the tensor must be contiguous")
        self.output, self._bound = tensor, True

    def upload(self, stream=None):
        capi.check(capi.lib().dalib200JpegUpload(self.jpeg.handle, capi.stream_handle(stream)))

    def launch(self, stream=None):
        lib, s = capi.lib(), capi.stream_handle(stream)
        capi.check(lib.dalib200JpegLaunch(self.jpeg.handle, self._dec_ptrs, s))
        if any(self.planar):
            srcs = (capi.PlanarImage * self.n)()
            for i in range(self.n):
                if self.planar[i]:
                    capi.check(lib.dalib200JpegPlanGetPlanes(self.jpeg.handle, i, C.byref(srcs[i])))
            capi.check(lib.dalib200ResampleLaunchPlanar(self.resample_planar.handle, srcs, self._res_ptrs, s))
        if self._rest:
            dec = capi.ptr_array([self._dec_ptrs[i] for i in self._rest])
            res = capi.ptr_array([self._res_ptrs[i] for i in self._rest])
            capi.check(lib.dalib200ResampleLaunch(self.resample.handle, dec, res, s))
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


class VideoPipelineC3:
    """warp_affine(LINEAR, fill 0, same size) -> hsv(u8) -> crop_mirror_normalize(fp16, CHW) over independent HWC frames
    (BASELINE configs[2]; FHWC sequences are flattened to frames, frames of a sequence share their parameters)."""

    def __init__(self, nframes, hw=(720, 1280), mean=IMAGENET_MEAN, std=IMAGENET_STD, device=None):
        import torch
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.hw = nframes, tuple(hw)
        self.warp = capi.Plan("Warp", nframes)
        self.hsv = capi.Plan("Pointwise", nframes)
        self.cmn = capi.Plan("Cmn", nframes)
        self.mean, self.inv_std = cmn_norm_args(mean, std)
        h, w = self.hw
        self.warped = torch.empty((nframes, h, w, 3), dtype=torch.uint8, device=self.device)
        self.twisted = torch.empty((nframes, h, w, 3), dtype=torch.uint8, device=self.device)
        self.output = torch.empty((nframes, 3, h, w), dtype=torch.float16, device=self.device)

    def setup(self, inv_matrices, hsv_params, mirror):
        """inv_matrices[i]: 2x3 dst->src; hsv_params[i] = (hue, saturation, value); mirror[i] in {0, 1}."""
        lib, n = capi.lib(), self.n
        h, w = self.hw
        ws = (capi.WarpSample * n)()
        cs = (capi.ColorSample * n)()
        cm = (capi.CmnSample * n)()
        M, T = np.empty(9, np.float32), np.empty(3, np.float32)
        last = None
        for i in range(n):
            s = ws[i]
            s.in_h, s.in_w, s.channels, s.out_h, s.out_w = h, w, 3, h, w
            s.matrix[:] = [float(v) for v in np.asarray(inv_matrices[i], np.float32).reshape(6)]
            hp = tuple(float(v) for v in hsv_params[i])
            if hp != last:
                lib.dalib200ColorTwistMatrix(C.c_float(hp[0]), C.c_float(hp[1]), C.c_float(hp[2]), C.c_float(1.0), C.c_float(1.0),
                                             C.c_float(128.0), M.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p))
                last = hp
            cs[i].num_pixels = h * w
            cs[i].matrix[:] = [float(v) for v in M]
            cs[i].offset[:] = [0.0, 0.0, 0.0]        # Hsv has no offset term (color_twist.h:156-170)
            c = cm[i]
            c.in_h, c.in_w, c.channels = h, w, 3
            c.anchor_y, c.anchor_x, c.crop_h, c.crop_w = 0, 0, h, w
            c.mirror = int(mirror[i])
            for k in range(3):
                c.mean[k] = float(self.mean[k]); c.inv_std[k] = float(self.inv_std[k]); c.fill[k] = 0.0
            c.mean[3] = 0.0; c.inv_std[3] = 1.0; c.fill[3] = 0.0
        capi.check(lib.dalib200WarpPlanSetup(self.warp.handle, n, ws, 1, 1, C.c_float(0.0), capi.UINT8))
        capi.check(lib.dalib200LinearTransformSetup(self.hsv.handle, n, cs, capi.UINT8))
        capi.check(lib.dalib200CmnPlanSetup(self.cmn.handle, n, cm, capi.FLOAT16, capi.LAYOUT_CHW, 3))
        fb = h * w * 3
        self._w_ptrs = capi.ptr_array([self.warped.data_ptr() + i * fb for i in range(n)])
        self._t_ptrs = capi.ptr_array([self.twisted.data_ptr() + i * fb for i in range(n)])
        self._o_ptrs = capi.ptr_array([self.output.data_ptr() + i * fb * 2 for i in range(n)])

    def launch(self, frames, stream=None):
        """frames: uint8 CUDA tensor [n, H, W, 3]."""
        lib, s = capi.lib(), capi.stream_handle(stream)
        fb = self.hw[0] * self.hw[1] * 3
        in_ptrs = capi.ptr_array([frames.data_ptr() + i * fb for i in range(self.n)])
        capi.check(lib.dalib200WarpLaunch(self.warp.handle, in_ptrs, self._w_ptrs, s))
        capi.check(lib.dalib200PointwiseLaunch(self.hsv.handle, self._w_ptrs, self._t_ptrs, s))
        capi.check(lib.dalib200CmnLaunch(self.cmn.handle, self._t_ptrs, self._o_ptrs, s))
        return self.output


class AudioPipelineC4:
    """spectrogram(nfft, window_length, window_step, power 2, centred, reflect) -> mel_filter_bank (BASELINE configs[3])."""

    def __init__(self, nclips, clip_len, nfft=1024, window_length=1024, window_step=256, nfilter=128, sample_rate=16000.0,
                 freq_high=8000.0, device=None, fused=False, keep_spectrogram=True):
        import torch
        # fused: STFT -> mel in one kernel (dalib200SpectrogramMelLaunch); with keep_spectrogram=False the spectrogram is never
        # written to HBM (it is an intermediate of the chain), which is what the executor does when nothing else consumes it
        self.fused, self.keep_spectrogram = bool(fused), bool(keep_spectrogram)
        self.torch = torch
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lib = capi.lib()
        self.n, self.clip_len, self.nfilter = nclips, clip_len, nfilter
        self.spec = capi.Plan("Spectrogram", nclips)
        self.mel = capi.Plan("Mel", nclips)
        args = capi.SpectrogramArgs(nfft, window_length, window_step, 2, 1, 1, 1)
        lens = (C.c_int64 * nclips)(*[clip_len] * nclips)
        capi.check(lib.dalib200SpectrogramPlanSetup(self.spec.handle, C.byref(args), None, nclips, lens))
        self.nbin = nfft // 2 + 1
        self.nwin = int(lib.dalib200SpectrogramNumWindows(self.spec.handle, 0))
        margs = capi.MelArgs(nfilter, sample_rate, 0.0, freq_high, 0, 1)
        nw = (C.c_int64 * nclips)(*[self.nwin] * nclips)
        capi.check(lib.dalib200MelPlanSetup(self.mel.handle, C.byref(margs), self.nbin, nclips, nw))
        self.spectra = torch.empty((nclips, self.nbin, self.nwin), dtype=torch.float32, device=self.device)
        self.output = torch.empty((nclips, nfilter, self.nwin), dtype=torch.float32, device=self.device)
        sb, ob = self.nbin * self.nwin * 4, nfilter * self.nwin * 4
        self._s_ptrs = capi.ptr_array([self.spectra.data_ptr() + i * sb for i in range(nclips)])
        self._o_ptrs = capi.ptr_array([self.output.data_ptr() + i * ob for i in range(nclips)])

    def launch(self, clips, stream=None):
        """clips: float32 CUDA tensor [n, clip_len]."""
        lib, s = capi.lib(), capi.stream_handle(stream)
        in_ptrs = capi.ptr_array([clips.data_ptr() + i * self.clip_len * 4 for i in range(self.n)])
        if self.fused:
            if not lib.dalib200SpectrogramMelSupported(self.spec.handle, self.mel.handle):
                raise capi.DaliB200Error("the fused STFT -> mel kernel needs nfft = 1024 and the (f, t) layout")
            capi.check(lib.dalib200SpectrogramMelLaunch(self.spec.handle, self.mel.handle, in_ptrs,
                                                        self._s_ptrs if self.keep_spectrogram else None, self._o_ptrs, s))
            return self.output
        capi.check(lib.dalib200SpectrogramLaunch(self.spec.handle, in_ptrs, self._s_ptrs, s))
        capi.check(lib.dalib200MelLaunch(self.mel.handle, self._s_ptrs, self._o_ptrs, s))
        return self.output
