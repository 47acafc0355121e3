"""Helpers for the -m gpu parity tests: they drive the CUDA path through the C-ABI (ctypes) with torch
tensors as device buffers and return numpy arrays for comparison with the oracle."""
import ctypes as C

import numpy as np

from dali_b200 import capi


def _torch():
    import torch
    return torch


def to_dev(arrs):
    torch = _torch()
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrs]


def cmn(imgs, anchors, crops, mirrors, mean, inv_std, out_dtype=np.float32, layout="CHW", out_channels=None, fill=None, plan=None):
    torch = _torch()
    n = len(imgs)
    oc = out_channels or imgs[0].shape[2]
    samples = (capi.CmnSample * n)()
    for i, im in enumerate(imgs):
        s = samples[i]
        s.in_h, s.in_w, s.channels = im.shape
        s.anchor_y, s.anchor_x = anchors[i]
        s.crop_h, s.crop_w = crops[i]
        s.mirror = int(mirrors[i])
        s.mean[:] = list(capi.np_f32(mean, 4)); s.inv_std[:] = list(capi.np_f32(inv_std, 4))
        if np.asarray(mean).size < 4:
            for c in range(np.asarray(mean).size, 4):
                s.mean[c] = 0.0; s.inv_std[c] = 1.0
        s.fill[:] = list(capi.np_f32(fill, 4))
    plan = plan or capi.Plan("Cmn", max(n, 1))
    dt = capi.FLOAT if np.dtype(out_dtype) == np.float32 else capi.FLOAT16
    capi.check(capi.lib().dalib200CmnPlanSetup(plan.handle, n, samples, dt, capi.LAYOUT_CHW if layout == "CHW" else capi.LAYOUT_HWC, oc))
    din = to_dev(imgs)
    tdt = torch.float32 if dt == capi.FLOAT else torch.float16
    outs = [torch.empty((oc, c[0], c[1]) if layout == "CHW" else (c[0], c[1], oc), dtype=tdt, device="cuda") for c in crops]
    capi.check(capi.lib().dalib200CmnLaunch(plan.handle, capi.ptr_array(din), capi.ptr_array(outs), capi.stream_handle()))
    torch.cuda.synchronize()
    return [o.cpu().numpy() for o in outs]


def fill_resample_sample(s, shape, out_hw, min_filter, mag_filter, roi):
    s.in_h, s.in_w, s.channels = shape
    s.out_h, s.out_w = out_hw
    def pair(f):
        return [capi.make_filter(f[0]), capi.make_filter(f[1])] if isinstance(f, list) else [capi.make_filter(f), capi.make_filter(f)]
    mn, mg = pair(min_filter), pair(mag_filter)
    for d in range(2):
        s.min_filter[d] = mn[d]; s.mag_filter[d] = mg[d]
        s.use_roi[d] = 0
        if roi is not None and roi[0][d] is not None:
            s.use_roi[d] = 1; s.roi_start[d] = roi[0][d]; s.roi_end[d] = roi[1][d]


def resample(imgs, out_hws, min_filter=(capi.FILTER_LINEAR, 1, 0.0), mag_filter=(capi.FILTER_LINEAR, 1, 0.0), out_dtype=None,
             rois=None, want_order=False, plan=None):
    torch = _torch()
    n = len(imgs)
    samples = (capi.ResampleSample * n)()
    for i, im in enumerate(imgs):
        fill_resample_sample(samples[i], im.shape, out_hws[i], min_filter, mag_filter, rois[i] if rois else None)
    in_dt = capi.UINT8 if imgs[0].dtype == np.uint8 else capi.FLOAT
    out_dtype = np.dtype(out_dtype or imgs[0].dtype)
    out_dt = capi.UINT8 if out_dtype == np.uint8 else capi.FLOAT
    plan = plan or capi.Plan("Resample", max(n, 1))
    capi.check(capi.lib().dalib200ResamplePlanSetup(plan.handle, n, samples, in_dt, out_dt))
    din = to_dev(imgs)
    outs = [torch.empty((hw[0], hw[1], im.shape[2]), dtype=torch.uint8 if out_dt == capi.UINT8 else torch.float32, device="cuda")
            for hw, im in zip(out_hws, imgs)]
    capi.check(capi.lib().dalib200ResampleLaunch(plan.handle, capi.ptr_array(din), capi.ptr_array(outs), capi.stream_handle()))
    torch.cuda.synchronize()
    res = [o.cpu().numpy() for o in outs]
    if want_order:
        return res, [capi.lib().dalib200ResamplePlanGetOrder(plan.handle, i) for i in range(n)]
    return res
