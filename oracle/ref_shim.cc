// oracle/ref_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" shim around the reference's OWN CPU kernels, compiled from the sources where
// they lie under /root/reference (never copied into this repo) by oracle/Makefile into
// oracle/_ref/libdali_ref_cpu.so.  It exists to (a) validate the plain-C restatement in
// oracle/*.c, (b) generate the golden vectors under tests/golden, and (c) serve as the
// "reference" CPU baseline in bench.py --impl reference.
//
// Reference entry points wrapped:
//   dali/kernels/imgproc/resample/separable_cpu.h:124-249      SeparableResampleCPU
//   dali/kernels/slice/slice_flip_normalize_permute_pad_cpu.h:318-362
//   dali/kernels/imgproc/warp_cpu.h:45-178                     WarpCPU<AffineMapping2D,...>
//   dali/kernels/imgproc/pointwise/linear_transformation_cpu.h:30-78
//   dali/kernels/imgproc/color_manipulation/color_space_conversion_impl.h:64-227
//   include/dali/core/float16.h / include/dali/util/half.hpp   (host float->half)
//   dali/kernels/signal/window/extract_windows_cpu.cc, dali/kernels/audio/mel_scale/*_cpu.cc
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

#include "dali/core/mm/malloc_resource.h"
#include "dali/core/mm/default_resources.h"
#include "dali/core/cuda_stream_pool.h"
#include "dali/core/float16.h"
#include "dali/core/geom/mat.h"
#include "dali/core/geom/transform.h"
#include "dali/kernels/imgproc/resample/separable_cpu.h"
#include "dali/kernels/slice/slice_flip_normalize_permute_pad_cpu.h"
#include "dali/kernels/imgproc/warp_cpu.h"
#include "dali/kernels/imgproc/warp/affine.h"
#include "dali/kernels/imgproc/pointwise/linear_transformation_cpu.h"
#include "dali/kernels/imgproc/color_manipulation/color_space_conversion_impl.h"

// ---------------------------------------------------------------------------------------------
// Link-time stand-ins for the bits of libdali_core the CPU kernels touch only nominally.
namespace dali {
namespace mm {
template <>
host_memory_resource *GetDefaultResource<memory_kind::host>() {
  static malloc_memory_resource r;
  return &r;
}
template <>
device_async_resource *GetDefaultResource<memory_kind::device>() { abort(); }
device_async_resource *GetDefaultDeviceResource(int) { abort(); }
}  // namespace mm
CUDAStreamPool &CUDAStreamPool::instance() { abort(); }
CUDAStreamLease CUDAStreamPool::Get(int) { abort(); }
void CUDAStreamPool::Put(CUDAStream &&, int) { abort(); }
void CUDAStream::DestroyHandle(cudaStream_t) {}
}  // namespace dali

using namespace dali;            // NOLINT
using namespace dali::kernels;   // NOLINT

namespace {
struct MallocScratch : public Scratchpad {
  std::vector<void *> blocks;
  void *Alloc(mm::memory_kind_id, size_t bytes, size_t alignment) override {
    void *p = nullptr;
    if (alignment < sizeof(void *)) alignment = sizeof(void *);
    if (posix_memalign(&p, alignment, bytes ? bytes : 1)) abort();
    blocks.push_back(p);
    return p;
  }
  ~MallocScratch() { for (auto *p : blocks) free(p); }
};

struct FilterDescC { int type; int antialias; float radius; };

ResamplingParams2D MakeParams(int outH, int outW, const FilterDescC *minf, const FilterDescC *magf,
                              const int *use_roi, const float *roi_start, const float *roi_end) {
  ResamplingParams2D p;
  int outs[2] = { outH, outW };
  for (int d = 0; d < 2; d++) {
    p[d].output_size = outs[d];
    p[d].min_filter = FilterDesc(static_cast<ResamplingFilterType>(minf[d].type), minf[d].antialias != 0,
                                 minf[d].radius);
    p[d].mag_filter = FilterDesc(static_cast<ResamplingFilterType>(magf[d].type), magf[d].antialias != 0,
                                 magf[d].radius);
    if (use_roi && use_roi[d]) p[d].roi = ResamplingParams::ROI(roi_start[d], roi_end[d]);
  }
  return p;
}

template <typename Out, typename In>
int RunResample(const void *in, int H, int W, int C, void *out, int outH, int outW,
                const FilterDescC *minf, const FilterDescC *magf, const int *use_roi,
                const float *roi_start, const float *roi_end, int *order0) {
  resampling::SeparableResampleCPU<Out, In, 2> k;
  KernelContext ctx;
  MallocScratch scratch;
  ctx.scratchpad = &scratch;
  auto params = MakeParams(outH, outW, minf, magf, use_roi, roi_start, roi_end);
  InTensorCPU<In, 3> tin(static_cast<const In *>(in), TensorShape<3>(H, W, C));
  OutTensorCPU<Out, 3> tout(static_cast<Out *>(out), TensorShape<3>(outH, outW, C));
  k.Setup(ctx, tin, params);
  if (order0) *order0 = k.setup.desc.order[0];
  k.Run(ctx, tout, tin, params);
  return 0;
}

template <typename Out, typename In>
int RunResample3D(const void *in, const int *ish, int C, void *out, const int *osh,
                  const FilterDescC *minf, const FilterDescC *magf, const int *use_roi,
                  const float *roi_start, const float *roi_end, int *order) {
  resampling::SeparableResampleCPU<Out, In, 3> k;
  KernelContext ctx;
  MallocScratch scratch;
  ctx.scratchpad = &scratch;
  ResamplingParams3D p;
  for (int d = 0; d < 3; d++) {
    p[d].output_size = osh[d];
    p[d].min_filter = FilterDesc(static_cast<ResamplingFilterType>(minf[d].type), minf[d].antialias != 0, minf[d].radius);
    p[d].mag_filter = FilterDesc(static_cast<ResamplingFilterType>(magf[d].type), magf[d].antialias != 0, magf[d].radius);
    if (use_roi && use_roi[d]) p[d].roi = ResamplingParams::ROI(roi_start[d], roi_end[d]);
  }
  InTensorCPU<In, 4> tin(static_cast<const In *>(in), TensorShape<4>(ish[0], ish[1], ish[2], C));
  OutTensorCPU<Out, 4> tout(static_cast<Out *>(out), TensorShape<4>(osh[0], osh[1], osh[2], C));
  k.Setup(ctx, tin, p);
  if (order) for (int i = 0; i < 3; i++) order[i] = k.setup.desc.order[i];
  k.Run(ctx, tout, tin, p);
  return 0;
}
}  // namespace

extern "C" {

// 3-D (DHWC) volumes: shape arrays [0] = depth, [1] = height, [2] = width (ResamplingParams3D order); order[3] in vec numbering (0 = x)
int ref_resample_dhwc(const void *in, int in_dtype, const int *in_shape, int C, void *out, int out_dtype, const int *out_shape,
                      const FilterDescC *minf, const FilterDescC *magf, const int *use_roi, const float *roi_start,
                      const float *roi_end, int *order) {
  try {
    if (in_dtype == 0 && out_dtype == 0)
      return RunResample3D<uint8_t, uint8_t>(in, in_shape, C, out, out_shape, minf, magf, use_roi, roi_start, roi_end, order);
    if (in_dtype == 0 && out_dtype == 1)
      return RunResample3D<float, uint8_t>(in, in_shape, C, out, out_shape, minf, magf, use_roi, roi_start, roi_end, order);
    if (in_dtype == 1 && out_dtype == 1)
      return RunResample3D<float, float>(in, in_shape, C, out, out_shape, minf, magf, use_roi, roi_start, roi_end, order);
    return -2;
  } catch (...) { return -1; }
}

// dtype: 0 = u8, 1 = f32 ; arrays indexed [0] = y, [1] = x (the reference's params order)
int ref_resample_hwc(const void *in, int in_dtype, int H, int W, int C, void *out, int out_dtype,
                     int outH, int outW, const FilterDescC *minf, const FilterDescC *magf,
                     const int *use_roi, const float *roi_start, const float *roi_end, int *order0) {
  try {
    if (in_dtype == 0 && out_dtype == 0)
      return RunResample<uint8_t, uint8_t>(in, H, W, C, out, outH, outW, minf, magf, use_roi, roi_start, roi_end, order0);
    if (in_dtype == 0 && out_dtype == 1)
      return RunResample<float, uint8_t>(in, H, W, C, out, outH, outW, minf, magf, use_roi, roi_start, roi_end, order0);
    if (in_dtype == 1 && out_dtype == 1)
      return RunResample<float, float>(in, H, W, C, out, outH, outW, minf, magf, use_roi, roi_start, roi_end, order0);
    return -2;
  } catch (...) { return -1; }
}

uint16_t ref_float2half(float x) {
  float16 h(x);
  uint16_t r;
  memcpy(&r, &h, 2);
  return r;
}

// out_dtype: 1 = f32, 2 = f16 ; layout_chw: 0 HWC, 1 CHW.  Window may exceed the image (padding).
int ref_cmn(const uint8_t *in, int H, int W, int C, void *out, int out_dtype, int layout_chw, int out_c,
            int ay, int ax, int ch, int cw, int mirror, const float *mean, const float *inv_std,
            const float *fill) {
  try {
    TensorShape<3> in_sh(H, W, C), sl_sh(ch, cw, out_c);
    SliceFlipNormalizePermutePadArgs<3> args(sl_sh, in_sh);
    args.anchor = TensorShape<3>(ay, ax, 0);
    args.flip[1] = mirror != 0;
    if (layout_chw) args.permuted_dims = { 2, 0, 1 };
    args.channel_dim = 2;
    if (mean) {
      args.mean.clear(); args.inv_stddev.clear();
      for (int c = 0; c < C; c++) { args.mean.push_back(mean[c]); args.inv_stddev.push_back(inv_std[c]); }
      // per-channel args must match the OUTPUT channel count (padding channels get mean 0 / scale 1)
      for (int c = C; c < out_c; c++) { args.mean.push_back(0.0f); args.inv_stddev.push_back(1.0f); }
    }
    args.fill_values.clear();
    for (int c = 0; c < out_c; c++) args.fill_values.push_back(fill ? fill[c] : 0.0f);
    KernelContext ctx;
    InTensorCPU<uint8_t, 3> tin(in, in_sh);
    TensorShape<3> out_sh = layout_chw ? TensorShape<3>(out_c, ch, cw) : TensorShape<3>(ch, cw, out_c);
    if (out_dtype == 1) {
      SliceFlipNormalizePermutePadCpu<float, uint8_t, 3> k;
      k.Setup(ctx, tin, args);
      k.Run(ctx, OutTensorCPU<float, 3>(static_cast<float *>(out), out_sh), tin, args);
    } else {
      SliceFlipNormalizePermutePadCpu<float16, uint8_t, 3> k;
      k.Setup(ctx, tin, args);
      k.Run(ctx, OutTensorCPU<float16, 3>(static_cast<float16 *>(out), out_sh), tin, args);
    }
    return 0;
  } catch (...) { return -1; }
}

// M: 2x3 row-major dst->src; interp 0 NN / 1 linear; border 0 clamp / 1 constant(fill[0..C))
int ref_warp_affine(const uint8_t *in, int H, int W, int C, void *out, int out_dtype, int outH, int outW,
                    const float *M, int interp, int border, const float *fill) {
  try {
    if (C != 3 && C != 1) return -2;
    KernelContext ctx;
    mat2x3 m;
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) m(i, j) = M[i * 3 + j];
    InTensorCPU<uint8_t, 3> tin(in, TensorShape<3>(H, W, C));
    TensorShape<2> osz(outH, outW);
    auto interp_t = interp ? DALI_INTERP_LINEAR : DALI_INTERP_NN;
    TensorShape<3> osh(outH, outW, C);
    auto run = [&](auto out_tag, auto border_val) {
      using Out = decltype(out_tag);
      using B = decltype(border_val);
      WarpCPU<AffineMapping2D, 2, Out, uint8_t, B> k;
      k.Setup(ctx, tin, m, osz, interp_t, border_val);
      k.Run(ctx, OutTensorCPU<Out, 3>(static_cast<Out *>(out), osh), tin, m, osz, interp_t, border_val);
    };
    if (border == 0) {
      if (out_dtype == 0) run(uint8_t(), BorderClamp()); else run(float(), BorderClamp());
    } else {
      // the operator passes a scalar fill_value converted to the input type (warp.h:268)
      uint8_t b = ConvertSat<uint8_t>(fill[0]);
      if (out_dtype == 0) run(uint8_t(), b); else run(float(), b);
    }
    return 0;
  } catch (...) { return -1; }
}

void ref_affine_inv(const float *M, float *Minv) {
  mat2x3 m;
  for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) m(i, j) = M[i * 3 + j];
  auto r = affine_mat_inv(m);
  for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) Minv[i * 3 + j] = r(i, j);
}

int ref_linear_transform(const uint8_t *in, int H, int W, void *out, int out_dtype, const float *M, const float *T) {
  try {
    KernelContext ctx;
    mat3 m; vec3 t;
    for (int i = 0; i < 3; i++) { t[i] = T[i]; for (int j = 0; j < 3; j++) m(i, j) = M[i * 3 + j]; }
    InTensorCPU<uint8_t, 3> tin(in, TensorShape<3>(H, W, 3));
    TensorShape<3> osh(H, W, 3);
    if (out_dtype == 0) {
      LinearTransformationCpu<uint8_t, uint8_t, 3, 3, 3> k;
      k.Setup(ctx, tin, m, t);
      k.Run(ctx, OutTensorCPU<uint8_t, 3>(static_cast<uint8_t *>(out), osh), tin, m, t);
    } else {
      LinearTransformationCpu<float, uint8_t, 3, 3, 3> k;
      k.Setup(ctx, tin, m, t);
      k.Run(ctx, OutTensorCPU<float, 3>(static_cast<float *>(out), osh), tin, m, t);
    }
    return 0;
  } catch (...) { return -1; }
}

// The matrix composition of dali/operators/image/color/color_twist.h:50-83,156-170, evaluated with
// the reference's own mat3 type (the operator header itself drags in the whole pipeline, so the
// seven-factor product is spelled out here with the same operand order and the same helpers).
void ref_color_twist_matrix(float hue, float saturation, float value, float brightness, float contrast,
                            float half_range, float *M, float *T) {
  const mat3 Rgb2Yiq = {{ {.299f, .587f, .114f}, {.596f, -.274f, -.321f}, {.211f, -.523f, .311f} }};
  const mat3 Yiq2Rgb = inverse(Rgb2Yiq);
  const float h_rad = hue * M_PI / 180;
  mat3 hm = mat3::eye();
  hm(1, 1) = cos(h_rad); hm(2, 2) = cos(h_rad); hm(1, 2) = sin(h_rad); hm(2, 1) = -sin(h_rad);
  mat3 sm = mat3::eye();
  sm(1, 1) = saturation; sm(2, 2) = saturation;
  mat3 r = mat3(brightness) * mat3(contrast) * Yiq2Rgb * hm * sm * mat3(value) * Rgb2Yiq;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) M[i * 3 + j] = r(i, j);
  float t = (half_range - half_range * contrast) * brightness;
  T[0] = T[1] = T[2] = t;
}

// in-tree BT.601 colour formulas (the OpenCV-backed RGB<->BGR<->GRAY cases are not covered here)
// type: 0 RGB, 2 GRAY, 3 YCbCr
int ref_csc_bt601(const uint8_t *in, size_t npix, uint8_t *out, int in_type, int out_type) {
  using namespace dali::kernels::color;  // NOLINT
  for (size_t p = 0; p < npix; p++) {
    if (in_type == 0 && out_type == 3) {
      vec<3, uint8_t> rgb = { in[3 * p], in[3 * p + 1], in[3 * p + 2] };
      out[3 * p] = itu_r_bt_601::rgb_to_y<uint8_t>(rgb);
      out[3 * p + 1] = itu_r_bt_601::rgb_to_cb<uint8_t>(rgb);
      out[3 * p + 2] = itu_r_bt_601::rgb_to_cr<uint8_t>(rgb);
    } else if (in_type == 3 && out_type == 0) {
      vec<3, uint8_t> y = { in[3 * p], in[3 * p + 1], in[3 * p + 2] };
      auto rgb = itu_r_bt_601::ycbcr_to_rgb<uint8_t>(y);
      out[3 * p] = rgb[0]; out[3 * p + 1] = rgb[1]; out[3 * p + 2] = rgb[2];
    } else if (in_type == 2 && out_type == 3) {
      out[3 * p] = itu_r_bt_601::gray_to_y<uint8_t>(in[p]); out[3 * p + 1] = 128; out[3 * p + 2] = 128;
    } else if (in_type == 3 && out_type == 2) {
      out[p] = itu_r_bt_601::y_to_gray<uint8_t>(in[3 * p]);
    } else if (in_type == 0 && out_type == 2) {   // GPU-kernel formula (jpeg luma), for information
      vec<3, uint8_t> rgb = { in[3 * p], in[3 * p + 1], in[3 * p + 2] };
      out[p] = rgb_to_gray<uint8_t>(rgb);
    } else {
      return -2;
    }
  }
  return 0;
}

}  // extern "C"
