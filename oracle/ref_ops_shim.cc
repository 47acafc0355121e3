// oracle/ref_ops_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" shim around further pieces of the reference, compiled from /root/reference where they lie (oracle/Makefile):
//   dali/operators/image/crop/random_crop_generator_util.{h,cc} + dali/core/random/philox.{h,cc}
//        RandomCropGenerator with the per-sample Philox states of RandomCropAttr (random_crop_attr.h:36-72)
//   dali/kernels/imgproc/color_manipulation/color_space_conversion_impl.h + include/dali/core/convert.h
//        the per-pixel functors the image decoder's post-conversion uses (dali/operators/imgcodec/util/convert.h:118-205:
//        ConvertPixelDType / ConvertPixel<Out, In, out_format, in_format>); convert.h itself needs the absent nvimgcodec headers,
//        so the thin dispatch (format switch, BGR = reversed channel order, :277-289) is restated here around the reference's functors.
#include <cstdint>
#include <cstring>
#include <vector>

#include "dali/core/convert.h"
#include "dali/core/geom/vec.h"
#include "dali/core/random/philox.h"
#include "dali/kernels/imgproc/color_manipulation/color_space_conversion_impl.h"
#include "dali/operators/image/crop/random_crop_generator_util.h"
#include "dali/core/geom/transform.h"
#include "dali/core/math_util.h"
#include "dali/operators/image/remap/rotate_params.h"
#include "dali/operators/image/color/brightness_contrast.h"

using namespace dali;  // NOLINT

// in: npix pixels, in_c = 3 (RGB) or 1 (GRAY) u8; out_type: 0 RGB, 1 BGR, 2 GRAY, 3 YCbCr; out_float: 0 = u8, 1 = f32
template <typename Out>
static void ConvertPixels(const uint8_t *in, size_t npix, int in_c, int out_type, Out *out) {
  using namespace kernels::color;  // NOLINT
  for (size_t i = 0; i < npix; i++) {
    if (in_c == 1) {
      const uint8_t g = in[i];
      if (out_type == 2) { out[i] = ConvertSatNorm<Out>(g); }
      else if (out_type == 3) { auto v = itu_r_bt_601::gray_to_ycbcr<Out, uint8_t>(g); for (int c = 0; c < 3; c++) out[3 * i + c] = v[c]; }
      else { const Out v = ConvertSatNorm<Out>(g); out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = v; }
    } else {
      vec<3, uint8_t> rgb{in[3 * i], in[3 * i + 1], in[3 * i + 2]};
      if (out_type == 2) { out[i] = rgb_to_gray<Out, uint8_t>(rgb); }
      else if (out_type == 3) { auto v = itu_r_bt_601::rgb_to_ycbcr<Out, uint8_t>(rgb); for (int c = 0; c < 3; c++) out[3 * i + c] = v[c]; }
      else {
        for (int c = 0; c < 3; c++) out[3 * i + (out_type == 1 ? 2 - c : c)] = ConvertSatNorm<Out>(rgb[c]);
      }
    }
  }
}

extern "C" {

// windows[k] = (anchor_y, anchor_x, h, w) of the k-th call for sample `sample_idx` of an operator constructed with `seed`
int ref_random_crop(int64_t seed, int sample_idx, int H, int W, float ar_lo, float ar_hi, float area_lo, float area_hi, int num_attempts,
                    int ncalls, int *windows) {
  const uint64_t key = static_cast<uint64_t>(seed) ^ 0x12345678abcdefeull;          // kRandomCropSeedModifier
  RandomCropGenerator gen({ar_lo, ar_hi}, {area_lo, area_hi}, Philox4x32_10::State(key, 65537ull * sample_idx, 0), num_attempts);
  for (int k = 0; k < ncalls; k++) {
    CropWindow w = gen.GenerateCropWindow(TensorShape<>{H, W});
    windows[4 * k] = static_cast<int>(w.anchor[0]); windows[4 * k + 1] = static_cast<int>(w.anchor[1]);
    windows[4 * k + 2] = static_cast<int>(w.shape[0]); windows[4 * k + 3] = static_cast<int>(w.shape[1]);
  }
  return 0;
}

int ref_decoder_convert(const uint8_t *in, size_t npix, int in_c, int out_type, int out_float, void *out) {
  if (out_float) ConvertPixels<float>(in, npix, in_c, out_type, static_cast<float *>(out));
  else ConvertPixels<uint8_t>(in, npix, in_c, out_type, static_cast<uint8_t *>(out));
  return 0;
}


// Rotate: output canvas (rotate_params.h:36-55 RotatedCanvasSize + the parity vote of InferSize :279-297 for one frame) and the
// destination -> source matrix of AdjustParams (:222-236), built from the reference's own geom/transform.h functions.
int ref_rotate_params(float angle_deg, int in_h, int in_w, int keep_size, const float *size_hw, int *out_hw, float *M) {
  float neg = -angle_deg;                                  // SetParams(): 2-D angles are negated
  int oh, ow;
  if (size_hw) { oh = std::max<int>(static_cast<int>(std::roundf(size_hw[0])), 1); ow = std::max<int>(static_cast<int>(std::roundf(size_hw[1])), 1); }
  else if (keep_size) { oh = in_h; ow = in_w; }
  else {
    ivec2 shape, parity;
    std::tie(shape, parity) = RotatedCanvasSize(TensorShape<2>(in_h, in_w), deg2rad(neg));
    ivec2 acc_shape = shape, acc_parity = parity;
    const int num_frames = 1;
    acc_shape += (acc_shape % 2) ^ (2 * acc_parity > num_frames);
    ow = acc_shape[0]; oh = acc_shape[1];
  }
  ivec2 in_size(in_w, in_h), out_size(ow, oh);
  float a = deg2rad(neg);
  mat3 T = translation(in_size * 0.5f) * rotation2D(-a) * translation(-out_size * 0.5f);
  auto P = sub<2, 3>(T);
  for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) M[r * 3 + c] = P(r, c);
  out_hw[0] = oh; out_hw[1] = ow;
  return 0;
}

// BrightnessContrast: OpArgsToKernelArgs<Out, uint8_t> (brightness_contrast.h:84-103; protected member, built with
// -fno-access-control) and the element functor of the CPU kernel (multiply_add.h:47-60)
struct BcProbe : BrightnessContrastOp<CPUBackend> {
  using BrightnessContrastOp<CPUBackend>::OpArgsToKernelArgs;
};
int ref_brightness_contrast(const uint8_t *in, size_t n, int out_float, float brightness, float shift, float contrast, float center, void *out) {
  float addend, multiplier;
  BcProbe *probe = nullptr;                                // OpArgsToKernelArgs touches no member
  if (out_float) probe->OpArgsToKernelArgs<float, uint8_t>(addend, multiplier, brightness, shift, contrast, center);
  else probe->OpArgsToKernelArgs<uint8_t, uint8_t>(addend, multiplier, brightness, shift, contrast, center);
  for (size_t i = 0; i < n; i++) {
    if (out_float) static_cast<float *>(out)[i] = ConvertSat<float>(in[i] * multiplier + addend);
    else static_cast<uint8_t *>(out)[i] = ConvertSat<uint8_t>(in[i] * multiplier + addend);
  }
  return 0;
}

}  // extern "C"
