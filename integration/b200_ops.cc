// integration/b200_ops.cc -- the operator plugin a DALI maintainer adds to run the hot path on libdali_b200.so.
//
// Written against the REFERENCE's own headers (dali/pipeline/operator/operator.h, crop_attr.h, resize_attr.h, resampling_attr.h, ...):
// the argument handling is the reference's (CropAttr, ResizeAttr, ResamplingFilterAttr, OpSpec::GetArgument), only the kernel call
// changes -- to the C-ABI of include/dali_b200.h.  Built as a normal DALI plugin and loaded with
// nvidia.dali.plugin_manager.load_library; the operators register under the `b200` prefix (fn.b200.resize, ...), because the stock
// names are taken in a stock build (operator_factory.h:55-58).
// tests/test_integration_cpu.py compiles this file with `g++ -std=c++20 -fsyntax-only` against /root/reference where it exists.
#include <string>
#include <vector>

#include "dali/core/static_switch.h"
#include "dali/kernels/imgproc/resample/params.h"
#include "dali/operators/image/crop/crop_attr.h"
#include "dali/operators/image/crop/random_crop_attr.h"
#include "dali/operators/image/resize/resampling_attr.h"
#include "dali/operators/image/resize/resize_attr.h"
#include "dali/operators/audio/mfcc/mfcc.h"
#include "dali/operators/audio/nonsilence_op.h"
#include "dali/operators/audio/resample.h"
#include "dali/operators/generic/flip.h"
#include "dali/operators/generic/slice/slice_attr.h"
#include "dali/operators/image/remap/rotate_params.h"
#include "dali/operators/image/color/brightness_contrast.h"
#include "dali/operators/image/color/color_twist.h"
#include "dali/operators/signal/decibel/to_decibels_op.h"
#include "dali/pipeline/operator/operator.h"

#include "dali_b200.h"

namespace b200 {
using namespace dali;  // NOLINT

inline void Check(int rc, const char *what) {
  if (rc != DALIB200_SUCCESS) DALI_FAIL(make_string(what, ": ", dalib200GetLastError()));
}

template <typename TL>
inline std::vector<const void *> InPtrs(const TL &tl) {
  std::vector<const void *> p(tl.num_samples());
  for (int i = 0; i < tl.num_samples(); i++) p[i] = tl.raw_tensor(i);
  return p;
}
template <typename TL>
inline std::vector<void *> OutPtrs(TL &tl) {
  std::vector<void *> p(tl.num_samples());
  for (int i = 0; i < tl.num_samples(); i++) p[i] = tl.raw_mutable_tensor(i);
  return p;
}

// ------------------------------------------------------------------------------------------------ decoders.image (mixed)
class ImageDecoder : public Operator<MixedBackend> {
 public:
  explicit ImageDecoder(const OpSpec &spec) : Operator<MixedBackend>(spec) {
    prm_.output_type = static_cast<int>(spec.GetArgument<DALIImageType>("output_type"));
    prm_.dtype = spec.GetArgument<DALIDataType>("dtype") == DALI_FLOAT ? DALIB200_FLOAT : DALIB200_UINT8;
    prm_.fancy_upsampling = spec.GetArgument<bool>("jpeg_fancy_upsampling");
    prm_.adjust_orientation = spec.GetArgument<bool>("adjust_orientation");
    Check(dalib200JpegPlanCreate(&plan_, max_batch_size_), "decoders.image");
  }
  ~ImageDecoder() override { dalib200JpegPlanDestroy(plan_); }
  bool HasContiguousOutputs() const override { return true; }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<CPUBackend>(0);
    const int n = in.num_samples();
    std::vector<const uint8_t *> ptrs(n);
    std::vector<size_t> lens(n);
    for (int i = 0; i < n; i++) {
      ptrs[i] = static_cast<const uint8_t *>(in.raw_tensor(i));
      lens[i] = static_cast<size_t>(in.tensor_shape(i).num_elements());
    }
    Check(dalib200JpegPlanSetupEx(plan_, n, ptrs.data(), lens.data(), &prm_, nullptr), "decoders.image");
    out.resize(1);
    out[0].type = prm_.dtype == DALIB200_FLOAT ? DALI_FLOAT : DALI_UINT8;
    out[0].shape.resize(n, 3);
    for (int i = 0; i < n; i++) {
      int32_t hwc[3];
      Check(dalib200JpegPlanGetOutputShape(plan_, i, hwc), "decoders.image");
      out[0].shape.set_tensor_shape(i, TensorShape<>{hwc[0], hwc[1], hwc[2]});
    }
    return true;                                     // the executor allocates the output (exec_node_task.cc:296-307)
  }
  void RunImpl(Workspace &ws) override {
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout("HWC");
    auto op = OutPtrs(out);
    Check(dalib200JpegUpload(plan_, ws.stream()), "decoders.image");
    Check(dalib200JpegLaunch(plan_, op.data(), ws.stream()), "decoders.image");      // enqueue only, no host sync
  }

 private:
  dalib200JpegPlan *plan_ = nullptr;
  dalib200JpegParams prm_{};
};

// ------------------------------------------------------------------------------------------------ decoders.image_crop / image_random_crop
// Region-of-interest decode: the window comes from the reference's OWN attribute classes (CropAttr: crop / crop_pos_x / crop_pos_y ...;
// RandomCropAttr: the Philox-based RandomCropGenerator, so a given seed yields the reference's windows), in the coordinates of the
// oriented image (imgcodec.h:26-44); the library decodes only the MCUs the window and the upsampling taps touch.
template <typename WindowAttr>
class ImageDecoderRoi : public Operator<MixedBackend> {
 public:
  explicit ImageDecoderRoi(const OpSpec &spec) : Operator<MixedBackend>(spec), attr_(spec) {
    prm_.output_type = static_cast<int>(spec.GetArgument<DALIImageType>("output_type"));
    prm_.dtype = spec.GetArgument<DALIDataType>("dtype") == DALI_FLOAT ? DALIB200_FLOAT : DALIB200_UINT8;
    prm_.fancy_upsampling = spec.GetArgument<bool>("jpeg_fancy_upsampling");
    prm_.adjust_orientation = spec.GetArgument<bool>("adjust_orientation");
    Check(dalib200JpegPlanCreate(&plan_, max_batch_size_), "decoders.image_crop");
  }
  ~ImageDecoderRoi() override { dalib200JpegPlanDestroy(plan_); }
  bool HasContiguousOutputs() const override { return true; }

 protected:
  virtual void AcquireWindowArgs(const Workspace &ws) {}
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<CPUBackend>(0);
    const int n = in.num_samples();
    AcquireWindowArgs(ws);
    std::vector<const uint8_t *> ptrs(n);
    std::vector<size_t> lens(n);
    std::vector<dalib200JpegRoi> rois(n);
    for (int i = 0; i < n; i++) {
      ptrs[i] = static_cast<const uint8_t *>(in.raw_tensor(i));
      lens[i] = static_cast<size_t>(in.tensor_shape(i).num_elements());
      dalib200JpegInfo info;
      Check(dalib200JpegGetInfo(ptrs[i], lens[i], &info), "decoders.image_crop");
      int64_t H = info.height, W = info.width;
      if (prm_.adjust_orientation && info.orientation >= 5) std::swap(H, W);       // image_decoder.h:678-681
      CropWindow win = attr_.GetCropWindowGenerator(i)(TensorShape<>{H, W}, "HW");
      win.EnforceInRange(TensorShape<>{H, W});                                      // roi_image_decoder.h:36-40
      rois[i] = { 1, static_cast<int32_t>(win.anchor[1]), static_cast<int32_t>(win.anchor[0]),
                  static_cast<int32_t>(win.anchor[1] + win.shape[1]), static_cast<int32_t>(win.anchor[0] + win.shape[0]), 0 };
    }
    Check(dalib200JpegPlanSetupEx(plan_, n, ptrs.data(), lens.data(), &prm_, rois.data()), "decoders.image_crop");
    out.resize(1);
    out[0].type = prm_.dtype == DALIB200_FLOAT ? DALI_FLOAT : DALI_UINT8;
    out[0].shape.resize(n, 3);
    for (int i = 0; i < n; i++) {
      int32_t hwc[3];
      Check(dalib200JpegPlanGetOutputShape(plan_, i, hwc), "decoders.image_crop");
      out[0].shape.set_tensor_shape(i, TensorShape<>{hwc[0], hwc[1], hwc[2]});
    }
    return true;
  }
  void RunImpl(Workspace &ws) override {
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout("HWC");
    auto op = OutPtrs(out);
    Check(dalib200JpegUpload(plan_, ws.stream()), "decoders.image_crop");
    Check(dalib200JpegLaunch(plan_, op.data(), ws.stream()), "decoders.image_crop");
  }

  WindowAttr attr_;
  dalib200JpegPlan *plan_ = nullptr;
  dalib200JpegParams prm_{};
};

class ImageDecoderCrop : public ImageDecoderRoi<CropAttr> {
 public:
  using ImageDecoderRoi<CropAttr>::ImageDecoderRoi;
 protected:
  void AcquireWindowArgs(const Workspace &ws) override { attr_.ProcessArguments(spec_, ws); }      // crop_attr.cc:100-245
};
using ImageDecoderRandomCrop = ImageDecoderRoi<RandomCropAttr>;
class ImageDecoderSlice : public ImageDecoderRoi<SliceAttr> {                                     // roi_image_decoder.h:63-78
 public:
  using ImageDecoderRoi<SliceAttr>::ImageDecoderRoi;
 protected:
  void AcquireWindowArgs(const Workspace &ws) override { attr_.ProcessArguments(spec_, ws); }      // slice_attr.h:361-381
};

// ------------------------------------------------------------------------------------------------ Resize
// Images and volumes: the reference's ResizeAttr parses the layout (spatial_ndim_ 2 or 3, first_spatial_dim_), computes sizes and ROIs
// (resize_attr.cc:102-259); the dimensions in front of the spatial ones are frames, those behind them channels (resize_op_impl.h:56-101).
class Resize : public Operator<GPUBackend> {
 public:
  explicit Resize(const OpSpec &spec) : Operator<GPUBackend>(spec) {
    Check(dalib200ResamplePlanCreate(&plan_, max_batch_size_ * 64), "Resize");
    Check(dalib200Resample3DPlanCreate(&plan3_, max_batch_size_ * 16), "Resize");
  }
  ~Resize() override { dalib200ResamplePlanDestroy(plan_); dalib200Resample3DPlanDestroy(plan3_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    const int n = in.num_samples();
    const auto &shape = in.shape();
    resize_attr_.PrepareResizeParams(spec_, ws, shape, in.GetLayout());            // resize_attr.cc:178-259 (reference code)
    resampling_attr_.PrepareFilterParams(spec_, ws, n);                             // resampling_attr.cc:76-133
    const int sd = resize_attr_.spatial_ndim_, fs = resize_attr_.first_spatial_dim_, nd = shape.sample_dim();
    DALI_ENFORCE(sd == 2 || sd == 3, "Resize: 2 or 3 spatial dimensions expected");
    std::vector<kernels::ResamplingParams> rp(static_cast<size_t>(n) * sd);
    resampling_attr_.GetResamplingParams(make_span(rp), make_cspan(resize_attr_.params_));
    out.resize(1);
    out[0].type = resampling_attr_.GetOutputType(in.type());
    resize_attr_.GetResizedShape(out[0].shape, shape);
    samples_.clear(); samples3_.clear(); frame_sample_.clear(); frame_in_elems_.clear(); frame_out_elems_.clear();
    for (int i = 0; i < n; i++) {
      auto sh = shape.tensor_shape_span(i);
      auto osh = out[0].shape.tensor_shape_span(i);
      int64_t frames = 1, channels = 1, in_vol = 1, out_vol = 1;
      for (int d = 0; d < fs; d++) frames *= sh[d];
      for (int d = fs + sd; d < nd; d++) channels *= sh[d];
      for (int d = fs; d < nd; d++) { in_vol *= sh[d]; out_vol *= osh[d]; }
      dalib200ResampleSample s2{};
      dalib200Resample3DSample s3{};
      for (int d = 0; d < sd; d++) {
        const auto &p = rp[static_cast<size_t>(sd) * i + d];
        const dalib200FilterDesc fmin = { static_cast<int>(p.min_filter.type), p.min_filter.antialias, p.min_filter.radius };
        const dalib200FilterDesc fmag = { static_cast<int>(p.mag_filter.type), p.mag_filter.antialias, p.mag_filter.radius };
        if (sd == 2) {
          s2.use_roi[d] = p.roi.use_roi; s2.roi_start[d] = p.roi.start; s2.roi_end[d] = p.roi.end; s2.min_filter[d] = fmin; s2.mag_filter[d] = fmag;
        } else {
          s3.in_shape[d] = static_cast<int>(sh[fs + d]); s3.out_shape[d] = p.output_size;
          s3.use_roi[d] = p.roi.use_roi; s3.roi_start[d] = p.roi.start; s3.roi_end[d] = p.roi.end; s3.min_filter[d] = fmin; s3.mag_filter[d] = fmag;
        }
      }
      if (sd == 2) {
        s2.in_h = static_cast<int>(sh[fs]); s2.in_w = static_cast<int>(sh[fs + 1]); s2.channels = static_cast<int>(channels);
        s2.out_h = rp[2 * i].output_size; s2.out_w = rp[2 * i + 1].output_size;
      } else {
        s3.channels = static_cast<int>(channels);
      }
      for (int64_t k = 0; k < frames; k++) {
        if (sd == 2) samples_.push_back(s2); else samples3_.push_back(s3);
        frame_sample_.push_back(i); frame_in_elems_.push_back(k * in_vol); frame_out_elems_.push_back(k * out_vol);
      }
    }
    volumes_ = sd == 3;
    const int it = in.type() == DALI_UINT8 ? DALIB200_UINT8 : DALIB200_FLOAT, ot = out[0].type == DALI_UINT8 ? DALIB200_UINT8 : DALIB200_FLOAT;
    if (volumes_) Check(dalib200Resample3DPlanSetup(plan3_, static_cast<int>(samples3_.size()), samples3_.data(), it, ot), "Resize");
    else Check(dalib200ResamplePlanSetup(plan_, static_cast<int>(samples_.size()), samples_.data(), it, ot), "Resize");
    in_esize_ = TypeTable::GetTypeInfo(in.type()).size(); out_esize_ = TypeTable::GetTypeInfo(out[0].type).size();
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    std::vector<const void *> ip(frame_sample_.size());
    std::vector<void *> op(frame_sample_.size());
    for (size_t k = 0; k < frame_sample_.size(); k++) {
      ip[k] = static_cast<const uint8_t *>(in.raw_tensor(frame_sample_[k])) + frame_in_elems_[k] * in_esize_;
      op[k] = static_cast<uint8_t *>(out.raw_mutable_tensor(frame_sample_[k])) + frame_out_elems_[k] * out_esize_;
    }
    if (volumes_) Check(dalib200Resample3DLaunch(plan3_, ip.data(), op.data(), ws.stream()), "Resize");
    else Check(dalib200ResampleLaunch(plan_, ip.data(), op.data(), ws.stream()), "Resize");
  }

 private:
  ResizeAttr resize_attr_;
  ResamplingFilterAttr resampling_attr_;
  dalib200ResamplePlan *plan_ = nullptr;
  dalib200Resample3DPlan *plan3_ = nullptr;
  bool volumes_ = false;
  size_t in_esize_ = 1, out_esize_ = 1;
  std::vector<dalib200ResampleSample> samples_;
  std::vector<dalib200Resample3DSample> samples3_;
  std::vector<int> frame_sample_;
  std::vector<int64_t> frame_in_elems_, frame_out_elems_;
};

// ------------------------------------------------------------------------------------------------ CropMirrorNormalize
class CropMirrorNormalize : public Operator<GPUBackend> {
 public:
  explicit CropMirrorNormalize(const OpSpec &spec) : Operator<GPUBackend>(spec), crop_attr_(spec) {
    out_type_ = spec.GetArgument<DALIDataType>("dtype");
    pad_output_ = spec.GetArgument<bool>("pad_output");
    mean_ = spec.GetRepeatedArgument<float>("mean");
    std_ = spec.GetRepeatedArgument<float>("std");
    scale_ = spec.GetArgument<float>("scale"); shift_ = spec.GetArgument<float>("shift");
    Check(dalib200CmnPlanCreate(&plan_, max_batch_size_), "CropMirrorNormalize");
  }
  ~CropMirrorNormalize() override { dalib200CmnPlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    const int n = in.num_samples();
    crop_attr_.ProcessArguments(spec_, ws);                                        // crop_attr.cc:100-245 (reference code)
    samples_.resize(n);
    out.resize(1);
    out[0].type = out_type_;
    out[0].shape.resize(n, 3);
    int out_c = 3;
    for (int i = 0; i < n; i++) {
      auto sh = in.tensor_shape(i);                                                // HWC
      auto win = crop_attr_.GetCropWindowGenerator(i)(sh, "HWC");
      auto &s = samples_[i];
      s.in_h = static_cast<int>(sh[0]); s.in_w = static_cast<int>(sh[1]); s.channels = static_cast<int>(sh[2]);
      s.anchor_y = static_cast<int>(win.anchor[0]); s.anchor_x = static_cast<int>(win.anchor[1]);
      s.crop_h = static_cast<int>(win.shape[0]); s.crop_w = static_cast<int>(win.shape[1]);
      s.mirror = spec_.GetArgument<int>("mirror", &ws, i);
      out_c = s.channels;
      if (pad_output_) { out_c = 1; while (out_c < s.channels) out_c *= 2; }       // crop_mirror_normalize.h:69-77
      for (int d = 0; d < 4; d++) {
        if (d < s.channels) {                                                      // crop_mirror_normalize.h:135-141
          const double mean_val = mean_[d % mean_.size()], std_val = std_[d % std_.size()];
          s.mean[d] = static_cast<float>(std::fma(-static_cast<double>(shift_), std_val / scale_, mean_val));
          s.inv_std[d] = static_cast<float>(scale_ / std_val);
        } else { s.mean[d] = 0.0f; s.inv_std[d] = 1.0f; }
        s.fill[d] = 0.0f;
      }
      out[0].shape.set_tensor_shape(i, TensorShape<>{out_c, win.shape[0], win.shape[1]});
    }
    Check(dalib200CmnPlanSetup(plan_, n, samples_.data(), out_type_ == DALI_FLOAT16 ? DALIB200_FLOAT16 : DALIB200_FLOAT,
                               DALIB200_LAYOUT_CHW, out_c), "CropMirrorNormalize");
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout("CHW");
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200CmnLaunch(plan_, ip.data(), op.data(), ws.stream()), "CropMirrorNormalize");
  }

 private:
  CropAttr crop_attr_;
  dalib200CmnPlan *plan_ = nullptr;
  std::vector<dalib200CmnSample> samples_;
  std::vector<float> mean_, std_;
  float scale_ = 1, shift_ = 0;
  DALIDataType out_type_ = DALI_FLOAT;
  bool pad_output_ = false;
};

// ------------------------------------------------------------------------------------------------ WarpAffine
class WarpAffine : public Operator<GPUBackend> {
 public:
  explicit WarpAffine(const OpSpec &spec) : Operator<GPUBackend>(spec) {
    interp_ = spec.GetArgument<DALIInterpType>("interp_type") == DALI_INTERP_LINEAR;
    invert_ = !spec.GetArgument<bool>("inverse_map");
    use_fill_ = spec.TryGetArgument(fill_, "fill_value");
    Check(dalib200WarpPlanCreate(&plan_, max_batch_size_), "WarpAffine");
  }
  ~WarpAffine() override { dalib200WarpPlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    const int n = in.num_samples();
    samples_.resize(n);
    out.resize(1);
    out[0].type = DALI_UINT8;
    out[0].shape = in.shape();
    for (int i = 0; i < n; i++) {
      auto sh = in.tensor_shape(i);
      std::vector<float> m;
      GetGeneralizedArg<float>(make_span(m = std::vector<float>(6)), "matrix", i, spec_, ws);
      auto &s = samples_[i];
      s.in_h = s.out_h = static_cast<int>(sh[0]); s.in_w = s.out_w = static_cast<int>(sh[1]); s.channels = static_cast<int>(sh[2]);
      if (invert_) dalib200AffineInverse(m.data(), s.matrix);                      // warp_affine_params.h:50-83
      else for (int k = 0; k < 6; k++) s.matrix[k] = m[k];
    }
    Check(dalib200WarpPlanSetup(plan_, n, samples_.data(), interp_, use_fill_, fill_, DALIB200_UINT8), "WarpAffine");
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200WarpLaunch(plan_, ip.data(), op.data(), ws.stream()), "WarpAffine");
  }

 private:
  dalib200WarpPlan *plan_ = nullptr;
  std::vector<dalib200WarpSample> samples_;
  bool interp_ = true, invert_ = false, use_fill_ = false;
  float fill_ = 0;
};

// ------------------------------------------------------------------------------------------------ Hsv / ColorSpaceConversion
class PointwiseBase : public Operator<GPUBackend> {
 public:
  explicit PointwiseBase(const OpSpec &spec) : Operator<GPUBackend>(spec) {
    Check(dalib200PointwisePlanCreate(&plan_, max_batch_size_), "pointwise");
  }
  ~PointwiseBase() override { dalib200PointwisePlanDestroy(plan_); }

 protected:
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200PointwiseLaunch(plan_, ip.data(), op.data(), ws.stream()), "pointwise");
  }
  dalib200PointwisePlan *plan_ = nullptr;
};

class Hsv : public PointwiseBase {
 public:
  explicit Hsv(const OpSpec &spec) : PointwiseBase(spec) {}

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    const int n = in.num_samples();
    std::vector<dalib200ColorSample> cs(n);
    for (int i = 0; i < n; i++) {
      cs[i].num_pixels = in.tensor_shape(i).num_elements() / 3;
      dalib200ColorTwistMatrix(spec_.GetArgument<float>("hue", &ws, i), spec_.GetArgument<float>("saturation", &ws, i),
                               spec_.GetArgument<float>("value", &ws, i), 1.0f, 1.0f, 128.0f, cs[i].matrix, cs[i].offset);
    }
    Check(dalib200LinearTransformSetup(plan_, n, cs.data(), DALIB200_UINT8), "Hsv");
    out.resize(1);
    out[0].shape = in.shape(); out[0].type = DALI_UINT8;
    return true;
  }
};

class ColorSpaceConversion : public PointwiseBase {
 public:
  explicit ColorSpaceConversion(const OpSpec &spec) : PointwiseBase(spec) {
    in_t_ = static_cast<int>(spec.GetArgument<DALIImageType>("image_type"));
    out_t_ = static_cast<int>(spec.GetArgument<DALIImageType>("output_type"));
  }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    const int n = in.num_samples(), nd = in.shape().sample_dim();
    const int ic = in_t_ == DALI_GRAY ? 1 : 3, oc = out_t_ == DALI_GRAY ? 1 : 3;
    std::vector<int64_t> npx(n);
    out.resize(1);
    out[0].type = DALI_UINT8;
    out[0].shape = in.shape();
    for (int i = 0; i < n; i++) {
      npx[i] = in.tensor_shape(i).num_elements() / ic;
      auto sh = in.tensor_shape(i);
      sh[nd - 1] = oc;
      out[0].shape.set_tensor_shape(i, sh);
    }
    Check(dalib200ColorSpaceSetup(plan_, n, npx.data(), in_t_, out_t_), "ColorSpaceConversion");
    return true;
  }

 private:
  int in_t_ = 0, out_t_ = 0;
};

// ------------------------------------------------------------------------------------------------ Spectrogram / MelFilterBank
class Spectrogram : public Operator<GPUBackend> {
 public:
  explicit Spectrogram(const OpSpec &spec) : Operator<GPUBackend>(spec) {
    args_.window_length = spec.GetArgument<int>("window_length");
    args_.window_step = spec.GetArgument<int>("window_step");
    args_.power = spec.GetArgument<int>("power");
    args_.nfft = spec.HasArgument("nfft") ? spec.GetArgument<int>("nfft") : args_.window_length;
    args_.center = spec.GetArgument<bool>("center_windows");
    args_.reflect = spec.GetArgument<bool>("reflect_padding");
    args_.layout_ft = spec.GetArgument<TensorLayout>("layout") == TensorLayout("ft");
    if (spec.HasArgument("window_fn")) window_ = spec.GetRepeatedArgument<float>("window_fn");
    Check(dalib200SpectrogramPlanCreate(&plan_, max_batch_size_), "Spectrogram");
  }
  ~Spectrogram() override { dalib200SpectrogramPlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    const int n = in.num_samples();
    std::vector<int64_t> lens(n);
    for (int i = 0; i < n; i++) lens[i] = in.tensor_shape(i).num_elements();
    Check(dalib200SpectrogramPlanSetup(plan_, &args_, window_.empty() ? nullptr : window_.data(), n, lens.data()), "Spectrogram");
    out.resize(1);
    out[0].type = DALI_FLOAT;
    out[0].shape.resize(n, 2);
    const int64_t nbin = args_.nfft / 2 + 1;
    for (int i = 0; i < n; i++) {
      const int64_t nw = dalib200SpectrogramNumWindows(plan_, i);
      out[0].shape.set_tensor_shape(i, args_.layout_ft ? TensorShape<>{nbin, nw} : TensorShape<>{nw, nbin});
    }
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(args_.layout_ft ? "ft" : "tf");
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200SpectrogramLaunch(plan_, ip.data(), op.data(), ws.stream()), "Spectrogram");
  }

 private:
  dalib200SpectrogramPlan *plan_ = nullptr;
  dalib200SpectrogramArgs args_{};
  std::vector<float> window_;
};

class MelFilterBank : public Operator<GPUBackend> {
 public:
  explicit MelFilterBank(const OpSpec &spec) : Operator<GPUBackend>(spec) {
    args_.nfilter = spec.GetArgument<int>("nfilter");
    args_.sample_rate = spec.GetArgument<float>("sample_rate");
    args_.freq_low = spec.GetArgument<float>("freq_low"); args_.freq_high = spec.GetArgument<float>("freq_high");
    args_.normalize = spec.GetArgument<bool>("normalize");
    args_.htk = spec.GetArgument<std::string>("mel_formula") == "htk";
    Check(dalib200MelPlanCreate(&plan_, max_batch_size_), "MelFilterBank");
  }
  ~MelFilterBank() override { dalib200MelPlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    const int n = in.num_samples();
    std::vector<int64_t> nwin(n);
    const int nbin = n ? static_cast<int>(in.tensor_shape(0)[0]) : 2;
    out.resize(1);
    out[0].type = DALI_FLOAT;
    out[0].shape.resize(n, 2);
    for (int i = 0; i < n; i++) {
      nwin[i] = in.tensor_shape(i)[1];
      out[0].shape.set_tensor_shape(i, TensorShape<>{args_.nfilter, nwin[i]});
    }
    Check(dalib200MelPlanSetup(plan_, &args_, nbin, n, nwin.data()), "MelFilterBank");
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout("ft");
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200MelLaunch(plan_, ip.data(), op.data(), ws.stream()), "MelFilterBank");
  }

 private:
  dalib200MelPlan *plan_ = nullptr;
  dalib200MelArgs args_{};
};

// ------------------------------------------------------------------------------------------------ AudioResample / NonsilentRegion
// These two derive from the reference's OWN operator bases: argument handling, shape inference and error messages are the reference's
// code (audio::ResampleBase::SetupImpl / CalculateShapeAndArgs, NonsilenceOperator::SetupImpl / AcquireArgs); only RunImpl changes.
class AudioResample : public audio::ResampleBase<GPUBackend> {
 public:
  using Base = audio::ResampleBase<GPUBackend>;
  explicit AudioResample(const OpSpec &spec) : Base(spec) { Check(dalib200SignalPlanCreate(&plan_, max_batch_size_), "AudioResample"); }
  ~AudioResample() override { dalib200SignalPlanDestroy(plan_); }

 protected:
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    DALI_ENFORCE(in.type() == DALI_FLOAT && out.type() == DALI_FLOAT, "b200 AudioResample: float input and output only");
    out.SetLayout(in.GetLayout());
    const int n = in.num_samples();
    std::vector<dalib200AudioResampleSample> s(n);
    for (int i = 0; i < n; i++) {
      const auto ish = in.tensor_shape(i), osh = out.tensor_shape(i);
      s[i].in_rate = args_[i].in_rate; s[i].out_rate = args_[i].out_rate;          // filled by ResampleBase::CalculateShapeAndArgs
      s[i].in_length = ish[0]; s[i].out_length = osh[0];
      s[i].channels = ish.sample_dim() > 1 ? static_cast<int>(ish[1]) : 1;
    }
    Check(dalib200AudioResampleSetup(plan_, n, s.data(), quality_), "AudioResample");
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200SignalLaunch(plan_, ip.data(), op.data(), ws.stream()), "AudioResample");
  }

 private:
  dalib200SignalPlan *plan_ = nullptr;
};

class NonsilentRegion : public NonsilenceOperator<GPUBackend> {
 public:
  explicit NonsilentRegion(const OpSpec &spec) : NonsilenceOperator<GPUBackend>(spec) {
    Check(dalib200SignalPlanCreate(&plan_, max_batch_size_), "NonsilentRegion");
  }
  ~NonsilentRegion() override { dalib200SignalPlanDestroy(plan_); }

 protected:
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &begin = ws.Output<GPUBackend>(0);
    auto &length = ws.Output<GPUBackend>(1);
    DALI_ENFORCE(in.type() == DALI_FLOAT, "b200 NonsilentRegion: float input only");
    const int n = in.num_samples();
    std::vector<int64_t> len(n);
    std::vector<dalib200NonsilentSample> a(n);
    for (int i = 0; i < n; i++) {
      len[i] = in.tensor_shape(i).num_elements();
      a[i].cutoff_db = cutoff_db_[i];                                                // filled by NonsilenceOperator::AcquireArgs
      a[i].use_reference_power = reference_max_ ? 0 : 1;
      a[i].reference_power = reference_max_ ? 0.0f : reference_power_[i];
    }
    Check(dalib200NonsilentSetup(plan_, n, len.data(), a.data(), window_length_, reset_interval_), "NonsilentRegion");
    auto ip = InPtrs(in);
    auto bp = OutPtrs(begin);
    auto lp = OutPtrs(length);
    Check(dalib200NonsilentLaunch(plan_, ip.data(), bp.data(), lp.data(), ws.stream()), "NonsilentRegion");
  }

 private:
  dalib200SignalPlan *plan_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ ColorTwist / BrightnessContrast
// Both sit on the reference's OWN operator bases: ColorTwistBase composes the 3x3 matrix and offset per sample (color_twist.h:128-170),
// BrightnessContrastOp acquires the arguments and folds them into multiplier / addend (brightness_contrast.h:78-131); only RunImpl
// -- the kernel call -- is replaced.
class ColorTwist : public ColorTwistBase<GPUBackend> {
 public:
  explicit ColorTwist(const OpSpec &spec) : ColorTwistBase<GPUBackend>(spec) {
    Check(dalib200PointwisePlanCreate(&plan_, max_batch_size_), "ColorTwist");
  }
  ~ColorTwist() override { dalib200PointwisePlanDestroy(plan_); }

 protected:
  using SequenceOperator<GPUBackend, StatelessOperator>::RunImpl;
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    DALI_ENFORCE(in.type() == DALI_UINT8 && (output_type_ == DALI_UINT8 || output_type_ == DALI_FLOAT),
                 "b200.color_twist: uint8 input, uint8 or float output");
    const int n = in.num_samples();
    std::vector<dalib200ColorSample> cs(n);
    for (int i = 0; i < n; i++) {
      cs[i].num_pixels = in.tensor_shape(i).num_elements() / 3;
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) cs[i].matrix[3 * r + c] = tmatrices_[i](r, c);
        cs[i].offset[r] = toffsets_[i][r];
      }
    }
    Check(dalib200LinearTransformSetup(plan_, n, cs.data(), output_type_ == DALI_FLOAT ? DALIB200_FLOAT : DALIB200_UINT8), "ColorTwist");
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200PointwiseLaunch(plan_, ip.data(), op.data(), ws.stream()), "ColorTwist");
  }

 private:
  dalib200PointwisePlan *plan_ = nullptr;
};

class BrightnessContrast : public BrightnessContrastOp<GPUBackend> {
 public:
  explicit BrightnessContrast(const OpSpec &spec) : BrightnessContrastOp<GPUBackend>(spec) {
    Check(dalib200GenericPlanCreate(&plan_, max_batch_size_), "BrightnessContrast");
  }
  ~BrightnessContrast() override { dalib200GenericPlanDestroy(plan_); }

 protected:
  using Base::RunImpl;
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    DALI_ENFORCE(input_type_ == DALI_UINT8 && (output_type_ == DALI_UINT8 || output_type_ == DALI_FLOAT),
                 "b200.brightness_contrast: uint8 input, uint8 or float output");
    const int n = in.num_samples();
    const auto &center = GetContrastCenter<uint8_t>(ws, n);
    std::vector<int64_t> vol(n);
    std::vector<float> mul(n), add(n);
    for (int i = 0; i < n; i++) {
      vol[i] = in.tensor_shape(i).num_elements();
      if (output_type_ == DALI_FLOAT) OpArgsToKernelArgs<float, uint8_t>(add[i], mul[i], brightness_[i], brightness_shift_[i], contrast_[i], center[i]);
      else OpArgsToKernelArgs<uint8_t, uint8_t>(add[i], mul[i], brightness_[i], brightness_shift_[i], contrast_[i], center[i]);
    }
    Check(dalib200MultiplyAddSetup(plan_, n, vol.data(), mul.data(), add.data(), output_type_ == DALI_FLOAT ? DALIB200_FLOAT : DALIB200_UINT8),
          "BrightnessContrast");
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200GenericLaunch(plan_, ip.data(), op.data(), ws.stream()), "BrightnessContrast");
  }

 private:
  dalib200GenericPlan *plan_ = nullptr;
};

// ------------------------------------------------------------------------------------------------ Flip / Crop (window copy)
class WindowCopyBase : public Operator<GPUBackend> {
 public:
  explicit WindowCopyBase(const OpSpec &spec) : Operator<GPUBackend>(spec) {
    Check(dalib200GenericPlanCreate(&plan_, max_batch_size_), "window copy");
  }
  ~WindowCopyBase() override { dalib200GenericPlanDestroy(plan_); }

 protected:
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200GenericLaunch(plan_, ip.data(), op.data(), ws.stream()), "window copy");
  }
  dalib200GenericPlan *plan_ = nullptr;
  std::vector<dalib200WindowSample> win_;
};

class Flip : public WindowCopyBase {
 public:
  using WindowCopyBase::WindowCopyBase;

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    DALI_ENFORCE(in.type() == DALI_UINT8 && in.shape().sample_dim() == 3, "b200.flip: uint8 HWC images");
    const int n = in.num_samples();
    win_.assign(n, dalib200WindowSample{});
    for (int i = 0; i < n; i++) {
      auto sh = in.shape().tensor_shape_span(i);
      auto &w = win_[i];
      w.in_h = w.out_h = static_cast<int>(sh[0]); w.in_w = w.out_w = static_cast<int>(sh[1]); w.channels = static_cast<int>(sh[2]);
      w.flip_x = spec_.GetArgument<int>("horizontal", &ws, i) != 0;              // flip.h:43-49
      w.flip_y = spec_.GetArgument<int>("vertical", &ws, i) != 0;
    }
    Check(dalib200WindowCopySetup(plan_, n, win_.data()), "Flip");
    out.resize(1);
    out[0].shape = in.shape(); out[0].type = DALI_UINT8;
    return true;
  }
};

class Crop : public WindowCopyBase {
 public:
  explicit Crop(const OpSpec &spec) : WindowCopyBase(spec), crop_attr_(spec) {}

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    DALI_ENFORCE(in.type() == DALI_UINT8 && in.shape().sample_dim() == 3, "b200.crop: uint8 HWC images");
    const int n = in.num_samples();
    crop_attr_.ProcessArguments(spec_, ws);                                         // crop_attr.cc:100-245 (reference code)
    win_.assign(n, dalib200WindowSample{});
    out.resize(1);
    out[0].type = DALI_UINT8;
    out[0].shape.resize(n, 3);
    for (int i = 0; i < n; i++) {
      auto sh = in.shape().tensor_shape_span(i);
      const CropWindow cw = crop_attr_.GetCropWindowGenerator(i)(TensorShape<>{sh[0], sh[1]}, "HW");
      auto &w = win_[i];
      w.in_h = static_cast<int>(sh[0]); w.in_w = static_cast<int>(sh[1]); w.channels = static_cast<int>(sh[2]);
      w.anchor_y = static_cast<int>(cw.anchor[0]); w.anchor_x = static_cast<int>(cw.anchor[1]);
      w.out_h = static_cast<int>(cw.shape[0]); w.out_w = static_cast<int>(cw.shape[1]);
      out[0].shape.set_tensor_shape(i, TensorShape<>{cw.shape[0], cw.shape[1], sh[2]});
    }
    Check(dalib200WindowCopySetup(plan_, n, win_.data()), "Crop");
    return true;
  }

 private:
  CropAttr crop_attr_;
};

class Slice : public WindowCopyBase {
 public:
  explicit Slice(const OpSpec &spec) : WindowCopyBase(spec), slice_attr_(spec) {}

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    DALI_ENFORCE(in.type() == DALI_UINT8 && in.shape().sample_dim() == 3, "b200.slice: uint8 HWC images");
    const int n = in.num_samples();
    slice_attr_.ProcessArguments(spec_, ws);                                        // slice_attr.h:361-381 (reference code)
    auto layout = in.GetLayout();
    if (layout.empty()) layout = "HWC";
    win_.assign(n, dalib200WindowSample{});
    out.resize(1);
    out[0].type = DALI_UINT8;
    out[0].shape.resize(n, 3);
    for (int i = 0; i < n; i++) {
      const CropWindow cw = slice_attr_.GetCropWindowGenerator(i)(in.shape()[i], layout);     // slice_base.h:68-72
      auto sh = in.shape().tensor_shape_span(i);
      DALI_ENFORCE(cw.anchor[2] == 0 && cw.shape[2] == sh[2], "b200.slice: the channel axis cannot be sliced");
      auto &w = win_[i];
      w.in_h = static_cast<int>(sh[0]); w.in_w = static_cast<int>(sh[1]); w.channels = static_cast<int>(sh[2]);
      w.anchor_y = static_cast<int>(cw.anchor[0]); w.anchor_x = static_cast<int>(cw.anchor[1]);
      w.out_h = static_cast<int>(cw.shape[0]); w.out_w = static_cast<int>(cw.shape[1]);
      out[0].shape.set_tensor_shape(i, TensorShape<>{cw.shape[0], cw.shape[1], sh[2]});
    }
    Check(dalib200WindowCopySetup(plan_, n, win_.data()), "Slice");
    return true;
  }

 private:
  SliceAttr slice_attr_;
};

// ------------------------------------------------------------------------------------------------ Rotate
// Output canvas and the destination -> source matrix from the reference's own geometry helpers (rotate_params.h:36-55
// RotatedCanvasSize, :222-236 AdjustParams; include/dali/core/geom/transform.h), applied by the warp kernel.
class Rotate : public Operator<GPUBackend> {
 public:
  explicit Rotate(const OpSpec &spec) : Operator<GPUBackend>(spec) {
    interp_ = spec.GetArgument<DALIInterpType>("interp_type") == DALI_INTERP_LINEAR;
    keep_size_ = spec.GetArgument<bool>("keep_size");
    use_fill_ = spec.TryGetArgument(fill_, "fill_value");
    Check(dalib200WarpPlanCreate(&plan_, max_batch_size_), "Rotate");
  }
  ~Rotate() override { dalib200WarpPlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    DALI_ENFORCE(in.type() == DALI_UINT8 && in.shape().sample_dim() == 3, "b200.rotate: uint8 HWC images");
    const int n = in.num_samples();
    const bool has_size = spec_.ArgumentDefined("size");
    samples_.resize(n);
    out.resize(1);
    out[0].type = DALI_UINT8;
    out[0].shape.resize(n, 3);
    for (int i = 0; i < n; i++) {
      auto sh = in.shape().tensor_shape_span(i);
      const int in_h = static_cast<int>(sh[0]), in_w = static_cast<int>(sh[1]);
      const float a = deg2rad(-spec_.GetArgument<float>("angle", &ws, i));           // 2-D angles are negated (rotate_params.h SetParams)
      int oh, ow;
      if (has_size) {
        std::vector<float> sz(2);
        GetGeneralizedArg<float>(make_span(sz), "size", i, spec_, ws);
        oh = std::max<int>(static_cast<int>(std::roundf(sz[0])), 1); ow = std::max<int>(static_cast<int>(std::roundf(sz[1])), 1);
      } else if (keep_size_) {
        oh = in_h; ow = in_w;
      } else {
        ivec2 shape, parity;
        std::tie(shape, parity) = RotatedCanvasSize(TensorShape<2>(in_h, in_w), a);
        shape += (shape % 2) ^ (2 * parity > 1);                                      // the parity vote of InferSize for one frame
        ow = shape[0]; oh = shape[1];
      }
      const ivec2 in_size(in_w, in_h), out_size(ow, oh);
      const mat3 T = translation(in_size * 0.5f) * rotation2D(-a) * translation(-out_size * 0.5f);
      auto &s = samples_[i];
      s.in_h = in_h; s.in_w = in_w; s.channels = static_cast<int>(sh[2]); s.out_h = oh; s.out_w = ow;
      for (int r = 0; r < 2; r++) for (int c = 0; c < 3; c++) s.matrix[3 * r + c] = T(r, c);
      out[0].shape.set_tensor_shape(i, TensorShape<>{oh, ow, sh[2]});
    }
    Check(dalib200WarpPlanSetup(plan_, n, samples_.data(), interp_, use_fill_, fill_, DALIB200_UINT8), "Rotate");
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200WarpLaunch(plan_, ip.data(), op.data(), ws.stream()), "Rotate");
  }

 private:
  dalib200WarpPlan *plan_ = nullptr;
  std::vector<dalib200WarpSample> samples_;
  bool interp_ = true, keep_size_ = false, use_fill_ = false;
  float fill_ = 0;
};

// ------------------------------------------------------------------------------------------------ RandomResizedCrop
// Window from the reference's RandomCropAttr (Philox generator, random_crop_attr.h:36-72), filters from ResamplingFilterAttr; the
// window becomes the resampling ROI exactly as in random_resized_crop.h:105-112.
class RandomResizedCrop : public OperatorWithRandomCrop<Operator<GPUBackend>> {
 public:
  explicit RandomResizedCrop(const OpSpec &spec) : OperatorWithRandomCrop<Operator<GPUBackend>>(spec) {
    GetSingleOrRepeatedArg(spec, size_, "size", 2);
    Check(dalib200ResamplePlanCreate(&plan_, max_batch_size_), "RandomResizedCrop");
  }
  ~RandomResizedCrop() override { dalib200ResamplePlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    DALI_ENFORCE(in.shape().sample_dim() == 3, "b200.random_resized_crop: HWC images");
    const int n = in.num_samples();
    resampling_attr_.PrepareFilterParams(spec_, ws, n);
    std::vector<kernels::ResamplingParams2D> rp(n);
    for (int i = 0; i < n; i++) {
      auto sh = in.shape().tensor_shape_span(i);
      const CropWindow wnd = this->GetCropWindowGenerator(i)(TensorShape<>{sh[0], sh[1]}, "HW");
      for (int d = 0; d < 2; d++) {
        rp[i][d].output_size = size_[d];
        rp[i][d].roi = kernels::ResamplingParams::ROI(wnd.anchor[d], wnd.anchor[d] + wnd.shape[d]);
      }
    }
    resampling_attr_.ApplyFilterParams(make_span(rp));
    samples_.resize(n);
    out.resize(1);
    out[0].type = resampling_attr_.GetOutputType(in.type());
    out[0].shape.resize(n, 3);
    for (int i = 0; i < n; i++) {
      auto sh = in.shape().tensor_shape_span(i);
      auto &s = samples_[i];
      s.in_h = static_cast<int>(sh[0]); s.in_w = static_cast<int>(sh[1]); s.channels = static_cast<int>(sh[2]);
      s.out_h = size_[0]; s.out_w = size_[1];
      for (int d = 0; d < 2; d++) {
        const auto &p = rp[i][d];
        s.use_roi[d] = p.roi.use_roi; s.roi_start[d] = p.roi.start; s.roi_end[d] = p.roi.end;
        s.min_filter[d] = { static_cast<int>(p.min_filter.type), p.min_filter.antialias, p.min_filter.radius };
        s.mag_filter[d] = { static_cast<int>(p.mag_filter.type), p.mag_filter.antialias, p.mag_filter.radius };
      }
      out[0].shape.set_tensor_shape(i, TensorShape<>{size_[0], size_[1], sh[2]});
    }
    Check(dalib200ResamplePlanSetup(plan_, n, samples_.data(), in.type() == DALI_UINT8 ? DALIB200_UINT8 : DALIB200_FLOAT,
                                    out[0].type == DALI_UINT8 ? DALIB200_UINT8 : DALIB200_FLOAT), "RandomResizedCrop");
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200ResampleLaunch(plan_, ip.data(), op.data(), ws.stream()), "RandomResizedCrop");
  }

 private:
  std::vector<int> size_;
  ResamplingFilterAttr resampling_attr_;
  dalib200ResamplePlan *plan_ = nullptr;
  std::vector<dalib200ResampleSample> samples_;
};

// ------------------------------------------------------------------------------------------------ ToDecibels / MFCC
// On the reference's operator classes: their constructors / GetArguments read and validate the arguments
// (to_decibels_op.h:38-50, mfcc.h:95-118); Setup and Run go to the signal plan of the library.
class ToDecibels : public ::dali::ToDecibels<GPUBackend> {
 public:
  explicit ToDecibels(const OpSpec &spec) : ::dali::ToDecibels<GPUBackend>(spec) {
    db_.multiplier = args_.multiplier;
    db_.ref_max = args_.ref_max;
    db_.reference = args_.ref_max ? 1.0f : args_.s_ref;
    db_.cutoff_db = spec.GetArgument<float>("cutoff_db");
    Check(dalib200SignalPlanCreate(&plan_, max_batch_size_), "ToDecibels");
  }
  ~ToDecibels() override { dalib200SignalPlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    DALI_ENFORCE(in.type() == DALI_FLOAT, "b200.to_decibels: float input");
    const int n = in.num_samples();
    std::vector<int64_t> vol(n);
    for (int i = 0; i < n; i++) vol[i] = in.tensor_shape(i).num_elements();
    Check(dalib200ToDecibelsSetup(plan_, &db_, n, vol.data()), "ToDecibels");
    out.resize(1);
    out[0].shape = in.shape(); out[0].type = DALI_FLOAT;
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200SignalLaunch(plan_, ip.data(), op.data(), ws.stream()), "ToDecibels");
  }

 private:
  dalib200ToDecibelsArgs db_{};
  dalib200SignalPlan *plan_ = nullptr;
};

class MFCC : public ::dali::MFCC<GPUBackend> {
 public:
  explicit MFCC(const OpSpec &spec) : ::dali::MFCC<GPUBackend>(spec) {
    Check(dalib200SignalPlanCreate(&plan_, max_batch_size_), "MFCC");
  }
  ~MFCC() override { dalib200SignalPlanDestroy(plan_); }

 protected:
  bool SetupImpl(std::vector<OutputDesc> &out, const Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    GetArguments(ws);                                                               // mfcc.h:95-118 (reference code)
    DALI_ENFORCE(in.type() == DALI_FLOAT && in.shape().sample_dim() == 2 && axis_ == 0,
                 "b200.mfcc: float (features, frames) input, transform along axis 0");
    const int n = in.num_samples();
    dalib200MfccArgs a{ args_[0].ndct, args_[0].dct_type, args_[0].normalize ? 1 : 0, lifter_ };
    std::vector<int64_t> shp(2 * static_cast<size_t>(n));
    for (int i = 0; i < n; i++) {
      auto sh = in.shape().tensor_shape_span(i);
      shp[2 * i] = sh[0]; shp[2 * i + 1] = sh[1];
    }
    Check(dalib200MfccSetup(plan_, &a, n, shp.data()), "MFCC");
    out.resize(1);
    out[0].type = DALI_FLOAT;
    out[0].shape.resize(n, 2);
    for (int i = 0; i < n; i++)
      out[0].shape.set_tensor_shape(i, TensorShape<>{std::min<int64_t>(a.n_mfcc, shp[2 * i]), shp[2 * i + 1]});
    return true;
  }
  void RunImpl(Workspace &ws) override {
    const auto &in = ws.Input<GPUBackend>(0);
    auto &out = ws.Output<GPUBackend>(0);
    out.SetLayout(in.GetLayout());
    auto ip = InPtrs(in);
    auto op = OutPtrs(out);
    Check(dalib200SignalLaunch(plan_, ip.data(), op.data(), ws.stream()), "MFCC");
  }

 private:
  dalib200SignalPlan *plan_ = nullptr;
};

}  // namespace b200

namespace dali {       // the registration macros expect the dali namespace (operator.h:327-333)

DALI_SCHEMA(b200__decoders__Image).NumInput(1).NumOutput(1).AddParent("decoders__Image");
DALI_SCHEMA(b200__Resize).NumInput(1).NumOutput(1).AddParent("Resize");
DALI_SCHEMA(b200__CropMirrorNormalize).NumInput(1).NumOutput(1).AddParent("CropMirrorNormalize");
DALI_SCHEMA(b200__WarpAffine).NumInput(1).NumOutput(1).AddParent("WarpAffine");
DALI_SCHEMA(b200__Hsv).NumInput(1).NumOutput(1).AddParent("Hsv");
DALI_SCHEMA(b200__ColorSpaceConversion).NumInput(1).NumOutput(1).AddParent("ColorSpaceConversion");
DALI_SCHEMA(b200__Spectrogram).NumInput(1).NumOutput(1).AddParent("Spectrogram");
DALI_SCHEMA(b200__MelFilterBank).NumInput(1).NumOutput(1).AddParent("MelFilterBank");

DALI_REGISTER_OPERATOR(b200__decoders__Image, b200::ImageDecoder, Mixed);
DALI_REGISTER_OPERATOR(b200__Resize, b200::Resize, GPU);
DALI_REGISTER_OPERATOR(b200__CropMirrorNormalize, b200::CropMirrorNormalize, GPU);
DALI_REGISTER_OPERATOR(b200__WarpAffine, b200::WarpAffine, GPU);
DALI_REGISTER_OPERATOR(b200__Hsv, b200::Hsv, GPU);
DALI_REGISTER_OPERATOR(b200__ColorSpaceConversion, b200::ColorSpaceConversion, GPU);
DALI_REGISTER_OPERATOR(b200__Spectrogram, b200::Spectrogram, GPU);
DALI_REGISTER_OPERATOR(b200__MelFilterBank, b200::MelFilterBank, GPU);
DALI_SCHEMA(b200__decoders__ImageCrop).NumInput(1).NumOutput(1).AddParent("decoders__ImageCrop");
DALI_SCHEMA(b200__decoders__ImageRandomCrop).NumInput(1).NumOutput(1).AddParent("decoders__ImageRandomCrop");
DALI_REGISTER_OPERATOR(b200__decoders__ImageCrop, b200::ImageDecoderCrop, Mixed);
DALI_REGISTER_OPERATOR(b200__decoders__ImageRandomCrop, b200::ImageDecoderRandomCrop, Mixed);
DALI_SCHEMA(b200__AudioResample).NumInput(1).NumOutput(1).AddParent("AudioResample");
DALI_SCHEMA(b200__NonsilentRegion).NumInput(1).NumOutput(2).AddParent("NonsilentRegion");
DALI_REGISTER_OPERATOR(b200__AudioResample, b200::AudioResample, GPU);
DALI_REGISTER_OPERATOR(b200__NonsilentRegion, b200::NonsilentRegion, GPU);
DALI_SCHEMA(b200__decoders__ImageSlice).NumInput(1, 3).NumOutput(1).AddParent("decoders__ImageSlice");
DALI_SCHEMA(b200__Slice).NumInput(1, 3).NumOutput(1).AddParent("Slice");
DALI_SCHEMA(b200__Rotate).NumInput(1).NumOutput(1).AddParent("Rotate");
DALI_REGISTER_OPERATOR(b200__decoders__ImageSlice, b200::ImageDecoderSlice, Mixed);
DALI_REGISTER_OPERATOR(b200__Slice, b200::Slice, GPU);
DALI_REGISTER_OPERATOR(b200__Rotate, b200::Rotate, GPU);
DALI_SCHEMA(b200__ColorTwist).NumInput(1).NumOutput(1).AddParent("ColorTwist");
DALI_SCHEMA(b200__BrightnessContrast).NumInput(1).NumOutput(1).AddParent("BrightnessContrast");
DALI_SCHEMA(b200__Flip).NumInput(1).NumOutput(1).AddParent("Flip");
DALI_SCHEMA(b200__Crop).NumInput(1).NumOutput(1).AddParent("Crop");
DALI_SCHEMA(b200__RandomResizedCrop).NumInput(1).NumOutput(1).AddParent("RandomResizedCrop");
DALI_SCHEMA(b200__ToDecibels).NumInput(1).NumOutput(1).AddParent("ToDecibels");
DALI_SCHEMA(b200__MFCC).NumInput(1).NumOutput(1).AddParent("MFCC");
DALI_REGISTER_OPERATOR(b200__ColorTwist, b200::ColorTwist, GPU);
DALI_REGISTER_OPERATOR(b200__BrightnessContrast, b200::BrightnessContrast, GPU);
DALI_REGISTER_OPERATOR(b200__Flip, b200::Flip, GPU);
DALI_REGISTER_OPERATOR(b200__Crop, b200::Crop, GPU);
DALI_REGISTER_OPERATOR(b200__RandomResizedCrop, b200::RandomResizedCrop, GPU);
DALI_REGISTER_OPERATOR(b200__ToDecibels, b200::ToDecibels, GPU);
DALI_REGISTER_OPERATOR(b200__MFCC, b200::MFCC, GPU);

}  // namespace dali
