// oracle/ref_audio_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// extern "C" shim over the reference's CPU audio kernels, compiled in place from /root/reference:
//   dali/kernels/signal/window/extract_windows_cpu.cc:97-146  (ExtractWindowsCpu)
//   dali/kernels/audio/mel_scale/mel_filter_bank_cpu.cc:149-237 (MelFilterBankCpu<float>)
//   dali/kernels/signal/window/window_functions.h:26-33        (HannWindow)
// The FFT stage (Fft1DImplFfts -> third_party/ffts) is NOT built: FFTS needs its cmake-generated
// configuration and a run-time code generator; see oracle/audio_oracle.c header.
#include <cstdint>
#include <vector>
#include "dali/kernels/signal/window/extract_windows_cpu.cc"   // NOLINT  (explicit instantiations live here)
#include "dali/kernels/audio/mel_scale/mel_filter_bank_cpu.cc"  // NOLINT
#include "dali/kernels/signal/window/window_functions.h"

using namespace dali;           // NOLINT
using namespace dali::kernels;  // NOLINT

extern "C" {

void ref_hann_window(float *w, int n) {
  signal::HannWindow(make_span(w, n));
}

// out: [nwin][win_len]
int ref_extract_windows(const float *in, int64_t n, const float *wfn, int win_len, int step,
                        int center_offset, int padding, float *out) {
  try {
    signal::ExtractWindowsCpu<float, float, 1, false> k;
    signal::ExtractWindowsArgs args;
    args.window_length = win_len; args.window_center = center_offset; args.window_step = step;
    args.axis = 0;
    args.padding = padding == 0 ? signal::Padding::None : padding == 1 ? signal::Padding::Zero : signal::Padding::Reflect;
    KernelContext ctx;
    InTensorCPU<float, 1> tin(in, TensorShape<1>(n));
    InTensorCPU<float, 1> twin(wfn, TensorShape<1>(win_len));
    auto req = k.Setup(ctx, tin, twin, args);
    auto sh = req.output_shapes[0][0];
    OutTensorCPU<float, 2> tout(out, TensorShape<2>(sh[0], sh[1]));
    k.Run(ctx, tout, tin, twin, args);
    return 0;
  } catch (...) { return -1; }
}

// in [nbin][nwin] -> out [nfilter][nwin]
int ref_mel_filter_bank_ft(const float *in, int nbin, int64_t nwin, float *out, int nfilter, float sample_rate,
                           float freq_low, float freq_high, int htk, int normalize) {
  try {
    audio::MelFilterBankCpu<float> k;
    audio::MelFilterBankArgs args;
    args.sample_rate = sample_rate; args.freq_low = freq_low; args.freq_high = freq_high;
    args.nfilter = nfilter; args.axis = 0; args.nfft = -1;
    args.mel_formula = htk ? audio::MelScaleFormula::HTK : audio::MelScaleFormula::Slaney;
    args.normalize = normalize != 0;
    KernelContext ctx;
    InTensorCPU<float> tin(in, TensorShape<>(nbin, nwin));
    OutTensorCPU<float> tout(out, TensorShape<>(nfilter, nwin));
    k.Setup(ctx, tin, args);
    k.Run(ctx, tout, tin);
    return 0;
  } catch (...) { return -1; }
}

}  // extern "C"

// ---- audio tail (round 2): ToDecibelsCpu, Dct1DCpu, MFCC liftering coefficients
#include "dali/kernels/signal/decibel/to_decibels_cpu.cc"   // NOLINT
#include "dali/kernels/signal/dct/dct_cpu.cc"               // NOLINT
#include "dali/operators/audio/mfcc/mfcc.h"                 // LifterCoeffs::CalculateCoeffs (private; built with -fno-access-control)

#include "dali/kernels/signal/moving_mean_square.cc"       // NOLINT
#include "dali/kernels/signal/resampling_cpu.h"
#include "dali/kernels/signal/resampling_cpu.cc"            // NOLINT
#include "dali/operators/audio/resampling_params.h"
#include "dali/kernels/signal/decibel/decibel_calculator.h"

extern "C" {

// NonsilentRegion for one float sample: the reference's MovingMeanSquareCpu and DecibelToMagnitude, with the thresholding and
// window adjustment of dali/operators/audio/nonsilence_op.h:60-130 (that header pulls in the whole operator framework, so its
// 25 lines of index logic are restated here around the reference kernels).
int ref_nonsilent_region(const float *in, int64_t n, float cutoff_db, float reference_power, int use_reference_power, int window_length,
                         int reset_interval, int32_t *begin, int32_t *length) {
  try {
    signal::MovingMeanSquareCpu<float> mms;
    signal::MovingMeanSquareArgs args{std::min<int>(window_length, (int)n), reset_interval};
    KernelContext ctx;
    InTensorCPU<float, 1> tin(in, TensorShape<1>(n));
    std::vector<float> buf(n);
    OutTensorCPU<float, 1> tout(buf.data(), TensorShape<1>(n));
    mms.Setup(ctx, tin, args);
    mms.Run(ctx, tout, tin, args);
    float ref = reference_power;
    if (!use_reference_power) { ref = buf[0]; for (int64_t i = 1; i < n; i++) ref = std::max(ref, buf[i]); }
    signal::DecibelToMagnitude<float> db2mag(10.f, ref);
    const float cutoff = db2mag(cutoff_db);
    int64_t end = n, b = n;
    for (int64_t i = 0; i < end; i++) if (buf[i] >= cutoff) { b = i; break; }
    int64_t first = 0, second = 0;
    if (b != end) {
      for (int64_t i = end - 1; i >= b; i--) if (buf[i] >= cutoff) { end = i; break; }
      first = b; second = end - b + 1;
    }
    if (first != 0 && second != 0) {
      const int new_start = std::max<int>((int)first - (args.window_size - 1), 0);
      second += first - new_start; first = new_start;
    }
    *begin = (int32_t)first; *length = (int32_t)second;
    return 0;
  } catch (...) { return -1; }
}

// AudioResample, float -> float: ResamplerCPU (windowed sinc from ResamplingParams::FromQuality) over [n_in][channels]
int ref_audio_resample(const float *in, int64_t n_in, int channels, double in_rate, double out_rate, int64_t n_out, float quality, float *out) {
  try {
    auto params = dali::audio::ResamplingParams::FromQuality(quality);
    signal::resampling::ResamplerCPU R;
    R.Initialize(params.lobes, params.lookup_size);
    R.Resample(out, 0, n_out, out_rate, in, n_in, in_rate, channels);
    return 0;
  } catch (...) { return -1; }
}

int64_t ref_resampled_length(int64_t in_length, double in_rate, double out_rate) {
  return signal::resampling::resampled_length(in_length, in_rate, out_rate);
}

int ref_to_decibels(const float *in, int64_t n, float *out, float multiplier, float reference, float cutoff_db, int ref_max) {
  try {
    signal::ToDecibelsCpu<float> k;
    signal::ToDecibelsArgs<float> args;
    args.multiplier = multiplier;
    args.ref_max = ref_max != 0;
    if (!args.ref_max) args.s_ref = reference;
    args.min_ratio = std::pow(10.0f, cutoff_db / args.multiplier);          // to_decibels_op.h:47-50
    if (args.min_ratio == 0) args.min_ratio = std::nextafter(0.0f, 1.0f);
    KernelContext ctx;
    InTensorCPU<float, DynamicDimensions> tin(in, TensorShape<>(n));
    OutTensorCPU<float, DynamicDimensions> tout(out, TensorShape<>(n));
    k.Setup(ctx, tin, args);
    k.Run(ctx, tout, tin, args);
    return 0;
  } catch (...) { return -1; }
}

// in [nfeat][ncols] -> out [ndct][ncols] (axis 0), optional liftering (mfcc.cc:52-72)
int ref_mfcc(const float *in, int nfeat, int64_t ncols, float *out, int n_mfcc, int dct_type, int normalize, float lifter) {
  try {
    signal::dct::Dct1DCpu<float, float, 2> k;
    signal::dct::DctArgs args;
    args.ndct = n_mfcc; args.dct_type = dct_type; args.normalize = normalize != 0;
    KernelContext ctx;
    InTensorCPU<float, 2> tin(in, TensorShape<2>(nfeat, ncols));
    auto req = k.Setup(ctx, tin, args, 0);
    auto sh = req.output_shapes[0][0];
    OutTensorCPU<float, 2> tout(out, TensorShape<2>(sh[0], sh[1]));
    k.Run(ctx, tout, tin, args, 0);
    if (lifter != 0.0f) {
      dali::detail::LifterCoeffs<dali::CPUBackend> lc;
      lc.lifter_ = lifter;
      std::vector<float> c(sh[0]);
      lc.CalculateCoeffs(c.data(), 0, sh[0]);
      for (int64_t r = 0; r < sh[0]; r++)
        for (int64_t t = 0; t < sh[1]; t++) out[r * sh[1] + t] = c[r] * out[r * sh[1] + t];
    }
    return static_cast<int>(sh[0]);
  } catch (...) { return -1; }
}

}  // extern "C"
