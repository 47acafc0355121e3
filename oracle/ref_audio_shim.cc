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
