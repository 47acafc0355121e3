/*
 * oracle/audio_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's Spectrogram -> MelFilterBank path.
 *   windows : dali/kernels/signal/window/extract_windows_cpu.cc:97-146, extract_windows_args.h:41-45,
 *             Hann: dali/kernels/signal/window/window_functions.h:26-33
 *   FFT     : dali/kernels/signal/fft/fft_cpu_impl_ffts.cc:96-166 (window centred in the nfft buffer,
 *             first nfft/2+1 bins, |X|^2 = std::norm or |X| = std::abs, fft_cpu_impl_utils.h:53-67).
 *             The reference calls the vendored FFTS library (third_party/ffts) in fp32; FFTS needs its
 *             own cmake-configured feature macros and a run-time x86 code generator, so it is NOT
 *             rebuilt here.  The oracle evaluates the DFT in DOUBLE precision (radix-2 when nfft is a
 *             power of two, direct O(n^2) otherwise) and rounds once to fp32: parity for the STFT is
 *             therefore a stated relative tolerance (tests use 2e-4 of the frame maximum, the bound
 *             of the reference's own GPU-vs-CPU test dali/kernels/signal/fft/stft_gpu_test.cu:246).
 *   mel     : dali/kernels/audio/mel_scale/mel_scale.h:28-131, mel_filter_bank_cpu.cc:40-134.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

void oracle_hann_window(float *w, int N) {
  double a = (2 * M_PI / N);
  for (int t = 0; t < N; t++) w[t] = (float)(0.5 * (1.0 - cos(a * (t + 0.5))));
}

static int64_t reflect101(int64_t idx, int64_t size) {   /* include/dali/core/boundary.h:144-156 */
  if (size < 2) return size - 1;
  for (;;) {
    if (idx < 0) idx = -idx;
    else if (idx >= size) idx = 2 * size - 2 - idx;
    else break;
  }
  return idx;
}

int64_t oracle_num_windows(int64_t length, int window_length, int step, int centered) {
  if (!centered) length -= window_length;
  return length / step + 1;
}

/* out: [nwin][win_len] (horizontal windows) */
void oracle_extract_windows(const float *in, int64_t n, const float *wfn, int win_len, int step,
                            int center_offset, int padding /*0 none,1 zero,2 reflect*/, float *out) {
  int64_t nwin = oracle_num_windows(n, win_len, step, padding != 0);
  if (padding == 0) center_offset = 0;
  for (int64_t w = 0; w < nwin; w++) {
    int64_t start = w * step - center_offset;
    for (int t = 0; t < win_len; t++) {
      int64_t i = start + t;
      float v;
      if (i < 0 || i >= n) {
        if (padding == 2) v = wfn[t] * in[reflect101(i, n)];
        else v = 0;
      } else v = wfn[t] * in[i];
      out[w * win_len + t] = v;
    }
  }
}

static void dft_double(const double *x, int n, double *re, double *im) {
  if ((n & (n - 1)) == 0) {
    /* iterative radix-2, complex input with zero imaginary part */
    double *ar = (double *)malloc(sizeof(double) * n), *ai = (double *)calloc(n, sizeof(double));
    int bits = 0; while ((1 << bits) < n) bits++;
    for (int i = 0; i < n; i++) {
      int r = 0; for (int b = 0; b < bits; b++) if (i & (1 << b)) r |= 1 << (bits - 1 - b);
      ar[r] = x[i];
    }
    for (int len = 2; len <= n; len <<= 1) {
      for (int i = 0; i < n; i += len)
        for (int k = 0; k < len / 2; k++) {
          double ang = -2 * M_PI * k / len, wr = cos(ang), wi = sin(ang);
          int a = i + k, b = i + k + len / 2;
          double tr = ar[b] * wr - ai[b] * wi, ti = ar[b] * wi + ai[b] * wr;
          ar[b] = ar[a] - tr; ai[b] = ai[a] - ti;
          ar[a] += tr; ai[a] += ti;
        }
    }
    for (int k = 0; k <= n / 2; k++) { re[k] = ar[k]; im[k] = ai[k]; }
    free(ar); free(ai);
  } else {
    for (int k = 0; k <= n / 2; k++) {
      double sr = 0, si = 0;
      for (int t = 0; t < n; t++) {
        double ang = -2 * M_PI * (double)(((int64_t)k * t) % n) / n;
        sr += x[t] * cos(ang); si += x[t] * sin(ang);
      }
      re[k] = sr; im[k] = si;
    }
  }
}

/* Full spectrogram of one 1-D signal.  out: layout_ft ? [nbin][nwin] : [nwin][nbin], nbin = nfft/2+1.
 * power: 1 magnitude, 2 power.  (dali/operators/signal/fft/spectrogram.cc:152-296) */
int oracle_spectrogram(const float *in, int64_t n, const float *wfn, int win_len, int step, int nfft,
                       int power, int center, int reflect, int layout_ft, float *out) {
  int padding = center ? (reflect ? 2 : 1) : 0;
  int center_offset = center ? win_len / 2 : 0;
  int64_t nwin = oracle_num_windows(n, win_len, step, center);
  if (nwin <= 0 || nfft < win_len) return -1;
  int nbin = nfft / 2 + 1;
  float *wins = (float *)malloc(sizeof(float) * nwin * win_len);
  oracle_extract_windows(in, n, wfn, win_len, step, center_offset, padding, wins);
  double *buf = (double *)calloc(nfft, sizeof(double));
  double *re = (double *)malloc(sizeof(double) * nbin), *im = (double *)malloc(sizeof(double) * nbin);
  int in_win_start = win_len < nfft ? (nfft - win_len) / 2 : 0;
  for (int64_t w = 0; w < nwin; w++) {
    memset(buf, 0, sizeof(double) * nfft);
    for (int t = 0; t < win_len; t++) buf[in_win_start + t] = wins[w * win_len + t];
    dft_double(buf, nfft, re, im);
    for (int k = 0; k < nbin; k++) {
      double p = re[k] * re[k] + im[k] * im[k];
      float v = (float)(power == 2 ? p : sqrt(p));
      if (layout_ft) out[(int64_t)k * nwin + w] = v; else out[w * nbin + k] = v;
    }
  }
  free(wins); free(buf); free(re); free(im);
  return 0;
}

/* ------------------------------------------------------------------ mel filter bank (T = float) */
static float slaney_hz_to_mel(float hz) {
  const float fsp = (float)(200.0 / 3.0), min_log_hz = 1000.0f, min_log_mel = (1000.0f - 0.0f) / fsp;
  const float step_log = 0.068751777f;
  if (hz >= min_log_hz) return min_log_mel + logf(hz / min_log_hz) / step_log;
  return (hz - 0.0f) / fsp;
}
static float slaney_mel_to_hz(float mel) {
  const float fsp = (float)(200.0 / 3.0), min_log_hz = 1000.0f, min_log_mel = (1000.0f - 0.0f) / fsp;
  const float step_log = 0.068751777f;
  if (mel >= min_log_mel) return min_log_hz * expf(step_log * (mel - min_log_mel));
  return 0.0f + mel * fsp;
}
static float htk_hz_to_mel(float hz) { return 1127.0f * logf(1.0f + hz / 700.0f); }
static float htk_mel_to_hz(float mel) { return 700.0f * (expf(mel / 1127.0f) - 1.0f); }

typedef struct {
  int nfilter, fftbin_size, fftbin_start, fftbin_end, normalize;
  float *weights_down, *norm_factors;
  int *intervals;
} mel_tables;

static void mel_build(mel_tables *m, int nfilter, float sample_rate, float freq_low, float freq_high,
                      int nfft, int htk, int normalize) {
  float (*h2m)(float) = htk ? htk_hz_to_mel : slaney_hz_to_mel;
  float (*m2h)(float) = htk ? htk_mel_to_hz : slaney_mel_to_hz;
  double mel_low = h2m(freq_low), mel_high = h2m(freq_high);
  double hz_step = (double)sample_rate / nfft;
  double mel_delta = (mel_high - mel_low) / (nfilter + 1);
  m->nfilter = nfilter; m->normalize = normalize;
  m->fftbin_size = nfft / 2 + 1;
  double inv_hz_step = 1.0 / hz_step;
  m->fftbin_start = (int)ceil(freq_low * inv_hz_step);
  m->fftbin_end = (int)ceil(freq_high * inv_hz_step);
  if (m->fftbin_end > m->fftbin_size) m->fftbin_end = m->fftbin_size;
  m->weights_down = (float *)calloc(m->fftbin_size, sizeof(float));
  m->norm_factors = (float *)malloc(sizeof(float) * nfilter);
  for (int i = 0; i < nfilter; i++) m->norm_factors[i] = 1.0f;
  double mel0 = mel_low, mel1 = mel_low + mel_delta;
  int fftbin = m->fftbin_start;
  double f = fftbin * hz_step;
  for (int interval = 0; interval <= nfilter; interval++, mel0 = mel1, mel1 += mel_delta) {
    if (interval == nfilter) mel1 = mel_high;
    double f0 = m2h((float)mel0), f1 = m2h((float)mel1);
    if (normalize && interval < nfilter) {
      double f2 = m2h((float)(mel1 + mel_delta));
      m->norm_factors[interval] = (float)(2.0 / (f2 - f0));
    }
    double slope = 1. / (f1 - f0);
    for (; fftbin < m->fftbin_end && f < f1; fftbin++, f = fftbin * hz_step)
      m->weights_down[fftbin] = (float)((f1 - f) * slope);
  }
  m->intervals = (int *)malloc(sizeof(int) * m->fftbin_size);
  for (int i = 0; i < m->fftbin_size; i++) m->intervals[i] = -1;
  fftbin = m->fftbin_start; f = fftbin * hz_step;
  double mel = mel_low + mel_delta;
  for (int interval = 0; interval < nfilter + 1; interval++, mel += mel_delta) {
    double freq = m2h((float)(interval == nfilter ? mel_high : mel));
    for (; fftbin < m->fftbin_end && f < freq; fftbin++, f = fftbin * hz_step) m->intervals[fftbin] = interval;
  }
}

/* in: [nbin][nwin] ("ft"), out: [nfilter][nwin].  mel_filter_bank_cpu.cc:77-111 */
int oracle_mel_filter_bank_ft(const float *in, int nbin, int64_t nwin, float *out, int nfilter,
                              float sample_rate, float freq_low, float freq_high, int htk, int normalize) {
  mel_tables m;
  int nfft = 2 * (nbin - 1);
  if (freq_high <= 0) freq_high = sample_rate / 2;
  mel_build(&m, nfilter, sample_rate, freq_low, freq_high, nfft, htk, normalize);
  memset(out, 0, sizeof(float) * nfilter * nwin);
  for (int b = m.fftbin_start; b < m.fftbin_end; b++) {
    const float *in_row = in + (int64_t)b * nwin;
    int up = m.intervals[b], down = up - 1;
    float w_up = 1.0f - m.weights_down[b], w_down = m.weights_down[b];
    if (down >= 0) {
      if (normalize) w_down *= m.norm_factors[down];
      float *o = out + (int64_t)down * nwin;
      for (int64_t t = 0; t < nwin; t++) { float p = w_down * in_row[t]; o[t] = o[t] + p; }
    }
    if (up >= 0 && up < nfilter) {
      if (normalize) w_up *= m.norm_factors[up];
      float *o = out + (int64_t)up * nwin;
      for (int64_t t = 0; t < nwin; t++) { float p = w_up * in_row[t]; o[t] = o[t] + p; }
    }
  }
  free(m.weights_down); free(m.norm_factors); free(m.intervals);
  return 0;
}

/* Dense [nfilter][nbin] weight matrix equivalent to the banded walk above (used to test the
 * tensor-core GEMM formulation); follows mel_filter_bank_test.cc:23-79 in spirit. */
int oracle_mel_weights(float *W, int nbin, int nfilter, float sample_rate, float freq_low, float freq_high,
                       int htk, int normalize) {
  mel_tables m;
  int nfft = 2 * (nbin - 1);
  if (freq_high <= 0) freq_high = sample_rate / 2;
  mel_build(&m, nfilter, sample_rate, freq_low, freq_high, nfft, htk, normalize);
  memset(W, 0, sizeof(float) * nfilter * nbin);
  for (int b = m.fftbin_start; b < m.fftbin_end; b++) {
    int up = m.intervals[b], down = up - 1;
    float w_up = 1.0f - m.weights_down[b], w_down = m.weights_down[b];
    if (down >= 0) { if (normalize) w_down *= m.norm_factors[down]; W[down * nbin + b] = w_down; }
    if (up >= 0 && up < nfilter) { if (normalize) w_up *= m.norm_factors[up]; W[up * nbin + b] = w_up; }
  }
  free(m.weights_down); free(m.norm_factors); free(m.intervals);
  return 0;
}
