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

/* ================================================================================================================================
 * Audio tail: plain-C restatements, pinned bit for bit against the compiled reference kernels (oracle/_ref) by
 * tests/test_audio_tail_ref_cpu.py.
 * ================================================================================================================================ */

/* ToDecibels: dali/kernels/signal/decibel/to_decibels_cpu.cc:47-72 + decibel_calculator.h:25-56
 *   out = (mul * log10(2)) * log2(max(min_ratio, in * (1 / s_ref))),  s_ref = max over the sample when ref_max (0 -> 1) */
void oracle_to_decibels(const float *in, int64_t n, float *out, float multiplier, float reference, float cutoff_db, int ref_max) {
  float min_ratio = powf(10.0f, cutoff_db / multiplier);              /* to_decibels_op.h:47-50 */
  if (min_ratio == 0) min_ratio = nextafterf(0.0f, 1.0f);
  float s_ref = reference;
  if (ref_max) {
    s_ref = 0.0f;
    for (int64_t i = 0; i < n; i++) if (in[i] > s_ref) s_ref = in[i];
    if (s_ref == 0.0f) s_ref = 1.0f;                                   /* avoid division by 0 */
  }
  const float kLog2Factor = 0.3010299956639812f;                      /* std::log10(2.0) */
  const float mul_log2 = multiplier * kLog2Factor;
  const float inv_s_ref = 1.0f / s_ref;
  for (int64_t i = 0; i < n; i++) {
    float r = in[i] * inv_s_ref;
    if (r < min_ratio) r = min_ratio;
    out[i] = mul_log2 * log2f(r);
  }
}

/* DCT along axis 0 of [nfeat][ncols]: dali/kernels/signal/dct/table.h:27-112 (cosine tables in double, stored as float) and
 * dct_cpu.cc:76-115 (out[k] = sum_n in[n] * table[k][n], n ascending, float mul + add); liftering mfcc.h:36-41, mfcc.cc:52-72. */
static void dct_table(float *table, int64_t n, int ndct, int type, int normalize) {
  int64_t idx = 0;
  if (type == 1) {
    const double phase_mul = M_PI / (n - 1);
    for (int64_t k = 0; k < ndct; k++) {
      table[idx++] = 0.5f;
      for (int64_t i = 1; i < n - 1; i++) table[idx++] = (float)cos(phase_mul * k * i);
      table[idx++] = k % 2 == 0 ? 0.5f : -0.5f;
    }
  } else if (type == 2) {
    const double phase_mul = M_PI / n;
    double f0 = 1, fi = 1;
    if (normalize) { fi = sqrt(2.0 / n); f0 = 1.0 / sqrt((double)n); }
    for (int64_t k = 0; k < ndct; k++) {
      const double nf = k == 0 ? f0 : fi;
      for (int64_t i = 0; i < n; i++) table[idx++] = (float)(nf * cos(phase_mul * (i + 0.5) * k));
    }
  } else if (type == 3) {
    const double phase_mul = M_PI / n;
    double f0 = 0.5, fi = 1;
    if (normalize) { fi = sqrt(2.0 / n); f0 = 1.0 / sqrt((double)n); }
    for (int64_t k = 0; k < ndct; k++) {
      table[idx++] = (float)f0;
      for (int64_t i = 1; i < n; i++) table[idx++] = (float)(fi * cos(phase_mul * i * (k + 0.5)));
    }
  } else {
    const double phase_mul = M_PI / n;
    const double f = normalize ? sqrt(2.0 / n) : 1.0;
    for (int64_t k = 0; k < ndct; k++)
      for (int64_t i = 0; i < n; i++) table[idx++] = (float)(f * cos(phase_mul * (i + 0.5) * (k + 0.5)));
  }
}

int oracle_mfcc(const float *in, int nfeat, int64_t ncols, float *out, int n_mfcc, int dct_type, int normalize, float lifter) {
  const int ndct = n_mfcc < nfeat ? n_mfcc : nfeat;
  float *table = (float *)malloc(sizeof(float) * (size_t)ndct * nfeat);
  if (!table) return -1;
  dct_table(table, nfeat, ndct, dct_type, normalize);
  for (int k = 0; k < ndct; k++) {
    float coeff = 1.0f;
    if (lifter != 0.0f) coeff = 1.0f + lifter / 2 * sinf((float)M_PI / lifter * (k + 1));   /* all-float, see audio_tail.cu */
    for (int64_t c = 0; c < ncols; c++) {
      float acc = 0.0f;
      for (int nn = 0; nn < nfeat; nn++) acc += table[(size_t)k * nfeat + nn] * in[(size_t)nn * ncols + c];
      out[(size_t)k * ncols + c] = lifter != 0.0f ? acc * coeff : acc;
    }
  }
  free(table);
  return ndct;
}

/* NonsilentRegion: moving mean square with a running float sum (dali/kernels/signal/moving_mean_square.cc:55-77), threshold
 * s_ref * pow(10, cutoff_db / 10) (decibel_calculator.h:60-73), first / last sample at or above it and the window adjustment
 * (dali/operators/audio/nonsilence_op.h:60-130). */
void oracle_nonsilent_region(const float *in, int64_t n, float cutoff_db, float reference_power, int use_reference_power, int window_length,
                             int reset_interval, int32_t *begin, int32_t *length) {
  const int win = window_length < n ? window_length : (int)n;
  const int64_t interval = reset_interval == -1 ? n : reset_interval;
  float *mms = (float *)malloc(sizeof(float) * (size_t)n);
  const float mean_factor = 1.0f / win;
  int64_t win_begin = -(int64_t)win + 1;
  for (int64_t out_pos = 0; out_pos < n;) {
    float sumsq = 0;
    for (int64_t pos = win_begin > 0 ? win_begin : 0; pos < out_pos; pos++) sumsq += in[pos] * in[pos];
    const int64_t interval_end = out_pos + interval < n ? out_pos + interval : n;
    for (; out_pos < interval_end; out_pos++, win_begin++) {
      sumsq += in[out_pos] * in[out_pos];
      mms[out_pos] = sumsq * mean_factor;
      if (win_begin >= 0) sumsq -= in[win_begin] * in[win_begin];
    }
  }
  float ref = reference_power;
  if (!use_reference_power) { ref = mms[0]; for (int64_t i = 1; i < n; i++) if (mms[i] > ref) ref = mms[i]; }
  const float cutoff = ref * powf(10.0f, cutoff_db * (1.0f / 10.0f));
  int64_t end = n, b = n, first = 0, second = 0;
  for (int64_t i = 0; i < end; i++) if (mms[i] >= cutoff) { b = i; break; }
  if (b != end) {
    for (int64_t i = end - 1; i >= b; i--) if (mms[i] >= cutoff) { end = i; break; }
    first = b; second = end - b + 1;
  }
  if (first != 0 && second != 0) {
    int64_t ns = first - (win - 1); if (ns < 0) ns = 0;
    second += first - ns; first = ns;
  }
  *begin = (int32_t)first; *length = (int32_t)second;
  free(mms);
}

/* AudioResample, float -> float: dali/kernels/signal/resampling.h:36-100 (Hann-windowed sinc, linearly interpolated lookup) and
 * resampling_cpu.cc:120-230 (one channel: four SSE partial sums over taps i0 + l + 4k combined as (f0 + f2) + (f1 + f3), then the
 * scalar tail; several channels: taps in order).  quality -> lobes: dali/operators/audio/resampling_params.h:27-30. */
typedef struct { float scale, center; int lobes, size; float *lookup; } rs_window;

static float rs_sincf(float x) {                                       /* math_util.h:188-193 */
  x *= M_PI;
  if (fabsf(x) < 1e-5f) return 1.0f - x * x * (1.0f / 6);
  return sinf(x) / x;
}
static float rs_eval(const rs_window *w, float x) {                    /* resampling.h:59-66 */
  float fi = x * w->scale + w->center;
  float floori = floorf(fi);
  float di = fi - floori;
  int i = (int)floori;
  return w->lookup[i] + di * (w->lookup[i + 1] - w->lookup[i]);
}
static float rs_eval_trunc(const rs_window *w, float x) {              /* resampling_cpu.cc:86-99 (cvttps: truncation) */
  float fi = x * w->scale + w->center;
  int i = (int)fi;
  float di = fi - (float)i;
  return w->lookup[i] + di * (w->lookup[i + 1] - w->lookup[i]);
}

int oracle_audio_resample(const float *in, int64_t n_in, int channels, double in_rate, double out_rate, int64_t n_out, float quality,
                          float *out) {
  rs_window w;
  {
    const double q = quality;
    w.lobes = (int)round(0.007 * q * q - 0.09 * q + 3);
    const int coeffs = w.lobes * 64 + 1;
    const float scale = 2.0f * w.lobes / (coeffs - 1);
    const float scale_envelope = 2.0f / coeffs;
    const int center = (int)((coeffs - 1) * 0.5f);
    w.size = coeffs + 5;
    w.lookup = (float *)calloc((size_t)w.size, sizeof(float));
    if (!w.lookup) return -1;
    for (int i = 0; i < coeffs; i++) {
      const float x = (i - center) * scale, y = (i - center) * scale_envelope;
      w.lookup[i + 1] = (float)(rs_sincf(x) * (0.5 * (1 + cos((double)y * M_PI))));
    }
    w.center = (float)(center + 1);
    w.scale = 1 / scale;
  }
  const int64_t block = 1 << 8;
  const double scale = in_rate / out_rate;
  const float fscale = (float)scale;
  for (int64_t out_block = 0; out_block < n_out; out_block += block) {
    const int64_t block_end = out_block + block < n_out ? out_block + block : n_out;
    const double in_block_f = out_block * scale;
    const int64_t in_block_i = (int64_t)floor(in_block_f);
    float in_pos = (float)(in_block_f - in_block_i);
    const float *inb = in + in_block_i * channels;
    for (int64_t out_pos = out_block; out_pos < block_end; out_pos++, in_pos += fscale) {
      const int xc = (int)ceilf(in_pos);
      int i0 = xc - w.lobes, i1 = xc + w.lobes;
      if (i0 + in_block_i < 0) i0 = (int)-in_block_i;
      if (i1 + in_block_i > n_in) i1 = (int)(n_in - in_block_i);
      if (channels == 1) {
        int i = i0;
        float f4[4] = {0, 0, 0, 0}, x4[4];
        for (int l = 0; l < 4; l++) x4[l] = (float)(i + l) - in_pos;
        for (; i + 3 < i1; i += 4)
          for (int l = 0; l < 4; l++) {
            const float wv = rs_eval_trunc(&w, x4[l]);
            f4[l] = f4[l] + inb[i + l] * wv;
            x4[l] = x4[l] + 4;
          }
        float f = (f4[0] + f4[2]) + (f4[1] + f4[3]);
        float x = (float)i - in_pos;
        for (; i < i1; i++, x++) f += inb[i] * rs_eval(&w, x);
        out[out_pos] = f;
      } else {
        float tmp[16];
        for (int c = 0; c < channels; c++) tmp[c] = 0;
        float x = (float)i0 - in_pos;
        for (int i = i0; i < i1; i++, x++) {
          const float wv = rs_eval(&w, x);
          for (int c = 0; c < channels; c++) tmp[c] += inb[(int64_t)i * channels + c] * wv;
        }
        for (int c = 0; c < channels; c++) out[out_pos * channels + c] = tmp[c];
      }
    }
  }
  free(w.lookup);
  return 0;
}
