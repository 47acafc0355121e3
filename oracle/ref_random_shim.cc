// oracle/ref_random_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// extern "C" shim around the reference's random number operators, compiled from /root/reference where the pieces lie
// (oracle/Makefile):
//   include/dali/core/random/philox.h + dali/core/random/philox.cc       the generator, with its skipahead / skipahead_sequence
//   dali/operators/random/random_dist.h                                   bernoulli_dist, uniform_real_dist, uniform_discrete_dist
//   include/dali/core/convert.h                                           ConvertSat
// The operator classes themselves (rng_base.h RNGBase / OperatorWithRng, coin_flip.h, uniform_distribution.h) hang off the
// operator framework and cannot be compiled alone; the few lines that address the generator are restated here around the reference's
// own classes, in the reference's own terms:
//   master.init(seed, 0, 0)                                               rng_base.h:118-119
//   per sample: rng = master; rng.skipahead_sequence(sample * kSkipaheadPerSample)          :108-112
//   per element: r = rng; r.skipahead(element * kSkipaheadPerElement); ConvertSat<T>(dist.Generate(r))   rng_base_cpu.h:49-58
//   after the batch: master.skipahead_sequence(batch size)                                   rng_base.h:104,143-145
//   coin_flip: bernoulli_dist(probability), default int32 (coin_flip.h:28-42,62-70)
//   uniform: continuous = uniform_real_dist<FloatType>(T(start), T(end)) with FloatType = double for integral T of >= 4 bytes and for
//   8-byte T, else float; discrete = uniform_discrete_dist<float>(values) (uniform_distribution.h:30-67,142-158)
#include <cstdint>
#include <type_traits>

#include "dali/core/convert.h"
#include "dali/core/random/philox.h"
#include "dali/operators/random/random_dist.h"

using namespace dali;  // NOLINT

namespace {
constexpr int kPerElement = 257;      // rng_base.h:46
constexpr int kPerSample = 65537;     // rng_base.h:55

template <typename T, typename Dist>
void RunBatches(int64_t seed, int iterations, int batch, int64_t volume, Dist dist, T *out) {
  Philox4x32_10 master;
  master.init(seed, 0, 0);
  for (int it = 0; it < iterations; it++) {
    for (int s = 0; s < batch; s++) {
      Philox4x32_10 rng = master;
      rng.skipahead_sequence(s * kPerSample);
      for (int64_t p = 0; p < volume; p++) {
        Philox4x32_10 r = rng;
        r.skipahead(p * kPerElement);
        *out++ = ConvertSat<T>(dist(r));
      }
    }
    master.skipahead_sequence(batch);
  }
}

template <typename T>
void Uniform(int64_t seed, int iterations, int batch, int64_t volume, float start, float end, const float *values, int64_t nvalues, void *out) {
  using F = std::conditional_t<((std::is_integral_v<T> && sizeof(T) >= 4) || sizeof(T) > 4), double, float>;
  if (values) RunBatches<T>(seed, iterations, batch, volume, random::uniform_discrete_dist<float>(values, nvalues), static_cast<T *>(out));
  else RunBatches<T>(seed, iterations, batch, volume, random::uniform_real_dist<F>(T(start), T(end)), static_cast<T *>(out));
}
}  // namespace

extern "C" {

// out: iterations x batch x volume elements; dtype: 0 = uint8, 6 = int32 (DALIDataType codes)
int ref_random_coin_flip(int64_t seed, int iterations, int batch, int64_t volume, float probability, int dtype, void *out) {
  random::bernoulli_dist dist(probability);
  if (dtype == 6) RunBatches<int32_t>(seed, iterations, batch, volume, dist, static_cast<int32_t *>(out));
  else if (dtype == 0) RunBatches<uint8_t>(seed, iterations, batch, volume, dist, static_cast<uint8_t *>(out));
  else return 1;
  return 0;
}

int ref_random_uniform(int64_t seed, int iterations, int batch, int64_t volume, float start, float end, const float *values, int64_t nvalues,
                       int dtype, void *out) {
  switch (dtype) {
    case 0: Uniform<uint8_t>(seed, iterations, batch, volume, start, end, values, nvalues, out); break;
    case 5: Uniform<int16_t>(seed, iterations, batch, volume, start, end, values, nvalues, out); break;
    case 6: Uniform<int32_t>(seed, iterations, batch, volume, start, end, values, nvalues, out); break;
    case 7: Uniform<int64_t>(seed, iterations, batch, volume, start, end, values, nvalues, out); break;
    case 9: Uniform<float>(seed, iterations, batch, volume, start, end, values, nvalues, out); break;
    case 10: Uniform<double>(seed, iterations, batch, volume, start, end, values, nvalues, out); break;
    default: return 1;
  }
  return 0;
}

// raw generator output after init(key, sequence, offset): the known-answer hook for the Philox restatement
int ref_philox(uint64_t key, uint64_t sequence, uint64_t offset, int n, uint32_t *out) {
  Philox4x32_10 r;
  r.init(key, sequence, offset);
  for (int i = 0; i < n; i++) out[i] = r();
  return 0;
}

}  // extern "C"
