// dali_b200/host/random_crop.h -- the random crop window of decoders.image_random_crop / random_resized_crop.
//
// Restates, so that a given `seed` yields the SAME windows as the reference:
//   * the counter-based generator Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) with
//     the state layout of include/dali/core/random/philox.h:30-52 (128-bit counter = [sequence : offset / 4], phase =
//     offset % 4) and the round function of dali/core/random/philox.cc:33-88;
//   * RandomCropGenerator::GenerateCropWindowImpl (dali/operators/image/crop/random_crop_generator_util.cc:36-105): area and
//     log-aspect drawn with the C++ standard library's uniform_real_distribution<float>, the anchor with
//     uniform_int_distribution<int> -- the same libstdc++ templates the reference instantiates, fed by the same bit stream;
//   * the per-sample generator states of RandomCropAttr (random_crop_attr.h:36-72): key = seed ^ kRandomCropSeedModifier,
//     sequence = kSkipaheadPerSample * sample_index (dali/operators/random/rng_base.h:55).
#ifndef DALI_B200_HOST_RANDOM_CROP_H_
#define DALI_B200_HOST_RANDOM_CROP_H_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <random>
#include <utility>
#include <vector>

namespace dali {

class Philox4x32_10 {
 public:
  using result_type = uint32_t;
  static constexpr uint32_t min() { return 0; }
  static constexpr uint32_t max() { return 0xffffffffu; }

  Philox4x32_10() { Recalc(); }
  Philox4x32_10(uint64_t key, uint64_t sequence, uint64_t offset) : key_(key), phase_(static_cast<int>(offset & 3)) {
    ctr_[0] = offset >> 2; ctr_[1] = sequence;
    Recalc();
  }
  uint32_t operator()() {
    const uint32_t r = out_[phase_++];
    if (phase_ >= 4) {
      phase_ = 0;
      if (++ctr_[0] == 0) ctr_[1]++;
      Recalc();
    }
    return r;
  }

 private:
  void Recalc() {
    uint32_t x = static_cast<uint32_t>(ctr_[0]), y = static_cast<uint32_t>(ctr_[0] >> 32);
    uint32_t z = static_cast<uint32_t>(ctr_[1]), w = static_cast<uint32_t>(ctr_[1] >> 32);
    uint32_t kx = static_cast<uint32_t>(key_), ky = static_cast<uint32_t>(key_ >> 32);
    for (int round = 0; round < 10; round++) {
      const uint64_t m0 = 0xD2511F53ull * x, m1 = 0xCD9E8D57ull * z;
      const uint32_t nx = static_cast<uint32_t>(m1 >> 32) ^ y ^ kx, ny = static_cast<uint32_t>(m1);
      const uint32_t nz = static_cast<uint32_t>(m0 >> 32) ^ w ^ ky, nw = static_cast<uint32_t>(m0);
      x = nx; y = ny; z = nz; w = nw;
      kx += 0x9E3779B9u; ky += 0xBB67AE85u;
    }
    out_[0] = x; out_[1] = y; out_[2] = z; out_[3] = w;
  }
  uint64_t key_ = 0, ctr_[2] = {0, 0};
  int phase_ = 0;
  uint32_t out_[4] = {0, 0, 0, 0};
};

struct CropWindow2D { int anchor[2] = {0, 0}, shape[2] = {0, 0}; };     // (y, x) order

class RandomCropGenerator {
 public:
  RandomCropGenerator(float ar_lo, float ar_hi, float area_lo, float area_hi, uint64_t key, uint64_t sequence, int num_attempts)
      : ar_lo_(ar_lo), ar_hi_(ar_hi), log_ar_(std::log(ar_lo), std::log(ar_hi)), area_(area_lo, area_hi), rng_(key, sequence, 0),
        num_attempts_(num_attempts) {}

  CropWindow2D Generate(int H, int W) {
    CropWindow2D crop;
    if (W <= 0 || H <= 0) return crop;
    const float min_wh = ar_lo_, max_wh = ar_hi_, max_hw = 1 / ar_lo_;
    const float min_area = W * H * area_.a();
    const int maxW = std::max<int>(1, H * max_wh), maxH = std::max<int>(1, W * max_hw);
    if (H * maxW < min_area) {                        // image too wide
      crop.shape[0] = H; crop.shape[1] = maxW;
    } else if (W * maxH < min_area) {                 // image too tall
      crop.shape[0] = maxH; crop.shape[1] = W;
    } else {
      int attempts_left = num_attempts_;
      for (; attempts_left > 0; attempts_left--) {
        const float scale = area_(rng_);
        const size_t original_area = static_cast<size_t>(H * W);
        const float target_area = scale * original_area;
        float ratio = std::exp(log_ar_(rng_));
        int w = static_cast<int>(std::roundf(sqrtf(target_area * ratio)));
        int h = static_cast<int>(std::roundf(sqrtf(target_area / ratio)));
        if (w < 1) w = 1;
        if (h < 1) h = 1;
        crop.shape[0] = h; crop.shape[1] = w;
        ratio = static_cast<float>(w) / h;
        if (w <= W && h <= H && ratio >= min_wh && ratio <= max_wh) break;
      }
      if (attempts_left <= 0) {
        const float max_area = area_.b() * W * H;
        const float ratio = static_cast<float>(W) / H;
        if (ratio > max_wh) { crop.shape[0] = H; crop.shape[1] = maxW; }
        else if (ratio < min_wh) { crop.shape[0] = maxH; crop.shape[1] = W; }
        else { crop.shape[0] = H; crop.shape[1] = W; }
        const float scale = std::min(1.0f, max_area / (crop.shape[0] * crop.shape[1]));
        crop.shape[0] = std::max<int>(1, crop.shape[0] * std::sqrt(scale));
        crop.shape[1] = std::max<int>(1, crop.shape[1] * std::sqrt(scale));
      }
    }
    crop.anchor[0] = std::uniform_int_distribution<int>(0, H - crop.shape[0])(rng_);
    crop.anchor[1] = std::uniform_int_distribution<int>(0, W - crop.shape[1])(rng_);
    return crop;
  }

 private:
  float ar_lo_, ar_hi_;
  std::uniform_real_distribution<float> log_ar_, area_;
  Philox4x32_10 rng_;
  int num_attempts_;
};

constexpr uint64_t kRandomCropSeedModifier = 0x12345678abcdefeull;
constexpr int kSkipaheadPerSample = 65537;

inline std::vector<RandomCropGenerator> MakeRandomCropGenerators(int max_batch, int64_t seed, const float aspect[2], const float area[2],
                                                                 int num_attempts) {
  std::vector<RandomCropGenerator> g;
  g.reserve(max_batch);
  const uint64_t key = static_cast<uint64_t>(seed) ^ kRandomCropSeedModifier;
  for (int i = 0; i < max_batch; i++)
    g.emplace_back(aspect[0], aspect[1], area[0], area[1], key, static_cast<uint64_t>(kSkipaheadPerSample) * i, num_attempts);
  return g;
}

}  // namespace dali

#endif  // DALI_B200_HOST_RANDOM_CROP_H_
