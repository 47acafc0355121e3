// dali_b200/csrc/jpeg_prog.h -- launch interface of the progressive-JPEG entropy stage (jpeg_prog.cu), called by dalib200JpegLaunch.
#ifndef DALI_B200_CSRC_JPEG_PROG_H_
#define DALI_B200_CSRC_JPEG_PROG_H_
#include <cuda_runtime.h>
#include <vector>
#include "jpeg_prog_core.h"

namespace dalib200 {

struct ProgLaunch {
  const ProgImage *d_images; int nimages;
  const ProgScan *d_scans;                   // sorted by wave
  const ProgHuff *d_huff;
  const int64_t *d_first_blk;                // prefix sum of the blocks of the progressive images (nimages entries)
  int64_t total_blocks;
  const std::vector<int> *wave_begin;        // host: scans [wave_begin[w], wave_begin[w + 1]) run in launch w
  const std::vector<ProgImage> *h_images;    // host copy (coefficient ranges to clear)
  const uint8_t *d_raw; int16_t *d_coef; int16_t *d_dc; int32_t *d_status;
};
// clears the coefficient blocks of the progressive images, runs the scans wave by wave and writes the DC differences the shared
// dc_scan stage expects.  Returns a DALIB200 status.
int LaunchProgressive(const ProgLaunch &a, cudaStream_t s);

}  // namespace dalib200
#endif  // DALI_B200_CSRC_JPEG_PROG_H_
