// dali_b200/csrc/resample_axis.h -- the per-axis host setup of the separable resampler (filter choice, scale / origin, ROI footprint,
// index + coefficient tables), shared by the 2-D plan (resample.cu, where it is implemented) and the 3-D plan (resample3d_plan.h).
// Plain C++: no CUDA types, so the 3-D planner can be compiled by a host compiler for the CPU emulation test (tools/emul).
#ifndef DALI_B200_CSRC_RESAMPLE_AXIS_H_
#define DALI_B200_CSRC_RESAMPLE_AXIS_H_
#include <stdint.h>
#include "../../include/dali_b200.h"

namespace dalib200 {

// One axis of one sample after SetupAxis (resampling_setup.cc:27-122): POD view of resample.cu's AxisSetup.
struct AxisShared {
  int in_size, out_size;
  int ftype;                      // DALIB200_FILTER_* actually used (after the min/mag choice and the linear <-> triangular swap)
  int support;                    // >= 1 (NN -> 1)
  int roi_lo, roi_hi;             // source footprint, clamped to the input extent
  float origin, scale;            // source coordinate of output 0 / step per output element (origin NOT yet shifted by roi_lo)
  const float *coeffs; int num_coeffs; float anchor, fscale;   // the filter (a static table; valid for the life of the process)
};

int AxisSetupShared(AxisShared *a, int in_size, int out_size, bool use_roi, float roi_start, float roi_end,
                    dalib200FilterDesc min_filter, dalib200FilterDesc mag_filter);
// FIR table of one axis: idx[out_size], coef[out_size * support] (resampling_impl_cpu.cc:22-47); `origin` is the (possibly shifted) origin
void AxisFirTableShared(const AxisShared *a, float origin, int32_t *idx, float *coef);
// which output columns of a horizontal pass the reference stores through its 16-lane SSE path (resampling_impl_cpu.h:126-336):
// flags[ow] bytes, 1 = round half to even
void HorzSimdFlagsShared(const int32_t *idx, int ow, int iw, int support, uint8_t *flags);

}  // namespace dalib200
#endif  // DALI_B200_CSRC_RESAMPLE_AXIS_H_
