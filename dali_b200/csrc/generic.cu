// dali_b200/csrc/generic.cu -- the remaining per-pixel / geometry operators of SURVEY.md 8f rank 4 that do not need a kernel
// family of their own:
//   * multiply-add  out = ConvertSat<Out>(in * multiplier + addend)        brightness_contrast
//                   (dali/kernels/imgproc/pointwise/multiply_add.h:47-60; mul and add rounded separately like the CPU kernel)
//   * window copy   crop / slice / flip of interleaved u8 images with optional horizontal / vertical flip and out-of-bounds fill
//                   (dali/kernels/slice/slice_cpu.h, dali/operators/generic/flip.h: pure index arithmetic -> bit-exact)
// u8 HWC inputs; one launch per batch over a per-sample descriptor list.
#include "common.cuh"
#include <algorithm>
#include <cstring>

namespace dalib200 {

struct GenDesc {
  const uint8_t *in; void *out;
  int32_t in_h, in_w, c;
  int32_t anchor_y, anchor_x, out_h, out_w;     // window copy
  int32_t flip_x, flip_y;
  float mul, add;                               // multiply-add
  uint8_t fill[4];
  int64_t n;                                    // output elements
  int64_t first_item;
};

constexpr int kGenItem = 8192;                  // output elements per work item

__device__ __forceinline__ int find_gen(const GenDesc *d, int n, int64_t v) {
  int lo = 0, hi = n - 1;
  while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (d[mid].first_item <= v) lo = mid; else hi = mid - 1; }
  return lo;
}

template <typename Out>
__global__ void __launch_bounds__(256) multiply_add_kernel(const GenDesc *__restrict__ descs, int n, int64_t total_items) {
  for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int s = find_gen(descs, n, item);
    const GenDesc &d = descs[s];
    const int64_t e0 = (item - d.first_item) * kGenItem, e1 = min(d.n, e0 + kGenItem);
    Out *out = static_cast<Out *>(d.out);
    const bool vec = (reinterpret_cast<uintptr_t>(d.in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & (4 * sizeof(Out) - 1)) == 0;
    for (int64_t e = e0 + 4 * threadIdx.x; e < e1; e += 4 * blockDim.x) {
      if (vec && e + 4 <= e1) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t *>(d.in + e));
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; k++) v[k] = add_rn(mul_rn(u8_to_float((w >> (8 * k)) & 0xFFu), d.mul), d.add);
        if (sizeof(Out) == 1) {
          uint32_t o = 0;
#pragma unroll
          for (int k = 0; k < 4; k++) o |= (uint32_t)sat_u8_half_away(v[k]) << (8 * k);
          *reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(out) + e) = o;
        } else {
          *reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + e) = make_float4(v[0], v[1], v[2], v[3]);
        }
      } else {
        for (int64_t q = e; q < min(e1, e + 4); q++) {
          const float v = add_rn(mul_rn((float)d.in[q], d.mul), d.add);
          if (sizeof(Out) == 1) reinterpret_cast<uint8_t *>(out)[q] = sat_u8_half_away(v);
          else reinterpret_cast<float *>(out)[q] = v;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) window_copy_kernel(const GenDesc *__restrict__ descs, int n, int64_t total_items) {
  for (int64_t item = blockIdx.x; item < total_items; item += gridDim.x) {
    const int s = find_gen(descs, n, item);
    const GenDesc &d = descs[s];
    const int64_t e0 = (item - d.first_item) * kGenItem, e1 = min(d.n, e0 + kGenItem);
    uint8_t *out = static_cast<uint8_t *>(d.out);
    const uint32_t row = (uint32_t)d.out_w * (uint32_t)d.c;          // WindowCopySetup rejects samples of 2^31 elements or more
    for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
      const int oy = (int)((uint32_t)e / row);
      const int r = (int)((uint32_t)e - (uint32_t)oy * row);
      const int ox = r / d.c, ch = r - ox * d.c;
      const int sy = d.anchor_y + (d.flip_y ? d.out_h - 1 - oy : oy), sx = d.anchor_x + (d.flip_x ? d.out_w - 1 - ox : ox);
      uint8_t v = d.fill[min(ch, 3)];
      if (sy >= 0 && sy < d.in_h && sx >= 0 && sx < d.in_w) v = __ldg(d.in + ((int64_t)sy * d.in_w + sx) * d.c + ch);
      out[e] = v;
    }
  }
}

}  // namespace dalib200

using namespace dalib200;  // NOLINT

struct dalib200GenericPlan {
  int max_batch = 0, n = 0, kind = 0, out_dtype = DALIB200_UINT8;      // kind 1 = multiply-add, 2 = window copy
  std::vector<GenDesc> descs;
  int64_t total_items = 0;
  DescArena arena;
  cudaEvent_t uploaded = nullptr;
  bool pending = false;
};

extern "C" {

int dalib200GenericPlanCreate(dalib200GenericPlan **plan, int max_batch) try {
  DB_CHECK_ARG(plan && max_batch > 0, "GenericPlanCreate: bad arguments");
  auto *p = new dalib200GenericPlan();
  p->max_batch = max_batch;
  if (cudaEventCreateWithFlags(&p->uploaded, cudaEventDisableTiming) != cudaSuccess) {
    SetLastError("GenericPlanCreate: cudaEventCreate failed"); delete p; return DALIB200_ERROR_CUDA;
  }
  *plan = p;
  return DALIB200_SUCCESS;
} DB_API_CATCH

int dalib200GenericPlanDestroy(dalib200GenericPlan *p) try {
  if (!p) return DALIB200_SUCCESS;
  if (p->uploaded) { cudaEventSynchronize(p->uploaded); cudaEventDestroy(p->uploaded); }
  p->arena.Free();
  delete p;
  return DALIB200_SUCCESS;
} DB_API_CATCH

int dalib200MultiplyAddSetup(dalib200GenericPlan *p, int n, const int64_t *volumes, const float *multipliers, const float *addends, int out_dtype) try {
  DB_CHECK_ARG(p && n >= 0 && n <= p->max_batch && (n == 0 || (volumes && multipliers && addends)), "MultiplyAddSetup: bad arguments");
  DB_CHECK_ARG(out_dtype == DALIB200_UINT8 || out_dtype == DALIB200_FLOAT, "MultiplyAddSetup: output type must be UINT8 or FLOAT");
  p->kind = 1; p->n = n; p->out_dtype = out_dtype;
  p->descs.assign(n, GenDesc());
  int64_t items = 0;
  for (int i = 0; i < n; i++) {
    GenDesc &d = p->descs[i];
    memset(&d, 0, sizeof(d));
    DB_CHECK_ARG(volumes[i] >= 0, "MultiplyAddSetup: sample %d has a negative size", i);
    d.n = volumes[i]; d.mul = multipliers[i]; d.add = addends[i];
    d.first_item = items;
    items += (d.n + kGenItem - 1) / kGenItem;
  }
  p->total_items = items;
  return DALIB200_SUCCESS;
} DB_API_CATCH

int dalib200WindowCopySetup(dalib200GenericPlan *p, int n, const dalib200WindowSample *samples) try {
  DB_CHECK_ARG(p && n >= 0 && n <= p->max_batch && (n == 0 || samples), "WindowCopySetup: bad arguments");
  p->kind = 2; p->n = n; p->out_dtype = DALIB200_UINT8;
  p->descs.assign(n, GenDesc());
  int64_t items = 0;
  for (int i = 0; i < n; i++) {
    const dalib200WindowSample &w = samples[i];
    DB_CHECK_ARG(w.in_h >= 0 && w.in_w >= 0 && w.channels >= 1 && w.out_h >= 0 && w.out_w >= 0, "WindowCopySetup: sample %d: bad shape", i);
    GenDesc &d = p->descs[i];
    memset(&d, 0, sizeof(d));
    d.in_h = w.in_h; d.in_w = w.in_w; d.c = w.channels;
    d.anchor_y = w.anchor_y; d.anchor_x = w.anchor_x; d.out_h = w.out_h; d.out_w = w.out_w;
    d.flip_x = w.flip_x != 0; d.flip_y = w.flip_y != 0;
    for (int k = 0; k < 4; k++) d.fill[k] = w.fill[k];
    DB_CHECK_ARG(ElementsFit31(w.out_h, w.out_w, w.channels) && ElementsFit31(w.in_h, w.in_w, w.channels),
                 "WindowCopySetup: sample %d: images of 2^31 elements or more are not supported", i);
    d.n = (int64_t)w.out_h * w.out_w * w.channels;
    d.first_item = items;
    items += (d.n + kGenItem - 1) / kGenItem;
  }
  p->total_items = items;
  return DALIB200_SUCCESS;
} DB_API_CATCH

int dalib200GenericLaunch(dalib200GenericPlan *p, const void *const *in_ptrs, void *const *out_ptrs, dalib200Stream_t stream) try {
  DB_CHECK_ARG(p && p->kind != 0 && (p->n == 0 || (in_ptrs && out_ptrs)), "GenericLaunch: call a ...Setup function first");
  if (p->n == 0 || p->total_items == 0) return DALIB200_SUCCESS;
  if (p->pending) { DB_CUDA(cudaEventSynchronize(p->uploaded)); p->pending = false; }
  int rc = p->arena.Reserve(sizeof(GenDesc) * p->n);
  if (rc) return rc;
  GenDesc *h = reinterpret_cast<GenDesc *>(p->arena.host);
  for (int i = 0; i < p->n; i++) { h[i] = p->descs[i]; h[i].in = static_cast<const uint8_t *>(in_ptrs[i]); h[i].out = out_ptrs[i]; }
  if ((rc = p->arena.Upload(sizeof(GenDesc) * p->n, stream))) return rc;
  DB_CUDA(cudaEventRecord(p->uploaded, stream));
  p->pending = true;
  const GenDesc *d = reinterpret_cast<const GenDesc *>(p->arena.dev);
  const int grid = (int)std::min<int64_t>(p->total_items, (int64_t)NumSMs() * 16);
  if (p->kind == 1) {
    ProfScope ps_("multiply_add", stream);
    if (p->out_dtype == DALIB200_UINT8) multiply_add_kernel<uint8_t><<<grid, 256, 0, stream>>>(d, p->n, p->total_items);
    else multiply_add_kernel<float><<<grid, 256, 0, stream>>>(d, p->n, p->total_items);
  } else {
    ProfScope ps_("window_copy", stream);
    window_copy_kernel<<<grid, 256, 0, stream>>>(d, p->n, p->total_items);
  }
  CountLaunch();
  DB_CUDA(cudaGetLastError());
  return DALIB200_SUCCESS;
} DB_API_CATCH

}  // extern "C"
