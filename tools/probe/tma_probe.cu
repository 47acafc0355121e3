// Probe: which tensor-map shapes does UTMALDG accept?  Each variant runs in a forked child (a faulting kernel poisons the context).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <unistd.h>
#include <sys/wait.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, uint32_t bytes, uint32_t *out) {
  extern __shared__ __align__(1024) uint8_t box[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   :: "r"(smem_u32(box)), "l"(&tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   :: "r"(smem_u32(box)), "l"(&tmap), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" :: "r"(smem_u32(&bar)) : "memory");
  uint32_t acc = 0;
  for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) acc += reinterpret_cast<uint32_t *>(box)[i];
  atomicAdd(out, acc);
}

__global__ void probe_g(const CUtensorMap *tmap, int c0, int c1, uint32_t bytes, uint32_t *out) {
  extern __shared__ __align__(1024) uint8_t box[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_u32(box)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW2:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D2;\n\tbra W2;\n\tD2:\n\t}" :: "r"(smem_u32(&bar)) : "memory");
  uint32_t acc = 0;
  for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) acc += reinterpret_cast<uint32_t *>(box)[i];
  atomicAdd(out, acc);
}

#include <dlfcn.h>
int run_g(int use_dlsym) {
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (use_dlsym) {
    void *h = dlopen("libcuda.so.1", RTLD_NOW);
    fp = h ? dlsym(h, "cuTensorMapEncodeTiled") : nullptr;
  } else {
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  }
  if (!fp) { printf("no entry point\n"); return 2; }
  auto enc = (EncodeTiledFn)fp;
  const int W = 640, H = 360;
  size_t pitch = (size_t)W * 4;
  uint8_t *d; cudaMalloc(&d, pitch * H);
  std::vector<uint32_t> h((size_t)W * H);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)(i * 2654435761u);
  cudaMemcpy(d, h.data(), pitch * H, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[2] = { W, H }; cuuint64_t strides[1] = { pitch };
  cuuint32_t box[2] = { 32, 16 }, es[2] = { 1, 1 };
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
  const uint64_t *tw = reinterpret_cast<const uint64_t *>(&tm);
  printf("[desc %016llx %016llx %016llx %016llx] ", (unsigned long long)tw[0], (unsigned long long)tw[1], (unsigned long long)tw[2], (unsigned long long)tw[3]);
  CUtensorMap *dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  uint32_t *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  probe_g<<<1, 128, 32 * 4 * 16 + 1024>>>(dtm, 8, 4, 32 * 4 * 16, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 4; }
  uint32_t got; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
  uint32_t want = 0;
  for (int y = 0; y < 16; y++) for (int x = 0; x < 32; x++) want += h[(size_t)(4 + y) * W + 8 + x];
  printf("%s (sum %08x vs %08x)\n", got == want ? "OK" : "MISMATCH", got, want);
  return 0;
}

int run(int rank, int box_w_words, int box_rows, int W, int H, int N, CUtensorMapDataType dt, int esize) {
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no entry point\n"); return 2; }
  auto enc = (EncodeTiledFn)fp;
  size_t pitch = (size_t)W * 3, frame = pitch * H;
  uint8_t *d; cudaMalloc(&d, frame * N);
  std::vector<uint8_t> h(frame * N);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 11));
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  const int ew = 4 / esize;      // elements per 32-bit word
  cuuint64_t dims[3] = { (cuuint64_t)(pitch / esize), (cuuint64_t)H, (cuuint64_t)N };
  cuuint64_t strides[2] = { pitch, frame };
  cuuint32_t box[3] = { (cuuint32_t)(box_w_words * ew), (cuuint32_t)box_rows, 1 }, es[3] = { 1, 1, 1 };
  if (rank == 2) dims[1] = (cuuint64_t)H * N;
  CUresult r = enc(&tm, dt, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
  uint32_t *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  const uint32_t bytes = box_w_words * 4 * box_rows;
  const int c0w = 5, c1 = 3, c2 = N - 1;
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10);
  cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 << 10);
  if (rank == 3) probe<3><<<1, 128, bytes + 1024>>>(tm, c0w * ew, c1, c2, bytes, out);
  else probe<2><<<1, 128, bytes + 1024>>>(tm, c0w * ew, c1 + c2 * H, 0, bytes, out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 4; }
  uint32_t got; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
  uint32_t want = 0;
  for (int r2 = 0; r2 < box_rows; r2++)
    for (int w = 0; w < box_w_words; w++) {
      size_t off = (size_t)c2 * frame + (size_t)(c1 + r2) * pitch + (size_t)(c0w + w) * 4;
      uint32_t v = 0;
      if (c1 + r2 < H && (size_t)(c0w + w) * 4 < pitch) memcpy(&v, &h[off], 4);
      want += v;
    }
  printf("%s (sum %08x vs %08x)\n", got == want ? "OK" : "MISMATCH", got, want);
  return got == want ? 0 : 5;
}

int main() {
  struct V { int rank, bw, br, W, H, N; CUtensorMapDataType dt; int es; const char *name; } vs[] = {
    {2, 32, 16, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "2d u32 box 128Bx16"},
    {2, 64, 16, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "2d u32 box 256Bx16"},
    {2, 112, 16, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "2d u32 box 448Bx16"},
    {2, 112, 48, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "2d u32 box 448Bx48"},
    {3, 32, 16, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "3d u32 box 128Bx16"},
    {3, 64, 48, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "3d u32 box 256Bx48"},
    {3, 112, 48, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "3d u32 box 448Bx48"},
    {3, 112, 48, 640, 360, 1, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "3d u32 box 448Bx48 n=1"},
    {3, 112, 48, 1280, 720, 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, "3d u32 box 448Bx48 720p"},
    {3, 56, 48, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, "3d u8 box 224Bx48"},
    {3, 112, 48, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, "3d u16 box 448Bx48"},
    {3, 112, 48, 640, 360, 2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, "3d f32 box 448Bx48"},
  };
  for (int dl = 0; dl < 2; dl++) {
    printf("%-28s : ", dl ? "gmem descriptor, dlsym" : "gmem descriptor, entrypoint"); fflush(stdout);
    pid_t pid = fork();
    if (pid == 0) { int rc = run_g(dl); fflush(stdout); _exit(rc); }
    int st; waitpid(pid, &st, 0);
  }
  { int drv = 0, rt = 0; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); printf("driver %d runtime %d\n", drv, rt); }
  for (auto &v : vs) {
    printf("%-28s : ", v.name); fflush(stdout);
    pid_t pid = fork();
    if (pid == 0) { int rc = run(v.rank, v.bw, v.br, v.W, v.H, v.N, v.dt, v.es); fflush(stdout); _exit(rc); }
    int st; waitpid(pid, &st, 0);
  }
  return 0;
}
