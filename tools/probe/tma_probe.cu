// Probe: which tensor-map forms does UTMALDG accept on this driver / toolchain pair?  Every variant runs in its own process
// (a faulting kernel poisons the context):  tma_probe <where: 0 param | 1 global> <l2: 0 none | 1 128B> <rank 2|3> <box words> <box rows>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK>
__device__ __forceinline__ void body(const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bytes, uint32_t *out) {
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
                   :: "r"(smem_u32(box)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(&bar)) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   :: "r"(smem_u32(box)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
  }
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" :: "r"(smem_u32(&bar)) : "memory");
  uint32_t acc = 0;
  for (uint32_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) acc += reinterpret_cast<uint32_t *>(box)[i];
  atomicAdd(out, acc);
}
template <int RANK> __global__ void probe_param(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, uint32_t bytes, uint32_t *out) { body<RANK>(&tmap, c0, c1, c2, bytes, out); }
template <int RANK> __global__ void probe_global(const CUtensorMap *tmap, int c0, int c1, int c2, uint32_t bytes, uint32_t *out) { body<RANK>(tmap, c0, c1, c2, bytes, out); }

int main(int argc, char **argv) {
  if (argc < 8) return 1;
  const int where = atoi(argv[1]), l2 = atoi(argv[2]), rank = atoi(argv[3]), bw = atoi(argv[4]), br = atoi(argv[5]);
  const int c0 = atoi(argv[6]); const size_t pitch = (size_t)atoi(argv[7]);
  printf("%s l2=%s %dd box %4dB x %2d c0=%d pitch=%zu : ", where ? "global" : "param ", l2 ? "128B" : "none", rank, bw * 4, br, c0, pitch); fflush(stdout);
  void *fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no entry point\n"); return 2; }
  auto enc = (EncodeTiledFn)fp;
  const int H = 360, N = 2;
  size_t frame = pitch * H;
  uint8_t *d; cudaMalloc(&d, frame * N);
  std::vector<uint8_t> h(frame * N);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 11));
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[3] = { (cuuint64_t)(pitch / 4), (cuuint64_t)H, (cuuint64_t)N };
  cuuint64_t strides[2] = { pitch, frame };
  cuuint32_t box[3] = { (cuuint32_t)bw, (cuuint32_t)br, 1 }, es[3] = { 1, 1, 1 };
  if (rank == 2) dims[1] = (cuuint64_t)H * N;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   l2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
  CUtensorMap *dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
  uint32_t *out; cudaMalloc(&out, 4); cudaMemset(out, 0, 4);
  const uint32_t bytes = bw * 4 * br;
  const int c1 = 3, c2 = N - 1;
  const int y = rank == 3 ? c1 : c1 + c2 * H;
  if (where == 0) {
    if (rank == 3) probe_param<3><<<1, 128, bytes + 1024>>>(tm, c0, y, c2, bytes, out); else probe_param<2><<<1, 128, bytes + 1024>>>(tm, c0, y, 0, bytes, out);
  } else {
    if (rank == 3) probe_global<3><<<1, 128, bytes + 1024>>>(dtm, c0, y, c2, bytes, out); else probe_global<2><<<1, 128, bytes + 1024>>>(dtm, c0, y, 0, bytes, out);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("kernel: %s\n", cudaGetErrorString(e)); return 4; }
  uint32_t got; cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost);
  uint32_t want = 0;
  for (int r2 = 0; r2 < br; r2++)
    for (int w = 0; w < bw; w++) {
      uint32_t v = 0;
      if (c1 + r2 < H && (size_t)(c0 + w) * 4 < pitch) memcpy(&v, &h[(size_t)c2 * frame + (size_t)(c1 + r2) * pitch + (size_t)(c0 + w) * 4], 4);
      want += v;
    }
  printf("%s\n", got == want ? "OK" : "MISMATCH");
  return got == want ? 0 : 5;
}
