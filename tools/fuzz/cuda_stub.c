// LD_PRELOAD stub: lets the HOST side of the JPEG plan (parse, unit / restart scan, descriptor build) run without a GPU under ASAN.
#include <stdlib.h>
#include <string.h>
typedef int cudaError_t;
cudaError_t cudaEventCreateWithFlags(void **e, unsigned f) { *e = malloc(8); return 0; }
cudaError_t cudaEventCreate(void **e) { *e = malloc(8); return 0; }
cudaError_t cudaEventDestroy(void *e) { free(e); return 0; }
cudaError_t cudaEventSynchronize(void *e) { return 0; }
cudaError_t cudaEventRecord(void *e, void *s) { return 0; }
cudaError_t cudaEventQuery(void *e) { return 0; }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned f) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
cudaError_t cudaFree(void *p) { free(p); return 0; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
cudaError_t cudaSetDevice(int d) { return 0; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
cudaError_t cudaDeviceGetAttribute(int *v, int attr, int dev) { *v = 148; return 0; }
cudaError_t cudaGetLastError(void) { return 0; }
cudaError_t cudaPeekAtLastError(void) { return 0; }
cudaError_t cudaStreamSynchronize(void *s) { return 0; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, void *s) { return 0; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int kind, void *st) { return 0; }
/* launch side: kernels are never run -- the host code in front of them (descriptor build, staging copies, arena growth) is what is fuzzed */
cudaError_t cudaLaunchKernel(const void *f, ...) { return 0; }
cudaError_t cudaFuncSetAttribute(const void *f, int attr, int v) { return 0; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *f, int b, size_t s) { *n = 2; return 0; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *n, const void *f, int b, size_t s, unsigned fl) { *n = 2; return 0; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, int kind) { return 0; }
cudaError_t cudaMemset(void *p, int v, size_t n) { return 0; }
cudaError_t cudaDeviceSynchronize(void) { return 0; }
cudaError_t cudaMemcpyBatchAsync(void **d, void **s, size_t *sz, size_t cnt, void *attrs, size_t *idx, size_t na, size_t *fail, void *st) { return 0; }
struct stubPtrAttr { int type; int device; void *devp; void *hostp; };
cudaError_t cudaPointerGetAttributes(struct stubPtrAttr *a, const void *p) { a->type = 0; a->device = 0; a->devp = 0; a->hostp = (void *)p; return 0; }
cudaError_t cudaGetDriverEntryPoint(const char *sym, void **fn, unsigned long long flags, int *status) { *fn = 0; if (status) *status = 1; return 0; }
cudaError_t cudaEventElapsedTime(float *ms, void *a, void *b) { *ms = 0; return 0; }
/* host library (executor): streams are opaque handles */
cudaError_t cudaStreamCreateWithFlags(void **s, unsigned f) { *s = malloc(8); return 0; }
cudaError_t cudaStreamCreate(void **s) { *s = malloc(8); return 0; }
cudaError_t cudaStreamDestroy(void *s) { free(s); return 0; }
cudaError_t cudaStreamWaitEvent(void *s, void *e, unsigned f) { return 0; }
