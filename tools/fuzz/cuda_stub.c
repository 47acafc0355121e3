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
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned f) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
cudaError_t cudaMallocHost(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
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
