// simt_cudart.cpp — TEST INFRASTRUCTURE (see simt_emu.h): the slice of the CUDA runtime API that rendiation_b200/csrc calls,
// over host memory.  "Device" memory is malloc'ed, copies are memcpy, streams and events are synchronous (a kernel launch
// returns when the emulated grid has finished), one "device" with a handful of SMs so that launch grids stay small.
#include <chrono>

#include "simt_emu.h"

namespace {
struct FakeEvent {
  double ms = 0;
};
double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
int device_count() {
  const char *e = std::getenv("RDN_SIMT_DEVICES");
  const int n = e ? std::atoi(e) : 1;
  return n < 1 ? 1 : n;
}
void *host_alloc(size_t n) {
  const size_t bytes = ((n ? n : 1) + 255u) & ~static_cast<size_t>(255u);
  return std::aligned_alloc(256, bytes);
}
}  // namespace

extern "C" {

cudaError_t cudaMalloc(void **p, size_t n) { *p = host_alloc(n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMalloc(p, n); }
cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }  // (everything is host memory here)
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
// (every caller buffer counts as pageable: the host-buffer path always runs its staging pipeline here)
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *) { std::memset(a, 0, sizeof(*a)); a->type = cudaMemoryTypeUnregistered; return cudaSuccess; }
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { if (n) std::memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(dst, src, n, k); }
cudaError_t cudaMemcpyPeer(void *dst, int, const void *src, int, size_t n) { return cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice); }
cudaError_t cudaMemset(void *p, int v, size_t n) { if (n) std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }

cudaError_t cudaSetDevice(int d) { return d >= 0 && d < device_count() ? cudaSuccess : cudaErrorInvalidDevice; }
cudaError_t cudaGetDeviceCount(int *n) { *n = device_count(); return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *prop, int) {
  std::memset(prop, 0, sizeof(*prop));
  std::snprintf(prop->name, sizeof(prop->name), "SIMT emulator (CPU, test infrastructure)");
  prop->multiProcessorCount = 4;
  prop->l2CacheSize = 1 << 20;
  prop->totalGlobalMem = static_cast<size_t>(8) << 30;
  prop->major = 10;
  prop->minor = 0;
  prop->warpSize = 32;
  return cudaSuccess;
}
cudaError_t cudaDeviceCanAccessPeer(int *can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime error"; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = reinterpret_cast<cudaStream_t>(host_alloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }

cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(new FakeEvent); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete reinterpret_cast<FakeEvent *>(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { reinterpret_cast<FakeEvent *>(e)->ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = static_cast<float>(reinterpret_cast<FakeEvent *>(b)->ms - reinterpret_cast<FakeEvent *>(a)->ms);
  return cudaSuccess;
}

cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, const void *, int, size_t) { *n = 2; return cudaSuccess; }
cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int *n, const void *, int, size_t, unsigned) { *n = 2; return cudaSuccess; }
// (reached only through the header's template wrappers; the emulated launches never use it)
cudaError_t cudaLaunchKernelExC(const cudaLaunchConfig_t *, const void *, void **) { return cudaErrorNotSupported; }

}  // extern "C"
