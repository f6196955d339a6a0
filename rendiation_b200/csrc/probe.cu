// probe.cu — measurement hook, no reference counterpart: read bandwidth of an L2-resident working set, the second roofline
// denominator SURVEY.md §8(d) asks for beside the HBM copy peak (the traversal's working set — top of the BVH, the tiles'
// nodes — lives in L2, not in HBM).
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.h"

namespace rdn {

namespace {

// every thread streams uint4 loads over the whole buffer, `passes` times; ld.global.cg (L2 only) so L1 does not serve repeats
__global__ void __launch_bounds__(256) k_l2_read(const uint4 *__restrict__ buf, uint64_t n_vec, int passes, unsigned long long *sink) {
  const uint64_t stride = static_cast<uint64_t>(gridDim.x) * blockDim.x;
  uint32_t acc = 0;
  for (int p = 0; p < passes; ++p) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n_vec; i += stride) {
      const uint4 v = __ldcg(buf + i);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x9E3779B9u) atomicAdd(sink, 1ull);  // keeps the loads alive
}

}  // namespace

int measure_l2_read_gbs(uint64_t bytes, int passes, int sm_count, double *out_gbs) {
  uint4 *buf = nullptr;
  unsigned long long *sink = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t err = cudaMalloc(&buf, bytes);
  if (err == cudaSuccess) err = cudaMalloc(&sink, 8);
  if (err == cudaSuccess) err = cudaMemset(buf, 0x5A, bytes);
  if (err == cudaSuccess) err = cudaMemset(sink, 0, 8);
  if (err == cudaSuccess) err = cudaEventCreate(&e0);
  if (err == cudaSuccess) err = cudaEventCreate(&e1);
  if (err == cudaSuccess) {
    const uint64_t n_vec = bytes / sizeof(uint4);
    const unsigned grid = static_cast<unsigned>(sm_count) * 8u;
    k_l2_read<<<grid, 256>>>(buf, n_vec, 2, sink);  // warm: the buffer becomes L2 resident
    cudaEventRecord(e0);
    k_l2_read<<<grid, 256>>>(buf, n_vec, passes, sink);
    cudaEventRecord(e1);
    err = cudaEventSynchronize(e1);
    if (err == cudaSuccess) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      *out_gbs = static_cast<double>(n_vec * sizeof(uint4)) * passes / (ms * 1e-3) / 1e9;
    }
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (buf) cudaFree(buf);
  if (sink) cudaFree(sink);
  return err == cudaSuccess ? 0 : -1;
}

}  // namespace rdn
