// wavefront.cu — device side of the one-call wavefront executor (rdn_rt_trace_ray, capi.cu).
//
// The reference records, per launch tile and per round, one compute dispatch per task group — ray generation, trace, every
// closest-hit and miss shader — each followed by the five-pass compaction of the group's alive list, and reads every group's size
// back to decide its indirect dispatch (wavefront_compute/mod.rs:111-196, trace_task.rs:152-360, task-graph/src/runtime/mod.rs:419-461,
// task_group.rs:176-278).  Here a round is: one traversal launch over the wave (its size stays on the device), one kernel that turns
// the hit records into a task code per ray (sbt.cu), one stable compaction per shader into that shader's task list, the caller's
// stage kernels over those lists, one compaction of the spawn flags and a gather into the next wave.  Nothing comes back to the host
// between rounds.  The kernels of this file are the glue: wave gather, device-sized marks, per-round counters, and the three stage
// helpers a closest-hit / miss / ray-generation stage is usually made of.
#include <cuda_runtime.h>

#include "kernels.h"

namespace rdn {

namespace {

unsigned grid_for(uint64_t n) {
  const uint64_t b = (n + 255) / 256;
  return static_cast<unsigned>(b < 1 ? 1 : (b > 148ull * 32 ? 148ull * 32 : b));
}

// next wave: ray k = the ray slot idx[k] asked for; its launch index is inherited from the task that spawned it (slot i of the wave
// that was traced; in round 0 slot i IS launch index i)
__global__ void __launch_bounds__(256) k_wave_gather(const rdn_ray *__restrict__ next_rays, const uint32_t *__restrict__ launch_in,
                                                     const uint32_t *__restrict__ idx, const uint64_t *__restrict__ count,
                                                     rdn_ray *__restrict__ rays_out, uint32_t *__restrict__ launch_out) {
  const uint64_t n = *count;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t slot = idx[k];
    const float4 *src = reinterpret_cast<const float4 *>(next_rays + slot);
    float4 *dst = reinterpret_cast<float4 *>(rays_out + k);
    dst[0] = src[0]; dst[1] = src[1];
    launch_out[k] = launch_in ? launch_in[slot] : slot;
  }
}

// keep[k] = this ray's task code is `code` (rays at and beyond the device-side wave size never match), iota[k] = k
__global__ void __launch_bounds__(256) k_wave_mark(const uint32_t *__restrict__ task, const uint64_t *__restrict__ wave_size, uint64_t n_max,
                                                   uint32_t code, uint8_t *__restrict__ keep, uint32_t *__restrict__ iota) {
  const uint64_t n = *wave_size;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n_max; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    keep[k] = (k < n && task[k] == code) ? 1 : 0;
    iota[k] = static_cast<uint32_t>(k);
  }
}

// spawn flags beyond the wave are not spawns
__global__ void __launch_bounds__(256) k_wave_clip_spawn(uint8_t *__restrict__ spawn, const uint64_t *__restrict__ wave_size, uint64_t n_max,
                                                         uint32_t *__restrict__ iota) {
  const uint64_t n = *wave_size;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n_max; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    if (k >= n) spawn[k] = 0;
    iota[k] = static_cast<uint32_t>(k);
  }
}

__global__ void k_wave_record(const uint64_t *const *__restrict__ counters, uint32_t n, uint64_t *__restrict__ row) {
  if (threadIdx.x < n) row[threadIdx.x] = *counters[threadIdx.x];
}
__global__ void k_store_u64(uint64_t *dst, uint64_t v) { *dst = v; }

// ---- stage helpers
__global__ void __launch_bounds__(256) k_stage_spawn_all(uint8_t *__restrict__ spawn, const uint64_t *__restrict__ count) {
  const uint64_t n = *count;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) spawn[k] = 1;
}
__global__ void __launch_bounds__(256) k_stage_store_f32(const uint32_t *__restrict__ tasks, const uint64_t *__restrict__ count,
                                                         const uint32_t *__restrict__ launch_index, float value, float *__restrict__ dst) {
  const uint64_t n = *count;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint32_t slot = tasks ? tasks[k] : static_cast<uint32_t>(k);
    dst[launch_index ? launch_index[slot] : slot] = value;
  }
}

}  // namespace

void launch_wave_gather(const rdn_ray *d_next_rays, const uint32_t *d_launch_in, const uint32_t *d_idx, const uint64_t *d_count, uint64_t n_max,
                        rdn_ray *d_rays_out, uint32_t *d_launch_out, cudaStream_t stream) {
  if (n_max) k_wave_gather<<<grid_for(n_max), 256, 0, stream>>>(d_next_rays, d_launch_in, d_idx, d_count, d_rays_out, d_launch_out);
}
void launch_wave_mark(const uint32_t *d_task, const uint64_t *d_wave_size, uint64_t n_max, uint32_t code, uint8_t *d_keep, uint32_t *d_iota,
                      cudaStream_t stream) {
  if (n_max) k_wave_mark<<<grid_for(n_max), 256, 0, stream>>>(d_task, d_wave_size, n_max, code, d_keep, d_iota);
}
void launch_wave_clip_spawn(uint8_t *d_spawn, const uint64_t *d_wave_size, uint64_t n_max, uint32_t *d_iota, cudaStream_t stream) {
  if (n_max) k_wave_clip_spawn<<<grid_for(n_max), 256, 0, stream>>>(d_spawn, d_wave_size, n_max, d_iota);
}
void launch_wave_record(const uint64_t *const *d_counters, uint32_t n, uint64_t *d_row, cudaStream_t stream) {
  if (n) k_wave_record<<<1, 64, 0, stream>>>(d_counters, n, d_row);
}
void launch_store_u64(uint64_t *d_dst, uint64_t v, cudaStream_t stream) { k_store_u64<<<1, 1, 0, stream>>>(d_dst, v); }
void launch_stage_spawn_all(uint8_t *d_spawn, const uint64_t *d_count, uint64_t n_max, cudaStream_t stream) {
  if (n_max) k_stage_spawn_all<<<grid_for(n_max), 256, 0, stream>>>(d_spawn, d_count);
}
void launch_stage_store_f32(const uint32_t *d_tasks, const uint64_t *d_count, uint64_t n_max, const uint32_t *d_launch_index, float value,
                            float *d_dst, cudaStream_t stream) {
  if (n_max) k_stage_store_f32<<<grid_for(n_max), 256, 0, stream>>>(d_tasks, d_count, d_launch_index, value, d_dst);
}

}  // namespace rdn
