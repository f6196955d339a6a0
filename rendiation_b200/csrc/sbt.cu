// sbt.cu — shader-binding-table dispatch after a trace (first slice).
//
// What the reference does per ray once traversal has finished (TraceTaskImpl::device_poll,
// shader/ray-tracing/src/backend/wavefront_compute/trace_task.rs:206-268):
//   hit  (and not RAY_FLAG_SKIP_CLOSEST_HIT_SHADER):
//        hit_group = ray.sbt_ray_config.offset + ray.sbt_ray_config.stride * geometry_id + instance_sbt_offset   (api/ctx.rs:53-55)
//        shader    = ray_hit[hit_group_start + hit_group].closest_hit            (sbt.rs:252-254); u32::MAX: nothing is spawned
//        spawn_dynamic(closest task of that shader, RayClosestHitCtxPayload)
//   miss: shader   = ray_miss[miss_start + ray.miss_index]                        (sbt.rs:264-268); u32::MAX: nothing is spawned
//        spawn_dynamic(miss task of that shader, RayMissHitCtxPayload)
// The reference appends each spawned task to its shader's task pool through an atomic bump allocator and polls the pools
// one compute dispatch per shader.  Here the same decision is one kernel over the hit records (k_sbt_dispatch: a task code per
// ray) and the "pools" are the ray indices grouped by shader in ray order (launch_sbt_group: one stable compaction per shader,
// the single-pass decoupled look-back scan of compact.cu), so a closest-hit / miss shader stage runs over a dense, coherent
// index list.  A one-pass multi-bucket partition would save the repeated reads; shader counts are small (the AO and path
// tracing pipelines of the reference bind 1-2 closest-hit and 1-2 miss shaders).
#include <cuda_runtime.h>

#include "kernels.h"

namespace rdn {

namespace {

__global__ void __launch_bounds__(256) k_sbt_dispatch(const InstanceRecord *__restrict__ instances, uint32_t n_instances,
                                                      const SbtHitGroup *__restrict__ hit_groups, uint32_t n_hit_groups,
                                                      const uint32_t *__restrict__ miss_shaders, uint32_t n_miss, rdn_sbt_ray_config cfg,
                                                      const rdn_hit *__restrict__ hits, uint64_t n_max, uint32_t *__restrict__ task,
                                                      const uint64_t *__restrict__ n_ptr) {
  const uint64_t n = n_ptr ? (*n_ptr < n_max ? *n_ptr : n_max) : n_max;  // (a wave sized on the device)
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const uint4 ids = __ldg(reinterpret_cast<const uint4 *>(hits + k) + 1);  // geometry_id, instance_id, instance_custom_id, hit_kind
    uint32_t code = RDN_TASK_NONE;
    if (ids.y != RDN_INVALID_ID) {
      if (!(cfg.ray_flags & RDN_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER) && ids.y < n_instances) {
        const uint32_t hit_group = cfg.sbt_ray_offset + cfg.sbt_ray_stride * ids.x + __ldg(&instances[ids.y].sbt_offset);
        if (hit_group < n_hit_groups) {
          const uint32_t shader = __ldg(&hit_groups[hit_group].closest_hit);
          if (shader != RDN_SBT_NO_SHADER) code = shader & ~RDN_TASK_MISS_BIT;
        }
      }
    } else if (cfg.miss_index < n_miss) {
      const uint32_t shader = __ldg(miss_shaders + cfg.miss_index);
      if (shader != RDN_SBT_NO_SHADER) code = shader | RDN_TASK_MISS_BIT;
    }
    task[k] = code;
  }
}

__global__ void __launch_bounds__(256) k_sbt_mark(const uint32_t *__restrict__ task, uint64_t n, uint32_t code, uint8_t *__restrict__ keep,
                                                  uint32_t *__restrict__ iota) {
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    keep[k] = task[k] == code ? 1 : 0;
    iota[k] = static_cast<uint32_t>(k);
  }
}

// queue[offsets[b] + i] = segment[i] for i < *count; offsets[b + 1] = offsets[b] + *count (offsets[b] was written by the previous
// bucket's launch on the same stream, offsets[0] by the memset in front of the first)
__global__ void __launch_bounds__(256) k_sbt_append(const uint32_t *__restrict__ segment, const uint64_t *__restrict__ count, uint32_t bucket,
                                                    uint32_t *__restrict__ queue, uint64_t *__restrict__ offsets) {
  const uint64_t base = offsets[bucket], m = *count;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < m; k += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    queue[base + k] = segment[k];
  if (blockIdx.x == 0 && threadIdx.x == 0) offsets[bucket + 1] = base + m;
}

unsigned grid_for(uint64_t n) {
  const uint64_t b = (n + 255) / 256;
  return static_cast<unsigned>(b < 1 ? 1 : (b > 148ull * 32 ? 148ull * 32 : b));
}

}  // namespace

void launch_sbt_dispatch(const SceneDev &scene, const SbtHitGroup *d_hit_groups, uint32_t n_hit_groups, const uint32_t *d_miss, uint32_t n_miss,
                         const rdn_sbt_ray_config &cfg, const rdn_hit *d_hits, uint64_t n, uint32_t *d_task, cudaStream_t stream) {
  if (n) k_sbt_dispatch<<<grid_for(n), 256, 0, stream>>>(scene.instances, scene.n_instances, d_hit_groups, n_hit_groups, d_miss, n_miss, cfg, d_hits, n, d_task, nullptr);
}
void launch_sbt_dispatch_n(const SceneDev &scene, const SbtHitGroup *d_hit_groups, uint32_t n_hit_groups, const uint32_t *d_miss, uint32_t n_miss,
                           const rdn_sbt_ray_config &cfg, const rdn_hit *d_hits, const uint64_t *d_n, uint64_t n_max, uint32_t *d_task,
                           cudaStream_t stream) {
  if (n_max) k_sbt_dispatch<<<grid_for(n_max), 256, 0, stream>>>(scene.instances, scene.n_instances, d_hit_groups, n_hit_groups, d_miss, n_miss, cfg, d_hits, n_max, d_task, d_n);
}

void launch_sbt_group(const uint32_t *d_task, uint64_t n, uint32_t n_closest, uint32_t n_miss, uint8_t *d_keep, uint32_t *d_iota,
                      uint32_t *d_segment, uint64_t *d_count, unsigned long long *d_status, uint32_t *d_queue, uint64_t *d_offsets,
                      cudaStream_t stream) {
  cudaMemsetAsync(d_offsets, 0, sizeof(uint64_t) * (static_cast<size_t>(n_closest) + n_miss + 1), stream);
  if (n == 0) return;
  for (uint32_t b = 0; b < n_closest + n_miss; ++b) {
    const uint32_t code = b < n_closest ? b : ((b - n_closest) | RDN_TASK_MISS_BIT);
    k_sbt_mark<<<grid_for(n), 256, 0, stream>>>(d_task, n, code, d_keep, d_iota);
    launch_compact_u32(d_iota, d_keep, n, d_segment, d_count, d_status, stream);
    k_sbt_append<<<grid_for(n), 256, 0, stream>>>(d_segment, d_count, b, d_queue, d_offsets);
  }
}

}  // namespace rdn
