// kernels.h — launch wrappers of the CUDA kernels (traverse.cu, compact.cu) used by the C ABI (capi.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/rdn_rt.h"
#include "layout.h"

namespace rdn {

struct SbtHitGroup {  // DeviceHitGroupShaderRecord, sbt.rs:62-69
  uint32_t closest_hit, any_hit, intersection;
};

// device pointers into one blob
struct SceneDev {
  const uint32_t *tlas_binding;
  const TlasRoot *tlas_root;
  const DeviceBVHNode *tlas_bvh_forest;
  const TlasBounding *tlas_bounding;
  const InstanceRecord *instances;
  const BlasMeta *blas_meta;
  const GeometryMeta *geometry_meta;
  const DeviceBVHNode *tri_bvh_forest;
  const TriRecord *triangles;
  const SlotInfo *slot_info;
  const WideNode *wide_nodes;
  const uint32_t *prim_to_slot;
  const uint32_t *irregular_instances;
  const LeafBox *irregular_leaf_boxes;
  const Wide4Node *wide4_nodes;
  uint32_t n_tlas_binding, n_tlas_root, n_blas_meta, n_instances;
  // any-hit stage of the launch (filled per launch by capi.cu when rdn_launch.any_hit != RDN_ANYHIT_NONE): the scene's programs and,
  // for RDN_ANYHIT_FROM_SBT, the hit groups of the bound table
  const rdn_anyhit_program *anyhit_programs;
  const struct SbtHitGroup *sbt_hit_groups;
  uint32_t n_anyhit_programs, n_sbt_hit_groups;
};

// per-device scratch owned by the scene
struct TraceScratch {
  unsigned long long *work_counter;  // next ray fetch index (persistent kernel)
  uint32_t *tie_count;               // rays queued for exact tie resolution
  uint32_t *tie_unresolved;          // clamped re-walks that found nothing and were repeated over the ray's whole range (an irregular
                                     // candidate the flattener's classification did not catch; not an error, reported in the stats)
  uint32_t *stack_overflow;          // safety net: traversal stack overflow (must stay 0)
  uint32_t *tie_cursor;              // next queued tie to resolve (in-kernel drain)
  uint32_t *tie_total;               // ties of completed launches since the host last cleared it (in-kernel drain resets tie_count)
  uint32_t *epoch_done;              // ordered launches completed on this scratch set (gate of the next launch on the set)
  uint32_t *gate_timeout;            // safety net: launches that gave up waiting at that gate (must stay 0)
  uint32_t *blocks_done;             // CTAs of the running ordered kernel that have left; the last one re-arms work_counter
  uint32_t *tie_queue;               // ray indices, capacity >= rays of the launch
  float *tie_best;                   // closest distance found by the ordered kernel, per queued ray
  unsigned long long *counters;      // 6 x u64 (rdn_counters)
};

// relative slack of the ordered kernel's pruning bound and near-tie detection (see DESIGN.md "Exactness")
constexpr float TIE_EPS = 1e-5f;

// The reference's threaded pre-order walk, one ray per thread (NaiveSahBvhCpu::traverse on the device).
// d_n (optional): the ray count lives on the device (what a compaction left there), n is its upper bound; ray lists only.
void launch_trace_reference(const SceneDev &scene, const rdn_launch &launch, const rdn_ray *d_rays, uint64_t n, rdn_hit *d_hits,
                            const TraceScratch &scratch, bool count_visits, int sm_count, cudaStream_t stream,
                            const unsigned long long *d_n = nullptr);

// Re-walk of the rays queued by the ordered kernel (grid-stride over the device-side queue, no host sync in between).
void launch_resolve_ties(const SceneDev &scene, const rdn_launch &launch, const rdn_ray *d_rays, rdn_hit *d_hits,
                         const TraceScratch &scratch, int sm_count, cudaStream_t stream);

// Ordered (near child first) persistent-thread traversal with tie detection.  Near-tie rays are re-walked in the
// reference's order inside the kernel (*ties_resolved_in_kernel = true), or — RDN_ORDERED_VARIANT=9, irregular TLASes — queued in
// scratch for launch_resolve_ties (false).  Returns the launch's status.  Needs *work_counter == 0 at launch; its last CTA resets it on the way out.
// `tlas` = the TlasRoot of tlas_binding[launch.tlas_idx] (resolved by the caller from its host copy; wide_root REF_EMPTY: every
// ray misses).  Rays whose original range meets one of the TLAS's irregular instances are queued for an unclamped
// reference-order walk instead of being traversed; a TLAS marked IRREGULAR_ROUTE_ALL must not come here at all.
int ordered_tie_mode();  // 0: queue + launch_resolve_ties; 1/2: re-walk by the finishing lane; 3: queue drained inside the kernel
cudaError_t launch_trace_ordered(const SceneDev &scene, const rdn_launch &launch, const TlasRoot &tlas, const rdn_ray *d_rays, uint64_t n,
                                 rdn_hit *d_hits, const TraceScratch &scratch, int sm_count, cudaStream_t stream, bool allow_overlap,
                                 uint32_t wait_epoch,  // = ordered launches issued before this one on `scratch`
                                 bool *ties_resolved_in_kernel, const unsigned long long *d_n = nullptr, struct TileHistory *history = nullptr);
// Tile history of a grid (traverse.cu k_build_tile_lists): TILE_CLASSES lists of tile indices, capacity n_tiles each, and their lengths.
struct TileMeta {
  uint32_t count[8];
  unsigned long long sum, ran;  // durations of the passes the lists were built from (what the next build measures its classes by)
};
struct TileHistory {
  const uint32_t *lists;   // what the launch reads (null: grid order) ...
  const TileMeta *meta;
  uint32_t *cost;          // ... where it notes its pass durations (null: nowhere) ...
  TileMeta *meta_clear;    // ... and the description it zeroes for the build behind it (null: none follows)
  uint32_t n_tiles;        // ceil(width / 8) * ceil(height / 4)
  bool used;               // out: the launch took the instantiation that does all this (plain grids only)
};
void launch_build_tile_lists(uint32_t *d_cost, uint32_t *d_lists, TileMeta *d_meta, const TileMeta *d_previous, uint32_t n_tiles, cudaStream_t stream);
// can the any-hit stage of this launch stop a traversal (END_SEARCH)?  Then the answer depends on the visiting order and the launch
// takes the reference-order kernel.  (host copy of the programs / hit groups)
bool any_hit_can_end_search(const rdn_launch &launch, const rdn_anyhit_program *programs, uint32_t n_programs, const SbtHitGroup *groups,
                            uint32_t n_groups);

// Stable stream compaction of u32 (single pass, decoupled look-back); d_status needs compact_status_words(n) u64.
uint64_t compact_status_words(uint64_t n);
void launch_compact_u32(const uint32_t *d_in, const uint8_t *d_keep, uint64_t n, uint32_t *d_out, uint64_t *d_out_n,
                        unsigned long long *d_status, cudaStream_t stream);

// ray generation and the closest-hit -> bounce step on the device (raygen.cu)
void launch_gen_pinhole_rays(const rdn_pinhole &p, rdn_ray *d_rays, cudaStream_t stream);
void launch_gen_pinhole_rays_batch(const rdn_pinhole *d_params, const uint64_t *d_offsets, uint32_t n_params, uint64_t max_rays_per_param,
                                   rdn_ray *d_rays, cudaStream_t stream);  // n_params <= 65535 per call
void launch_gen_camera_rays(const rdn_camera &p, rdn_ray *d_rays, cudaStream_t stream);
void launch_mark_hits(const rdn_hit *d_hits, uint64_t n, uint8_t *d_keep, uint32_t *d_iota, cudaStream_t stream);
// AO frame accumulation (feature/ao.rs:187-232); d_payload: n_pixels floats, 1.0f before the first sample (the kernel re-arms it)
void launch_fill_f32(float *d_dst, uint64_t n, float v, cudaStream_t stream);
void launch_ao_accumulate(const rdn_hit *d_secondary_hits, const uint32_t *d_src_index, const uint64_t *d_n_secondary, uint64_t n_pixels,
                          uint32_t sample_count, uint32_t max_sample, float *d_payload, float *d_ao_buffer, cudaStream_t stream);
// d_src_index[0 .. *d_n_src) = indices of the source rays (stable compaction of the hits); n_max bounds the grid
void launch_gen_bounce_rays(const SceneDev &scene, const rdn_bounce &p, const rdn_ray *d_rays_in, const rdn_hit *d_hits,
                            const uint32_t *d_src_index, const uint64_t *d_n_src, uint64_t n_max, rdn_ray *d_rays_out, cudaStream_t stream);

// brute-force mesh picking (pick.cu)
struct PickMeshDev {
  const float *positions;   // 3 floats per vertex
  const uint32_t *indices;  // nullptr: non-indexed
  uint64_t n_prims;         // (count + step - stride) / step, access.rs:142-150
  uint32_t topology;        // rdn_topology
};
// d_best_key / d_first_hit: one slot per ray, 0xFF-filled before the first call (the finish kernel re-arms them)
void launch_pick_nearest(const PickMeshDev &mesh, const rdn_ray *d_rays, uint64_t n_rays, float tolerance, uint32_t face_side,
                         unsigned long long *d_best_key, uint32_t *d_first_hit, rdn_mesh_hit *d_out, int sm_count, cudaStream_t stream);
void launch_pick_all_mark(const PickMeshDev &mesh, const rdn_ray *d_ray, float tolerance, uint32_t face_side, uint8_t *d_keep, uint32_t *d_iota,
                          rdn_mesh_hit *d_records, int sm_count, cudaStream_t stream);
void launch_pick_all_gather(const uint32_t *d_index, const uint64_t *d_n_kept, uint64_t capacity, const rdn_mesh_hit *d_records, rdn_mesh_hit *d_out,
                            int sm_count, cudaStream_t stream);

// binned-SAH build on the device (build_device.cu): 0 = built (node for node the host builder's tree),
// 1 = not supported there (buckets > 4, a long degenerate range, ...: use the host builder), < 0 = error
struct FlattenBVH;
struct TreeBuildOption;
struct Box3;
int build_bvh_sah_device(const Box3 *boxes, uint64_t n, uint32_t n_buckets, const TreeBuildOption &option, int device, FlattenBVH &out,
                         std::string &err);

// shader binding table dispatch (sbt.cu)

void launch_sbt_dispatch(const SceneDev &scene, const SbtHitGroup *d_hit_groups, uint32_t n_hit_groups, const uint32_t *d_miss, uint32_t n_miss,
                         const rdn_sbt_ray_config &cfg, const rdn_hit *d_hits, uint64_t n, uint32_t *d_task, cudaStream_t stream);
// d_keep: n bytes, d_iota / d_segment: n u32, d_count: one u64, d_status: compact_status_words(n) u64
void launch_sbt_group(const uint32_t *d_task, uint64_t n, uint32_t n_closest, uint32_t n_miss, uint8_t *d_keep, uint32_t *d_iota,
                      uint32_t *d_segment, uint64_t *d_count, unsigned long long *d_status, uint32_t *d_queue, uint64_t *d_offsets,
                      cudaStream_t stream);

// one-call wavefront executor (wavefront.cu): glue kernels between the rounds of rdn_rt_trace_ray
void launch_wave_gather(const rdn_ray *d_next_rays, const uint32_t *d_launch_in, const uint32_t *d_idx, const uint64_t *d_count, uint64_t n_max,
                        rdn_ray *d_rays_out, uint32_t *d_launch_out, cudaStream_t stream);
void launch_wave_mark(const uint32_t *d_task, const uint64_t *d_wave_size, uint64_t n_max, uint32_t code, uint8_t *d_keep, uint32_t *d_iota,
                      cudaStream_t stream);
void launch_wave_clip_spawn(uint8_t *d_spawn, const uint64_t *d_wave_size, uint64_t n_max, uint32_t *d_iota, cudaStream_t stream);
void launch_wave_record(const uint64_t *const *d_counters, uint32_t n, uint64_t *d_row, cudaStream_t stream);
void launch_store_u64(uint64_t *d_dst, uint64_t v, cudaStream_t stream);
void launch_stage_spawn_all(uint8_t *d_spawn, const uint64_t *d_count, uint64_t n_max, cudaStream_t stream);
void launch_stage_store_f32(const uint32_t *d_tasks, const uint64_t *d_count, uint64_t n_max, const uint32_t *d_launch_index, float value,
                            float *d_dst, cudaStream_t stream);
// the closest-hit -> next ray step as a stage: source rays are the tasks of `d_tasks`, ray k goes to slot d_tasks[k] of d_rays_out
// with d_spawn[slot] = 1; sample indices come from the launch index when d_launch_index is given
void launch_stage_bounce_rays(const SceneDev &scene, const rdn_bounce &p, const rdn_ray *d_rays_in, const rdn_hit *d_hits, const uint32_t *d_tasks,
                              const uint64_t *d_count, uint64_t n_max, const uint32_t *d_launch_index, rdn_ray *d_rays_out, uint8_t *d_spawn,
                              cudaStream_t stream);
// device-sized variant of launch_sbt_dispatch: rays [0, *d_n) of an upper bound n_max
void launch_sbt_dispatch_n(const SceneDev &scene, const SbtHitGroup *d_hit_groups, uint32_t n_hit_groups, const uint32_t *d_miss, uint32_t n_miss,
                           const rdn_sbt_ray_config &cfg, const rdn_hit *d_hits, const uint64_t *d_n, uint64_t n_max, uint32_t *d_task,
                           cudaStream_t stream);
void launch_ao_resolve(float *d_payload, uint64_t n_pixels, uint32_t sample_count, uint32_t max_sample, float *d_ao_buffer, cudaStream_t stream);

// measurement hook (probe.cu): read bandwidth of an L2-resident buffer of `bytes` on the current device, GB/s
int measure_l2_read_gbs(uint64_t bytes, int passes, int sm_count, double *out_gbs);

// path A: intersect_nearest_bvh over a FlattenBVH (content/mesh/core/src/feature/bvh.rs:57-86); the kernel walks
// the tree in the reference's order (right child first, no distance pruning) so equal-distance ties resolve identically.
constexpr int PATHA_MAX_DEPTH = 128;
struct PathANode {  // 32 B: box + (left_count | leaf range)
  float bmin[3]; uint32_t a;   // inner: right child index; leaf: first slot
  float bmax[3]; uint32_t b;   // inner: 0xFFFFFFFF;        leaf: one past the last slot (leaf iff b != 0xFFFFFFFF)
};
struct PathATri {  // 48 B: the triangle of a slot (sorted_primitive_index order), pre-gathered through indices -> positions
  float a[3]; uint32_t prim;   // original primitive index (MeshBufferHitPoint::primitive_index)
  float b[3]; uint32_t pad0;
  float c[3]; uint32_t pad1;
};
void launch_patha_nearest(const PathANode *d_nodes, const PathATri *d_tris, const rdn_ray *d_rays, uint64_t n, uint32_t face_side,
                          rdn_mesh_hit *d_out, cudaStream_t stream);

// intersect_list_bvh: d_out == nullptr -> per-ray hit counts into d_counts; else hits written at d_offsets[ray] in visiting order
void launch_patha_list(const PathANode *d_nodes, const PathATri *d_tris, const rdn_ray *d_rays, uint64_t n, uint32_t face_side,
                       uint32_t *d_counts, const uint64_t *d_offsets, rdn_mesh_hit *d_out, cudaStream_t stream);

}  // namespace rdn
