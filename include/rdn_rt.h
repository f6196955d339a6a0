/* rdn_rt.h — C ABI of the B200-native BVH closest-hit path (drop-in boundary).
 *
 * What a thin Rust `-sys` crate (or any FFI) binds in place of rendiation's software ray-tracing
 * geometry backend and its space-query helpers.  Every entry point cites the reference interface
 * it replaces (paths relative to the rendiation repo root).
 *
 * Conventions (following the reference's own C API style,
 * application/viewer-content-api/src/c_api/viewer.rs:295-336, but with status codes instead of
 * panics — the reference aborts on error, Cargo.toml:161-162):
 *   - every function returns 0 on success, a negative rdn_status on error; never throws/aborts;
 *   - rdn_rt_last_error() returns a thread-local message for the last failure;
 *   - the caller owns every array passed in or out; build inputs are copied by the callee
 *     (`source.to_vec()`, geometry/naive/mod.rs:98-119); the scene object owns all device memory;
 *   - create/delete/bind take a write lock, trace a read lock (the Arc<RwLock<..>> discipline of
 *     geometry/naive/mod.rs:497-536), so a scene may be shared between threads.
 */
#ifndef RDN_RT_H
#define RDN_RT_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rdn_status {
  RDN_OK = 0,
  RDN_ERR_INVALID_ARGUMENT = -1,
  RDN_ERR_CUDA = -2,
  RDN_ERR_INVALID_HANDLE = -3,   /* the reference panics: unwrap() on a deleted BLAS, naive/mod.rs:273-275 */
  RDN_ERR_BUILD = -4,            /* the reference panics inside the builder (index out of bounds) */
  RDN_ERR_NOT_COMMITTED = -5,
  RDN_ERR_CAPACITY = -6          /* scene exceeds the 32-bit reference encoding of the flattened layout */
} rdn_status;

/* ---- ray flags: RayFlagConfigRaw, shader/ray-tracing/src/api/ty.rs:102-114 ---- */
#define RDN_RAY_FLAG_NONE 0x00u
#define RDN_RAY_FLAG_FORCE_OPAQUE 0x01u
#define RDN_RAY_FLAG_FORCE_NON_OPAQUE 0x02u
#define RDN_RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH 0x04u
#define RDN_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER 0x08u
#define RDN_RAY_FLAG_CULL_BACK_FACING_TRIANGLES 0x10u
#define RDN_RAY_FLAG_CULL_FRONT_FACING_TRIANGLES 0x20u
#define RDN_RAY_FLAG_CULL_OPAQUE 0x40u
#define RDN_RAY_FLAG_CULL_NON_OPAQUE 0x80u
#define RDN_RAY_FLAG_SKIP_TRIANGLES 0x100u
#define RDN_RAY_FLAG_SKIP_PROCEDURAL_PRIMITIVES 0x200u
/* ---- GeometryInstanceFlags, api/ty.rs:143-151 ---- */
#define RDN_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE 0x1u
#define RDN_GEOMETRY_INSTANCE_TRIANGLE_FLIP_FACING 0x2u
#define RDN_GEOMETRY_INSTANCE_FORCE_OPAQUE 0x4u
#define RDN_GEOMETRY_INSTANCE_FORCE_NO_OPAQUE 0x8u
/* ---- GeometryFlags, api/ty.rs:155-159 ---- */
#define RDN_GEOMETRY_FLAG_OPAQUE 0x1u
#define RDN_GEOMETRY_FLAG_NO_DUPLICATE_ANYHIT_INVOCATION 0x2u
/* ---- RayHitKind, api/ty.rs:138-140 ---- */
#define RDN_HIT_KIND_FRONT_FACING_TRIANGLE 0xFEu
#define RDN_HIT_KIND_BACK_FACING_TRIANGLE 0xFFu
#define RDN_INVALID_ID 0xFFFFFFFFu

/* One ray: the reference's `Ray` (geometry/mod.rs:33-42: origin, flags, direction, mask) with the two
 * u32 slots carrying the per-ray range of ShaderRayTraceCallStoragePayload.range
 * (wavefront_compute/ctx.rs:15-29).  32 B = 2 x float4. */
typedef struct rdn_ray { float ox, oy, oz, tmin, dx, dy, dz, tmax; } rdn_ray;

/* Launch-uniform part of ShaderRayTraceCallStoragePayload (ctx.rs:15-29): tlas_idx, ray_flags, cull_mask, the SBT ray
 * configuration and miss index.
 * grid_width: optional hint — rays form a row-major 2D launch of this width (launch_size.x); 0 = plain list.
 * The result is identical either way; the hint lets the kernel walk rays in 8x4 pixel tiles and, for device-resident launches,
 * start the tiles that took longest in the previous launch over a grid of the same size first.
 * any_hit: what the any-hit stage of the pipeline decides for candidate hits of NON-OPAQUE geometry (traverse_cpu.rs:164-192):
 * RDN_ANYHIT_NONE = every candidate is accepted (the reference without an any-hit shader), k + 1 = program k of
 * rdn_rt_set_any_hit_programs for all non-opaque geometry, RDN_ANYHIT_FROM_SBT = the any_hit handle of the candidate's hit group
 * in the table bound with rdn_rt_bind_sbt (hit group = sbt_ray_offset + sbt_ray_stride * geometry_id + the instance's record
 * offset, api/ctx.rs:53-55; trace_task.rs:189-203), RDN_SBT_NO_SHADER there = accepted.
 * A zero-filled tail (any_hit, sbt_ray_offset, sbt_ray_stride, miss_index) is the plain closest-hit query. */
typedef struct rdn_launch {
  uint32_t ray_flags, cull_mask, tlas_idx, grid_width;
  uint32_t any_hit, sbt_ray_offset, sbt_ray_stride, miss_index;
} rdn_launch;
#define RDN_ANYHIT_NONE 0u
#define RDN_ANYHIT_FROM_SBT 0xFFFFFFFFu

/* Closest-hit record = the fields of RayClosestHitCtx a consumer reads (api/ctx.rs:13-55;
 * storage forms wavefront_compute/ctx.rs:34-57): hit_distance (world), bary_coord (u,v) with
 * p = v0 + u*e1 + v*e2, primitive_id (original triangle index of the geometry), geometry_id,
 * instance_id (slot in BVH-sorted tlas_data), instance_custom_id, hit_kind.
 * Miss: instance_id == primitive_id == RDN_INVALID_ID, t == ray.tmax, hit_kind == 0. */
typedef struct rdn_hit {
  float t, u, v;
  uint32_t primitive_id, geometry_id, instance_id, instance_custom_id, hit_kind;
} rdn_hit;

/* BottomLevelAccelerationStructureBuildSource, api/backend.rs:104-118.
 * kind 0 = Triangles{positions, indices (NULL => non-indexed)}, kind 1 = AABBs (accepted and ignored,
 * exactly as the reference's naive builder does, geometry/naive/mod.rs:201-237). */
typedef struct rdn_blas_geometry {
  const float *positions;   /* 3 floats per vertex (or 6 per AABB) */
  uint64_t n_positions;     /* vertex (or AABB) count */
  const uint32_t *indices;  /* may be NULL */
  uint64_t n_indices;
  uint32_t flags;           /* RDN_GEOMETRY_FLAG_* */
  uint32_t kind;
} rdn_blas_geometry;

/* TopLevelAccelerationStructureSourceInstance, api/backend.rs:160-167.  transform is the reference's
 * column-major Mat4 (a1..d4, math/algebra/src/mat/mat4.rs:9-14). */
typedef struct rdn_instance {
  float transform[16];
  uint32_t instance_custom_index, mask, instance_shader_binding_table_record_offset, flags, blas_handle;
} rdn_instance;

/* The reference's traversal counters (naive/traverse_cpu.rs:37-41) + instances entered + candidates the
 * reference's RayRange::update_far asserts would have aborted on (traverse_cpu.rs:272-276). */
typedef struct rdn_counters { uint64_t bvh_visit, bvh_hit, tri_visit, tri_hit, inst_visit, ref_abort; } rdn_counters;

/* traversal order selection for rdn_rt_trace_closest_device */
typedef enum rdn_trace_mode {
  RDN_TRACE_AUTO = 0,             /* ordered traversal + exact tie resolution (default) */
  RDN_TRACE_REFERENCE_ORDER = 1,  /* the reference's threaded pre-order walk, one ray per thread */
  /* OR-ed into the order selection: this launch may start while the last long rays of the PREVIOUS rdn_rt_trace_closest_device
   * call on the same stream are still being walked (programmatic dependent launch; +45 % on back-to-back frames).
   * The caller promises that (1) nothing else was put into that stream since that call — a kernel of the caller's own that
   * produces this launch's rays would not be waited for — and (2) this launch's d_rays / d_hits do not alias the previous launch's
   * d_hits, nor its d_hits the previous d_rays (stragglers of the previous launch still read and write them).  The library checks
   * what it can see (same stream, no other entry point of this library in between, the four address ranges) and launches in plain
   * stream order when the check fails or the flag is absent. */
  RDN_TRACE_OVERLAP_PREVIOUS = 0x100
} rdn_trace_mode;

typedef struct rdn_trace_stats {
  uint64_t rays, tie_rays;       /* rays re-walked in reference order: near-ties + rays whose range meets an irregular instance */
  uint32_t kernel_launches;      /* kernels this call launched */
  float kernel_ms;               /* device time of the traversal kernels (CUDA events on the call's stream) */
  uint64_t whole_range_rewalks;  /* near-tie re-walks that had to be repeated over the ray's whole range (the closest candidate
                                    lay outside its boxes: an irregular triangle the build-time classification did not flag) */
} rdn_trace_stats;

/* What the flattener found (rdn_rt_scene_build_stats).  "Irregular" triangles / instances are those whose hits need not
 * lie inside their bounding boxes — needle triangles, singular / non-affine / ill-conditioned instance transforms, and the
 * reference's per-geometry blas_box indexing (naive/mod.rs:239,273) — so that NaiveSahBvhCpu::traverse's answer depends on
 * its visiting order; rays that can reach them are walked in that order (DESIGN.md "Exactness"). */
typedef struct rdn_build_stats {
  uint64_t balance_fallbacks, balance_fallbacks_gt10;  /* SAH -> BalanceTree fallbacks (strategy.rs:230-233); ... over > 10 primitives */
  uint64_t irregular_triangles, irregular_instances;
  uint64_t reference_routed_tlas;                      /* TLASes whose every ray takes the reference-order kernel */
  double bvh_build_ms, flatten_ms, upload_ms;          /* wall clock of the last commit: tree builds / flattening / blob + upload */
  uint64_t build_threads;                              /* worker threads the tree builder used (RDN_BUILD_THREADS caps it) */
  uint64_t device_built_trees;                         /* geometry trees built by the device SAH builder (RDN_COMMIT_DEVICE_BUILD=1) */
  uint64_t tlas_only_commits;                          /* commits so far that kept every BLAS array and patched the TLAS part of the device
                                                          blobs in place (rdn_rt_tlas_update) */
  uint64_t kernels_enqueued;                           /* traversal-path kernels this scene has put into streams so far (ordered /
                                                          reference-order / tie / tile-list kernels): a caller that wants to know what a
                                                          loop launched reads it before and after */
} rdn_build_stats;

typedef struct rdn_rt_scene rdn_rt_scene;   /* opaque: NaiveSahBVHSystem (geometry/naive/mod.rs:495-610) */

/* ---- scene lifetime ----
 * n_devices CUDA devices; the flattened scene is replicated to each on commit and host-buffer traces
 * are sharded across them by ray tile.  (One process per GPU uses n_devices = 1.) */
int rdn_rt_scene_create(int n_devices, const int *device_ids, rdn_rt_scene **out);
void rdn_rt_scene_destroy(rdn_rt_scene *scene);

/* ---- GPUAccelerationStructureSystemProvider (api/backend.rs:120-142) ---- */
int rdn_rt_blas_create(rdn_rt_scene *scene, const rdn_blas_geometry *geometries, uint32_t n, uint32_t *out_handle);
                                                            /* create_bottom_level_acceleration_structure */
int rdn_rt_blas_destroy(rdn_rt_scene *scene, uint32_t handle); /* delete_bottom_level_acceleration_structure */
int rdn_rt_tlas_create(rdn_rt_scene *scene, const rdn_instance *instances, uint32_t n, uint32_t *out_handle);
                                                            /* create_top_level_acceleration_structure */
int rdn_rt_tlas_destroy(rdn_rt_scene *scene, uint32_t handle); /* delete_top_level_acceleration_structure */
/* Replace the instances of a live TLAS (moving objects: same handle, new transforms).  No reference counterpart — its build is
 * "todo incremental change" (naive/mod.rs:121) and any mutation invalidates everything (:546-549).  The next commit after TLAS-only
 * mutations (this call, TLAS create / destroy, bind_tlas) rebuilds the TLAS trees alone — the reference's build_tlas on the new
 * instances, same result as a from-scratch build — keeps every BLAS array, and, when no array changes its length, patches the
 * device blobs in place instead of uploading the scene again. */
int rdn_rt_tlas_update(rdn_rt_scene *scene, uint32_t handle, const rdn_instance *instances, uint32_t n);
int rdn_rt_bind_tlas(rdn_rt_scene *scene, const uint32_t *handles, uint32_t n);   /* bind_tlas */
uint32_t rdn_rt_bind_tlas_max_len(const rdn_rt_scene *scene);                     /* bind_tlas_max_len */

/* Lazy build -> flatten -> upload -> replicate; what create_comp_instance triggers through
 * get_or_build_gpu_data (geometry/naive/mod.rs:521-536,561-567).  Idempotent; traces call it implicitly. */
int rdn_rt_commit(rdn_rt_scene *scene);

/* ---- ...InvocationTraversable::traverse, batched (geometry/mod.rs:16-25; CPU twin
 *      NaiveSahBvhCpu::traverse, naive/traverse_cpu.rs:52-245) ---- */
/* host buffers: H2D, traversal, D2H pipelined inside the call; sharded over the scene's devices */
int rdn_rt_trace_closest(rdn_rt_scene *scene, const rdn_launch *launch, const rdn_ray *rays, uint64_t n,
                         rdn_hit *out_hits);
/* Page-locked host memory for the call above (and every other host-buffer entry point).  The call takes any host pointer: ordinary
 * (pageable) memory is staged chunk by chunk through page-locked buffers of the library by several host threads (about two thirds
 * of the page-locked rate, bound by the host's memory bandwidth; left to the driver it was a fifth), but ray
 * and hit buffers that live for more than a frame should be page-locked: allocated with rdn_rt_host_alloc (freed with
 * rdn_rt_host_free), or — memory the caller owns, e.g. a Vec — registered with rdn_rt_host_register for as long as it lives and
 * unregistered before it is freed.  The library never registers caller memory by itself: it cannot see the free. */
int rdn_rt_host_alloc(uint64_t bytes, void **out);
void rdn_rt_host_free(void *ptr);
int rdn_rt_host_register(void *ptr, uint64_t bytes);
int rdn_rt_host_unregister(void *ptr);
/* device-resident buffers on device `device_index` (index into the scene's device list); asynchronous on
 * `cuda_stream` (a cudaStream_t, 0 = default stream) unless `stats` is non-NULL (then it synchronises).
 * d_rays and d_hits must be 32-byte aligned (any cudaMalloc pointer plus a whole number of records is). */
int rdn_rt_trace_closest_device(rdn_rt_scene *scene, int device_index, const rdn_launch *launch,
                                const rdn_ray *d_rays, uint64_t n, rdn_hit *d_hits, void *cuda_stream,
                                int mode, rdn_trace_stats *stats);
/* The same for a wave whose size a previous kernel left ON THE DEVICE (the output count of a compaction / bounce step): the
 * launch covers rays [0, min(*d_n, n_max)), *d_n is read by the kernel when it runs, and nothing comes back to the host — the
 * reference reads the size of every wave back after its compaction (task-graph/src/runtime/task_group.rs:259-277).  d_rays and
 * d_hits hold n_max records (<= 2^31); records at and beyond *d_n are left untouched.  Ray lists only (launch->grid_width is ignored). */
int rdn_rt_trace_closest_device_n(rdn_rt_scene *scene, int device_index, const rdn_launch *launch, const rdn_ray *d_rays,
                                  const uint64_t *d_n, uint64_t n_max, rdn_hit *d_hits, void *cuda_stream, int mode);
/* Safety-net flags of the asynchronous device path.  A launch that overflowed the 120-entry traversal stack (deeper than the
 * builder's depth limits allow) or gave up waiting for its predecessor wrote unreliable records; the synchronous paths
 * (host buffers, `stats`) return the error themselves, the asynchronous one is polled: rdn_rt_poll_errors waits for
 * `cuda_stream`, returns the flags raised on `device_index` since the last poll in *out_flags (may be NULL) and clears them;
 * return value RDN_OK, RDN_ERR_CAPACITY (stack overflow) or RDN_ERR_CUDA (gate timeout). */
#define RDN_ERROR_FLAG_STACK_OVERFLOW 1u
#define RDN_ERROR_FLAG_GATE_TIMEOUT 2u
int rdn_rt_poll_errors(rdn_rt_scene *scene, int device_index, void *cuda_stream, uint32_t *out_flags);
/* ---- any-hit (row a19 / f4): RayAnyHitBehavior bits (api/ty.rs:134-136) and the any-hit shaders of a pipeline as DATA.
 *      The reference's any-hit shader is arbitrary EDSL code run inside the traversal for every candidate of non-opaque geometry;
 *      a precompiled kernel cannot take device code through a C ABI, so a shader is one of a few stateless programs over the fields
 *      of the reference's `Hit` (geometry_idx, primitive_idx, distance — traverse_cpu.rs:43-49): `behavior` where its predicate
 *      holds, `otherwise` where not.  ACCEPT_HIT commits the candidate (the range shrinks, the result is replaced), END_SEARCH
 *      stops the whole traversal — with or without ACCEPT_HIT, exactly as traverse_cpu.rs:177-192.  Launches whose programs can
 *      return END_SEARCH have an order-dependent answer and take the reference-order kernel, like ACCEPT_FIRST_HIT rays. */
#define RDN_ANYHIT_BEHAVIOR_ACCEPT_HIT 1u
#define RDN_ANYHIT_BEHAVIOR_END_SEARCH 2u
typedef enum rdn_anyhit_kind {
  RDN_ANYHIT_CONSTANT = 0,        /* always `behavior` (the reference's tests: TEST_ANYHIT_BEHAVIOR, naive/test.rs:7) */
  RDN_ANYHIT_PRIMITIVE_MASK = 1,  /* (primitive_idx & mask) == value: a cut-out pattern over the triangles */
  RDN_ANYHIT_MIN_DISTANCE = 2     /* distance >= distance: candidates nearer than that are seen through */
} rdn_anyhit_kind;
typedef struct rdn_anyhit_program { uint32_t kind, behavior, otherwise, mask, value; float distance; uint32_t pad0, pad1; } rdn_anyhit_program;
/* the programs an SBT's any_hit handles (and rdn_launch.any_hit) name by index; copied; n = 0 removes them */
int rdn_rt_set_any_hit_programs(rdn_rt_scene *scene, const rdn_anyhit_program *programs, uint32_t n);
/* reference-order walk with the reference's visit counters (host buffers; for parity / bytes model) */
int rdn_rt_trace_counted(rdn_rt_scene *scene, const rdn_launch *launch, const rdn_ray *rays, uint64_t n,
                         rdn_hit *out_hits, rdn_counters *out_counters);

/* ---- f4 (first slice): shader binding table and the dispatch that follows a trace.
 *      Replaces ShaderBindingTableProvider (shader/ray-tracing/src/api/backend.rs:90-101), GPURayTracingDeviceProvider::create_sbt
 *      (api/backend.rs:74-79; wavefront_compute/mod.rs:78-93, sbt.rs:13-50,164-230) and the closest-hit / miss shader selection of
 *      TraceTaskImpl::device_poll (wavefront_compute/trace_task.rs:206-268, api/ctx.rs:53-55, sbt.rs:252-273).
 *      A shader is named by its ShaderHandle value (u32); RDN_SBT_NO_SHADER = the reference's None / u32::MAX.
 *      The table holds max_geometry_count_in_blas * max_tlas_offset * ray_type_count hit groups (all empty at creation),
 *      ray_type_count miss shaders and one ray generation shader; hit group (geometry_idx, tlas_offset, ray_ty_idx) sits at
 *      ray_ty_idx + geometry_idx * ray_type_count + tlas_offset, as in the reference. */
#define RDN_SBT_NO_SHADER 0xFFFFFFFFu
#define RDN_TASK_NONE 0xFFFFFFFFu        /* nothing is spawned for this ray */
#define RDN_TASK_MISS_BIT 0x80000000u    /* task code = miss shader | RDN_TASK_MISS_BIT, or the closest-hit shader */
typedef struct rdn_sbt rdn_sbt;
typedef struct rdn_sbt_ray_config {      /* the launch-uniform part of ShaderRayTraceCallStoragePayload that dispatch reads */
  uint32_t ray_flags;                    /* RDN_RAY_FLAG_SKIP_CLOSEST_HIT_SHADER: hits spawn nothing */
  uint32_t sbt_ray_offset, sbt_ray_stride;   /* RaySBTConfig { offset, stride } */
  uint32_t miss_index;
} rdn_sbt_ray_config;
/* the executor's current_sbt (wavefront_compute/mod.rs:147-151): the table RDN_ANYHIT_FROM_SBT launches and rdn_rt_trace_ray read; NULL unbinds */
int rdn_rt_bind_sbt(rdn_rt_scene *scene, rdn_sbt *sbt);
int rdn_sbt_create(rdn_rt_scene *scene, uint32_t max_geometry_count_in_blas, uint32_t max_tlas_offset, uint32_t ray_type_count, rdn_sbt **out);
void rdn_sbt_destroy(rdn_sbt *sbt);
int rdn_sbt_config_ray_generation(rdn_sbt *sbt, uint32_t shader);
int rdn_sbt_config_hit_group(rdn_sbt *sbt, uint32_t geometry_idx, uint32_t tlas_offset, uint32_t ray_ty_idx, uint32_t closest_hit,
                             uint32_t any_hit, uint32_t intersection);
int rdn_sbt_config_missing(rdn_sbt *sbt, uint32_t ray_ty_idx, uint32_t shader);
int rdn_sbt_ray_generation(const rdn_sbt *sbt, uint32_t *out_shader);
/* d_task[i] = task code of ray i from its hit record (device arrays of scene device `device_index`, asynchronous on `cuda_stream`).
 * A hit group index beyond the table selects nothing (the reference reads out of bounds there). */
int rdn_rt_sbt_dispatch_device(rdn_rt_scene *scene, int device_index, rdn_sbt *sbt, const rdn_sbt_ray_config *config, const rdn_hit *d_hits,
                               uint64_t n, uint32_t *d_task, void *cuda_stream);
/* The per-shader task lists: ray indices grouped by task code, ray order kept inside a group.  Groups: closest-hit shaders
 * 0 .. n_closest_shaders-1, then miss shaders 0 .. n_miss_shaders-1; group g is d_queue[d_offsets[g] .. d_offsets[g+1]).
 * d_queue: n u32, d_offsets: n_closest_shaders + n_miss_shaders + 1 u64.  Rays with RDN_TASK_NONE (or a shader beyond the
 * counts) are in no group. */
int rdn_rt_sbt_group_device(rdn_rt_scene *scene, int device_index, rdn_sbt *sbt, const uint32_t *d_task, uint64_t n, uint32_t n_closest_shaders,
                            uint32_t n_miss_shaders, uint32_t *d_queue, uint64_t *d_offsets, void *cuda_stream);
/* both steps on host buffers (synchronous; device 0 of the scene); task / queue may be NULL when not wanted */
int rdn_rt_sbt_dispatch(rdn_rt_scene *scene, rdn_sbt *sbt, const rdn_sbt_ray_config *config, const rdn_hit *hits, uint64_t n,
                        uint32_t n_closest_shaders, uint32_t n_miss_shaders, uint32_t *task, uint32_t *queue, uint64_t *offsets);

/* ---- measurement hook (no reference counterpart): between _begin and _end every traversal kernel launched on the
 *      device is bracketed by CUDA events on its launching stream; _end waits for them and returns the summed device
 *      time per kernel (k_trace_ordered_rounds / k_resolve_ties / k_trace_reference). ---- */
typedef struct rdn_kernel_times {
  uint64_t ordered_launches, tie_launches, reference_launches;
  double ordered_ms, tie_ms, reference_ms;
} rdn_kernel_times;
int rdn_rt_kernel_timing_begin(rdn_rt_scene *scene, int device_index);
int rdn_rt_kernel_timing_end(rdn_rt_scene *scene, int device_index, rdn_kernel_times *out);

/* ---- the step on either side of the traversal, on the device: primary-ray generation and the
 *      closest-hit -> bounce-ray step, so a frame never moves rays or hits through the host.
 * All buffers are device pointers on device `device_index`; calls are asynchronous on `cuda_stream`. ---- */
/* Pinhole grid of the reference's own trace tests (geometry/naive/test.rs:259-264): for pixel (i, j) of a width x height
 * launch  x = (i + jitter_x) / width * 2 - 1 [* aspect],  y = 1 - (j + jitter_y) / height * 2,  d = normalize((x, y, -1) - origin).
 * Rays of the sub-rectangle (rect_x, rect_y, rect_w, rect_h) are written row-major (a launch tile, rdn multi-GPU sharding). */
typedef struct rdn_pinhole {
  uint32_t width, height, rect_x, rect_y, rect_w, rect_h;
  float origin[3], tmin, tmax, aspect, jitter_x, jitter_y;   /* aspect 1 = none; jitter 0.5 = pixel centre */
} rdn_pinhole;
int rdn_rt_gen_pinhole_rays_device(rdn_rt_scene *scene, int device_index, const rdn_pinhole *params, rdn_ray *d_rays,
                                   void *cuda_stream);
/* n_params rectangles (launch tiles x samples) in ONE kernel launch: the rays of params[k] follow those of params[k-1] in d_rays */
int rdn_rt_gen_pinhole_rays_batch_device(rdn_rt_scene *scene, int device_index, const rdn_pinhole *params, uint32_t n_params,
                                         rdn_ray *d_rays, void *cuda_stream);
/* DefaultRtxCameraInvocation::generate_ray (scene/rendering/gpu-ray-tracing/src/camera.rs:66-98):
 * uv = pixel / size + sampler.next_2d() / size with PCGRandomSampler seeded by xxhash32(pixel.x, pixel.y, sample_index)
 * (sampler.rs:11-72); target = unproject(view_projection_inv, uv, ndc_depth) (shader/library/src/lib.rs:18-28);
 * d = normalize(target - world_position).  view_projection_inv is the reference's column-major Mat4. */
typedef struct rdn_camera {
  float view_projection_inv[16], world_position[3], ndc_depth, tmin, tmax;
  uint32_t width, height, rect_x, rect_y, rect_w, rect_h, sample_index, pad;
} rdn_camera;
int rdn_rt_gen_camera_rays_device(rdn_rt_scene *scene, int device_index, const rdn_camera *params, rdn_ray *d_rays,
                                  void *cuda_stream);
/* One bounce ray per primary hit, stably compacted (misses produce no ray): origin = hit_world_position
 * (api/ctx.rs:209-211), normal = geometric normal turned towards the ray origin (bindless_mesh_bridge.rs:103-114),
 * direction by `mode`:
 *   0  cosine_sample_hemisphere_in_dir(normal, (van_der_corput, sobol)(index_base + source ray index)) with the given
 *      scrambles (math/statistics/src/distribution_map.rs:10-58, sampling/sobol.rs:40-68)
 *   1  tbn(normal) * sample_hemisphere_cos(hammersley_2d(sample_index, max_sample)), the AO secondary ray
 *      (feature/ao.rs:249-284, shader/library/src/sampling.rs:33-83)
 *   2  the shadow-test ray of the path tracer's closest-hit stage (feature/path_tracing/ray_hit.rs:20-45) towards a point light
 *      at `target` (lighting_bridge.rs:86-94): to_light = target - hit_world_position, distance = |to_light|,
 *      direction = to_light / distance, range [tmin, distance]; trace it with RDN_RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH
 * flags & RDN_BOUNCE_OFFSET_ORIGIN: the ray starts at offset_ray_hit(hit_world_position, geometric normal)
 * (scene/rendering/gpu-ray-tracing/src/ray_util.rs:6-40, the integer-offset self-intersection guard of Ray Tracing Gems ch. 6)
 * as ray_hit.rs:26,82 does; directions are computed from the un-offset position, as there.
 * d_rays_out and d_src_index need n slots; d_src_index[k] = index of the primary ray behind bounce ray k;
 * *d_out_n (device) = number of bounce rays.  Uses the scene's compaction scratch (rdn_rt_compact_u32_device). */
#define RDN_BOUNCE_OFFSET_ORIGIN 0x1u
typedef struct rdn_bounce {
  uint32_t mode, index_base, scramble0, scramble1, sample_index, max_sample;
  float tmin, tmax;
  uint32_t flags;
  float target[3];
} rdn_bounce;
int rdn_rt_gen_bounce_rays_device(rdn_rt_scene *scene, int device_index, const rdn_bounce *params, const rdn_ray *d_rays_in,
                                  const rdn_hit *d_hits, uint64_t n, rdn_ray *d_rays_out, uint32_t *d_src_index,
                                  uint64_t *d_out_n, void *cuda_stream);

/* One sample of the AO frame's accumulation (scene/rendering/gpu-ray-tracing/src/feature/ao.rs:187-232): payload(pixel) = 1 where the
 * primary ray misses (miss shader) or its AO test ray misses, 0 where the AO test ray hits (secondary closest-hit shader);
 * ao_buffer = (ao_buffer * sample_count + payload) / (sample_count + 1) while sample_count < max_sample (256 in the reference),
 * else unchanged.  d_secondary_hits / d_src_index / d_n_secondary: the hits of the compacted AO rays generated by
 * rdn_rt_gen_bounce_rays_device(mode 1) and traced with ACCEPT_FIRST_HIT_AND_END_SEARCH, and that call's d_src_index / d_out_n.
 * d_ao_buffer: n_pixels floats (the .x the reference stores in its Rgba32Float texture).  Asynchronous on cuda_stream. */
int rdn_rt_ao_accumulate_device(rdn_rt_scene *scene, int device_index, const rdn_hit *d_secondary_hits, const uint32_t *d_src_index,
                                const uint64_t *d_n_secondary, uint64_t n_pixels, uint32_t sample_count, uint32_t max_sample,
                                float *d_ao_buffer, void *cuda_stream);

/* ---- a20 / f4: the wavefront executor in ONE call.
 *      Replaces RayTracingEncoderProvider::trace_ray (shader/ray-tracing/src/api/backend.rs:48-55) as implemented by
 *      GPUWaveFrontComputeRaytracingEncoder::trace_ray (wavefront_compute/mod.rs:111-196): ray generation over the launch grid, then
 *      `execution_round_hint` rounds of { trace the wave (TraceTaskImpl::device_poll, trace_task.rs:152-360) -> pick the closest-hit or
 *      miss shader of every ray through the shader binding table (trace_task.rs:206-268) -> run each shader over its tasks -> the rays
 *      those tasks ask for, compacted in order, are the next wave (use_compact_alive_tasks, task-graph/src/runtime/task_group.rs:220-278) },
 *      all on the device: no size comes back to the host between rounds (the reference reads every task group's size back after each
 *      compaction to record its indirect dispatch).
 *      A shader stage is a callback that ENQUEUES the caller's kernels on the given stream (it must not synchronise); it sees its task
 *      list as device memory (rdn_wave).  A task that wants a ray traced in the next round writes it to d_next_rays[slot] and sets
 *      d_spawn[slot] = 1, slot = d_tasks[k] (round 0: slot = launch index) — one ray per task, the launch index is inherited.
 *      What of shader/task-graph is NOT reproduced, and what stands in for it: a task's states (task_pool.rs:85-107) collapse to
 *      membership — WAKEN = the ray is in the wave, FINISHED = its slot spawned nothing; the SLEEP / GO_TO_SLEEP states exist in the
 *      reference so that a parent task can wait for the payload of its child and copy it back (trace_task.rs:271-357) — here payloads
 *      live in caller memory indexed by launch index and stages update them in place, so nothing waits and nothing is copied back.
 *      Ray flags, cull mask, TLAS, SBT ray configuration and miss index are uniform per wave (round_launch[r], the last entry
 *      repeating) instead of per trace call. */
typedef struct rdn_wave {
  uint32_t round;                  /* 0 = ray generation; r >= 1: the stages after the r-th traversal */
  uint32_t shader;                 /* the ShaderHandle value this call runs (index into closest_hit / miss; the table's ray-gen shader) */
  const uint32_t *d_tasks;         /* this shader's tasks: indices into the traced wave, in wave order; NULL in round 0 (task k = launch index k) */
  const uint64_t *d_task_count;    /* how many, ON THE DEVICE */
  uint64_t max_tasks;              /* upper bound of *d_task_count: size grids by it */
  const rdn_ray *d_rays;           /* the wave that was traced (NULL in round 0) ... */
  const rdn_hit *d_hits;           /* ... and its hit records */
  const uint32_t *d_launch_index;  /* launch index (x + y * width) of every ray of the wave (NULL in round 0: identity) */
  uint32_t width, height;
  void *d_payload;                 /* rdn_trace_ray_desc.d_payload: caller memory, by convention indexed by launch index */
  rdn_ray *d_next_rays;            /* out: slot i = the ray task i of the wave wants traced next */
  uint8_t *d_spawn;                /* out: 1 where d_next_rays[slot] was written (zeroed by the library before the stages of a round) */
} rdn_wave;
typedef int (*rdn_stage_fn)(void *user, const rdn_wave *wave, void *cuda_stream);   /* returns 0, or a negative status that aborts the call */
typedef struct rdn_wave_counts {   /* one row per round, read back once when the call ends (rdn_trace_ray_desc.counts) */
  uint64_t wave;                   /* rays traced in this round (round 0: rays the ray generation spawned... see rdn_rt_trace_ray) */
  uint64_t closest_tasks, miss_tasks, no_task;   /* wave == closest_tasks + miss_tasks + no_task: every ray ends in exactly one list or none */
  uint64_t spawned;                /* rays the round's stages asked for = the next round's wave */
} rdn_wave_counts;
typedef struct rdn_trace_ray_desc {
  uint32_t width, height;                      /* launch size (z = 1, as the reference asserts, mod.rs:152) */
  uint32_t execution_round_hint;               /* GPURaytracingPipelineAndBindingSource::execution_round_hint: traversal rounds */
  uint32_t n_round_launch;
  const rdn_launch *round_launch;              /* the launch-uniform trace parameters of round r = round_launch[min(r - 1, n - 1)]; grid_width is ignored */
  rdn_stage_fn ray_generation; void *ray_generation_user;
  uint32_t n_closest_hit, n_miss;              /* shader handles 0 .. n - 1 of each kind (what the SBT's records name) */
  const rdn_stage_fn *closest_hit; void *const *closest_hit_user;   /* NULL entries = an empty stage */
  const rdn_stage_fn *miss; void *const *miss_user;
  void *d_payload;
  rdn_wave_counts *counts; uint32_t n_counts;  /* optional: rows 0 .. min(n_counts, rounds + 1) - 1; asking for them synchronises the stream at the end */
} rdn_trace_ray_desc;
/* `sbt`: the table whose hit groups / miss shaders pick the stages (also read by RDN_ANYHIT_FROM_SBT rounds).  Asynchronous on
 * `cuda_stream` unless desc->counts is set.  Rays whose shader handle is >= n_closest_hit / n_miss count as no_task. */
int rdn_rt_trace_ray(rdn_rt_scene *scene, int device_index, rdn_sbt *sbt, const rdn_trace_ray_desc *desc, void *cuda_stream);
/* What stages are usually made of, callable from inside a stage callback (asynchronous on the stream):
 *   spawn_all      d_spawn[k] = 1 for every task of a ray-generation stage that filled d_next_rays (e.g. with rdn_rt_gen_*_rays_device)
 *   bounce         the closest-hit -> next-ray step (rdn_rt_gen_bounce_rays_device's recipes, same `params`) for the stage's tasks,
 *                  low-discrepancy indices taken from the launch index
 *   store_f32      d_dst[launch index of the task] = value  (a payload write: "occluded" / "sky")
 *   ao_resolve     the AO ray-gen shader's accumulation over a payload buffer (feature/ao.rs:187-232); re-arms the payload to 1 */
int rdn_rt_stage_spawn_all(rdn_rt_scene *scene, int device_index, const rdn_wave *wave, void *cuda_stream);
int rdn_rt_stage_bounce(rdn_rt_scene *scene, int device_index, const rdn_bounce *params, const rdn_wave *wave, void *cuda_stream);
int rdn_rt_stage_store_f32(rdn_rt_scene *scene, int device_index, const rdn_wave *wave, float value, float *d_dst, void *cuda_stream);
int rdn_rt_ao_resolve_device(rdn_rt_scene *scene, int device_index, float *d_payload, uint64_t n_pixels, uint32_t sample_count,
                             uint32_t max_sample, float *d_ao_buffer, void *cuda_stream);


/* ---- wavefront active-list compaction: use_stream_compaction
 *      (shader/parallel-compute/src/stream_compaction.rs:3-45) as used by use_compact_alive_tasks
 *      (shader/task-graph/src/runtime/task_group.rs:220-278).  Stable; out has n slots, zero past *out_n. ---- */
int rdn_rt_compact_u32(rdn_rt_scene *scene, const uint32_t *in, const uint8_t *keep, uint64_t n,
                       uint32_t *out, uint64_t *out_n);
int rdn_rt_compact_u32_device(rdn_rt_scene *scene, int device_index, const uint32_t *d_in, const uint8_t *d_keep,
                              uint64_t n, uint32_t *d_out, uint64_t *d_out_n, void *cuda_stream);

/* ---- replication of the flattened scene (one contiguous device blob) for one-process-per-GPU drivers:
 *      rank 0 commits, broadcasts the blob (NCCL over NVLink), the other ranks adopt it. ---- */
/* device_index == -1 on a host-only scene (n_devices == 0): the blob is host memory (used to test replication without a GPU) */
int rdn_rt_scene_blob(rdn_rt_scene *scene, int device_index, void **out_device_ptr, uint64_t *out_bytes);
int rdn_rt_scene_adopt_blob(rdn_rt_scene *scene, int device_index, const void *d_blob, uint64_t bytes);

/* ---- flattened reference-layout arrays, copied to host, for cross-checking the flattener ---- */
typedef enum rdn_array_id {
  RDN_ARRAY_TLAS_BINDING = 0, RDN_ARRAY_TLAS_BVH_ROOT, RDN_ARRAY_TLAS_BVH_FOREST, RDN_ARRAY_TLAS_BOUNDING,
  RDN_ARRAY_INSTANCES, RDN_ARRAY_BLAS_META, RDN_ARRAY_GEOMETRY_META, RDN_ARRAY_TRI_BVH_FOREST,
  RDN_ARRAY_TRIANGLES, RDN_ARRAY_SLOT_INFO, RDN_ARRAY_WIDE_NODES, RDN_ARRAY_PRIM_TO_SLOT, RDN_ARRAY_IRREGULAR_INSTANCES,
  RDN_ARRAY_IRREGULAR_LEAF_BOXES, RDN_ARRAY_WIDE4_NODES, RDN_ARRAY_COUNT
} rdn_array_id;
int rdn_rt_scene_array(rdn_rt_scene *scene, int array_id, void *out, uint64_t capacity_bytes, uint64_t *out_bytes);
/* build statistics of the committed scene (commits first); a scene that adopted another rank's blob reports only what the
 * blob itself records (irregular_instances listed per TLAS, reference_routed_tlas) */
int rdn_rt_scene_build_stats(rdn_rt_scene *scene, rdn_build_stats *out);

/* ---- space-query surface: FlattenBVH::new (content/space/src/bvh/mod.rs:55-79) and
 *      intersect_nearest_bvh (content/mesh/core/src/feature/bvh.rs:57-86) ---- */
typedef struct rdn_flat_bvh rdn_flat_bvh;
typedef struct rdn_tree_build_option { uint64_t max_tree_depth, bin_size; } rdn_tree_build_option; /* utils.rs:20-32 */
typedef enum rdn_bvh_strategy { RDN_BVH_SAH = 0, RDN_BVH_BALANCE_TREE = 1 } rdn_bvh_strategy;
typedef struct rdn_flat_bvh_node {          /* FlattenBVHNode, content/space/src/bvh/node.rs:5-27 */
  float bounding_min[3], bounding_max[3];
  uint64_t primitive_start, primitive_end, self_index, left_count;
  int32_t has_child, split_axis;
} rdn_flat_bvh_node;
typedef enum rdn_face_side { RDN_FACE_FRONT = 0, RDN_FACE_BACK = 1, RDN_FACE_DOUBLE = 2 } rdn_face_side;
typedef struct rdn_mesh_view { const float *positions; uint64_t n_positions; const uint32_t *indices; uint64_t n_indices; } rdn_mesh_view;
/* MeshBufferHitPoint (content/mesh/core/src/feature/intersection.rs:42-45): hit == 0 => OptionalNearest::none */
typedef struct rdn_mesh_hit { float px, py, pz, distance; uint32_t primitive_index, hit, pad0, pad1; } rdn_mesh_hit;

int rdn_bvh_build(const float *boxes_min_max6, uint64_t n, int strategy, uint32_t sah_buckets,
                  const rdn_tree_build_option *option, rdn_flat_bvh **out);
/* the same tree, built on CUDA device `device` (SAH with up to 4 buckets: level-synchronous passes, see csrc/build_device.cu);
 * what the device build does not cover (more buckets, a degenerate range of more than 64 primitives) is built by the host
 * builder instead — rdn_bvh_built_on_device tells which it was */
int rdn_bvh_build_device(const float *boxes_min_max6, uint64_t n, uint32_t sah_buckets, const rdn_tree_build_option *option,
                         int device, rdn_flat_bvh **out);
int rdn_bvh_built_on_device(const rdn_flat_bvh *bvh);
void rdn_bvh_destroy(rdn_flat_bvh *bvh);
int rdn_bvh_nodes(const rdn_flat_bvh *bvh, const rdn_flat_bvh_node **out_nodes, uint64_t *out_n);
int rdn_bvh_sorted_primitive_index(const rdn_flat_bvh *bvh, const uint64_t **out_index, uint64_t *out_n);
/* build_bvh_for_abstract_mesh over an indexed triangle list (feature/bvh.rs:5-21) */
int rdn_bvh_build_for_mesh(const rdn_mesh_view *mesh, int strategy, uint32_t sah_buckets,
                           const rdn_tree_build_option *option, rdn_flat_bvh **out);
/* intersect_nearest_bvh for a batch of rays on CUDA device `device` (host buffers) */
int rdn_bvh_query_nearest(const rdn_flat_bvh *bvh, const rdn_mesh_view *mesh, const rdn_ray *rays, uint64_t n,
                          uint32_t face_side, int device, rdn_mesh_hit *out);

/* intersect_list_bvh (content/mesh/core/src/feature/bvh.rs:23-55) for a batch of rays: EVERY intersected primitive, as a CSR
 * list — the hits of ray i are out_hits[out_offsets[i] .. out_offsets[i+1]) in the reference's Vec order (right-first DFS, leaf
 * primitives in sorted_primitive_index order).  out_offsets has n + 1 slots; *out_total = number of hits.  Two-call protocol:
 * with out_hits == NULL (or capacity < total) only offsets and total are produced. */
int rdn_bvh_query_list(const rdn_flat_bvh *bvh, const rdn_mesh_view *mesh, const rdn_ray *rays, uint64_t n, uint32_t face_side,
                       int device, uint64_t *out_offsets, rdn_mesh_hit *out_hits, uint64_t capacity, uint64_t *out_total);
/* device-resident form: gather the mesh through the BVH's sorted_primitive_index and upload it with the nodes once, then
 * query any number of device ray batches (asynchronous on cuda_stream) */
int rdn_bvh_upload(rdn_flat_bvh *bvh, const rdn_mesh_view *mesh, int device);
int rdn_bvh_query_nearest_device(const rdn_flat_bvh *bvh, const rdn_ray *d_rays, uint64_t n, uint32_t face_side,
                                 rdn_mesh_hit *d_out, void *cuda_stream);

/* ---- mesh picking: AbstractMeshIntersectionExt::ray_intersect_nearest / ray_intersect_all
 *      (content/mesh/core/src/feature/intersection.rs:3-37) over an attribute mesh of any MeshPrimitiveTopology
 *      (content/mesh/core/src/primitive.rs:108-148; primitive k reads `stride` consecutive (indexed) vertices from step*k,
 *      container/attributes/access.rs:142-150,199-243) with MeshBufferIntersectConfig (container/attributes/picking.rs:4-27):
 *      points and line segments hit within `tolerance_local` (math/geometry/src/dimension3/intersection.rs:79-121, Ray3::
 *      distance_sq_to_segment ray3.rs:48-145), triangles by the GTE test with `triangle_face`.  The reference tests every
 *      primitive on one CPU thread per pick (scene/geometry-query/src/model.rs:245-255); here the primitives of a ray are spread
 *      over the GPU.  Ray directions are unit vectors (Ray3 holds a NormalizedVector); tmin / tmax of rdn_ray are ignored. ---- */
typedef enum rdn_topology {
  RDN_TOPOLOGY_POINT_LIST = 0, RDN_TOPOLOGY_LINE_LIST = 1, RDN_TOPOLOGY_LINE_STRIP = 2, RDN_TOPOLOGY_TRIANGLE_LIST = 3,
  RDN_TOPOLOGY_TRIANGLE_STRIP = 4
} rdn_topology;
typedef struct rdn_pick_config { float tolerance_local; uint32_t triangle_face; } rdn_pick_config;   /* MeshBufferIntersectConfig */
typedef struct rdn_pick_mesh rdn_pick_mesh;   /* a mesh resident on one CUDA device */
/* copies positions (and indices, when mesh->indices != NULL) to `device`; vertex indices are checked here (the reference's
 * index_get returns None and the primitive is skipped; an out-of-range index is an error instead) */
int rdn_pick_mesh_create(const rdn_mesh_view *mesh, uint32_t topology, int device, rdn_pick_mesh **out);
void rdn_pick_mesh_destroy(rdn_pick_mesh *mesh);
int rdn_pick_mesh_primitive_count(const rdn_pick_mesh *mesh, uint64_t *out_count);
/* ray_intersect_nearest for a batch of pick rays (host buffers): strict `<` refresh in primitive order, i.e. the smallest
 * primitive index among equally near hits */
int rdn_pick_mesh_nearest(rdn_pick_mesh *mesh, const rdn_pick_config *config, const rdn_ray *rays, uint64_t n, rdn_mesh_hit *out);
/* ray_intersect_all for one ray: every hit in primitive order; *out_total = number of hits, min(total, capacity) written */
int rdn_pick_mesh_all(rdn_pick_mesh *mesh, const rdn_pick_config *config, const rdn_ray *ray, rdn_mesh_hit *out, uint64_t capacity,
                      uint64_t *out_total);

/* ---- measurement hook (no reference counterpart): read bandwidth, in GB/s, of a buffer of `bytes` that has been made L2
 *      resident on the scene's device (uint4 loads that bypass L1, `passes` sweeps inside one kernel, CUDA events) — the L2
 *      denominator of the roofline next to the HBM copy peak.  Synchronous; allocates and frees the buffer. */
int rdn_rt_measure_l2_read_gbs(rdn_rt_scene *scene, int device_index, uint64_t bytes, int passes, double *out_gbs);

const char *rdn_rt_last_error(void);
const char *rdn_rt_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RDN_RT_H */
