/* ORACLE — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of rendiation's BVH closest-hit path, used as the
 * parity checker (tests/, __graft_entry__.smoke()) and as the timed CPU baseline
 * (bench.py cpu_baseline / --impl reference).  Nothing under rendiation_b200/
 * may include, link or call it.
 *
 * PARITY PINNING: the reference holds NO golden vectors for ray/box, ray/triangle,
 * intersect_nearest_bvh or NaiveSahBvhCpu::traverse (its tests only dump .pbm
 * images: shader/ray-tracing/src/backend/wavefront_compute/geometry/naive/test.rs:297),
 * and the reference cannot be compiled here (no rustc/cargo).  Traversal parity is
 * therefore "parity unpinned" upstream; what IS pinned is checked in tests/:
 *   - stream compaction / prefix scan / shuffle known answers
 *     (shader/parallel-compute/src/stream_compaction.rs:100-125, prefix_scan.rs:122-172,
 *      shuffle_move.rs:119-133)
 *   - Mat4 translate*scale*point (math/algebra/src/mat/mat4.rs:204-219)
 *   - tessellation counts (content/mesh/generator/src/builder/mod.rs:128-146)
 * plus self-consistency (brute force vs path A vs path B) and geometric invariants.
 */
#ifndef RDN_ORACLE_H
#define RDN_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#include "oracle_math.h"

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_INVALID_NEXT 0xFFFFFFFFu

/* ---------- records shared with the product ABI (same byte layout as include/rdn_rt.h) ---------- */
typedef struct { float ox, oy, oz, tmin, dx, dy, dz, tmax; } orc_ray;                 /* 32 B */
typedef struct { uint32_t ray_flags, cull_mask, tlas_idx, grid_width; } orc_launch;
typedef struct { float t, u, v; uint32_t primitive_id, geometry_id, instance_id, instance_custom_id, hit_kind; } orc_hit; /* 32 B */
typedef struct { float px, py, pz, distance; uint32_t primitive_index, hit; uint32_t pad0, pad1; } orc_mesh_hit;        /* path A */

/* any-hit shaders as data (oracle_scene.c "any-hit"): kind 0 constant, 1 (primitive_idx & mask) == value, 2 distance >= distance;
 * `behavior` when the predicate holds, `otherwise` when not (ANYHIT_BEHAVIOR_ACCEPT_HIT = 1, _END_SEARCH = 2, api/ty.rs:134-136) */
typedef struct { uint32_t kind, behavior, otherwise, mask, value; float distance; uint32_t pad0, pad1; } orc_anyhit_program;   /* 32 B */
typedef struct {
  uint32_t mode;              /* 0: every candidate accepted, 1: uniform_program for all non-opaque geometry, 2: through the hit groups */
  uint32_t uniform_program;
  const orc_anyhit_program *programs; uint32_t n_programs;
  const uint32_t *hit_group_any; uint32_t n_hit_groups;   /* any_hit handle of each SBT hit group (0xFFFFFFFF: none) */
  uint32_t sbt_ray_offset, sbt_ray_stride;
} orc_anyhit_setup;

/* the reference's four traversal counters (traverse_cpu.rs:37-41) + instances entered + would-abort events */
typedef struct { uint64_t bvh_visit, bvh_hit, tri_visit, tri_hit, inst_visit, ref_abort; } orc_counters;

/* ---------- FlattenBVH (content/space/src/bvh/{mod,node,strategy,apply}.rs, utils.rs) ---------- */
typedef struct {
  obox bounding;
  uint64_t start, end;      /* primitive_range into sorted_primitive_index */
  uint64_t self_index;
  uint64_t left_count;      /* valid iff has_child */
  int32_t has_child;
  int32_t split_axis;
} orc_bvh_node;

typedef struct {
  orc_bvh_node *nodes;
  uint64_t n_nodes, cap_nodes;
  uint64_t *sorted_primitive_index;
  uint64_t n_prims;
  uint64_t balance_fallbacks;        /* SAH degenerate -> BalanceTree fallbacks taken */
  uint64_t balance_fallbacks_gt10;   /* ... on ranges > 10 primitives (Rust select_nth_unstable order not reproducible) */
  int32_t error;                     /* nonzero: the reference would have panicked (bucket index out of range) */
} orc_bvh;

enum { ORC_STRATEGY_SAH = 0, ORC_STRATEGY_BALANCE = 1 };

orc_bvh *orc_bvh_build(const obox *boxes, uint64_t n, int strategy, uint32_t sah_buckets,
                       uint64_t max_tree_depth, uint64_t bin_size);
void orc_bvh_free(orc_bvh *b);
/* compute_bvh_next (naive/mod.rs:612-632): out[2*i] = hit, out[2*i+1] = miss */
void orc_bvh_compute_next(const orc_bvh *b, uint32_t *out_hit_miss);

/* ---------- path A: content/space BVH ray query through mesh-core ---------- */
enum { ORC_FACE_FRONT = 0, ORC_FACE_BACK = 1, ORC_FACE_DOUBLE = 2 };
int orc_ray_box_a(const float *ray_od6, const obox *box);
int orc_ray_triangle_a(const float *ray_od6, ov3 a, ov3 b, ov3 c, int face_side, float *out_pos3_dist);
/* intersect_nearest_bvh over an indexed triangle list */
void orc_patha_query_nearest(const orc_bvh *bvh, const float *positions, const uint32_t *indices,
                             const orc_ray *rays, uint64_t n_rays, int face_side, orc_mesh_hit *out,
                             int n_threads);
/* intersect_list_bvh (feature/bvh.rs:23-55): CSR list of every intersection per ray, in visiting order; returns the total */
uint64_t orc_patha_query_list(const orc_bvh *bvh, const float *positions, const uint32_t *indices, const orc_ray *rays,
                              uint64_t n_rays, int face_side, uint64_t *offsets, orc_mesh_hit *out, uint64_t capacity);
/* brute force ray_intersect_nearest (content/mesh/core/src/feature/intersection.rs:11-37) */
void orc_brute_query_nearest(const float *positions, const uint32_t *indices, uint64_t n_tris,
                             const orc_ray *rays, uint64_t n_rays, int face_side, orc_mesh_hit *out,
                             int n_threads);

/* ---------- path B: naive software TLAS/BLAS (shader/ray-tracing .../geometry/naive) ---------- */
typedef struct { ov3 aabb_min; uint32_t hit_next; ov3 aabb_max; uint32_t miss_next; uint32_t range_x, range_y, tail0, tail1; } orc_dev_node; /* 48 B */
typedef struct { om4 transform, transform_inv; uint32_t instance_custom_index, sbt_offset, flags, blas; } orc_dev_instance;          /* 144 B */
typedef struct { ov3 world_min; uint32_t mask; ov3 world_max; uint32_t flags; } orc_tlas_bounding;                                 /* 32 B */
typedef struct { uint32_t bvh_root_idx, geometry_idx, primitive_start, geometry_flags; } orc_geom_meta;                            /* 16 B */
typedef struct { uint32_t tri_root_x, tri_root_y; } orc_blas_meta;                                                                 /* 8 B */

typedef struct {
  float transform[16];           /* column-major a1..d4 */
  uint32_t instance_custom_index, mask, sbt_offset, flags, blas_handle;
} orc_instance_src;              /* TopLevelAccelerationStructureSourceInstance, api/backend.rs:160-167 */

typedef struct orc_scene orc_scene;
orc_scene *orc_scene_new(void);
void orc_scene_free(orc_scene *s);
/* one BLAS = n_geoms geometries; geometry g: positions[g] (3*n_pos floats), indices[g] (may be NULL), flags[g];
 * is_aabb[g] != 0 marks an AABB geometry (accepted, ignored by the naive builder, naive/mod.rs:201-237) */
uint32_t orc_scene_create_blas(orc_scene *s, uint32_t n_geoms, const float *const *positions, const uint64_t *n_pos,
                               const uint32_t *const *indices, const uint64_t *n_idx, const uint32_t *flags,
                               const uint8_t *is_aabb);
void orc_scene_delete_blas(orc_scene *s, uint32_t handle);
uint32_t orc_scene_create_tlas(orc_scene *s, const orc_instance_src *inst, uint32_t n);
void orc_scene_delete_tlas(orc_scene *s, uint32_t handle);
void orc_scene_bind_tlas(orc_scene *s, const uint32_t *handles, uint32_t n);
/* NaiveSahBvhSource::build; returns 0 ok, <0 where the reference would panic */
int orc_scene_build(orc_scene *s);

/* flattened arrays (valid after build) for cross-checking the product's flattener */
typedef struct {
  const uint32_t *tlas_binding; uint64_t n_tlas_binding;
  const uint32_t *tlas_bvh_root; uint64_t n_tlas_bvh_root;
  const orc_dev_node *tlas_bvh_forest; uint64_t n_tlas_bvh_forest;
  const orc_dev_instance *tlas_data; uint64_t n_tlas_data;
  const orc_tlas_bounding *tlas_bounding; uint64_t n_tlas_bounding;
  const orc_blas_meta *blas_meta_info; uint64_t n_blas_meta_info;
  const orc_geom_meta *tri_bvh_root; uint64_t n_tri_bvh_root;
  const orc_dev_node *tri_bvh_forest; uint64_t n_tri_bvh_forest;
  const uint32_t *indices_redirect; uint64_t n_indices_redirect;
  const uint32_t *indices; uint64_t n_indices;
  const float *vertices; uint64_t n_vertices;   /* vertex count (3 floats each) */
  uint64_t balance_fallbacks, balance_fallbacks_gt10;
} orc_scene_view;
void orc_scene_get_view(const orc_scene *s, orc_scene_view *out);

/* NaiveSahBvhCpu::traverse for a batch (any_hit == always ACCEPT, as TEST_ANYHIT_BEHAVIOR, naive/test.rs:7);
 * counters may be NULL; n_threads <= 1 runs on the calling thread */
/* the any-hit setup used by the traces that follow (NULL: every candidate accepted); copied */
int orc_scene_set_any_hit(orc_scene *s, const orc_anyhit_setup *setup);
int orc_scene_trace(const orc_scene *s, const orc_launch *launch, const orc_ray *rays, uint64_t n_rays,
                    orc_hit *out_hits, orc_counters *counters, int n_threads);

/* test tool, NOT the reference: same trees, no shrinking range, closest candidate kept (order-free model of a pruning traversal) */
int orc_scene_trace_unpruned(const orc_scene *s, const orc_launch *launch, const orc_ray *rays, uint64_t n_rays,
                             orc_hit *out_hits, int n_threads);

/* test tool, NOT the reference: scalar model of the ordered traversal + tie re-walk of csrc/traverse.cu (algorithm-level fuzzing) */
int orc_scene_trace_ordered_model(const orc_scene *s, const orc_launch *launch, const orc_ray *rays, uint64_t n_rays, orc_hit *out_hits,
                                  uint64_t *out_stats);

/* every (instance, triangle) pair of the bound TLAS whose triangle test passes with the ray's ORIGINAL range: no BVH, no pruning */
typedef struct { float distance, t_object, u, v, sign, scaling; uint32_t instance_id, geometry_id, primitive_id, slot, in_range, pad; } orc_candidate;
uint64_t orc_scene_candidates(const orc_scene *s, const orc_launch *launch, const orc_ray *ray, orc_candidate *out, uint64_t cap);

/* ---------- brute-force mesh picking (oracle_pick.c): points / lines with tolerance, triangles ---------- */
enum { ORC_TOPOLOGY_POINT_LIST = 0, ORC_TOPOLOGY_LINE_LIST = 1, ORC_TOPOLOGY_LINE_STRIP = 2, ORC_TOPOLOGY_TRIANGLE_LIST = 3, ORC_TOPOLOGY_TRIANGLE_STRIP = 4 };
int orc_ray_segment(const float *ray_od6, ov3 v0, ov3 v1, float tolerance, float *out_pos3_dist);
int orc_ray_point(const float *ray_od6, ov3 point, float tolerance, float *out_pos3_dist);
uint64_t orc_pick_primitive_count(uint64_t n_positions, uint64_t n_indices, int has_indices, int topology);
void orc_pick_nearest(const float *positions, uint64_t n_positions, const uint32_t *indices, uint64_t n_indices, int topology,
                      float tolerance, int face_side, const orc_ray *rays, uint64_t n_rays, orc_mesh_hit *out, int n_threads);
uint64_t orc_pick_all(const float *positions, uint64_t n_positions, const uint32_t *indices, uint64_t n_indices, int topology,
                      float tolerance, int face_side, const orc_ray *ray, orc_mesh_hit *out, uint64_t capacity);

/* ---------- parallel-compute restatements (scan / compaction / scatter) ---------- */
void orc_workgroup_inclusive_scan_u32(const uint32_t *in, uint64_t n, uint32_t workgroup, uint32_t *out);
void orc_inclusive_scan_u32(const uint32_t *in, uint64_t n, uint32_t *out);
/* use_stream_compaction: out has n slots, zero-filled past the returned size */
uint64_t orc_stream_compaction_u32(const uint32_t *in, const uint8_t *keep, uint64_t n, uint32_t *out);
/* shuffle_move: out[target[i]] = in[i] where moved[i] != 0 */
void orc_shuffle_move_u32(const uint32_t *in, const uint32_t *target, const uint8_t *moved, uint64_t n, uint32_t *out);

/* Mat4 helpers exported for the Mat4 KAT and scene builders: out = column-major 16 floats */
void orc_mat4_compose(const float *a16, const float *b16, float *out16);
void orc_mat4_inverse_or_identity(const float *m16, float *out16);
void orc_mat4_mul_vec4(const float *m16, const float *v4, float *out4);

#ifdef __cplusplus
}
#endif
#endif
