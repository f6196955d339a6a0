/* ORACLE — TEST INFRASTRUCTURE ONLY.  BLAS/TLAS assembly + CPU traversal (path B).
 *
 * Follows /root/reference/shader/ray-tracing/src/backend/wavefront_compute/geometry:
 *   naive/mod.rs:98-119    create/delete blas/tlas (handles = push index, never recycled)
 *   naive/mod.rs:122-260   build_blas  (SAH::new(4), depth 50, bin 2; blas_box pushed PER GEOMETRY — kept as is)
 *   naive/mod.rs:262-320   build_tlas  (SAH::new(4), depth 50, bin 10; FLIP_FACING xor when det(mat3) < 0)
 *   naive/mod.rs:322-493   build (flatten_bvh_to_gpu_node, forest offsets)
 *   naive/traverse_cpu.rs:52-319  NaiveSahBvhCpu::traverse, RayRange, TraverseBvhIteratorCpu
 *   naive/flag.rs:6-117    TraverseFlags
 *   mod.rs:44-64           intersect_ray_aabb_cpu
 *   mod.rs:105-155         intersect_ray_triangle_cpu
 *
 * Deliberate, documented deviations (the reference has no defined result there):
 *   - RayRange::update_far asserts (traverse_cpu.rs:272-276) abort the reference process; here the
 *     candidate is rejected and counted in counters.ref_abort (NaN distance from zero-area triangles,
 *     or distance rounding outside [near, far]).
 *   - out-of-range handles / deleted BLAS referenced by an instance panic in the reference
 *     (naive/mod.rs:273-275 unwrap); orc_scene_build returns a negative code instead.
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include "oracle.h"

/* ---- flag constants (api/ty.rs:102-159, naive/flag.rs:6-25) ---- */
#define F_FORCE_OPAQUE 0x01u
#define F_FORCE_NON_OPAQUE 0x02u
#define F_ACCEPT_FIRST_HIT_AND_END_SEARCH 0x04u
#define F_CULL_BACK_FACING_TRIANGLES 0x10u
#define F_CULL_FRONT_FACING_TRIANGLES 0x20u
#define F_CULL_OPAQUE 0x40u
#define F_CULL_NON_OPAQUE 0x80u
#define F_SKIP_TRIANGLES 0x100u
#define F_TRIANGLE_FLIP_FACING 0x400u
#define GI_TRIANGLE_FACING_CULL_DISABLE 0x1u
#define GI_TRIANGLE_FLIP_FACING 0x2u
#define GI_FORCE_OPAQUE 0x4u
#define GI_FORCE_NO_OPAQUE 0x8u
#define G_OPAQUE 0x1u
#define HIT_KIND_FRONT 0xFEu
#define HIT_KIND_BACK 0xFFu

typedef struct { float *positions; uint64_t n_pos; uint32_t *indices; uint64_t n_idx; int has_indices; uint32_t flags; uint8_t is_aabb; } geom_src;
typedef struct { geom_src *geoms; uint32_t n_geoms; int alive; } blas_src;
typedef struct { orc_instance_src *inst; uint32_t n; int alive; } tlas_src;

#define VEC(T) struct { T *p; uint64_t n, cap; }
#define VEC_PUSH(v, val) do { if ((v).n == (v).cap) { (v).cap = (v).cap ? (v).cap * 2 : 16; (v).p = realloc((v).p, (v).cap * sizeof(*(v).p)); } (v).p[(v).n++] = (val); } while (0)
#define VEC_FREE(v) do { free((v).p); (v).p = NULL; (v).n = (v).cap = 0; } while (0)

struct orc_scene {
  VEC(blas_src) blas;
  VEC(tlas_src) tlas;
  VEC(uint32_t) binding;
  /* NaiveSahBvhCpu */
  VEC(uint32_t) tlas_bvh_root;
  VEC(orc_dev_node) tlas_bvh_forest;
  VEC(orc_dev_instance) tlas_data;
  VEC(orc_tlas_bounding) tlas_bounding;
  VEC(orc_blas_meta) blas_meta_info;
  VEC(orc_geom_meta) tri_bvh_root;
  VEC(orc_dev_node) tri_bvh_forest;
  VEC(uint32_t) indices_redirect;
  VEC(uint32_t) indices;
  VEC(ov3) vertices;
  uint64_t balance_fallbacks, balance_fallbacks_gt10;
  int built;
  /* any-hit shaders of the bound pipeline, as data (orc_scene_set_any_hit) */
  orc_anyhit_setup anyhit;
  orc_anyhit_program *anyhit_programs;
  uint32_t *anyhit_groups;
};

orc_scene *orc_scene_new(void) { return (orc_scene *)calloc(1, sizeof(orc_scene)); }

static void free_built(orc_scene *s) {
  VEC_FREE(s->tlas_bvh_root); VEC_FREE(s->tlas_bvh_forest); VEC_FREE(s->tlas_data); VEC_FREE(s->tlas_bounding);
  VEC_FREE(s->blas_meta_info); VEC_FREE(s->tri_bvh_root); VEC_FREE(s->tri_bvh_forest);
  VEC_FREE(s->indices_redirect); VEC_FREE(s->indices); VEC_FREE(s->vertices);
  s->built = 0;
}
static void free_blas(blas_src *b) {
  for (uint32_t g = 0; g < b->n_geoms; g++) { free(b->geoms[g].positions); free(b->geoms[g].indices); }
  free(b->geoms); b->geoms = NULL; b->n_geoms = 0; b->alive = 0;
}
void orc_scene_free(orc_scene *s) {
  if (!s) return;
  free_built(s);
  for (uint64_t i = 0; i < s->blas.n; i++) free_blas(&s->blas.p[i]);
  for (uint64_t i = 0; i < s->tlas.n; i++) free(s->tlas.p[i].inst);
  VEC_FREE(s->blas); VEC_FREE(s->tlas); VEC_FREE(s->binding);
  free(s);
}

uint32_t orc_scene_create_blas(orc_scene *s, uint32_t n_geoms, const float *const *positions, const uint64_t *n_pos,
                               const uint32_t *const *indices, const uint64_t *n_idx, const uint32_t *flags,
                               const uint8_t *is_aabb) {
  blas_src b;
  b.n_geoms = n_geoms; b.alive = 1;
  b.geoms = (geom_src *)calloc(n_geoms ? n_geoms : 1, sizeof(geom_src));
  for (uint32_t g = 0; g < n_geoms; g++) {
    geom_src *q = &b.geoms[g];
    q->n_pos = n_pos[g];
    q->positions = (float *)malloc((q->n_pos ? q->n_pos : 1) * 3 * sizeof(float));
    memcpy(q->positions, positions[g], q->n_pos * 3 * sizeof(float));
    q->has_indices = indices && indices[g] != NULL;
    q->n_idx = q->has_indices ? n_idx[g] : 0;
    q->indices = NULL;
    if (q->has_indices) {
      q->indices = (uint32_t *)malloc((q->n_idx ? q->n_idx : 1) * sizeof(uint32_t));
      memcpy(q->indices, indices[g], q->n_idx * sizeof(uint32_t));
    }
    q->flags = flags[g];
    q->is_aabb = is_aabb ? is_aabb[g] : 0;
  }
  s->built = 0;
  VEC_PUSH(s->blas, b);
  return (uint32_t)(s->blas.n - 1);
}
void orc_scene_delete_blas(orc_scene *s, uint32_t h) { if (h < s->blas.n) { free_blas(&s->blas.p[h]); s->built = 0; } }
uint32_t orc_scene_create_tlas(orc_scene *s, const orc_instance_src *inst, uint32_t n) {
  tlas_src t; t.n = n; t.alive = 1;
  t.inst = (orc_instance_src *)malloc((n ? n : 1) * sizeof(orc_instance_src));
  memcpy(t.inst, inst, n * sizeof(orc_instance_src));
  s->built = 0;
  VEC_PUSH(s->tlas, t);
  return (uint32_t)(s->tlas.n - 1);
}
void orc_scene_delete_tlas(orc_scene *s, uint32_t h) {
  if (h < s->tlas.n) { free(s->tlas.p[h].inst); s->tlas.p[h].inst = NULL; s->tlas.p[h].n = 0; s->tlas.p[h].alive = 0; s->built = 0; }
}
void orc_scene_bind_tlas(orc_scene *s, const uint32_t *handles, uint32_t n) {
  s->binding.n = 0;
  for (uint32_t i = 0; i < n; i++) VEC_PUSH(s->binding, handles[i]);
  s->built = 0;
}

static orc_dev_node flatten_node(const orc_bvh_node *n, uint32_t hit, uint32_t miss, uint32_t next_offset, uint32_t primitive_offset) {
  orc_dev_node d;
  memset(&d, 0, sizeof(d));
  d.aabb_min = n->bounding.min;
  d.aabb_max = n->bounding.max;
  d.hit_next = hit != ORC_INVALID_NEXT ? hit + next_offset : ORC_INVALID_NEXT;
  d.miss_next = miss != ORC_INVALID_NEXT ? miss + next_offset : ORC_INVALID_NEXT;
  d.range_x = (uint32_t)n->start + primitive_offset;
  d.range_y = (uint32_t)n->end + primitive_offset;
  return d;
}

typedef struct { int some; obox box; } opt_box;

int orc_scene_build(orc_scene *s) {
  free_built(s);
  s->balance_fallbacks = s->balance_fallbacks_gt10 = 0;
  VEC(opt_box) blas_box; memset(&blas_box, 0, sizeof(blas_box));
  int rc = 0;

  /* ---- build_blas + the tri_bvh_forest part of build ---- */
  for (uint64_t bi = 0; bi < s->blas.n; bi++) {
    blas_src *blas = &s->blas.p[bi];
    if (!blas->alive) {
      orc_blas_meta z = {0, 0};
      VEC_PUSH(s->blas_meta_info, z);
      opt_box none; memset(&none, 0, sizeof(none));
      VEC_PUSH(blas_box, none);
      continue;
    }
    uint32_t tri_start = (uint32_t)s->tri_bvh_root.n;
    for (uint32_t g = 0; g < blas->n_geoms; g++) {
      geom_src *src = &blas->geoms[g];
      obox root_box = obox_empty();
      if (!src->is_aabb) {
        uint32_t primitive_start = (uint32_t)(s->indices.n / 3);
        uint32_t vertex_start = (uint32_t)s->vertices.n;
        uint64_t n_idx = src->has_indices ? src->n_idx : src->n_pos;
        uint64_t n_tri = n_idx / 3;
        for (uint64_t v = 0; v < src->n_pos; v++)
          VEC_PUSH(s->vertices, ov3_new(src->positions[3 * v], src->positions[3 * v + 1], src->positions[3 * v + 2]));
        obox *boxes = (obox *)malloc((n_tri ? n_tri : 1) * sizeof(obox));
        for (uint64_t t = 0; t < n_tri; t++) {
          obox b = obox_empty();
          for (int k = 0; k < 3; k++) {
            uint64_t idx = src->has_indices ? src->indices[3 * t + k] : (3 * t + k);
            if (idx >= src->n_pos) { rc = -2; idx = 0; } /* reference: index out of bounds panic */
            obox_expand_point(&b, ov3_new(src->positions[3 * idx], src->positions[3 * idx + 1], src->positions[3 * idx + 2]));
          }
          boxes[t] = b;
        }
        orc_bvh *bvh = orc_bvh_build(boxes, n_tri, ORC_STRATEGY_SAH, 4, 50, 2);
        free(boxes);
        if (bvh->error) rc = -3;
        s->balance_fallbacks += bvh->balance_fallbacks;
        s->balance_fallbacks_gt10 += bvh->balance_fallbacks_gt10;
        obox_expand_box(&root_box, bvh->nodes[0].bounding);
        uint32_t *next = (uint32_t *)malloc(bvh->n_nodes * 2 * sizeof(uint32_t));
        orc_bvh_compute_next(bvh, next);
        uint32_t raw_primitive_start = (uint32_t)(s->indices.n / 3);
        for (uint64_t i = 0; i < bvh->n_prims; i++)
          VEC_PUSH(s->indices_redirect, raw_primitive_start + (uint32_t)bvh->sorted_primitive_index[i]);
        for (uint64_t t = 0; t < n_tri; t++)
          for (int k = 0; k < 3; k++) {
            uint32_t idx = src->has_indices ? src->indices[3 * t + k] : (uint32_t)(3 * t + k);
            VEC_PUSH(s->indices, vertex_start + idx);
          }
        /* forest append (naive/mod.rs:368-384) */
        uint32_t bvh_start = (uint32_t)s->tri_bvh_forest.n;
        orc_geom_meta gm = {bvh_start, g, primitive_start, src->flags};
        VEC_PUSH(s->tri_bvh_root, gm);
        for (uint64_t i = 0; i < bvh->n_nodes; i++)
          VEC_PUSH(s->tri_bvh_forest, flatten_node(&bvh->nodes[i], next[2 * i], next[2 * i + 1], bvh_start, primitive_start));
        free(next);
        orc_bvh_free(bvh);
      }
      opt_box ob; ob.some = 1; ob.box = root_box;
      VEC_PUSH(blas_box, ob); /* per geometry, exactly as naive/mod.rs:239 */
    }
    orc_blas_meta m = {tri_start, (uint32_t)s->tri_bvh_root.n};
    VEC_PUSH(s->blas_meta_info, m);
  }

  /* ---- build_tlas per TLAS ---- */
  for (uint64_t ti = 0; ti < s->tlas.n; ti++) {
    tlas_src *tlas = &s->tlas.p[ti];
    if (!tlas->alive) { VEC_PUSH(s->tlas_bvh_root, ORC_INVALID_NEXT); continue; }
    uint32_t bvh_start = (uint32_t)s->tlas_bvh_forest.n;
    uint32_t primitive_start = (uint32_t)s->tlas_data.n;
    obox *aabbs = (obox *)malloc((tlas->n ? tlas->n : 1) * sizeof(obox));
    for (uint32_t i = 0; i < tlas->n; i++) {
      const orc_instance_src *src = &tlas->inst[i];
      om4 m; memcpy(&m, src->transform, sizeof(m));
      if (src->blas_handle >= blas_box.n || !blas_box.p[src->blas_handle].some) {
        rc = -4; aabbs[i] = obox_empty(); continue; /* reference: unwrap() on None / OOB panic */
      }
      aabbs[i] = obox_apply_matrix(blas_box.p[src->blas_handle].box, m);
    }
    orc_bvh *bvh = orc_bvh_build(aabbs, tlas->n, ORC_STRATEGY_SAH, 4, 50, 10);
    if (bvh->error) rc = -3;
    s->balance_fallbacks += bvh->balance_fallbacks;
    s->balance_fallbacks_gt10 += bvh->balance_fallbacks_gt10;
    uint32_t *next = (uint32_t *)malloc(bvh->n_nodes * 2 * sizeof(uint32_t));
    orc_bvh_compute_next(bvh, next);
    for (uint64_t k = 0; k < bvh->n_prims; k++) {
      uint64_t box_idx = bvh->sorted_primitive_index[k];
      const orc_instance_src *src = &tlas->inst[box_idx];
      om4 m; memcpy(&m, src->transform, sizeof(m));
      uint32_t flags = src->flags;
      if (om3_det(om4_to_mat3(m)) < 0.0f) flags ^= GI_TRIANGLE_FLIP_FACING;
      orc_dev_instance di;
      memset(&di, 0, sizeof(di));
      di.transform = m;
      di.transform_inv = om4_inverse_or_identity(m);
      di.instance_custom_index = src->instance_custom_index;
      di.sbt_offset = src->sbt_offset;
      di.flags = flags;
      di.blas = src->blas_handle;
      orc_tlas_bounding tb;
      tb.world_min = aabbs[box_idx].min; tb.world_max = aabbs[box_idx].max;
      tb.mask = src->mask; tb.flags = flags;
      VEC_PUSH(s->tlas_data, di);
      VEC_PUSH(s->tlas_bounding, tb);
    }
    VEC_PUSH(s->tlas_bvh_root, bvh_start);
    for (uint64_t i = 0; i < bvh->n_nodes; i++)
      VEC_PUSH(s->tlas_bvh_forest, flatten_node(&bvh->nodes[i], next[2 * i], next[2 * i + 1], bvh_start, primitive_start));
    free(next); free(aabbs);
    orc_bvh_free(bvh);
  }
  VEC_FREE(blas_box);
  s->built = (rc == 0);
  return rc;
}

void orc_scene_get_view(const orc_scene *s, orc_scene_view *o) {
  o->tlas_binding = s->binding.p; o->n_tlas_binding = s->binding.n;
  o->tlas_bvh_root = s->tlas_bvh_root.p; o->n_tlas_bvh_root = s->tlas_bvh_root.n;
  o->tlas_bvh_forest = s->tlas_bvh_forest.p; o->n_tlas_bvh_forest = s->tlas_bvh_forest.n;
  o->tlas_data = s->tlas_data.p; o->n_tlas_data = s->tlas_data.n;
  o->tlas_bounding = s->tlas_bounding.p; o->n_tlas_bounding = s->tlas_bounding.n;
  o->blas_meta_info = s->blas_meta_info.p; o->n_blas_meta_info = s->blas_meta_info.n;
  o->tri_bvh_root = s->tri_bvh_root.p; o->n_tri_bvh_root = s->tri_bvh_root.n;
  o->tri_bvh_forest = s->tri_bvh_forest.p; o->n_tri_bvh_forest = s->tri_bvh_forest.n;
  o->indices_redirect = s->indices_redirect.p; o->n_indices_redirect = s->indices_redirect.n;
  o->indices = s->indices.p; o->n_indices = s->indices.n;
  o->vertices = (const float *)s->vertices.p; o->n_vertices = s->vertices.n;
  o->balance_fallbacks = s->balance_fallbacks; o->balance_fallbacks_gt10 = s->balance_fallbacks_gt10;
}

/* ---------------- traversal ---------------- */

static inline int intersect_ray_aabb(ov3 o, ov3 d, float t_min, float t_max, ov3 bmin, ov3 bmax) {
  ov3 inv_d = ov3_div(ov3_new(1.0f, 1.0f, 1.0f), d);
  ov3 t0 = ov3_mul(ov3_sub(bmin, o), inv_d);
  ov3 t1 = ov3_mul(ov3_sub(bmax, o), inv_d);
  ov3 t_near = ov3_min(t0, t1);
  ov3 t_far = ov3_max(t0, t1);
  float t_near_max = ov3_max_channel(t_near);
  float t_far_min = ov3_min_channel(t_far);
  return t_near_max <= t_far_min && t_min < t_far_min && t_near_max < t_max;
}

/* returns (sign, t, u, v); sign == 0 -> miss */
static inline ov4 intersect_ray_triangle(ov3 origin, ov3 direction, float rx, float ry, ov3 v0, ov3 v1, ov3 v2,
                                         int cull_enable, int cull_back) {
  ov4 zero = {0, 0, 0, 0};
  ov3 e1 = ov3_sub(v1, v0);
  ov3 e2 = ov3_sub(v2, v0);
  ov3 normal = ov3_normalize(ov3_cross(e1, e2));
  float b = ov3_dot(normal, direction);
  float sign = of_signum(b);
  if (cull_enable) {
    int pass = cull_back != (b < 0.0f);
    if (!pass) return zero;
  }
  ov3 w0 = ov3_sub(origin, v0);
  float a = -ov3_dot(normal, w0);
  float t = a / b;
  if (t < rx || t > ry) return zero;
  ov3 p = ov3_add(origin, ov3_scale(direction, t));
  float uu = ov3_dot(e1, e1);
  float uv = ov3_dot(e1, e2);
  float vv = ov3_dot(e2, e2);
  ov3 w = ov3_sub(p, v0);
  float wu = ov3_dot(w, e1);
  float wv = ov3_dot(w, e2);
  float inverse_d = 1.0f / (uv * uv - uu * vv);
  float u = (uv * wv - vv * wu) * inverse_d;
  if (u < 0.0f || u > 1.0f) return zero;
  float v = (uv * wu - uu * wv) * inverse_d;
  if (v < 0.0f || (u + v) > 1.0f) return zero;
  ov4 r = {sign, t, u, v};
  return r;
}

static inline uint32_t merge_geometry_instance_flag(uint32_t f, uint32_t gi) {
  if (gi & GI_TRIANGLE_FACING_CULL_DISABLE) f &= ~(F_CULL_BACK_FACING_TRIANGLES | F_CULL_FRONT_FACING_TRIANGLES);
  if (gi & GI_TRIANGLE_FLIP_FACING) f ^= F_TRIANGLE_FLIP_FACING;
  if (gi & GI_FORCE_OPAQUE) f |= F_FORCE_OPAQUE;
  if (gi & GI_FORCE_NO_OPAQUE) f |= F_FORCE_NON_OPAQUE;
  return f;
}

/* threaded walk: returns next leaf index or INVALID */
static inline uint32_t bvh_iter_next(const orc_dev_node *bvh, uint32_t *curr_idx, ov3 o, ov3 d, float near, const float *far,
                                     float scaling, orc_counters *c) {
  while (*curr_idx != ORC_INVALID_NEXT) {
    c->bvh_visit++;
    const orc_dev_node *node = &bvh[*curr_idx];
    if (intersect_ray_aabb(o, d, near * scaling, *far * scaling, node->aabb_min, node->aabb_max)) {
      uint32_t curr = *curr_idx;
      *curr_idx = node->hit_next;
      if (node->hit_next == node->miss_next) { c->bvh_hit++; return curr; }
    } else {
      *curr_idx = node->miss_next;
    }
  }
  return ORC_INVALID_NEXT;
}

/* unpruned != 0: NOT the reference — the order-free model used to test the regularity classification (orc_scene_trace_unpruned):
 * every box / triangle range test uses the ray's ORIGINAL range and the closest accepted candidate is kept, i.e. the result every
 * traversal that prunes by its own closest hit converges to when hits lie inside their boxes. */
static uint32_t any_hit_eval(const orc_scene *s, uint32_t geometry_idx, uint32_t primitive_idx, float distance, uint32_t instance_sbt_offset);
static void traverse_clamped(const orc_scene *s, const orc_launch *L, const orc_ray *ray, orc_hit *out, orc_counters *c, int unpruned,
                             float near_walk, float far_init);
static void traverse_one(const orc_scene *s, const orc_launch *L, const orc_ray *ray, orc_hit *out, orc_counters *c, int unpruned) {
  traverse_clamped(s, L, ray, out, c, unpruned, ray->tmin, ray->tmax);
}
/* near_walk / far_init: the range the box and triangle range tests start from.  The reference's walk has near_walk = tmin and
 * far_init = tmax; the tie re-walk of the ordered traversal (orc_scene_trace_ordered_model, csrc/traverse.cu) clamps them around the
 * closest distance while the update_far asserts keep using the ray's own near. */
static void traverse_clamped(const orc_scene *s, const orc_launch *L, const orc_ray *ray, orc_hit *out, orc_counters *c, int unpruned,
                             float near_walk, float far_init) {
  out->t = ray->tmax; out->u = 0; out->v = 0;
  out->primitive_id = out->geometry_id = out->instance_id = out->instance_custom_id = 0xFFFFFFFFu;
  out->hit_kind = 0;

  const uint32_t flags0 = L->ray_flags;
  const float near = ray->tmin;
  float far = far_init;              /* the shared Rc<Cell<f32>> */
  float fixed_far = ray->tmax;
  float *pf = unpruned ? &fixed_far : &far;  /* what the box / triangle range tests see */
  const ov3 ro = ov3_new(ray->ox, ray->oy, ray->oz);
  const ov3 rd = ov3_new(ray->dx, ray->dy, ray->dz);

  if (L->tlas_idx >= s->binding.n) return;
  uint32_t handle = s->binding.p[L->tlas_idx];
  if (handle >= s->tlas_bvh_root.n) return;
  uint32_t tlas_cursor = s->tlas_bvh_root.p[handle];

  for (;;) {
    uint32_t leaf = bvh_iter_next(s->tlas_bvh_forest.p, &tlas_cursor, ro, rd, near_walk, pf, 1.0f, c);
    if (leaf == ORC_INVALID_NEXT) break;
    const orc_dev_node *node = &s->tlas_bvh_forest.p[leaf];
    for (uint32_t tlas_idx = node->range_x; tlas_idx < node->range_y; tlas_idx++) {
      const orc_tlas_bounding *tb = &s->tlas_bounding.p[tlas_idx];
      /* ORIGINAL ray.range, not the shrunken far (traverse_cpu.rs:80-86) */
      if (!intersect_ray_aabb(ro, rd, ray->tmin, ray->tmax, tb->world_min, tb->world_max)) continue;
      if ((L->cull_mask & tb->mask) == 0) continue;
      c->inst_visit++;
      const orc_dev_instance *td = &s->tlas_data.p[tlas_idx];
      uint32_t blas_idx = td->blas;
      uint32_t flags = merge_geometry_instance_flag(flags0, td->flags);

      ov4 o4 = {ro.x, ro.y, ro.z, 1.0f};
      o4 = om4_mul_v4(td->transform_inv, o4);
      ov3 bo = ov3_divs(ov3_new(o4.x, o4.y, o4.z), o4.w);
      ov3 bd = om3_mul_v3(om4_to_mat3(td->transform_inv), rd);
      float scaling = ov3_length(bd);
      bd = ov3_normalize(bd);

      if (blas_idx >= s->blas_meta_info.n) continue;
      const orc_blas_meta *bm = &s->blas_meta_info.p[blas_idx];
      if (flags & F_SKIP_TRIANGLES) continue;

      for (uint32_t tri_root_index = bm->tri_root_x; tri_root_index < bm->tri_root_y; tri_root_index++) {
        orc_geom_meta geometry = s->tri_bvh_root.p[tri_root_index];
        /* cull_geometry */
        int geometry_opaque = (geometry.geometry_flags & G_OPAQUE) != 0;
        int force_opaque = (flags & F_FORCE_OPAQUE) != 0, force_non_opaque = (flags & F_FORCE_NON_OPAQUE) != 0;
        int cull_opaque = (flags & F_CULL_OPAQUE) != 0, cull_non_opaque = (flags & F_CULL_NON_OPAQUE) != 0;
        int is_opaque = (geometry_opaque || force_opaque) && !force_non_opaque;
        int pass = (is_opaque && !cull_opaque) || (!is_opaque && !cull_non_opaque);
        if (!pass) continue;
        /* cull_triangle */
        int flip = (flags & F_TRIANGLE_FLIP_FACING) != 0;
        int cull_front = (flags & F_CULL_FRONT_FACING_TRIANGLES) != 0, cull_back_f = (flags & F_CULL_BACK_FACING_TRIANGLES) != 0;
        int cull_enable = cull_front || cull_back_f;
        int cull_back = (flip && cull_back_f) || (!flip && cull_front);

        uint32_t cursor = geometry.bvh_root_idx;
        for (;;) {
          uint32_t bl = bvh_iter_next(s->tri_bvh_forest.p, &cursor, bo, bd, near_walk, pf, scaling, c);
          if (bl == ORC_INVALID_NEXT) break;
          const orc_dev_node *bn = &s->tri_bvh_forest.p[bl];
          for (uint32_t slot = bn->range_x; slot < bn->range_y; slot++) {
            uint32_t tri_idx = s->indices_redirect.p[slot];
            uint32_t i0 = s->indices.p[(uint64_t)tri_idx * 3], i1 = s->indices.p[(uint64_t)tri_idx * 3 + 1], i2 = s->indices.p[(uint64_t)tri_idx * 3 + 2];
            ov3 v0 = s->vertices.p[i0], v1 = s->vertices.p[i1], v2 = s->vertices.p[i2];
            c->tri_visit++;
            ov4 isect = intersect_ray_triangle(bo, bd, near_walk * scaling, *pf * scaling, v0, v1, v2, cull_enable, cull_back);
            if (isect.x != 0.0f) {
              float distance = isect.y / scaling;
              uint32_t primitive_idx = tri_idx - geometry.primitive_start;
              c->tri_hit++;
              if (unpruned && !(near <= distance && distance <= ray->tmax && distance <= far)) continue;
              /* opaque -> ACCEPT; non-opaque -> any_hit() */
              uint32_t behavior = is_opaque ? 1u : any_hit_eval(s, geometry.geometry_idx, primitive_idx, distance, td->sbt_offset);
              if (behavior & 1u) {
                /* RayRange::update_far: assert!(near <= far); assert!(far <= self.far) */
                if (!(near <= distance) || !(distance <= far)) { c->ref_abort++; continue; }
                far = distance;
                out->t = distance; out->u = isect.z; out->v = isect.w;
                out->primitive_id = primitive_idx; out->geometry_id = geometry.geometry_idx;
                out->instance_id = tlas_idx; out->instance_custom_id = td->instance_custom_index;
                out->hit_kind = isect.x < 0.0f ? HIT_KIND_BACK : HIT_KIND_FRONT;
                if (flags & F_ACCEPT_FIRST_HIT_AND_END_SEARCH) behavior |= 2u;
              }
              if (behavior & 2u) return;
            }
          }
        }
      }
    }
  }
}

/* ---------- any-hit (traverse_cpu.rs:164-192) ----------
 * The reference calls the pipeline's any-hit shader for every candidate of NON-OPAQUE geometry: TraceTaskImpl::device_poll selects
 * it through the shader binding table — hit_group = sbt_ray_config.offset + stride * geometry_id + instance_sbt_offset
 * (api/ctx.rs:53-55), shader = sbt.get_any_handle(current_sbt, hit_group), no shader = ANYHIT_BEHAVIOR_ACCEPT_HIT
 * (trace_task.rs:189-203) — and honours the bits it returns: ACCEPT_HIT commits the candidate (range shrinks, result replaced),
 * END_SEARCH stops the whole traversal.  A shader is arbitrary code there; here it is one of a few stateless programs over the
 * fields of the reference's `Hit` (geometry_idx, primitive_idx, distance). */
static uint32_t any_hit_eval(const orc_scene *s, uint32_t geometry_idx, uint32_t primitive_idx, float distance, uint32_t instance_sbt_offset) {
  const orc_anyhit_setup *a = &s->anyhit;
  uint32_t program = 0xFFFFFFFFu;
  if (a->mode == 1) {
    program = a->uniform_program;
  } else if (a->mode == 2) {
    uint32_t group = a->sbt_ray_offset + a->sbt_ray_stride * geometry_idx + instance_sbt_offset;
    if (group < a->n_hit_groups) program = s->anyhit_groups[group];
  }
  if (program >= a->n_programs) return 1u; /* no shader: ACCEPT_HIT */
  const orc_anyhit_program *p = &s->anyhit_programs[program];
  int holds = 1;
  if (p->kind == 1) holds = (primitive_idx & p->mask) == p->value;
  else if (p->kind == 2) holds = distance >= p->distance;
  return holds ? p->behavior : p->otherwise;
}

int orc_scene_set_any_hit(orc_scene *s, const orc_anyhit_setup *setup) {
  free(s->anyhit_programs); free(s->anyhit_groups);
  s->anyhit_programs = NULL; s->anyhit_groups = NULL;
  memset(&s->anyhit, 0, sizeof(s->anyhit));
  if (!setup) return 0;
  s->anyhit = *setup;
  if (setup->n_programs) {
    s->anyhit_programs = (orc_anyhit_program *)malloc(sizeof(orc_anyhit_program) * setup->n_programs);
    memcpy(s->anyhit_programs, setup->programs, sizeof(orc_anyhit_program) * setup->n_programs);
  }
  if (setup->n_hit_groups) {
    s->anyhit_groups = (uint32_t *)malloc(sizeof(uint32_t) * setup->n_hit_groups);
    memcpy(s->anyhit_groups, setup->hit_group_any, sizeof(uint32_t) * setup->n_hit_groups);
  }
  s->anyhit.programs = s->anyhit_programs; s->anyhit.hit_group_any = s->anyhit_groups;
  return 0;
}

/* Brute-force candidate list of one ray (debugging / self-consistency): every instance of the bound TLAS that passes the
 * mask + original-range box test and every triangle of its BLAS, NO BVH pruning and NO live range: the triangle test runs with
 * the ray's original range.  The closest hit of the traversal is the minimum-distance entry (first / last of equals per the
 * reference's order).  Returns the number of candidates (may exceed cap; only cap are written). */
uint64_t orc_scene_candidates(const orc_scene *s, const orc_launch *L, const orc_ray *ray, orc_candidate *out, uint64_t cap) {
  uint64_t n = 0;
  const ov3 ro = ov3_new(ray->ox, ray->oy, ray->oz), rd = ov3_new(ray->dx, ray->dy, ray->dz);
  if (L->tlas_idx >= s->binding.n) return 0;
  uint32_t handle = s->binding.p[L->tlas_idx];
  if (handle >= s->tlas_bvh_root.n) return 0;
  for (uint32_t tn = s->tlas_bvh_root.p[handle]; tn != ORC_INVALID_NEXT;) {
    const orc_dev_node *node = &s->tlas_bvh_forest.p[tn];
    const int leaf = node->hit_next == node->miss_next;
    tn = node->hit_next;  /* pre-order over ALL nodes */
    if (!leaf) continue;
    for (uint32_t tlas_idx = node->range_x; tlas_idx < node->range_y; tlas_idx++) {
      const orc_tlas_bounding *tb = &s->tlas_bounding.p[tlas_idx];
      if (!intersect_ray_aabb(ro, rd, ray->tmin, ray->tmax, tb->world_min, tb->world_max)) continue;
      if ((L->cull_mask & tb->mask) == 0) continue;
      const orc_dev_instance *td = &s->tlas_data.p[tlas_idx];
      uint32_t flags = merge_geometry_instance_flag(L->ray_flags, td->flags);
      ov4 o4 = {ro.x, ro.y, ro.z, 1.0f};
      o4 = om4_mul_v4(td->transform_inv, o4);
      ov3 bo = ov3_divs(ov3_new(o4.x, o4.y, o4.z), o4.w);
      ov3 bd = om3_mul_v3(om4_to_mat3(td->transform_inv), rd);
      float scaling = ov3_length(bd);
      bd = ov3_normalize(bd);
      if (td->blas >= s->blas_meta_info.n || (flags & F_SKIP_TRIANGLES)) continue;
      const orc_blas_meta *bm = &s->blas_meta_info.p[td->blas];
      for (uint32_t g = bm->tri_root_x; g < bm->tri_root_y; g++) {
        orc_geom_meta geometry = s->tri_bvh_root.p[g];
        int geometry_opaque = (geometry.geometry_flags & G_OPAQUE) != 0;
        int is_opaque = (geometry_opaque || (flags & F_FORCE_OPAQUE)) && !(flags & F_FORCE_NON_OPAQUE);
        int pass = (is_opaque && !(flags & F_CULL_OPAQUE)) || (!is_opaque && !(flags & F_CULL_NON_OPAQUE));
        if (!pass) continue;
        int flip = (flags & F_TRIANGLE_FLIP_FACING) != 0;
        int cull_front = (flags & F_CULL_FRONT_FACING_TRIANGLES) != 0, cull_back_f = (flags & F_CULL_BACK_FACING_TRIANGLES) != 0;
        int cull_enable = cull_front || cull_back_f;
        int cull_back = (flip && cull_back_f) || (!flip && cull_front);
        for (uint32_t bn = geometry.bvh_root_idx; bn != ORC_INVALID_NEXT;) {
          const orc_dev_node *b = &s->tri_bvh_forest.p[bn];
          const int bleaf = b->hit_next == b->miss_next;
          bn = b->hit_next;
          if (!bleaf) continue;
          for (uint32_t slot = b->range_x; slot < b->range_y; slot++) {
            uint32_t tri_idx = s->indices_redirect.p[slot];
            ov3 v0 = s->vertices.p[s->indices.p[(uint64_t)tri_idx * 3]], v1 = s->vertices.p[s->indices.p[(uint64_t)tri_idx * 3 + 1]],
                v2 = s->vertices.p[s->indices.p[(uint64_t)tri_idx * 3 + 2]];
            ov4 isect = intersect_ray_triangle(bo, bd, ray->tmin * scaling, ray->tmax * scaling, v0, v1, v2, cull_enable, cull_back);
            if (isect.x == 0.0f) continue;
            float distance = isect.y / scaling;
            if (n < cap) {
              orc_candidate *c = &out[n];
              c->distance = distance; c->t_object = isect.y; c->u = isect.z; c->v = isect.w; c->sign = isect.x; c->scaling = scaling;
              c->instance_id = tlas_idx; c->geometry_id = geometry.geometry_idx; c->primitive_id = tri_idx - geometry.primitive_start;
              c->slot = slot; c->in_range = (ray->tmin <= distance) && (distance <= ray->tmax); c->pad = 0;
            }
            n++;
          }
        }
      }
    }
  }
  return n;
}


/* ---------------------------------------------------------------------------------------------------------------------------
 * NOT the reference: a scalar model of the ORDERED traversal the CUDA kernel performs (csrc/traverse.cu k_trace_ordered_rounds) —
 * near child first with a stack, pruning against bound = min(tmax, best + TIE_EPS |best|), the kernel's acceptance rule, near-tie
 * detection through the second-smallest candidate distance, re-walk in the reference's order clamped to best -/+ 3 TIE_EPS |best|,
 * whole-range walk when the clamped one finds nothing.  Test tool: lets the CPU suite fuzz the ALGORITHM (does it end on the
 * reference's record?) over far more rays and scenes than GPU time allows.  Rays the product classifies as able to reach irregular
 * content are the caller's business (the kernel hands them to the reference-order walk). */
#define ORC_TIE_EPS 1e-5f
typedef struct { float best, second, u, v, sign; uint32_t slot_tri, inst, geom, prim; int found; float bound; } ordered_state;

static inline int slab_entry(ov3 o, ov3 inv_d, float t_min, float t_max, ov3 bmin, ov3 bmax, float *t_near_max) {
  ov3 t0 = ov3_mul(ov3_sub(bmin, o), inv_d);
  ov3 t1 = ov3_mul(ov3_sub(bmax, o), inv_d);
  ov3 t_near = ov3_min(t0, t1);
  ov3 t_far = ov3_max(t0, t1);
  *t_near_max = ov3_max_channel(t_near);
  float t_far_min = ov3_min_channel(t_far);
  return *t_near_max <= t_far_min && t_min < t_far_min && *t_near_max < t_max;
}

/* near-first walk of one threaded tree; leaf_fn(ctx, leaf node) may shrink *far_s */
typedef void (*ordered_leaf_fn)(void *ctx, const orc_dev_node *leaf);
static void ordered_walk_tree(const orc_dev_node *forest, uint32_t root, ov3 o, ov3 d, float near_s, const float *far_s, ordered_leaf_fn fn, void *ctx) {
  if (root == ORC_INVALID_NEXT) return;
  ov3 inv = ov3_div(ov3_new(1.0f, 1.0f, 1.0f), d);
  float tn;
  if (!slab_entry(o, inv, near_s, *far_s, forest[root].aabb_min, forest[root].aabb_max, &tn)) return;  /* the pseudo root */
  uint32_t stack[256];
  int sp = 0;
  uint32_t cur = root;
  for (;;) {
    const orc_dev_node *node = &forest[cur];
    if (node->hit_next == node->miss_next) {
      fn(ctx, node);
    } else {
      uint32_t l = cur + 1, r = forest[l].miss_next;
      float n0, n1;
      int h0 = slab_entry(o, inv, near_s, *far_s, forest[l].aabb_min, forest[l].aabb_max, &n0);
      int h1 = slab_entry(o, inv, near_s, *far_s, forest[r].aabb_min, forest[r].aabb_max, &n1);
      if (h0 && h1) {
        int first0 = n0 <= n1;
        if (sp < 256) stack[sp++] = first0 ? r : l;
        cur = first0 ? l : r;
        continue;
      } else if (h0) { cur = l; continue; }
      else if (h1) { cur = r; continue; }
    }
    if (sp == 0) return;
    cur = stack[--sp];
  }
}

typedef struct {
  const orc_scene *s; const orc_launch *L; const orc_ray *ray; ordered_state *st;
  ov3 ro, rd;            /* world ray */
  ov3 bo, bd; float scaling, far_s; uint32_t flags, tlas_idx; orc_geom_meta geometry; int cull_enable, cull_back;  /* current instance / geometry */
  float far_world;       /* = bound, as the world-level range end */
} ordered_ctx;

static void ordered_triangle_leaf(void *p, const orc_dev_node *leaf) {
  ordered_ctx *x = (ordered_ctx *)p;
  const orc_scene *s = x->s;
  ordered_state *st = x->st;
  const float near = x->ray->tmin, far0 = x->ray->tmax;
  for (uint32_t slot = leaf->range_x; slot < leaf->range_y; slot++) {
    uint32_t tri_idx = s->indices_redirect.p[slot];
    ov3 v0 = s->vertices.p[s->indices.p[(uint64_t)tri_idx * 3]], v1 = s->vertices.p[s->indices.p[(uint64_t)tri_idx * 3 + 1]],
        v2 = s->vertices.p[s->indices.p[(uint64_t)tri_idx * 3 + 2]];
    ov4 isect = intersect_ray_triangle(x->bo, x->bd, near * x->scaling, x->far_s, v0, v1, v2, x->cull_enable, x->cull_back);
    if (isect.x == 0.0f) continue;
    float distance = isect.y / x->scaling;
    if (!(near <= distance) || !(distance <= far0)) continue;
    if (!st->found || distance < st->best) {
      if (st->found) st->second = fminf(st->second, st->best);
      st->found = 1;
      st->best = distance; st->u = isect.z; st->v = isect.w; st->sign = isect.x;
      st->inst = x->tlas_idx; st->geom = x->geometry.geometry_idx; st->prim = tri_idx - x->geometry.primitive_start;
      st->bound = fminf(far0, st->best + ORC_TIE_EPS * fabsf(st->best));
      x->far_s = st->bound * x->scaling;
      x->far_world = st->bound;
    } else {
      st->second = fminf(st->second, distance);
    }
  }
}

static void ordered_instance_leaf(void *p, const orc_dev_node *leaf) {
  ordered_ctx *x = (ordered_ctx *)p;
  const orc_scene *s = x->s;
  const orc_launch *L = x->L;
  ov3 inv = ov3_div(ov3_new(1.0f, 1.0f, 1.0f), x->rd);
  for (uint32_t tlas_idx = leaf->range_x; tlas_idx < leaf->range_y; tlas_idx++) {
    const orc_tlas_bounding *tb = &s->tlas_bounding.p[tlas_idx];
    float tn;
    if (!slab_entry(x->ro, inv, x->ray->tmin, x->st->bound, tb->world_min, tb->world_max, &tn)) continue;  /* the kernel prunes with its bound */
    if ((L->cull_mask & tb->mask) == 0) continue;
    const orc_dev_instance *td = &s->tlas_data.p[tlas_idx];
    uint32_t flags = merge_geometry_instance_flag(L->ray_flags, td->flags);
    if ((flags & F_SKIP_TRIANGLES) || td->blas >= s->blas_meta_info.n) continue;
    const orc_blas_meta *bm = &s->blas_meta_info.p[td->blas];
    ov4 o4 = {x->ro.x, x->ro.y, x->ro.z, 1.0f};
    o4 = om4_mul_v4(td->transform_inv, o4);
    x->bo = ov3_divs(ov3_new(o4.x, o4.y, o4.z), o4.w);
    ov3 bd0 = om3_mul_v3(om4_to_mat3(td->transform_inv), x->rd);
    x->scaling = ov3_length(bd0);
    x->bd = ov3_normalize(bd0);
    x->far_s = x->st->bound * x->scaling;
    x->flags = flags; x->tlas_idx = tlas_idx;
    int flip = (flags & F_TRIANGLE_FLIP_FACING) != 0;
    int cull_front = (flags & F_CULL_FRONT_FACING_TRIANGLES) != 0, cull_back_f = (flags & F_CULL_BACK_FACING_TRIANGLES) != 0;
    x->cull_enable = cull_front || cull_back_f;
    x->cull_back = (flip && cull_back_f) || (!flip && cull_front);
    for (uint32_t g = bm->tri_root_x; g < bm->tri_root_y; g++) {
      x->geometry = s->tri_bvh_root.p[g];
      int geometry_opaque = (x->geometry.geometry_flags & G_OPAQUE) != 0;
      int is_opaque = (geometry_opaque || (flags & F_FORCE_OPAQUE)) && !(flags & F_FORCE_NON_OPAQUE);
      int pass = (is_opaque && !(flags & F_CULL_OPAQUE)) || (!is_opaque && !(flags & F_CULL_NON_OPAQUE));
      if (!pass) continue;
      ordered_walk_tree(s->tri_bvh_forest.p, x->geometry.bvh_root_idx, x->bo, x->bd, x->ray->tmin * x->scaling, &x->far_s, ordered_triangle_leaf, x);
    }
  }
}

/* returns 0 = the ordered result stands, 1 = near-tie resolved by the clamped re-walk, 2 = ... by the whole-range walk */
static int traverse_ordered_model(const orc_scene *s, const orc_launch *L, const orc_ray *ray, orc_hit *out) {
  orc_counters c; memset(&c, 0, sizeof(c));
  if (L->ray_flags & F_ACCEPT_FIRST_HIT_AND_END_SEARCH) { traverse_one(s, L, ray, out, &c, 0); return 0; }  /* reference-order kernel */
  out->t = ray->tmax; out->u = 0; out->v = 0;
  out->primitive_id = out->geometry_id = out->instance_id = out->instance_custom_id = 0xFFFFFFFFu;
  out->hit_kind = 0;
  if (L->tlas_idx >= s->binding.n) return 0;
  uint32_t handle = s->binding.p[L->tlas_idx];
  if (handle >= s->tlas_bvh_root.n) return 0;
  ordered_state st; memset(&st, 0, sizeof(st));
  st.best = INFINITY; st.second = INFINITY; st.bound = ray->tmax;
  ordered_ctx x; memset(&x, 0, sizeof(x));
  x.s = s; x.L = L; x.ray = ray; x.st = &st;
  x.ro = ov3_new(ray->ox, ray->oy, ray->oz); x.rd = ov3_new(ray->dx, ray->dy, ray->dz);
  x.far_world = ray->tmax;
  ordered_walk_tree(s->tlas_bvh_forest.p, s->tlas_bvh_root.p[handle], x.ro, x.rd, ray->tmin, &x.far_world, ordered_instance_leaf, &x);
  if (!st.found) return 0;
  if (st.second <= st.best + ORC_TIE_EPS * fabsf(st.best)) {
    const float slack = 3.0f * ORC_TIE_EPS * fabsf(st.best);
    traverse_clamped(s, L, ray, out, &c, 0, fmaxf(ray->tmin, st.best - slack), fminf(ray->tmax, st.best + slack));
    if (out->instance_id != 0xFFFFFFFFu) return 1;
    traverse_one(s, L, ray, out, &c, 0);
    return 2;
  }
  out->t = st.best; out->u = st.u; out->v = st.v;
  out->primitive_id = st.prim; out->geometry_id = st.geom; out->instance_id = st.inst;
  out->instance_custom_id = s->tlas_data.p[st.inst].instance_custom_index;
  out->hit_kind = st.sign < 0.0f ? HIT_KIND_BACK : HIT_KIND_FRONT;
  return 0;
}

/* single-threaded batch; out_stats[0] = rays resolved by the clamped re-walk, [1] = by the whole-range walk */
int orc_scene_trace_ordered_model(const orc_scene *s, const orc_launch *launch, const orc_ray *rays, uint64_t n_rays, orc_hit *out_hits,
                                  uint64_t *out_stats) {
  if (!s->built) return -1;
  uint64_t ties = 0, whole = 0;
  for (uint64_t i = 0; i < n_rays; i++) {
    int r = traverse_ordered_model(s, launch, &rays[i], &out_hits[i]);
    ties += r == 1; whole += r == 2;
  }
  if (out_stats) { out_stats[0] = ties; out_stats[1] = whole; }
  return 0;
}

/* multi-thread driver: rays are handed out in chunks from an atomic cursor (hit rays cluster in image space, so static
 * ranges would leave most threads idle) */
#define ORC_CHUNK 2048u
typedef struct { const orc_scene *s; const orc_launch *L; const orc_ray *rays; orc_hit *hits; uint64_t n; atomic_ullong *cursor; orc_counters c; int unpruned; } trace_job;
static void *trace_worker(void *p) {
  trace_job *j = (trace_job *)p;
  for (;;) {
    uint64_t begin = atomic_fetch_add(j->cursor, ORC_CHUNK);
    if (begin >= j->n) break;
    uint64_t end = begin + ORC_CHUNK < j->n ? begin + ORC_CHUNK : j->n;
    for (uint64_t i = begin; i < end; i++) traverse_one(j->s, j->L, &j->rays[i], &j->hits[i], &j->c, j->unpruned);
  }
  return NULL;
}

static int scene_trace(const orc_scene *s, const orc_launch *launch, const orc_ray *rays, uint64_t n_rays,
                       orc_hit *out_hits, orc_counters *counters, int n_threads, int unpruned) {
  if (!s->built) return -1;
  if (n_threads < 1) n_threads = 1;
  trace_job *jobs = (trace_job *)calloc(n_threads, sizeof(trace_job));
  pthread_t *th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
  atomic_ullong cursor;
  atomic_init(&cursor, 0);
  for (int t = 0; t < n_threads; t++) {
    jobs[t].s = s; jobs[t].L = launch; jobs[t].rays = rays; jobs[t].hits = out_hits; jobs[t].n = n_rays; jobs[t].cursor = &cursor;
    jobs[t].unpruned = unpruned;
  }
  if (n_threads == 1) trace_worker(&jobs[0]);
  else {
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, trace_worker, &jobs[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  }
  if (counters) {
    memset(counters, 0, sizeof(*counters));
    for (int t = 0; t < n_threads; t++) {
      counters->bvh_visit += jobs[t].c.bvh_visit; counters->bvh_hit += jobs[t].c.bvh_hit;
      counters->tri_visit += jobs[t].c.tri_visit; counters->tri_hit += jobs[t].c.tri_hit;
      counters->inst_visit += jobs[t].c.inst_visit; counters->ref_abort += jobs[t].c.ref_abort;
    }
  }
  free(jobs); free(th);
  return 0;
}

int orc_scene_trace(const orc_scene *s, const orc_launch *launch, const orc_ray *rays, uint64_t n_rays,
                    orc_hit *out_hits, orc_counters *counters, int n_threads) {
  return scene_trace(s, launch, rays, n_rays, out_hits, counters, n_threads, 0);
}

/* NOT the reference: the closest candidate of a walk over the same trees that never shrinks its range (see traverse_one).  Where
 * it differs from orc_scene_trace beyond an exact tie, the reference's answer depends on its visiting order. */
int orc_scene_trace_unpruned(const orc_scene *s, const orc_launch *launch, const orc_ray *rays, uint64_t n_rays,
                             orc_hit *out_hits, int n_threads) {
  return scene_trace(s, launch, rays, n_rays, out_hits, NULL, n_threads, 1);
}

/* ---- Mat4 helpers for KATs / scene builders ---- */
void orc_mat4_compose(const float *a16, const float *b16, float *out16) {
  om4 a, b; memcpy(&a, a16, sizeof(a)); memcpy(&b, b16, sizeof(b));
  om4 r = om4_mul(a, b); memcpy(out16, &r, sizeof(r));
}
void orc_mat4_inverse_or_identity(const float *m16, float *out16) {
  om4 m; memcpy(&m, m16, sizeof(m));
  om4 r = om4_inverse_or_identity(m); memcpy(out16, &r, sizeof(r));
}
void orc_mat4_mul_vec4(const float *m16, const float *v4, float *out4) {
  om4 m; memcpy(&m, m16, sizeof(m));
  ov4 v = {v4[0], v4[1], v4[2], v4[3]};
  ov4 r = om4_mul_v4(m, v);
  out4[0] = r.x; out4[1] = r.y; out4[2] = r.z; out4[3] = r.w;
}
