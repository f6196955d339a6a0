/* ORACLE — TEST INFRASTRUCTURE ONLY.  Path A (content/space BVH query through mesh-core),
 * brute-force nearest, and the parallel-compute scan / compaction / scatter restatements.
 *
 * Follows /root/reference:
 *   math/geometry/src/dimension3/intersection.rs:3-77     Ray3 x Triangle (GTE form, FaceSide)
 *   math/geometry/src/dimension3/intersection.rs:128-207  Ray3 x Box3 (sign-selected slabs, NaN-aware tighten)
 *   math/geometry/src/intersect_util.rs:15-17,52-70       strict `<` nearest refresh (first of equal wins)
 *   math/geometry/src/hyper_ray.rs:11-13                  at(t) = origin + direction * t
 *   utility/abstract-tree/src/lib.rs:33-51                stack DFS; children pushed left,right => right popped first
 *   content/space/src/bvh/mod.rs:94-113                   traverse_branch_leaf_visitor (leaves are not box tested)
 *   content/mesh/core/src/feature/bvh.rs:57-86            intersect_nearest_bvh
 *   content/mesh/core/src/feature/intersection.rs:11-37   ray_intersect_nearest (brute force)
 *   shader/parallel-compute/src/prefix_scan.rs:64-101     per-workgroup Kogge-Stone inclusive scan
 *   shader/parallel-compute/src/lib.rs:398-429            two-level (global) scan
 *   shader/parallel-compute/src/stream_compaction.rs:3-45 use_stream_compaction
 *   shader/parallel-compute/src/shuffle_move.rs:27-45     shuffle_move scatter
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include "oracle.h"

int orc_ray_box_a(const float *r, const obox *box3) {
  ov3 origin = ov3_new(r[0], r[1], r[2]);
  ov3 dir = ov3_new(r[3], r[4], r[5]);
  float t_max, t_min, ty_min, ty_max, tz_min, tz_max;
  float inv_dir_x = 1.0f / dir.x;
  float inv_dir_y = 1.0f / dir.y;
  float inv_dir_z = 1.0f / dir.z;
  if (inv_dir_x >= 0.0f) {
    t_min = (box3->min.x - origin.x) * inv_dir_x;
    t_max = (box3->max.x - origin.x) * inv_dir_x;
  } else {
    t_min = (box3->max.x - origin.x) * inv_dir_x;
    t_max = (box3->min.x - origin.x) * inv_dir_x;
  }
  if (inv_dir_y >= 0.0f) {
    ty_min = (box3->min.y - origin.y) * inv_dir_y;
    ty_max = (box3->max.y - origin.y) * inv_dir_y;
  } else {
    ty_min = (box3->max.y - origin.y) * inv_dir_y;
    ty_max = (box3->min.y - origin.y) * inv_dir_y;
  }
  if ((t_min > ty_max) || (ty_min > t_max)) return 0;
  if (ty_min > t_min || isnan(t_min)) t_min = ty_min;
  if (ty_max < t_max || isnan(t_max)) t_max = ty_max;
  if (inv_dir_z >= 0.0f) {
    tz_min = (box3->min.z - origin.z) * inv_dir_z;
    tz_max = (box3->max.z - origin.z) * inv_dir_z;
  } else {
    tz_min = (box3->max.z - origin.z) * inv_dir_z;
    tz_max = (box3->min.z - origin.z) * inv_dir_z;
  }
  if ((t_min > tz_max) || (tz_min > t_max)) return 0;
  if (tz_min > t_min || isnan(t_min)) t_min = tz_min;
  if (tz_max < t_max || isnan(t_max)) t_max = tz_max;
  if (t_max < 0.0f) return 0;
  return 1;
}

int orc_ray_triangle_a(const float *r, ov3 a, ov3 b, ov3 c, int face_side, float *out) {
  ov3 origin = ov3_new(r[0], r[1], r[2]);
  ov3 dir = ov3_new(r[3], r[4], r[5]);
  if (face_side == ORC_FACE_BACK) { ov3 t = a; a = c; c = t; }  /* Triangle::flip swaps a and c */
  int backface_culling = face_side != ORC_FACE_DOUBLE;
  ov3 edge1 = ov3_sub(b, a);
  ov3 edge2 = ov3_sub(c, a);
  ov3 normal = ov3_cross(edge1, edge2);
  float DdN = ov3_dot(dir, normal);
  float sign;
  if (DdN > 0.0f) {
    if (backface_culling) return 0;
    sign = 1.0f;
  } else if (DdN < 0.0f) {
    sign = -1.0f;
    DdN = -DdN;
  } else {
    return 0;
  }
  ov3 diff = ov3_sub(origin, a);
  float DdQxE2 = sign * ov3_dot(dir, ov3_cross(diff, edge2));
  if (DdQxE2 < 0.0f) return 0;
  float DdE1xQ = sign * ov3_dot(dir, ov3_cross(edge1, diff));
  if (DdE1xQ < 0.0f) return 0;
  if (DdQxE2 + DdE1xQ > DdN) return 0;
  float QdN = -sign * ov3_dot(diff, normal);
  if (QdN < 0.0f) return 0;
  float t = QdN / DdN;
  ov3 p = ov3_add(origin, ov3_scale(dir, t));
  out[0] = p.x; out[1] = p.y; out[2] = p.z; out[3] = t;
  return 1;
}

static inline void tri_of(const float *positions, const uint32_t *indices, uint64_t prim, ov3 *a, ov3 *b, ov3 *c) {
  uint32_t i0 = indices[3 * prim], i1 = indices[3 * prim + 1], i2 = indices[3 * prim + 2];
  *a = ov3_new(positions[3 * i0], positions[3 * i0 + 1], positions[3 * i0 + 2]);
  *b = ov3_new(positions[3 * i1], positions[3 * i1 + 1], positions[3 * i1 + 2]);
  *c = ov3_new(positions[3 * i2], positions[3 * i2 + 1], positions[3 * i2 + 2]);
}

static inline void refresh_nearest(orc_mesh_hit *best, const float *pd, uint64_t prim) {
  if (!best->hit || pd[3] < best->distance) {
    best->px = pd[0]; best->py = pd[1]; best->pz = pd[2]; best->distance = pd[3];
    best->primitive_index = (uint32_t)prim; best->hit = 1;
  }
}

static void patha_one(const orc_bvh *bvh, const float *positions, const uint32_t *indices, const orc_ray *ray,
                      int face_side, orc_mesh_hit *out, uint64_t *stack) {
  float r[6] = {ray->ox, ray->oy, ray->oz, ray->dx, ray->dy, ray->dz};
  memset(out, 0, sizeof(*out));
  uint64_t sp = 0;
  stack[sp++] = 0;
  while (sp > 0) {
    const orc_bvh_node *node = &bvh->nodes[stack[--sp]];
    if (!node->has_child) {
      for (uint64_t k = node->start; k < node->end; k++) {
        uint64_t prim = bvh->sorted_primitive_index[k];
        ov3 a, b, c; tri_of(positions, indices, prim, &a, &b, &c);
        float pd[4];
        if (orc_ray_triangle_a(r, a, b, c, face_side, pd)) refresh_nearest(out, pd, prim);
      }
    } else if (orc_ray_box_a(r, &node->bounding)) {
      stack[sp++] = node->self_index + 1;                     /* left pushed first ... */
      stack[sp++] = node->self_index + node->left_count + 1;  /* ... right popped first */
    }
  }
}

typedef struct {
  const orc_bvh *bvh; const float *positions; const uint32_t *indices; uint64_t n_tris;
  const orc_ray *rays; orc_mesh_hit *out; int face_side; uint64_t n; atomic_ullong *cursor;
} qa_job;
#define ORC_QA_CHUNK 1024u

static void *patha_worker(void *p) {
  qa_job *j = (qa_job *)p;
  uint64_t *stack = (uint64_t *)malloc((j->bvh->n_nodes + 2) * sizeof(uint64_t));
  for (;;) {
    uint64_t begin = atomic_fetch_add(j->cursor, ORC_QA_CHUNK);
    if (begin >= j->n) break;
    uint64_t end = begin + ORC_QA_CHUNK < j->n ? begin + ORC_QA_CHUNK : j->n;
    for (uint64_t i = begin; i < end; i++) patha_one(j->bvh, j->positions, j->indices, &j->rays[i], j->face_side, &j->out[i], stack);
  }
  free(stack);
  return NULL;
}
static void *brute_worker(void *p) {
  qa_job *j = (qa_job *)p;
  for (;;) {
  uint64_t begin = atomic_fetch_add(j->cursor, ORC_QA_CHUNK);
  if (begin >= j->n) break;
  uint64_t end = begin + ORC_QA_CHUNK < j->n ? begin + ORC_QA_CHUNK : j->n;
  for (uint64_t i = begin; i < end; i++) {
    const orc_ray *ray = &j->rays[i];
    float r[6] = {ray->ox, ray->oy, ray->oz, ray->dx, ray->dy, ray->dz};
    orc_mesh_hit *out = &j->out[i];
    memset(out, 0, sizeof(*out));
    for (uint64_t prim = 0; prim < j->n_tris; prim++) {
      ov3 a, b, c; tri_of(j->positions, j->indices, prim, &a, &b, &c);
      float pd[4];
      if (orc_ray_triangle_a(r, a, b, c, j->face_side, pd)) refresh_nearest(out, pd, prim);
    }
  }
  }
  return NULL;
}

static void run_jobs(qa_job proto, uint64_t n_rays, int n_threads, void *(*fn)(void *)) {
  if (n_threads < 1) n_threads = 1;
  qa_job *jobs = (qa_job *)calloc(n_threads, sizeof(qa_job));
  pthread_t *th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
  atomic_ullong cursor;
  atomic_init(&cursor, 0);
  for (int t = 0; t < n_threads; t++) { jobs[t] = proto; jobs[t].n = n_rays; jobs[t].cursor = &cursor; }
  if (n_threads == 1) fn(&jobs[0]);
  else {
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, fn, &jobs[t]);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  }
  free(jobs); free(th);
}

void orc_patha_query_nearest(const orc_bvh *bvh, const float *positions, const uint32_t *indices,
                             const orc_ray *rays, uint64_t n_rays, int face_side, orc_mesh_hit *out, int n_threads) {
  qa_job j; memset(&j, 0, sizeof(j));
  j.bvh = bvh; j.positions = positions; j.indices = indices; j.rays = rays; j.out = out; j.face_side = face_side;
  run_jobs(j, n_rays, n_threads, patha_worker);
}

void orc_brute_query_nearest(const float *positions, const uint32_t *indices, uint64_t n_tris,
                             const orc_ray *rays, uint64_t n_rays, int face_side, orc_mesh_hit *out, int n_threads) {
  qa_job j; memset(&j, 0, sizeof(j));
  j.positions = positions; j.indices = indices; j.n_tris = n_tris; j.rays = rays; j.out = out; j.face_side = face_side;
  run_jobs(j, n_rays, n_threads, brute_worker);
}

/* intersect_list_bvh (content/mesh/core/src/feature/bvh.rs:23-55): every intersected primitive in visiting order (same
 * right-first DFS, primitives of a leaf in sorted_primitive_index order).  Pass out == NULL to count only.  Single thread:
 * the list of one ray is appended after the previous ray's (CSR: offsets[i] .. offsets[i+1]). */
uint64_t orc_patha_query_list(const orc_bvh *bvh, const float *positions, const uint32_t *indices, const orc_ray *rays,
                              uint64_t n_rays, int face_side, uint64_t *offsets, orc_mesh_hit *out, uint64_t capacity) {
  uint64_t *stack = (uint64_t *)malloc((bvh->n_nodes + 2) * sizeof(uint64_t));
  uint64_t total = 0;
  for (uint64_t i = 0; i < n_rays; i++) {
    const orc_ray *ray = &rays[i];
    float r[6] = {ray->ox, ray->oy, ray->oz, ray->dx, ray->dy, ray->dz};
    offsets[i] = total;
    uint64_t sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
      const orc_bvh_node *node = &bvh->nodes[stack[--sp]];
      if (!node->has_child) {
        for (uint64_t k = node->start; k < node->end; k++) {
          uint64_t prim = bvh->sorted_primitive_index[k];
          ov3 a, b, c; tri_of(positions, indices, prim, &a, &b, &c);
          float pd[4];
          if (orc_ray_triangle_a(r, a, b, c, face_side, pd)) {
            if (out && total < capacity) {
              orc_mesh_hit *h = &out[total];
              memset(h, 0, sizeof(*h));
              h->px = pd[0]; h->py = pd[1]; h->pz = pd[2]; h->distance = pd[3]; h->primitive_index = (uint32_t)prim; h->hit = 1;
            }
            total++;
          }
        }
      } else if (orc_ray_box_a(r, &node->bounding)) {
        stack[sp++] = node->self_index + 1;
        stack[sp++] = node->self_index + node->left_count + 1;
      }
    }
  }
  offsets[n_rays] = total;
  free(stack);
  return total;
}

/* ---- parallel-compute restatements ---- */

/* Kogge-Stone per workgroup: after log2(W) steps value[i] = sum of its workgroup prefix (prefix_scan.rs:64-101) */
void orc_workgroup_inclusive_scan_u32(const uint32_t *in, uint64_t n, uint32_t W, uint32_t *out) {
  uint32_t *a = (uint32_t *)malloc((W ? W : 1) * sizeof(uint32_t));
  uint32_t *b = (uint32_t *)malloc((W ? W : 1) * sizeof(uint32_t));
  for (uint64_t base = 0; base < n; base += W) {
    uint64_t m = n - base < W ? n - base : W;
    for (uint32_t i = 0; i < W; i++) a[i] = i < m ? in[base + i] : 0;
    for (uint32_t stride = 1; stride < W; stride <<= 1) {
      for (uint32_t i = 0; i < W; i++) b[i] = i >= stride ? a[i] + a[i - stride] : a[i];
      uint32_t *t = a; a = b; b = t;
    }
    for (uint64_t i = 0; i < m; i++) out[base + i] = a[i];
  }
  free(a); free(b);
}

void orc_inclusive_scan_u32(const uint32_t *in, uint64_t n, uint32_t *out) {
  uint32_t acc = 0;
  for (uint64_t i = 0; i < n; i++) { acc += in[i]; out[i] = acc; }
}

void orc_shuffle_move_u32(const uint32_t *in, const uint32_t *target, const uint8_t *moved, uint64_t n, uint32_t *out) {
  for (uint64_t i = 0; i < n; i++) if (moved[i]) out[target[i]] = in[i];
}

uint64_t orc_stream_compaction_u32(const uint32_t *in, const uint8_t *keep, uint64_t n, uint32_t *out) {
  uint32_t *flags = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t *incl = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  uint32_t *target = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  uint8_t *moved = (uint8_t *)malloc(n ? n : 1);
  for (uint64_t i = 0; i < n; i++) flags[i] = keep[i] ? 1u : 0u;
  orc_inclusive_scan_u32(flags, n, incl);
  for (uint64_t i = 0; i < n; i++) {
    uint32_t p_prev = i ? incl[i - 1] : 0u;
    target[i] = p_prev;
    moved[i] = incl[i] != p_prev;
  }
  memset(out, 0, n * sizeof(uint32_t));
  orc_shuffle_move_u32(in, target, moved, n, out);
  uint64_t size = n ? incl[n - 1] : 0;
  free(flags); free(incl); free(target); free(moved);
  return size;
}
