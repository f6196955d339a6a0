/* ORACLE — TEST INFRASTRUCTURE ONLY.  Brute-force mesh picking: the nearest / all hits of a ray over
 * every primitive of a point / line / triangle mesh, with the local-space tolerance of points and lines.
 *
 * Follows /root/reference:
 *   content/mesh/core/src/feature/intersection.rs:3-37            ray_intersect_iter / _all / _nearest (enumerate, strict `<`)
 *   content/mesh/core/src/container/attributes/picking.rs:4-27    MeshBufferIntersectConfig, dispatch on the primitive kind
 *   content/mesh/core/src/container/attributes/access.rs:142-150,199-243   primitive_count, primitive_at (step / stride)
 *   content/mesh/core/src/primitive.rs:108-148                    MeshPrimitiveTopology: stride and step
 *   math/geometry/src/dimension3/intersection.rs:79-121           Ray3 x LineSegment / Point with tolerance
 *   math/geometry/src/dimension3/ray3.rs:48-145                   distance_sq_to_segment (GTE DistRaySegment)
 *   math/geometry/src/dimension3/intersection.rs:3-77             Ray3 x Triangle (orc_ray_triangle_a)
 */
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include "oracle.h"

/* out = (position of the hit, distance); returns 0 = none */
int orc_ray_segment(const float *r, ov3 v0, ov3 v1, float tolerance, float *out) {
  ov3 origin = ov3_new(r[0], r[1], r[2]);
  ov3 direction = ov3_new(r[3], r[4], r[5]);
  ov3 seg_center = ov3_scale(ov3_add(v0, v1), 0.5f);
  ov3 seg_dir = ov3_normalize(ov3_sub(v1, v0));
  ov3 diff = ov3_sub(origin, seg_center);
  float seg_length = ov3_length(ov3_sub(v0, v1)) * 0.5f;
  float a01 = -ov3_dot(direction, seg_dir);
  float b0 = ov3_dot(diff, direction);
  float b1 = -ov3_dot(diff, seg_dir);
  float c = ov3_dot(diff, diff);
  float det = fabsf(1.0f - a01 * a01);
  float s0 = 0.0f, s1 = 0.0f, sq_dist;
  if (det > 0.0f) {
    s0 = a01 * b1 - b0;
    s1 = a01 * b0 - b1;
    float ext_det = seg_length * det;
    if (s0 >= 0.0f) {
      if (s1 >= -ext_det) {
        if (s1 <= ext_det) {                       /* region 0 */
          float inv_det = 1.0f / det;
          s0 *= inv_det;
          s1 *= inv_det;
          sq_dist = s0 * (s0 + a01 * s1 + 2.0f * b0) + s1 * (a01 * s0 + s1 + 2.0f * b1) + c;
        } else {                                   /* region 1 */
          s1 = seg_length;
          s0 = fmaxf(0.0f, -(a01 * s1 + b0));
          sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
        }
      } else {                                     /* region 5 */
        s1 = -seg_length;
        s0 = fmaxf(0.0f, -(a01 * s1 + b0));
        sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
      }
    } else if (s1 <= -ext_det) {                   /* region 4 */
      s0 = fmaxf(0.0f, -(-a01 * seg_length + b0));
      s1 = s0 > 0.0f ? -seg_length : fminf(fmaxf(-seg_length, -b1), seg_length);
      sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
    } else if (s1 <= ext_det) {                    /* region 3 */
      s0 = 0.0f;
      s1 = fminf(fmaxf(-seg_length, -b1), seg_length);
      sq_dist = s1 * (s1 + 2.0f * b1) + c;
    } else {                                       /* region 2 */
      s0 = fmaxf(0.0f, -(a01 * seg_length + b0));
      s1 = s0 > 0.0f ? seg_length : fminf(fmaxf(-seg_length, -b1), seg_length);
      sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
    }
  } else {                                         /* parallel */
    s1 = a01 > 0.0f ? -seg_length : seg_length;
    s0 = fmaxf(0.0f, -(a01 * s1 + b0));
    sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
  }
  if (sq_dist > tolerance * tolerance) return 0;
  ov3 inter_ray = ov3_add(ov3_scale(direction, s0), origin);
  out[0] = inter_ray.x; out[1] = inter_ray.y; out[2] = inter_ray.z;
  out[3] = ov3_length(ov3_sub(origin, inter_ray));
  return 1;
}

int orc_ray_point(const float *r, ov3 point, float tolerance, float *out) {
  ov3 origin = ov3_new(r[0], r[1], r[2]);
  ov3 direction = ov3_new(r[3], r[4], r[5]);
  ov3 oc = ov3_sub(point, origin);
  float tca = ov3_dot(oc, direction);
  if (tca < 0.0f) return 0;  /* behind the origin: no hit even when the origin is within tolerance */
  float dist_sq = ov3_dot(oc, oc) - tca * tca;
  if (dist_sq > tolerance * tolerance) return 0;
  out[0] = point.x; out[1] = point.y; out[2] = point.z;
  out[3] = ov3_length(ov3_sub(origin, point));
  return 1;
}

static const uint64_t TOPOLOGY_STRIDE[5] = {1, 2, 2, 3, 3}, TOPOLOGY_STEP[5] = {1, 2, 1, 3, 1};

uint64_t orc_pick_primitive_count(uint64_t n_positions, uint64_t n_indices, int has_indices, int topology) {
  uint64_t count = has_indices ? n_indices : n_positions;
  uint64_t step = TOPOLOGY_STEP[topology], stride = TOPOLOGY_STRIDE[topology];
  return count + step < stride ? 0 : (count + step - stride) / step;  /* (usize arithmetic would underflow / panic below) */
}

static inline ov3 vertex_at(const float *positions, const uint32_t *indices, uint64_t k) {
  uint64_t v = indices ? indices[k] : k;
  return ov3_new(positions[3 * v], positions[3 * v + 1], positions[3 * v + 2]);
}

static int pick_primitive(const float *positions, const uint32_t *indices, int topology, uint64_t prim, const float *r, float tolerance,
                          int face_side, float *pd) {
  uint64_t at = TOPOLOGY_STEP[topology] * prim;
  switch (topology) {
    case ORC_TOPOLOGY_POINT_LIST: return orc_ray_point(r, vertex_at(positions, indices, at), tolerance, pd);
    case ORC_TOPOLOGY_LINE_LIST:
    case ORC_TOPOLOGY_LINE_STRIP: return orc_ray_segment(r, vertex_at(positions, indices, at), vertex_at(positions, indices, at + 1), tolerance, pd);
    default: return orc_ray_triangle_a(r, vertex_at(positions, indices, at), vertex_at(positions, indices, at + 1), vertex_at(positions, indices, at + 2), face_side, pd);
  }
}

typedef struct {
  const float *positions; const uint32_t *indices; uint64_t n_prims; int topology; float tolerance; int face_side;
  const orc_ray *rays; orc_mesh_hit *out; uint64_t n; atomic_ullong *cursor;
} pick_job;

static void *pick_worker(void *p) {
  pick_job *j = (pick_job *)p;
  for (;;) {
    uint64_t begin = atomic_fetch_add(j->cursor, 16);
    if (begin >= j->n) break;
    uint64_t end = begin + 16 < j->n ? begin + 16 : j->n;
    for (uint64_t i = begin; i < end; i++) {
      const orc_ray *ray = &j->rays[i];
      float r[6] = {ray->ox, ray->oy, ray->oz, ray->dx, ray->dy, ray->dz};
      orc_mesh_hit *best = &j->out[i];
      memset(best, 0, sizeof(*best));
      for (uint64_t prim = 0; prim < j->n_prims; prim++) {
        float pd[4];
        if (!pick_primitive(j->positions, j->indices, j->topology, prim, r, j->tolerance, j->face_side, pd)) continue;
        if (!best->hit || pd[3] < best->distance) {  /* refresh_nearest: strict `<`, the first of equals stays */
          best->px = pd[0]; best->py = pd[1]; best->pz = pd[2]; best->distance = pd[3];
          best->primitive_index = (uint32_t)prim; best->hit = 1;
        }
      }
    }
  }
  return NULL;
}

void orc_pick_nearest(const float *positions, uint64_t n_positions, const uint32_t *indices, uint64_t n_indices, int topology,
                      float tolerance, int face_side, const orc_ray *rays, uint64_t n_rays, orc_mesh_hit *out, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  atomic_ullong cursor;
  atomic_init(&cursor, 0);
  pick_job proto = {positions, indices, orc_pick_primitive_count(n_positions, n_indices, indices != NULL, topology), topology, tolerance,
                    face_side, rays, out, n_rays, &cursor};
  pthread_t *th = (pthread_t *)calloc(n_threads, sizeof(pthread_t));
  for (int t = 1; t < n_threads; t++) pthread_create(&th[t], NULL, pick_worker, &proto);
  pick_worker(&proto);
  for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th);
}

/* ray_intersect_all for ONE ray: every hit in primitive order; returns the number of hits (only `capacity` are written) */
uint64_t orc_pick_all(const float *positions, uint64_t n_positions, const uint32_t *indices, uint64_t n_indices, int topology,
                      float tolerance, int face_side, const orc_ray *ray, orc_mesh_hit *out, uint64_t capacity) {
  float r[6] = {ray->ox, ray->oy, ray->oz, ray->dx, ray->dy, ray->dz};
  uint64_t n_prims = orc_pick_primitive_count(n_positions, n_indices, indices != NULL, topology), total = 0;
  for (uint64_t prim = 0; prim < n_prims; prim++) {
    float pd[4];
    if (!pick_primitive(positions, indices, topology, prim, r, tolerance, face_side, pd)) continue;
    if (out && total < capacity) {
      orc_mesh_hit *h = &out[total];
      memset(h, 0, sizeof(*h));
      h->px = pd[0]; h->py = pd[1]; h->pz = pd[2]; h->distance = pd[3]; h->primitive_index = (uint32_t)prim; h->hit = 1;
    }
    total++;
  }
  return total;
}
