// pick.cu — brute-force mesh picking on the device: the reference walks every primitive of a mesh
// for every pick ray on one CPU thread; here the primitives of a ray are spread over the whole GPU.
//   AbstractMeshIntersectionExt::ray_intersect_nearest / _all   content/mesh/core/src/feature/intersection.rs:3-37
//   MeshBufferIntersectConfig + dispatch on the primitive kind  content/mesh/core/src/container/attributes/picking.rs:4-27
//   primitive_count / primitive_at (step, stride)               content/mesh/core/src/container/attributes/access.rs:142-150,199-243
//   Ray3 x LineSegment / Point with tolerance                   math/geometry/src/dimension3/intersection.rs:79-121
//   Ray3::distance_sq_to_segment (GTE DistRaySegment)           math/geometry/src/dimension3/ray3.rs:48-145
//   Ray3 x Triangle (GTE form, FaceSide)                        math/geometry/src/dimension3/intersection.rs:3-77
// Arithmetic in the reference's order, compiled -fmad=false like everything else: distances and positions are bit-identical.
//
// nearest: every thread tests a strided share of the primitives and keeps the smallest (distance bits, primitive index) key —
// distances are >= 0, so their bit patterns order like the values and ties go to the smaller index, which is the reference's
// "first of equals stays" (strict `<` refresh in primitive order) — block reduction, one atomicMin per block and ray.  The
// reference's refresh never replaces a NaN distance, so a NaN hit wins iff it is the very first hit of the ray: the smallest
// hit index is tracked beside the key and a last kernel re-evaluates the winner and writes the record.
// all: per-primitive hit flags + records, stable compaction of the indices (compact.cu), gather.
#include <cuda_runtime.h>

#include "kernels.h"
#include "rdn_math.h"

namespace rdn {

namespace {

__device__ __forceinline__ Vec3 vertex_at(const float *__restrict__ positions, const uint32_t *__restrict__ indices, uint64_t k) {
  const uint64_t v = indices ? indices[k] : k;
  return Vec3{__ldg(positions + 3 * v), __ldg(positions + 3 * v + 1), __ldg(positions + 3 * v + 2)};
}

__device__ __forceinline__ bool ray_triangle_gte(Vec3 origin, Vec3 dir, Vec3 a, Vec3 b, Vec3 c, uint32_t face_side, Vec3 &pos, float &t) {
  if (face_side == RDN_FACE_BACK) { const Vec3 tmp = a; a = c; c = tmp; }
  const bool backface_culling = face_side != RDN_FACE_DOUBLE;
  const Vec3 edge1 = b - a, edge2 = c - a;
  const Vec3 normal = cross(edge1, edge2);
  float DdN = dot(dir, normal);
  float sign;
  if (DdN > 0.0f) {
    if (backface_culling) return false;
    sign = 1.0f;
  } else if (DdN < 0.0f) {
    sign = -1.0f;
    DdN = -DdN;
  } else {
    return false;
  }
  const Vec3 diff = origin - a;
  const float DdQxE2 = sign * dot(dir, cross(diff, edge2));
  if (DdQxE2 < 0.0f) return false;
  const float DdE1xQ = sign * dot(dir, cross(edge1, diff));
  if (DdE1xQ < 0.0f) return false;
  if (DdQxE2 + DdE1xQ > DdN) return false;
  const float QdN = -sign * dot(diff, normal);
  if (QdN < 0.0f) return false;
  t = QdN / DdN;
  pos = origin + dir * t;
  return true;
}

__device__ __forceinline__ bool ray_segment(Vec3 origin, Vec3 direction, Vec3 v0, Vec3 v1, float tolerance, Vec3 &pos, float &distance) {
  const Vec3 seg_center = (v0 + v1) * 0.5f;
  const Vec3 seg_dir = normalize(v1 - v0);
  const Vec3 diff = origin - seg_center;
  const float seg_length = length(v0 - v1) * 0.5f;
  const float a01 = -dot(direction, seg_dir);
  const float b0 = dot(diff, direction);
  const float b1 = -dot(diff, seg_dir);
  const float c = dot(diff, diff);
  const float det = fabsf(1.0f - a01 * a01);
  float s0 = 0.0f, s1 = 0.0f, sq_dist;
  if (det > 0.0f) {
    s0 = a01 * b1 - b0;
    s1 = a01 * b0 - b1;
    const float ext_det = seg_length * det;
    if (s0 >= 0.0f) {
      if (s1 >= -ext_det) {
        if (s1 <= ext_det) {  // region 0
          const float inv_det = 1.0f / det;
          s0 *= inv_det;
          s1 *= inv_det;
          sq_dist = s0 * (s0 + a01 * s1 + 2.0f * b0) + s1 * (a01 * s0 + s1 + 2.0f * b1) + c;
        } else {  // region 1
          s1 = seg_length;
          s0 = fmaxf(0.0f, -(a01 * s1 + b0));
          sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
        }
      } else {  // region 5
        s1 = -seg_length;
        s0 = fmaxf(0.0f, -(a01 * s1 + b0));
        sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
      }
    } else if (s1 <= -ext_det) {  // region 4
      s0 = fmaxf(0.0f, -(-a01 * seg_length + b0));
      s1 = s0 > 0.0f ? -seg_length : fminf(fmaxf(-seg_length, -b1), seg_length);
      sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
    } else if (s1 <= ext_det) {  // region 3
      s0 = 0.0f;
      s1 = fminf(fmaxf(-seg_length, -b1), seg_length);
      sq_dist = s1 * (s1 + 2.0f * b1) + c;
    } else {  // region 2
      s0 = fmaxf(0.0f, -(a01 * seg_length + b0));
      s1 = s0 > 0.0f ? seg_length : fminf(fmaxf(-seg_length, -b1), seg_length);
      sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
    }
  } else {  // parallel
    s1 = a01 > 0.0f ? -seg_length : seg_length;
    s0 = fmaxf(0.0f, -(a01 * s1 + b0));
    sq_dist = -s0 * s0 + s1 * (s1 + 2.0f * b1) + c;
  }
  if (sq_dist > tolerance * tolerance) return false;
  pos = direction * s0 + origin;
  distance = length(origin - pos);
  return true;
}

__device__ __forceinline__ bool ray_point(Vec3 origin, Vec3 direction, Vec3 point, float tolerance, Vec3 &pos, float &distance) {
  const Vec3 oc = point - origin;
  const float tca = dot(oc, direction);
  if (tca < 0.0f) return false;
  const float dist_sq = dot(oc, oc) - tca * tca;
  if (dist_sq > tolerance * tolerance) return false;
  pos = point;
  distance = length(origin - point);
  return true;
}

__device__ __forceinline__ uint32_t topology_step(uint32_t topology) {
  return topology == RDN_TOPOLOGY_LINE_LIST ? 2u : (topology == RDN_TOPOLOGY_TRIANGLE_LIST ? 3u : 1u);
}

__device__ __forceinline__ bool pick_primitive(const PickMeshDev &M, uint64_t prim, Vec3 origin, Vec3 dir, float tolerance, uint32_t face_side,
                                               Vec3 &pos, float &distance) {
  const uint64_t at = static_cast<uint64_t>(topology_step(M.topology)) * prim;
  if (M.topology == RDN_TOPOLOGY_POINT_LIST) return ray_point(origin, dir, vertex_at(M.positions, M.indices, at), tolerance, pos, distance);
  if (M.topology == RDN_TOPOLOGY_LINE_LIST || M.topology == RDN_TOPOLOGY_LINE_STRIP)
    return ray_segment(origin, dir, vertex_at(M.positions, M.indices, at), vertex_at(M.positions, M.indices, at + 1), tolerance, pos, distance);
  return ray_triangle_gte(origin, dir, vertex_at(M.positions, M.indices, at), vertex_at(M.positions, M.indices, at + 1),
                          vertex_at(M.positions, M.indices, at + 2), face_side, pos, distance);
}

constexpr unsigned long long KEY_NONE = 0xFFFFFFFFFFFFFFFFull;

// grid = (chunks, rays): block (x, y) tests primitives x*blockDim.x + threadIdx.x, + gridDim.x*blockDim.x, ... of ray y
__global__ void __launch_bounds__(256) k_pick_nearest(const PickMeshDev M, const rdn_ray *__restrict__ rays, float tolerance, uint32_t face_side,
                                                      unsigned long long *__restrict__ best_key, uint32_t *__restrict__ first_hit) {
  const uint64_t ri = blockIdx.y;
  const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays + ri));
  const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays + ri) + 1);
  const Vec3 origin = {r0.x, r0.y, r0.z}, dir = {r1.x, r1.y, r1.z};
  unsigned long long key = KEY_NONE;
  uint32_t first = 0xFFFFFFFFu;
  for (uint64_t prim = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; prim < M.n_prims;
       prim += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    Vec3 pos;
    float distance;
    if (!pick_primitive(M, prim, origin, dir, tolerance, face_side, pos, distance)) continue;
    const uint32_t p = static_cast<uint32_t>(prim);
    first = p < first ? p : first;
    if (distance == distance) {
      const unsigned long long k = (static_cast<unsigned long long>(__float_as_uint(distance)) << 32) | p;
      key = k < key ? k : key;
    }
  }
  // warp then block reduction (min of keys, min of first-hit indices)
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const unsigned long long ok = __shfl_down_sync(0xFFFFFFFFu, key, off);
    const uint32_t of = __shfl_down_sync(0xFFFFFFFFu, first, off);
    key = ok < key ? ok : key;
    first = of < first ? of : first;
  }
  __shared__ unsigned long long s_key[8];
  __shared__ uint32_t s_first[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_key[warp] = key; s_first[warp] = first; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w) {
      key = s_key[w] < key ? s_key[w] : key;
      first = s_first[w] < first ? s_first[w] : first;
    }
    if (key != KEY_NONE) atomicMin(best_key + ri, key);
    if (first != 0xFFFFFFFFu) atomicMin(first_hit + ri, first);
  }
}

__global__ void __launch_bounds__(128) k_pick_finish(const PickMeshDev M, const rdn_ray *__restrict__ rays, uint64_t n, float tolerance,
                                                     uint32_t face_side, unsigned long long *__restrict__ best_key,
                                                     uint32_t *__restrict__ first_hit, rdn_mesh_hit *__restrict__ out) {
  const uint64_t ri = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x;
  if (ri >= n) return;
  const uint32_t first = first_hit[ri];
  const unsigned long long key = best_key[ri];
  best_key[ri] = KEY_NONE;       // re-armed for the next call
  first_hit[ri] = 0xFFFFFFFFu;
  rdn_mesh_hit h;
  h.px = h.py = h.pz = h.distance = 0.f;
  h.primitive_index = 0; h.hit = 0; h.pad0 = 0; h.pad1 = 0;
  if (first != 0xFFFFFFFFu) {
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays + ri));
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays + ri) + 1);
    const Vec3 origin = {r0.x, r0.y, r0.z}, dir = {r1.x, r1.y, r1.z};
    Vec3 pos;
    float distance;
    pick_primitive(M, first, origin, dir, tolerance, face_side, pos, distance);
    uint32_t winner = first;
    if (distance == distance && key != KEY_NONE) {  // the first hit is an ordinary one: the smallest key wins
      winner = static_cast<uint32_t>(key & 0xFFFFFFFFull);
      if (winner != first) pick_primitive(M, winner, origin, dir, tolerance, face_side, pos, distance);
    }
    h.px = pos.x; h.py = pos.y; h.pz = pos.z; h.distance = distance; h.primitive_index = winner; h.hit = 1;
  }
  out[ri] = h;
}

// ray_intersect_all of ONE ray: flag + record per primitive
__global__ void __launch_bounds__(256) k_pick_all_mark(const PickMeshDev M, const rdn_ray *__restrict__ ray, float tolerance, uint32_t face_side,
                                                       uint8_t *__restrict__ keep, uint32_t *__restrict__ iota, rdn_mesh_hit *__restrict__ records) {
  const float4 r0 = __ldg(reinterpret_cast<const float4 *>(ray));
  const float4 r1 = __ldg(reinterpret_cast<const float4 *>(ray) + 1);
  const Vec3 origin = {r0.x, r0.y, r0.z}, dir = {r1.x, r1.y, r1.z};
  for (uint64_t prim = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; prim < M.n_prims;
       prim += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    Vec3 pos;
    float distance;
    const bool hit = pick_primitive(M, prim, origin, dir, tolerance, face_side, pos, distance);
    keep[prim] = hit ? 1 : 0;
    iota[prim] = static_cast<uint32_t>(prim);
    if (hit) {
      rdn_mesh_hit h;
      h.px = pos.x; h.py = pos.y; h.pz = pos.z; h.distance = distance;
      h.primitive_index = static_cast<uint32_t>(prim); h.hit = 1; h.pad0 = 0; h.pad1 = 0;
      records[prim] = h;
    }
  }
}

__global__ void __launch_bounds__(256) k_pick_all_gather(const uint32_t *__restrict__ index, const uint64_t *__restrict__ n_kept, uint64_t capacity,
                                                         const rdn_mesh_hit *__restrict__ records, rdn_mesh_hit *__restrict__ out) {
  const uint64_t n = *n_kept < capacity ? *n_kept : capacity;
  for (uint64_t k = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; k < n; k += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[k] = records[index[k]];
}

}  // namespace

void launch_pick_nearest(const PickMeshDev &mesh, const rdn_ray *d_rays, uint64_t n_rays, float tolerance, uint32_t face_side,
                         unsigned long long *d_best_key, uint32_t *d_first_hit, rdn_mesh_hit *d_out, int sm_count, cudaStream_t stream) {
  if (n_rays == 0) return;
  if (mesh.n_prims) {
    // enough blocks per ray to fill the GPU even for one ray; at most 65535 rays per launch in grid.y
    uint64_t chunks = (mesh.n_prims + 255) / 256;
    const uint64_t want = std::max<uint64_t>(1, static_cast<uint64_t>(sm_count) * 8 / std::min<uint64_t>(n_rays, static_cast<uint64_t>(sm_count) * 8));
    if (chunks > want) chunks = want;
    for (uint64_t off = 0; off < n_rays; off += 65535) {
      const uint64_t m = std::min<uint64_t>(65535, n_rays - off);
      k_pick_nearest<<<dim3(static_cast<unsigned>(chunks), static_cast<unsigned>(m)), 256, 0, stream>>>(mesh, d_rays + off, tolerance, face_side,
                                                                                                           d_best_key + off, d_first_hit + off);
    }
  }
  k_pick_finish<<<static_cast<unsigned>((n_rays + 127) / 128), 128, 0, stream>>>(mesh, d_rays, n_rays, tolerance, face_side, d_best_key, d_first_hit, d_out);
}

void launch_pick_all_mark(const PickMeshDev &mesh, const rdn_ray *d_ray, float tolerance, uint32_t face_side, uint8_t *d_keep, uint32_t *d_iota,
                          rdn_mesh_hit *d_records, int sm_count, cudaStream_t stream) {
  if (mesh.n_prims == 0) return;
  const uint64_t blocks = std::min<uint64_t>((mesh.n_prims + 255) / 256, static_cast<uint64_t>(sm_count) * 8);
  k_pick_all_mark<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(mesh, d_ray, tolerance, face_side, d_keep, d_iota, d_records);
}

void launch_pick_all_gather(const uint32_t *d_index, const uint64_t *d_n_kept, uint64_t capacity, const rdn_mesh_hit *d_records, rdn_mesh_hit *d_out,
                            int sm_count, cudaStream_t stream) {
  if (capacity == 0) return;
  const uint64_t blocks = std::min<uint64_t>((capacity + 255) / 256, static_cast<uint64_t>(sm_count) * 8);
  k_pick_all_gather<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(d_index, d_n_kept, capacity, d_records, d_out);
}

}  // namespace rdn
