// compact.cu — (1) stable stream compaction of the wavefront's active-index list, (2) the path-A space query.
//
// (1) replaces the reference's five-pass pipeline (use_stream_compaction,
//     shader/parallel-compute/src/stream_compaction.rs:3-45: Kogge-Stone workgroup scan prefix_scan.rs:64-101,
//     block-sum scan + add-back lib.rs:398-429, shuffle_move scatter shuffle_move.rs:27-45, 1-thread size
//     write-back task-graph/src/runtime/task_group.rs:259-277) with ONE pass: per-thread popcount of 8 keep flags,
//     warp-aggregated shuffle scan, block scan through shared memory, decoupled look-back across tiles, scatter.
//     Contract kept: order preserving, *out_n = tail of the inclusive scan, out[] zero past *out_n.  Unlike the
//     reference (valid only for n <= W^2, lib.rs:407-408) any n works.
// (2) intersect_nearest_bvh (content/mesh/core/src/feature/bvh.rs:57-86): one thread per ray walks the FlattenBVH
//     exactly as traverse_by_branch_leaf does (utility/abstract-tree/src/lib.rs:33-51: explicit stack, right child
//     first, leaves not box tested, no distance pruning) with the reference's Ray3 x Box3 / Ray3 x Triangle
//     arithmetic (math/geometry/src/dimension3/intersection.rs:3-77,128-207); strict `<` keeps the first of equal t.
//     Compiled with -fmad=false.
#include <cuda_runtime.h>

#include <cstdlib>

#include "kernels.h"
#include "rdn_math.h"

namespace rdn {

namespace {

constexpr uint32_t FULL_MASK = 0xFFFFFFFFu;
#ifndef RDN_COMPACT_CB
#define RDN_COMPACT_CB 256
#endif
#ifndef RDN_COMPACT_MINB
#define RDN_COMPACT_MINB 8   // 8 CTAs per SM (32 registers): +30 % over 5 (profiles/compact_r2o_configuration_sweep.log)
#endif
constexpr int CB = RDN_COMPACT_CB;        // threads per CTA
constexpr int ITEMS = 16;      // consecutive items per thread: 64 B of values + 16 B of flags per thread in flight before the look-back
constexpr int TILE = CB * ITEMS;
constexpr unsigned long long ST_AGGREGATE = 1ull << 62, ST_PREFIX = 2ull << 62, ST_FLAG_MASK = 3ull << 62;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  return *reinterpret_cast<const volatile unsigned long long *>(p);
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  *reinterpret_cast<volatile unsigned long long *>(p) = v;
}

// status[0] = dynamic tile counter, status[1 + tile] = (flag << 62) | value
// One pass.  Per tile of 4096 elements: (1) every thread loads its 16 flags and 16 values — all loads of the tile are in flight
// before anything waits; (2) warp shuffle scan + block scan of the kept counts; (3) the kept values are packed into shared memory
// in order (local offsets only), which overlaps with (4) warp 0's decoupled look-back for the tile's global offset; (5) the tile's
// kept values leave shared memory as one contiguous, coalesced run.  FAST: 16-byte aligned inputs (vector loads).
template <bool FAST>
__global__ void __launch_bounds__(CB, RDN_COMPACT_MINB) k_compact_u32(const uint32_t *__restrict__ in, const uint8_t *__restrict__ keep, uint64_t n,
                                                    uint32_t *__restrict__ out, uint64_t *__restrict__ out_n,
                                                    unsigned long long *__restrict__ status, uint64_t n_tiles) {
  __shared__ unsigned long long s_tile, s_prefix;
  __shared__ uint32_t s_warp_total[CB / 32];
  __shared__ uint32_t s_vals[TILE];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

#ifdef RDN_COMPACT_STATIC_TILES
  const uint64_t tile = blockIdx.x;
#else
  if (tid == 0) s_tile = atomicAdd(status, 1ull);  // tiles are claimed in launch order: look-back never waits on an unscheduled CTA
  __syncthreads();
  const uint64_t tile = s_tile;
#endif
  const uint64_t idx0 = tile * TILE + static_cast<uint64_t>(tid) * ITEMS;

  // ---- 16 keep flags -> bit mask, 16 values -> registers
  uint32_t mask = 0;
  uint32_t vals[ITEMS];
  if (FAST && idx0 + ITEMS <= n) {
    const uint4 k16 = __ldg(reinterpret_cast<const uint4 *>(keep + idx0));
    const uint4 *vp = reinterpret_cast<const uint4 *>(in + idx0);
    const uint4 v0 = __ldg(vp), v1 = __ldg(vp + 1), v2 = __ldg(vp + 2), v3 = __ldg(vp + 3);
    const uint32_t kw[4] = {k16.x, k16.y, k16.z, k16.w};
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int j = 0; j < 4; ++j) mask |= (((kw[w] >> (8 * j)) & 0xFFu) != 0 ? 1u : 0u) << (4 * w + j);
    vals[0] = v0.x; vals[1] = v0.y; vals[2] = v0.z; vals[3] = v0.w; vals[4] = v1.x; vals[5] = v1.y; vals[6] = v1.z; vals[7] = v1.w;
    vals[8] = v2.x; vals[9] = v2.y; vals[10] = v2.z; vals[11] = v2.w; vals[12] = v3.x; vals[13] = v3.y; vals[14] = v3.z; vals[15] = v3.w;
  } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const bool inside = idx0 + j < n;
      vals[j] = inside ? in[idx0 + j] : 0u;
      if (inside && keep[idx0 + j] != 0) mask |= 1u << j;
    }
  }
  const uint32_t count = __popc(mask);

  // ---- warp-aggregated inclusive scan of the per-thread counts
  uint32_t incl = count;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t up = __shfl_up_sync(FULL_MASK, incl, off);
    if (lane >= static_cast<uint32_t>(off)) incl += up;
  }
  if (lane == 31) s_warp_total[warp] = incl;
  __syncthreads();
  uint32_t warp_offset = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < CB / 32; ++w) {
    const uint32_t t = s_warp_total[w];
    if (w < static_cast<int>(warp)) warp_offset += t;
    block_total += t;
  }

  // ---- the tile's kept values, packed in order (block-local offsets; overlaps with the look-back below for warps 1..7)
  {
    uint32_t local = warp_offset + (incl - count);
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if (mask & (1u << j)) s_vals[local++] = vals[j];
  }

  // ---- decoupled look-back by warp 0: 32 predecessors per probe
  if (warp == 0) {
    unsigned long long exclusive = 0;
    if (tile == 0) {
      if (lane == 0) st_status(status + 1, ST_PREFIX | block_total);
    } else {
      if (lane == 0) st_status(status + 1 + tile, ST_AGGREGATE | block_total);
      long long look = static_cast<long long>(tile) - 1;
      for (;;) {
        const long long t = look - static_cast<long long>(lane);
        unsigned long long s = t >= 0 ? ld_status(status + 1 + t) : ST_PREFIX;
        while (__any_sync(FULL_MASK, (s & ST_FLAG_MASK) == 0)) {
          if ((s & ST_FLAG_MASK) == 0) s = ld_status(status + 1 + t);
        }
        const uint32_t has_prefix = __ballot_sync(FULL_MASK, (s & ST_FLAG_MASK) == ST_PREFIX);
        const int first = has_prefix ? __ffs(has_prefix) - 1 : 31;
        unsigned long long v = static_cast<int>(lane) <= first ? (s & ~ST_FLAG_MASK) : 0ull;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(FULL_MASK, v, off);
        exclusive += __shfl_sync(FULL_MASK, v, 0);
        if (has_prefix) break;
        look -= 32;
      }
      if (lane == 0) st_status(status + 1 + tile, ST_PREFIX | (exclusive + block_total));
    }
    if (lane == 0) {
      s_prefix = exclusive;
      if (tile == n_tiles - 1) *out_n = exclusive + block_total;
    }
  }
  __syncthreads();

  // ---- one contiguous run per tile (order preserving)
  const uint64_t base = s_prefix;
  for (uint32_t i = tid; i < block_total; i += CB) out[base + i] = s_vals[i];
}

// ---- large inputs: three streaming passes instead of one pass with a look-back chain.  The single-pass kernel is bound by the
// latency of that chain (every tile waits for the running prefix of the tiles before it: 2.9 TB/s at best on 132 M elements);
// counting first costs one more read of the flags (10 B per element instead of 9) and no tile ever waits for another.
__global__ void __launch_bounds__(CB) k_compact_count(const uint8_t *__restrict__ keep, uint64_t n, unsigned long long *__restrict__ counts) {
  __shared__ uint32_t s_warp_total[CB / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint64_t idx0 = static_cast<uint64_t>(blockIdx.x) * TILE + static_cast<uint64_t>(tid) * ITEMS;
  uint32_t count = 0;
  if (idx0 + ITEMS <= n && (reinterpret_cast<uintptr_t>(keep) & 15u) == 0) {
    const uint4 k16 = __ldg(reinterpret_cast<const uint4 *>(keep + idx0));
    const uint32_t kw[4] = {k16.x, k16.y, k16.z, k16.w};
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int j = 0; j < 4; ++j) count += ((kw[w] >> (8 * j)) & 0xFFu) != 0 ? 1u : 0u;
  } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j)
      if (idx0 + j < n && keep[idx0 + j] != 0) ++count;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) count += __shfl_down_sync(FULL_MASK, count, off);
  if (lane == 0) s_warp_total[warp] = count;
  __syncthreads();
  if (tid == 0) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < CB / 32; ++w) total += s_warp_total[w];
    counts[blockIdx.x] = total;
  }
}

// exclusive scan of the per-tile counts in place (one CTA: 32 K tiles for 132 M elements), total -> *out_n
__global__ void __launch_bounds__(1024) k_compact_scan_tiles(unsigned long long *__restrict__ counts, uint64_t n_tiles, uint64_t *__restrict__ out_n) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  constexpr int PER = 8;
  if (tid == 0) s_carry = 0ull;
  __syncthreads();
  for (uint64_t base = 0; base < n_tiles; base += 1024ull * PER) {
    unsigned long long v[PER], sum = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const uint64_t i = base + static_cast<uint64_t>(tid) * PER + j;
      v[j] = i < n_tiles ? counts[i] : 0ull;
      sum += v[j];
    }
    unsigned long long incl = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long up = __shfl_up_sync(FULL_MASK, incl, off);
      if (lane >= static_cast<uint32_t>(off)) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_offset = 0;
    for (uint32_t w = 0; w < warp; ++w) warp_offset += s_warp[w];
    unsigned long long run = s_carry + warp_offset + (incl - sum);
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const uint64_t i = base + static_cast<uint64_t>(tid) * PER + j;
      if (i < n_tiles) counts[i] = run;
      run += v[j];
    }
    __syncthreads();
    if (tid == 1023) s_carry = run;
    __syncthreads();
  }
  if (tid == 0) *out_n = s_carry;
}

// the tile's kept values to out[offsets[tile] ...): the body of k_compact_u32 without the look-back
template <bool FAST>
__global__ void __launch_bounds__(CB, RDN_COMPACT_MINB) k_compact_scatter(const uint32_t *__restrict__ in, const uint8_t *__restrict__ keep, uint64_t n,
                                                                          uint32_t *__restrict__ out, const unsigned long long *__restrict__ offsets) {
  __shared__ uint32_t s_warp_total[CB / 32];
  __shared__ uint32_t s_vals[TILE];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint64_t tile = blockIdx.x;
  const uint64_t idx0 = tile * TILE + static_cast<uint64_t>(tid) * ITEMS;
  uint32_t mask = 0;
  uint32_t vals[ITEMS];
  if (FAST && idx0 + ITEMS <= n) {
    const uint4 k16 = __ldg(reinterpret_cast<const uint4 *>(keep + idx0));
    const uint4 *vp = reinterpret_cast<const uint4 *>(in + idx0);
    const uint4 v0 = __ldg(vp), v1 = __ldg(vp + 1), v2 = __ldg(vp + 2), v3 = __ldg(vp + 3);
    const uint32_t kw[4] = {k16.x, k16.y, k16.z, k16.w};
#pragma unroll
    for (int w = 0; w < 4; ++w)
#pragma unroll
      for (int j = 0; j < 4; ++j) mask |= (((kw[w] >> (8 * j)) & 0xFFu) != 0 ? 1u : 0u) << (4 * w + j);
    vals[0] = v0.x; vals[1] = v0.y; vals[2] = v0.z; vals[3] = v0.w; vals[4] = v1.x; vals[5] = v1.y; vals[6] = v1.z; vals[7] = v1.w;
    vals[8] = v2.x; vals[9] = v2.y; vals[10] = v2.z; vals[11] = v2.w; vals[12] = v3.x; vals[13] = v3.y; vals[14] = v3.z; vals[15] = v3.w;
  } else {
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const bool inside = idx0 + j < n;
      vals[j] = inside ? in[idx0 + j] : 0u;
      if (inside && keep[idx0 + j] != 0) mask |= 1u << j;
    }
  }
  const uint32_t count = __popc(mask);
  uint32_t incl = count;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t up = __shfl_up_sync(FULL_MASK, incl, off);
    if (lane >= static_cast<uint32_t>(off)) incl += up;
  }
  if (lane == 31) s_warp_total[warp] = incl;
  __syncthreads();
  uint32_t warp_offset = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < CB / 32; ++w) {
    const uint32_t t = s_warp_total[w];
    if (w < static_cast<int>(warp)) warp_offset += t;
    block_total += t;
  }
  uint32_t local = warp_offset + (incl - count);
#pragma unroll
  for (int j = 0; j < ITEMS; ++j)
    if (mask & (1u << j)) s_vals[local++] = vals[j];
  __syncthreads();
  const uint64_t base = offsets[tile];
  for (uint32_t i = tid; i < block_total; i += CB) out[base + i] = s_vals[i];
}

__global__ void k_zero_tail_u32(uint32_t *__restrict__ out, const uint64_t *__restrict__ out_n, uint64_t n) {
  const uint64_t first = *out_n;
  for (uint64_t i = first + blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
    out[i] = 0u;
}

// ------------------------------------------------------------------------------------------------ path A
__device__ __forceinline__ bool ray_box_a(Vec3 origin, Vec3 dir, Vec3 bmin, Vec3 bmax) {
  float t_max, t_min, ty_min, ty_max, tz_min, tz_max;
  const float inv_dir_x = 1.0f / dir.x, inv_dir_y = 1.0f / dir.y, inv_dir_z = 1.0f / dir.z;
  if (inv_dir_x >= 0.0f) { t_min = (bmin.x - origin.x) * inv_dir_x; t_max = (bmax.x - origin.x) * inv_dir_x; }
  else { t_min = (bmax.x - origin.x) * inv_dir_x; t_max = (bmin.x - origin.x) * inv_dir_x; }
  if (inv_dir_y >= 0.0f) { ty_min = (bmin.y - origin.y) * inv_dir_y; ty_max = (bmax.y - origin.y) * inv_dir_y; }
  else { ty_min = (bmax.y - origin.y) * inv_dir_y; ty_max = (bmin.y - origin.y) * inv_dir_y; }
  if ((t_min > ty_max) || (ty_min > t_max)) return false;
  if (ty_min > t_min || isnan(t_min)) t_min = ty_min;
  if (ty_max < t_max || isnan(t_max)) t_max = ty_max;
  if (inv_dir_z >= 0.0f) { tz_min = (bmin.z - origin.z) * inv_dir_z; tz_max = (bmax.z - origin.z) * inv_dir_z; }
  else { tz_min = (bmax.z - origin.z) * inv_dir_z; tz_max = (bmin.z - origin.z) * inv_dir_z; }
  if ((t_min > tz_max) || (tz_min > t_max)) return false;
  if (tz_min > t_min || isnan(t_min)) t_min = tz_min;
  if (tz_max < t_max || isnan(t_max)) t_max = tz_max;
  if (t_max < 0.0f) return false;
  return true;
}

__device__ __forceinline__ bool ray_triangle_a(Vec3 origin, Vec3 dir, Vec3 a, Vec3 b, Vec3 c, uint32_t face_side, float &t) {
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
  return true;
}

__global__ void __launch_bounds__(128) k_patha_nearest(const PathANode *__restrict__ nodes, const PathATri *__restrict__ tris,
                                                       const rdn_ray *__restrict__ rays, uint64_t n, uint32_t face_side,
                                                       rdn_mesh_hit *__restrict__ out) {
  uint32_t stack[PATHA_MAX_DEPTH + 2];
  for (uint64_t ri = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; ri < n;
       ri += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays + ri));
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays + ri) + 1);
    const Vec3 origin = {r0.x, r0.y, r0.z}, dir = {r1.x, r1.y, r1.z};
    float best_t = 0.f;
    uint32_t best_prim = 0, have = 0;
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
      const uint32_t ni = stack[--sp];
      const float4 *np = reinterpret_cast<const float4 *>(nodes + ni);
      const float4 n0 = __ldg(np), n1 = __ldg(np + 1);
      const uint32_t a = __float_as_uint(n0.w), b = __float_as_uint(n1.w);
      if (b != 0xFFFFFFFFu) {
        for (uint32_t k = a; k < b; ++k) {  // slots in sorted_primitive_index order: one contiguous 48 B record each
          const float4 *tp = reinterpret_cast<const float4 *>(tris + k);
          const float4 qa = __ldg(tp), qb = __ldg(tp + 1), qc = __ldg(tp + 2);
          float t;
          if (ray_triangle_a(origin, dir, Vec3{qa.x, qa.y, qa.z}, Vec3{qb.x, qb.y, qb.z}, Vec3{qc.x, qc.y, qc.z}, face_side, t) &&
              (!have || t < best_t)) { best_t = t; best_prim = __float_as_uint(qa.w); have = 1; }
        }
      } else if (ray_box_a(origin, dir, Vec3{n0.x, n0.y, n0.z}, Vec3{n1.x, n1.y, n1.z})) {
        if (sp + 2 <= PATHA_MAX_DEPTH + 2) {
          stack[sp++] = ni + 1;  // left pushed first ...
          stack[sp++] = a;       // ... right popped first
        }
      }
    }
    const Vec3 p = origin + dir * best_t;  // HyperRay::at
    float4 *dst = reinterpret_cast<float4 *>(out + ri);
    dst[0] = have ? make_float4(p.x, p.y, p.z, best_t) : make_float4(0.f, 0.f, 0.f, 0.f);
    dst[1] = make_float4(__uint_as_float(have ? best_prim : 0u), __uint_as_float(have), 0.f, 0.f);
  }
}

// intersect_list_bvh (content/mesh/core/src/feature/bvh.rs:23-55): the same right-first DFS, but EVERY intersected primitive
// is reported, in visiting order.  FILL = false counts the hits of each ray; FILL = true writes them at offsets[ray] (exclusive
// scan of the counts), so the ragged result is a CSR list whose per-ray order is the reference's Vec order.
template <bool FILL>
__global__ void __launch_bounds__(128) k_patha_list(const PathANode *__restrict__ nodes, const PathATri *__restrict__ tris,
                                                    const rdn_ray *__restrict__ rays, uint64_t n, uint32_t face_side,
                                                    uint32_t *__restrict__ counts, const uint64_t *__restrict__ offsets,
                                                    rdn_mesh_hit *__restrict__ out) {
  uint32_t stack[PATHA_MAX_DEPTH + 2];
  for (uint64_t ri = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; ri < n;
       ri += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays + ri));
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays + ri) + 1);
    const Vec3 origin = {r0.x, r0.y, r0.z}, dir = {r1.x, r1.y, r1.z};
    uint32_t found = 0;
    uint64_t dst = FILL ? offsets[ri] : 0;
    int sp = 0;
    stack[sp++] = 0;
    while (sp > 0) {
      const uint32_t ni = stack[--sp];
      const float4 *np = reinterpret_cast<const float4 *>(nodes + ni);
      const float4 n0 = __ldg(np), n1 = __ldg(np + 1);
      const uint32_t a = __float_as_uint(n0.w), b = __float_as_uint(n1.w);
      if (b != 0xFFFFFFFFu) {
        for (uint32_t k = a; k < b; ++k) {
          const float4 *tp = reinterpret_cast<const float4 *>(tris + k);
          const float4 qa = __ldg(tp), qb = __ldg(tp + 1), qc = __ldg(tp + 2);
          float t;
          if (!ray_triangle_a(origin, dir, Vec3{qa.x, qa.y, qa.z}, Vec3{qb.x, qb.y, qb.z}, Vec3{qc.x, qc.y, qc.z}, face_side, t)) continue;
          if (FILL) {
            const Vec3 p = origin + dir * t;
            float4 *o = reinterpret_cast<float4 *>(out + dst);
            o[0] = make_float4(p.x, p.y, p.z, t);
            o[1] = make_float4(qa.w, __uint_as_float(1u), 0.f, 0.f);
            ++dst;
          }
          ++found;
        }
      } else if (ray_box_a(origin, dir, Vec3{n0.x, n0.y, n0.z}, Vec3{n1.x, n1.y, n1.z})) {
        if (sp + 2 <= PATHA_MAX_DEPTH + 2) {
          stack[sp++] = ni + 1;
          stack[sp++] = a;
        }
      }
    }
    if (!FILL) counts[ri] = found;
  }
}

}  // namespace

uint64_t compact_status_words(uint64_t n) { return 2 + (n + TILE - 1) / TILE; }

void launch_compact_u32(const uint32_t *d_in, const uint8_t *d_keep, uint64_t n, uint32_t *d_out, uint64_t *d_out_n,
                        unsigned long long *d_status, cudaStream_t stream) {
  const uint64_t n_tiles = (n + TILE - 1) / TILE;
  cudaMemsetAsync(d_status, 0, compact_status_words(n) * sizeof(unsigned long long), stream);
  if (n == 0) {
    cudaMemsetAsync(d_out_n, 0, sizeof(uint64_t), stream);
    return;
  }
  const bool vec_keep = ((reinterpret_cast<uintptr_t>(d_keep) | reinterpret_cast<uintptr_t>(d_in)) & 15u) == 0;
  const unsigned zb = static_cast<unsigned>(n_tiles < 1184 ? n_tiles : 1184);
  static const uint64_t streaming_min = []() { const char *e = getenv("RDN_COMPACT_STREAMING_MIN"); return e ? strtoull(e, nullptr, 10) : (1ull << 22); }();
  if (n >= streaming_min) {  // count -> scan of the tile counts -> scatter: no tile waits for another (the memset above is not needed here)
    unsigned long long *counts = d_status + 1;
    k_compact_count<<<static_cast<unsigned>(n_tiles), CB, 0, stream>>>(d_keep, n, counts);
    k_compact_scan_tiles<<<1, 1024, 0, stream>>>(counts, n_tiles, d_out_n);
    if (vec_keep)
      k_compact_scatter<true><<<static_cast<unsigned>(n_tiles), CB, 0, stream>>>(d_in, d_keep, n, d_out, counts);
    else
      k_compact_scatter<false><<<static_cast<unsigned>(n_tiles), CB, 0, stream>>>(d_in, d_keep, n, d_out, counts);
    k_zero_tail_u32<<<zb, 256, 0, stream>>>(d_out, d_out_n, n);
    return;
  }
  if (vec_keep)
    k_compact_u32<true><<<static_cast<unsigned>(n_tiles), CB, 0, stream>>>(d_in, d_keep, n, d_out, d_out_n, d_status, n_tiles);
  else
    k_compact_u32<false><<<static_cast<unsigned>(n_tiles), CB, 0, stream>>>(d_in, d_keep, n, d_out, d_out_n, d_status, n_tiles);
  k_zero_tail_u32<<<zb, 256, 0, stream>>>(d_out, d_out_n, n);
}

void launch_patha_nearest(const PathANode *d_nodes, const PathATri *d_tris, const rdn_ray *d_rays, uint64_t n, uint32_t face_side,
                          rdn_mesh_hit *d_out, cudaStream_t stream) {
  if (n == 0) return;
  uint64_t blocks = (n + 127) / 128;
  if (blocks > 148ull * 64) blocks = 148ull * 64;
  k_patha_nearest<<<static_cast<unsigned>(blocks), 128, 0, stream>>>(d_nodes, d_tris, d_rays, n, face_side, d_out);
}

void launch_patha_list(const PathANode *d_nodes, const PathATri *d_tris, const rdn_ray *d_rays, uint64_t n, uint32_t face_side,
                       uint32_t *d_counts, const uint64_t *d_offsets, rdn_mesh_hit *d_out, cudaStream_t stream) {
  if (n == 0) return;
  uint64_t blocks = (n + 127) / 128;
  if (blocks > 148ull * 64) blocks = 148ull * 64;
  if (d_out) k_patha_list<true><<<static_cast<unsigned>(blocks), 128, 0, stream>>>(d_nodes, d_tris, d_rays, n, face_side, nullptr, d_offsets, d_out);
  else k_patha_list<false><<<static_cast<unsigned>(blocks), 128, 0, stream>>>(d_nodes, d_tris, d_rays, n, face_side, d_counts, nullptr, nullptr);
}

}  // namespace rdn
