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

// A warp's 512 consecutive elements as four rows of 128: lane l holds elements [128 r + 4 l, + 4) of row r, so that every load
// instruction of the warp reads one contiguous run (512 B of values, 128 B of flags) — sixteen consecutive elements per thread
// would make each instruction touch thirty-two separate 64 B segments.  mask bit 4 r + c: element c of row r is kept.
template <bool FAST>
__device__ __forceinline__ void load_rows(const uint32_t *__restrict__ in, const uint8_t *__restrict__ keep, uint64_t n, uint64_t warp_base,
                                          uint32_t lane, uint32_t (&vals)[ITEMS], uint32_t &mask) {
  static_assert(ITEMS == 16, "four rows of four elements per lane");
  mask = 0;
  if (FAST && warp_base + 512 <= n) {
    uint4 v[4];
    uint32_t f[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const uint64_t idx = warp_base + 128u * r + 4u * lane;
      v[r] = __ldg(reinterpret_cast<const uint4 *>(in + idx));
      f[r] = __ldg(reinterpret_cast<const uint32_t *>(keep + idx));
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      vals[4 * r] = v[r].x; vals[4 * r + 1] = v[r].y; vals[4 * r + 2] = v[r].z; vals[4 * r + 3] = v[r].w;
#pragma unroll
      for (int c = 0; c < 4; ++c) mask |= (((f[r] >> (8 * c)) & 0xFFu) != 0 ? 1u : 0u) << (4 * r + c);
    }
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint64_t idx = warp_base + 128u * r + 4u * lane + c;
        const bool inside = idx < n;
        vals[4 * r + c] = inside ? in[idx] : 0u;
        if (inside && keep[idx] != 0) mask |= 1u << (4 * r + c);
      }
  }
}

// Ranks of a warp's kept elements in element order.  The four per-row counts of a lane travel through ONE shuffle scan, a byte each
// (a row of 128 elements cannot overflow its byte).  Returns the warp's total; first[r] = rank of the lane's first kept element of
// row r within the warp.
__device__ __forceinline__ uint32_t rank_rows(uint32_t mask, uint32_t lane, uint32_t (&first)[4]) {
  const uint32_t packed = __popc(mask & 0xFu) | (__popc(mask & 0xF0u) << 8) | (__popc(mask & 0xF00u) << 16) | (__popc(mask & 0xF000u) << 24);
  uint32_t incl = packed;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t up = __shfl_up_sync(FULL_MASK, incl, off);
    if (lane >= static_cast<uint32_t>(off)) incl += up;
  }
  const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
  const uint32_t excl = incl - packed;
  const uint32_t t0 = total & 0xFFu, t1 = (total >> 8) & 0xFFu, t2 = (total >> 16) & 0xFFu, t3 = total >> 24;
  first[0] = excl & 0xFFu;
  first[1] = t0 + ((excl >> 8) & 0xFFu);
  first[2] = t0 + t1 + ((excl >> 16) & 0xFFu);
  first[3] = t0 + t1 + t2 + (excl >> 24);
  return t0 + t1 + t2 + t3;
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
  // ---- the warp's 512 elements: 16 values and 16 keep flags per thread, every load in flight before anything waits
  uint32_t mask;
  uint32_t vals[ITEMS];
  load_rows<FAST>(in, keep, n, tile * TILE + 512ull * warp, lane, vals, mask);
  uint32_t first[4];
  const uint32_t warp_total = rank_rows(mask, lane, first);
  if (lane == 31) s_warp_total[warp] = warp_total;
  __syncthreads();
  uint32_t warp_offset = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < CB / 32; ++w) {
    const uint32_t t = s_warp_total[w];
    if (w < static_cast<int>(warp)) warp_offset += t;
    block_total += t;
  }

  // ---- the tile's kept values, packed in order (block-local offsets; overlaps with the look-back below for warps 1..7)
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    uint32_t local = warp_offset + first[r];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (mask & (1u << (4 * r + c))) s_vals[local++] = vals[4 * r + c];
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
constexpr int COUNT_TILES = 2;  // tiles per CTA of the counting pass: 32 B of flags per thread in flight
constexpr int GROUP = 64;       // tiles per group: the scan runs over group sums, a tile adds up the counts of its group before it
__global__ void __launch_bounds__(CB) k_compact_count(const uint8_t *__restrict__ keep, uint64_t n, unsigned long long *__restrict__ counts,
                                                      unsigned long long *__restrict__ groups, uint64_t n_tiles) {
  __shared__ uint32_t s_warp_total[COUNT_TILES][CB / 32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint64_t tile0 = static_cast<uint64_t>(blockIdx.x) * COUNT_TILES;
  const bool aligned = (reinterpret_cast<uintptr_t>(keep) & 15u) == 0;
  uint4 k16[COUNT_TILES];
  bool fast[COUNT_TILES];
#pragma unroll
  for (int t = 0; t < COUNT_TILES; ++t) {
    const uint64_t idx0 = (tile0 + t) * TILE + static_cast<uint64_t>(tid) * ITEMS;
    fast[t] = aligned && idx0 + ITEMS <= n;
    k16[t] = fast[t] ? __ldg(reinterpret_cast<const uint4 *>(keep + idx0)) : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int t = 0; t < COUNT_TILES; ++t) {
    const uint64_t idx0 = (tile0 + t) * TILE + static_cast<uint64_t>(tid) * ITEMS;
    uint32_t count = 0;
    if (fast[t]) {
      const uint32_t kw[4] = {k16[t].x, k16[t].y, k16[t].z, k16[t].w};
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
    if (lane == 0) s_warp_total[t][warp] = count;
  }
  __syncthreads();
  if (tid < COUNT_TILES && tile0 + tid < n_tiles) {
    uint32_t total = 0;
#pragma unroll
    for (int w = 0; w < CB / 32; ++w) total += s_warp_total[tid][w];
    counts[tile0 + tid] = total;
    atomicAdd(groups + (tile0 + tid) / GROUP, static_cast<unsigned long long>(total));
  }
}

// exclusive scan of the group sums in place (one CTA, one value per thread and round: 507 groups for 132 M elements), total -> *out_n
__global__ void __launch_bounds__(1024) k_compact_scan_groups(unsigned long long *__restrict__ groups, uint64_t n_groups, uint64_t *__restrict__ out_n) {
  __shared__ unsigned long long s_warp[32];
  __shared__ unsigned long long s_carry;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  if (tid == 0) s_carry = 0ull;
  __syncthreads();
  for (uint64_t base = 0; base < n_groups; base += 1024) {
    const uint64_t i = base + tid;
    const unsigned long long v = i < n_groups ? groups[i] : 0ull;
    unsigned long long incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long up = __shfl_up_sync(FULL_MASK, incl, off);
      if (lane >= static_cast<uint32_t>(off)) incl += up;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long w = lane < warp ? s_warp[lane] : 0ull;  // (32 warps: one value per lane)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(FULL_MASK, w, off);
    const unsigned long long carry = s_carry;
    if (i < n_groups) groups[i] = carry + w + (incl - v);
    __syncthreads();
    if (tid == 1023) s_carry = carry + w + incl;
    __syncthreads();
  }
  if (tid == 0) *out_n = s_carry;
}

// the tile's kept values to out[offsets[tile] ...): the body of k_compact_u32 without the look-back.  The tile also writes its share
// of the zeros behind the kept values — as many as it dropped, at total + (elements dropped by the tiles before it) — so the tail
// needs no kernel of its own.
template <bool FAST>
__global__ void __launch_bounds__(CB, RDN_COMPACT_MINB) k_compact_scatter(const uint32_t *__restrict__ in, const uint8_t *__restrict__ keep, uint64_t n,
                                                                          uint32_t *__restrict__ out, const unsigned long long *__restrict__ counts,
                                                                          const unsigned long long *__restrict__ group_prefix,
                                                                          const uint64_t *__restrict__ total_kept) {
  __shared__ uint32_t s_warp_total[CB / 32];
  __shared__ uint32_t s_vals[TILE];
  __shared__ unsigned long long s_base;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint64_t tile = blockIdx.x;
  // the tile's offset: the prefix of its group + the counts of the tiles of the group before it (the last warp adds them up while
  // the values are on their way)
  unsigned long long before = 0;
  if (warp == CB / 32 - 1) {
    const uint64_t g0 = tile / GROUP * GROUP;
    static_assert(GROUP == 64, "two counts per lane");
    if (g0 + lane < tile) before += counts[g0 + lane];
    if (g0 + 32 + lane < tile) before += counts[g0 + 32 + lane];
    if (lane == 0) before += group_prefix[tile / GROUP];
  }
  uint32_t mask;
  uint32_t vals[ITEMS];
  load_rows<FAST>(in, keep, n, tile * TILE + 512ull * warp, lane, vals, mask);
  const uint64_t total = *total_kept;
  if (warp == CB / 32 - 1) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) before += __shfl_xor_sync(FULL_MASK, before, off);
    if (lane == 0) s_base = before;
  }
  uint32_t first[4];
  const uint32_t warp_total = rank_rows(mask, lane, first);
  if (lane == 31) s_warp_total[warp] = warp_total;
  __syncthreads();
  const uint64_t base = s_base;
  uint32_t warp_offset = 0, block_total = 0;
#pragma unroll
  for (int w = 0; w < CB / 32; ++w) {
    const uint32_t t = s_warp_total[w];
    if (w < static_cast<int>(warp)) warp_offset += t;
    block_total += t;
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    uint32_t local = warp_offset + first[r];
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (mask & (1u << (4 * r + c))) s_vals[local++] = vals[4 * r + c];
  }
  __syncthreads();
  for (uint32_t i = tid; i < block_total; i += CB) out[base + i] = s_vals[i];
  const uint64_t tile_first = tile * TILE;
  const uint32_t tile_elems = static_cast<uint32_t>(n - tile_first < TILE ? n - tile_first : TILE);
  const uint64_t zero_at = total + (tile_first - base);
  for (uint32_t i = tid; i < tile_elems - block_total; i += CB) out[zero_at + i] = 0u;
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

uint64_t compact_status_words(uint64_t n) {  // tile counter, one word per tile, one per group of tiles
  const uint64_t n_tiles = (n + TILE - 1) / TILE;
  return 2 + n_tiles + (n_tiles + GROUP - 1) / GROUP + 1;
}

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
  if (n >= streaming_min) {  // count -> scan of the tile counts -> scatter: no tile waits for another 
    unsigned long long *counts = d_status + 1, *groups = counts + n_tiles;   // (the group sums start at zero: the memset above)
    const uint64_t n_groups = (n_tiles + GROUP - 1) / GROUP;
    k_compact_count<<<static_cast<unsigned>((n_tiles + COUNT_TILES - 1) / COUNT_TILES), CB, 0, stream>>>(d_keep, n, counts, groups, n_tiles);
    k_compact_scan_groups<<<1, 1024, 0, stream>>>(groups, n_groups, d_out_n);
    if (vec_keep)
      k_compact_scatter<true><<<static_cast<unsigned>(n_tiles), CB, 0, stream>>>(d_in, d_keep, n, d_out, counts, groups, d_out_n);
    else
      k_compact_scatter<false><<<static_cast<unsigned>(n_tiles), CB, 0, stream>>>(d_in, d_keep, n, d_out, counts, groups, d_out_n);
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
