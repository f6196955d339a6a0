// traverse.cu — BVH closest-hit traversal kernels for sm_100a.
//
// Compiled with -fmad=false (no FMA contraction) and IEEE div/sqrt: every ray/box and ray/triangle value is
// bit-identical to the reference's CPU arithmetic
//   intersect_ray_aabb_cpu      shader/ray-tracing/src/backend/wavefront_compute/geometry/mod.rs:44-64
//   intersect_ray_triangle_cpu  .../geometry/mod.rs:105-155
//   NaiveSahBvhCpu::traverse    .../geometry/naive/traverse_cpu.rs:52-319
//   TraverseFlags               .../geometry/naive/flag.rs:6-117
//
// Kernels:
//   k_trace_reference       the reference's stackless threaded pre-order walk, one ray per thread (a warp = an 8x4 pixel tile on
//                           grid launches).  Identical to the CPU query by construction (same visit order, same live-range
//                           pruning, "last accepted of equal t wins").  Used for ACCEPT_FIRST_HIT rays, for TLASes the
//                           flattener routes as irregular, and for the reference's visit counters; its walk
//                           (reference_walk) also serves every re-walk below.
//   k_trace_ordered_rounds  persistent-thread, near-child-first walk over the 64 B two-box nodes with a per-thread stack,
//                           256-bit node / triangle fetches through the read-only path, warp-synchronous rounds and whole-tile
//                           ray refill (ballot + one atomic per warp).  Box decisions use the reference's arithmetic on the
//                           reference's boxes; the pruning bound is inflated by TIE_EPS and any ray that saw a second
//                           candidate within TIE_EPS of the closest is queued (warp-aggregated append) and re-walked in the
//                           reference's order with its range clamped around the closest distance — so ids come out exactly
//                           as the CPU query's.  Rays that can reach an irregular instance (accel.cpp "regularity") are
//                           queued at refill instead of being traversed and re-walked over their whole range.  The queue is
//                           drained inside the kernel by warps that ran out of rays (last CTA sweeps the rest); consecutive
//                           launches on a stream overlap their tails (programmatic dependent launch, two scratch sets,
//                           epoch gate).  The round loop itself is ordered_rounds.inc.  Template switches select the shipped
//                           forms (tile history for grids, late work sharing for lists, any-hit stage) and the
//                           measured-and-rejected experiments kept for A/B runs (RDN_ORDERED_VARIANT, see
//                           launch_trace_ordered and DESIGN.md §5).
//   k_build_tile_lists      behind a grid launch that noted its pass durations: the tiles filed by class for the next launch.
//   k_resolve_ties          the queue walked by a separate kernel, one thread per entry: launches of a TLAS that lists
//                           irregular instances (their queue can be long), and RDN_ORDERED_VARIANT=9.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "rdn_math.h"

namespace rdn {

namespace {

constexpr uint32_t FULL_MASK = 0xFFFFFFFFu;
#ifndef RDN_STACK_MAX
#define RDN_STACK_MAX 120  // (the emulated test build of tests/simt lowers it to reach the overflow report)
#endif
constexpr int STACK_MAX = RDN_STACK_MAX;  // TLAS depth 50 + BLAS depth 50 (TreeBuildOption of naive/mod.rs:173-176,280-283) + bookkeeping
#ifndef RDN_ORDERED_BLOCK
#define RDN_ORDERED_BLOCK 128
#endif
#ifndef RDN_ORDERED_MINB
#define RDN_ORDERED_MINB 8
#endif
#ifndef RDN_ORDERED_K
#define RDN_ORDERED_K 3
#endif
constexpr int ORDERED_BLOCK = RDN_ORDERED_BLOCK;

// TraverseFlags bits (flag.rs:6-25)
constexpr uint32_t TF_FORCE_OPAQUE = 0x01, TF_FORCE_NON_OPAQUE = 0x02, TF_END_SEARCH = 0x04, TF_CULL_BACK = 0x10,
                   TF_CULL_FRONT = 0x20, TF_CULL_OPAQUE = 0x40, TF_CULL_NON_OPAQUE = 0x80, TF_SKIP_TRIANGLES = 0x100,
                   TF_FLIP_FACING = 0x400;

__device__ __forceinline__ uint32_t merge_geometry_instance_flag(uint32_t f, uint32_t gi) {
  if (gi & RDN_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE) f &= ~(TF_CULL_BACK | TF_CULL_FRONT);
  if (gi & RDN_GEOMETRY_INSTANCE_TRIANGLE_FLIP_FACING) f ^= TF_FLIP_FACING;
  if (gi & RDN_GEOMETRY_INSTANCE_FORCE_OPAQUE) f |= TF_FORCE_OPAQUE;
  if (gi & RDN_GEOMETRY_INSTANCE_FORCE_NO_OPAQUE) f |= TF_FORCE_NON_OPAQUE;
  return f;
}
// cull_geometry -> pass (is_opaque only selects the any-hit path, which is fixed to ACCEPT here)
__device__ __forceinline__ bool cull_geometry_pass(uint32_t f, uint32_t geometry_flags) {
  const bool geometry_opaque = (geometry_flags & RDN_GEOMETRY_FLAG_OPAQUE) != 0;
  const bool is_opaque = (geometry_opaque || (f & TF_FORCE_OPAQUE)) && !(f & TF_FORCE_NON_OPAQUE);
  return (is_opaque && !(f & TF_CULL_OPAQUE)) || (!is_opaque && !(f & TF_CULL_NON_OPAQUE));
}
// cull_triangle -> bit0 cull_enable, bit1 cull_back
__device__ __forceinline__ uint32_t cull_triangle_bits(uint32_t f) {
  const bool flip = (f & TF_FLIP_FACING) != 0, cull_front = (f & TF_CULL_FRONT) != 0, cull_back = (f & TF_CULL_BACK) != 0;
  const bool enable = cull_front || cull_back;
  const bool back = (flip && cull_back) || (!flip && cull_front);
  return (enable ? 1u : 0u) | (back ? 2u : 0u);
}

// is the geometry non-opaque for this ray (cull_geometry's is_opaque, flag.rs:62-78)?  Its candidates go through the any-hit stage.
__device__ __forceinline__ bool geometry_non_opaque(uint32_t f, uint32_t geometry_flags) {
  const bool geometry_opaque = (geometry_flags & RDN_GEOMETRY_FLAG_OPAQUE) != 0;
  return !((geometry_opaque || (f & TF_FORCE_OPAQUE)) && !(f & TF_FORCE_NON_OPAQUE));
}
// The any-hit stage for one candidate of non-opaque geometry (traverse_cpu.rs:164-176; shader selection trace_task.rs:189-203):
// RayAnyHitBehavior bits.  Out of line on purpose: opaque scenes never get here.
__device__ __noinline__ uint32_t any_hit_behavior(const SceneDev &S, const rdn_launch &L, uint32_t slot, uint32_t inst, float distance) {
  uint32_t program = RDN_SBT_NO_SHADER;
  const SlotInfo si = S.slot_info[slot];
  if (L.any_hit == RDN_ANYHIT_FROM_SBT) {
    const uint32_t group = L.sbt_ray_offset + L.sbt_ray_stride * si.geometry_idx + __ldg(&S.instances[inst].sbt_offset);
    if (group < S.n_sbt_hit_groups) program = __ldg(&S.sbt_hit_groups[group].any_hit);
  } else {
    program = L.any_hit - 1u;
  }
  if (program >= S.n_anyhit_programs) return RDN_ANYHIT_BEHAVIOR_ACCEPT_HIT;  // no shader: the candidate is accepted
  const rdn_anyhit_program p = S.anyhit_programs[program];
  bool holds = true;
  if (p.kind == RDN_ANYHIT_PRIMITIVE_MASK) holds = (si.primitive_id & p.mask) == p.value;
  else if (p.kind == RDN_ANYHIT_MIN_DISTANCE) holds = distance >= p.distance;
  return holds ? p.behavior : p.otherwise;
}

#if defined(RDN_DEBUG_STEPS) || defined(RDN_DEBUG_TIMELINE)
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
__device__ __forceinline__ Vec3 xyz(const float4 &q) { return Vec3{q.x, q.y, q.z}; }
__device__ __forceinline__ Vec3 recip3(Vec3 d) { return Vec3{1.0f / d.x, 1.0f / d.y, 1.0f / d.z}; }

// Region markers of the ordered kernel for the SIMT issue model of the emulated build (tests/simt, tools/issue_model.py): how often
// a warp issues each region, to be multiplied by the region's SASS instruction count.  They compile to nothing in the product.
#if defined(RDN_SIMT_EMU)
#define RDN_COST(region) ::simt::cost_mark(region)
#else
#define RDN_COST(region) ((void)0)
#endif
enum CostRegion {
  COST_PROLOGUE = 0, COST_REFILL, COST_RAY_LOAD, COST_OUTER, COST_ROUND, COST_NODE, COST_PHASE2, COST_LEAF, COST_TRI, COST_LEAF_END,
  COST_INSTANCE, COST_INSTANCE_ENTER, COST_EXIT_INSTANCE, COST_EMPTY, COST_GEOMETRY, COST_VOTE, COST_FINISH, COST_TIE, COST_EPILOGUE,
  COST_TRI_RANGE, COST_TRI_U, COST_TRI_V, COST_TRI_HIT, COST_INSTANCE_SKIP,  // inside a triangle iteration: past the facing test, past the range test, past u, accepted
  COST_SHARE, COST_MERGE,  // SHARE: deferred subtrees handed to idle lanes, partial results merged back
  COST_REGION_COUNT
};

// intersect_ray_aabb_cpu with inv_d = 1/d hoisted (same value every call)
__device__ __forceinline__ bool slab_test(Vec3 o, Vec3 inv_d, float t_min, float t_max, Vec3 bmin, Vec3 bmax, float &t_near_max) {
  const Vec3 t0 = (bmin - o) * inv_d;
  const Vec3 t1 = (bmax - o) * inv_d;
  const Vec3 t_near = vmin(t0, t1);
  const Vec3 t_far = vmax(t0, t1);
  t_near_max = fmaxf(fmaxf(t_near.x, t_near.y), t_near.z);
  const float t_far_min = fminf(fminf(t_far.x, t_far.y), t_far.z);
  return t_near_max <= t_far_min && t_min < t_far_min && t_near_max < t_max;
}

// intersect_ray_triangle_cpu with the ray-independent terms read from the TriRecord
__device__ __forceinline__ bool triangle_test(const float4 qn, const float4 qv0, const float4 qe1, const float4 qe2, Vec3 origin,
                                              Vec3 direction, float range_x, float range_y, uint32_t cull_bits, float &sign,
                                              float &t, float &u, float &v) {
  const Vec3 normal = xyz(qn), v0 = xyz(qv0), e1 = xyz(qe1), e2 = xyz(qe2);
  const float b = dot(normal, direction);
  sign = copysignf(1.0f, b);
  if (cull_bits & 1u) {
    const bool cull_back = (cull_bits & 2u) != 0;
    if (!(cull_back != (b < 0.0f))) return false;
  }
  RDN_COST(COST_TRI_RANGE);
  const Vec3 w0 = origin - v0;
  const float a = -dot(normal, w0);
  t = a / b;
  if (t < range_x || t > range_y) return false;
  RDN_COST(COST_TRI_U);
  const Vec3 p = origin + direction * t;
  const float uu = qv0.w, uv = qe1.w, vv = qe2.w, inverse_d = qn.w;
  const Vec3 w = p - v0;
  const float wu = dot(w, e1);
  const float wv = dot(w, e2);
  u = (uv * wv - vv * wu) * inverse_d;
  if (u < 0.0f || u > 1.0f) return false;
  RDN_COST(COST_TRI_V);
  v = (uv * wu - uu * wv) * inverse_d;
  if (v < 0.0f || (u + v) > 1.0f) return false;
  return true;
}

// world -> object ray (traverse_cpu.rs:99-104)
__device__ __forceinline__ void to_object_space(const InstanceRecord *rec, Vec3 ro, Vec3 rd, Vec3 &bo, Vec3 &bd, float &scaling) {
  const float4 *m = reinterpret_cast<const float4 *>(rec->transform_inv);
  const float4 ca = __ldg(m), cb = __ldg(m + 1), cc = __ldg(m + 2), cd = __ldg(m + 3);  // columns a,b,c,d
  const float x = ro.x * ca.x + ro.y * cb.x + ro.z * cc.x + 1.0f * cd.x;
  const float y = ro.x * ca.y + ro.y * cb.y + ro.z * cc.y + 1.0f * cd.y;
  const float z = ro.x * ca.z + ro.y * cb.z + ro.z * cc.z + 1.0f * cd.z;
  const float w = ro.x * ca.w + ro.y * cb.w + ro.z * cc.w + 1.0f * cd.w;
  bo = Vec3{x / w, y / w, z / w};
  const Vec3 d0 = Vec3{rd.x * ca.x + rd.y * cb.x + rd.z * cc.x, rd.x * ca.y + rd.y * cb.y + rd.z * cc.y,
                       rd.x * ca.z + rd.y * cb.z + rd.z * cc.z};
  scaling = length(d0);
  bd = normalize(d0);
}

__device__ __forceinline__ void store_hit(rdn_hit *dst, float t, float u, float v, uint32_t prim, uint32_t geom, uint32_t inst,
                                          uint32_t custom, uint32_t kind) {
  float4 *p = reinterpret_cast<float4 *>(dst);
  p[0] = make_float4(t, u, v, __uint_as_float(prim));
  p[1] = make_float4(__uint_as_float(geom), __uint_as_float(inst), __uint_as_float(custom), __uint_as_float(kind));
}

// 32 bytes in one instruction (Blackwell's 256-bit global loads / stores: LDG.E.256 / STG.E.256): a 64 B node or triangle record is
// two load instructions instead of four, a ray or a hit record one instead of two.  The address must be 32-byte aligned.
// (RDN_SIMT_EMU: the CPU emulation build of the test-suite, tests/simt/ — PTX-only operations get plain C++ stand-ins there)
#ifndef RDN_SIMT_EMU
__device__ __forceinline__ void ld_global_nc_256(const void *p, float4 &a, float4 &b) {
  asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
      : "l"(p));
}
__device__ __forceinline__ void st_global_256(void *p, float4 a, float4 b) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
               "f"(b.y), "f"(b.z), "f"(b.w)
               : "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
inline void ld_global_nc_256(const void *p, float4 &a, float4 &b) { a = static_cast<const float4 *>(p)[0]; b = static_cast<const float4 *>(p)[1]; }
inline void st_global_256(void *p, float4 a, float4 b) { static_cast<float4 *>(p)[0] = a; static_cast<float4 *>(p)[1] = b; }
inline void pdl_launch_dependents() {}
inline void pdl_wait() {}
#endif
template <bool WIDE>
__device__ __forceinline__ void load_pair(const void *p, float4 &a, float4 &b) {
  if (WIDE) {
    ld_global_nc_256(p, a, b);
  } else {
    a = __ldg(reinterpret_cast<const float4 *>(p));
    b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
  }
}
template <bool WIDE>
__device__ __forceinline__ void store_hit_as(rdn_hit *dst, float t, float u, float v, uint32_t prim, uint32_t geom, uint32_t inst,
                                             uint32_t custom, uint32_t kind) {
  if (WIDE) {
    st_global_256(dst, make_float4(t, u, v, __uint_as_float(prim)),
                  make_float4(__uint_as_float(geom), __uint_as_float(inst), __uint_as_float(custom), __uint_as_float(kind)));
  } else {
    store_hit(dst, t, u, v, prim, geom, inst, custom, kind);
  }
}

// ================================================================================================ reference order
struct WalkResult {
  float t, u, v;
  uint32_t slot, inst, kind;
};
struct WalkCounters {
  unsigned long long bvh_visit = 0, bvh_hit = 0, tri_visit = 0, tri_hit = 0, inst = 0, abort = 0;
};

// NaiveSahBvhCpu::traverse for one ray: stackless threaded pre-order walk over the reference-layout forests.
// Plain query: near_walk = near, far_init = far0.  Tie re-walk: the range used for box / triangle range tests is clamped to
// [near_walk, far_init] just around the closest distance (subtrees that end before it hold no hit at all, candidates
// beyond it cannot win), while the update_far asserts keep using the ray's own near.
template <bool COUNT>
__device__ __forceinline__ void reference_walk(const SceneDev &S, const rdn_launch &L, Vec3 ro, Vec3 rd, float near, float far0,
                                               float near_walk, float far_init, WalkResult &res, WalkCounters &ctr) {
  float far = far_init;  // the shared Rc<Cell<f32>> of RayRange
  res.t = far0; res.u = 0.f; res.v = 0.f;
  res.slot = RDN_INVALID_ID; res.inst = RDN_INVALID_ID; res.kind = 0;

  uint32_t tlas_cursor = 0xFFFFFFFFu;
  if (L.tlas_idx < S.n_tlas_binding) {
    const uint32_t handle = S.tlas_binding[L.tlas_idx];
    if (handle < S.n_tlas_root) tlas_cursor = S.tlas_root[handle].bvh_root_idx;
  }
  const Vec3 inv_rd = recip3(rd);
  bool end_search = false;

  while (tlas_cursor != 0xFFFFFFFFu && !end_search) {
    // ---- TraverseBvhIteratorCpu over the TLAS
    const float4 *np = reinterpret_cast<const float4 *>(S.tlas_bvh_forest + tlas_cursor);
    const float4 n0 = __ldg(np), n1 = __ldg(np + 1);
    if (COUNT) ctr.bvh_visit++;
    float tn;
    const uint32_t hit_next = __float_as_uint(n0.w), miss_next = __float_as_uint(n1.w);
    if (!slab_test(ro, inv_rd, near_walk * 1.0f, far * 1.0f, xyz(n0), xyz(n1), tn)) { tlas_cursor = miss_next; continue; }
    const uint32_t leaf_node = tlas_cursor;
    tlas_cursor = hit_next;
    if (hit_next != miss_next) continue;
    if (COUNT) ctr.bvh_hit++;
    const uint2 irange = *reinterpret_cast<const uint2 *>(S.tlas_bvh_forest[leaf_node].content_range);

    for (uint32_t tlas_idx = irange.x; tlas_idx < irange.y && !end_search; ++tlas_idx) {
      float4 b0, b1;
      load_pair<true>(S.tlas_bounding + tlas_idx, b0, b1);
      // ORIGINAL ray.range for the instance box (traverse_cpu.rs:80-86)
      if (!slab_test(ro, inv_rd, near, far0, xyz(b0), xyz(b1), tn)) continue;
      if ((L.cull_mask & __float_as_uint(b0.w)) == 0) continue;
      if (COUNT) ctr.inst++;
      const InstanceRecord *rec = S.instances + tlas_idx;
      const uint32_t flags = merge_geometry_instance_flag(L.ray_flags, rec->flags);
      Vec3 bo, bd;
      float scaling;
      to_object_space(rec, ro, rd, bo, bd, scaling);
      const Vec3 inv_bd = recip3(bd);
      const uint32_t blas_idx = rec->blas;
      if (blas_idx >= S.n_blas_meta) continue;
      if (flags & TF_SKIP_TRIANGLES) continue;
      const uint32_t g0 = S.blas_meta[blas_idx].tri_root_range[0], g1 = S.blas_meta[blas_idx].tri_root_range[1];
      const uint32_t cull_bits = cull_triangle_bits(flags);

      for (uint32_t g = g0; g < g1 && !end_search; ++g) {
        const GeometryMeta gm = S.geometry_meta[g];
        if (!cull_geometry_pass(flags, gm.geometry_flags)) continue;
        const bool ask_any_hit = L.any_hit != RDN_ANYHIT_NONE && geometry_non_opaque(flags, gm.geometry_flags);
        uint32_t cursor = gm.bvh_root_idx;
        while (cursor != 0xFFFFFFFFu && !end_search) {
          const float4 *bp = reinterpret_cast<const float4 *>(S.tri_bvh_forest + cursor);
          const float4 m0 = __ldg(bp), m1 = __ldg(bp + 1);
          if (COUNT) ctr.bvh_visit++;
          const uint32_t bh = __float_as_uint(m0.w), bm = __float_as_uint(m1.w);
          if (!slab_test(bo, inv_bd, near_walk * scaling, far * scaling, xyz(m0), xyz(m1), tn)) { cursor = bm; continue; }
          const uint32_t leaf = cursor;
          cursor = bh;
          if (bh != bm) continue;
          if (COUNT) ctr.bvh_hit++;
          const uint2 trange = *reinterpret_cast<const uint2 *>(S.tri_bvh_forest[leaf].content_range);
          for (uint32_t slot = trange.x; slot < trange.y; ++slot) {
            const float4 *tp = reinterpret_cast<const float4 *>(S.triangles + slot);
            float4 qn, qv0, qe1, qe2;
            load_pair<true>(tp, qn, qv0);
            load_pair<true>(tp + 2, qe1, qe2);
            if (COUNT) ctr.tri_visit++;
            float sign, t, u, v;
            if (!triangle_test(qn, qv0, qe1, qe2, bo, bd, near_walk * scaling, far * scaling, cull_bits, sign, t, u, v)) continue;
            const float distance = t / scaling;
            if (COUNT) ctr.tri_hit++;
            uint32_t behavior = RDN_ANYHIT_BEHAVIOR_ACCEPT_HIT;  // opaque -> commit; non-opaque -> the any-hit stage
            if (ask_any_hit) behavior = any_hit_behavior(S, L, slot, tlas_idx, distance);
            if (behavior & RDN_ANYHIT_BEHAVIOR_ACCEPT_HIT) {
              // RayRange::update_far asserts: the reference aborts; the candidate is rejected here
              if (!(near <= distance) || !(distance <= far)) { if (COUNT) ctr.abort++; continue; }
              far = distance;
              res.t = distance; res.u = u; res.v = v;
              res.slot = slot; res.inst = tlas_idx;
              res.kind = sign < 0.0f ? RDN_HIT_KIND_BACK_FACING_TRIANGLE : RDN_HIT_KIND_FRONT_FACING_TRIANGLE;
              if (flags & TF_END_SEARCH) behavior |= RDN_ANYHIT_BEHAVIOR_END_SEARCH;
            }
            if (behavior & RDN_ANYHIT_BEHAVIOR_END_SEARCH) { end_search = true; break; }
          }
        }
      }
    }
  }
}

__device__ __forceinline__ void store_walk_result(const SceneDev &S, rdn_hit *dst, const WalkResult &r, float far0) {
  if (r.slot != RDN_INVALID_ID) {
    const SlotInfo si = S.slot_info[r.slot];
    store_hit(dst, r.t, r.u, r.v, si.primitive_id, si.geometry_idx, r.inst, S.instances[r.inst].instance_custom_index, r.kind);
  } else {
    store_hit(dst, far0, 0.f, 0.f, RDN_INVALID_ID, RDN_INVALID_ID, RDN_INVALID_ID, RDN_INVALID_ID, 0);
  }
}

template <bool COUNT>
__global__ void __launch_bounds__(128) k_trace_reference(const SceneDev S, const rdn_launch L, const rdn_ray *__restrict__ rays,
                                                         uint64_t n, rdn_hit *__restrict__ hits, const TraceScratch scratch,
                                                         uint32_t tiles_x, uint32_t width, uint32_t height, uint64_t n_fetch_max,
                                                         const unsigned long long *n_ptr) {
  WalkCounters ctr;
  const uint64_t n_fetch = n_ptr ? (__ldg(n_ptr) < n_fetch_max ? __ldg(n_ptr) : n_fetch_max) : n_fetch_max;  // device-side wave size
  for (uint64_t f = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; f < n_fetch;
       f += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
    // a warp walks an 8x4 pixel tile of a grid launch (neighbours on screen walk the same nodes), else 32 consecutive rays
    uint64_t ri = f;
    if (tiles_x) {
      const uint32_t tile = static_cast<uint32_t>(f >> 5), in_tile = static_cast<uint32_t>(f) & 31u;
      const uint32_t ty = tile / tiles_x, tx = tile - ty * tiles_x;
      const uint32_t x = tx * 8u + (in_tile & 7u), y = ty * 4u + (in_tile >> 3);
      if (x >= width || y >= height) continue;
      ri = static_cast<uint64_t>(y) * width + x;
    }
    const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays + ri));
    const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays + ri) + 1);
    WalkResult res;
    reference_walk<COUNT>(S, L, xyz(r0), xyz(r1), r0.w, r1.w, r0.w, r1.w, res, ctr);
    store_walk_result(S, hits + ri, res, r1.w);
  }
  if (COUNT) {
    // warp-reduce, one atomic per warp per counter
    unsigned long long vals[6] = {ctr.bvh_visit, ctr.bvh_hit, ctr.tri_visit, ctr.tri_hit, ctr.inst, ctr.abort};
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      unsigned long long x = vals[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(FULL_MASK, x, off);
      if ((threadIdx.x & 31) == 0 && x) atomicAdd(scratch.counters + i, x);
    }
  }
}

// Exact re-walk of a ray the ordered kernel queued, in the reference's order.
//   best is a distance: a near-tie (a second candidate within TIE_EPS of the closest hit).  The walk's range is clamped to
//     best -/+ 3*TIE_EPS*|best|: subtrees that end before the clamp hold no hit at all and candidates beyond it cannot win, so
//     from the first shared accept on the clamped walk is in the same state as the reference's unclamped one and ends on the
//     reference's answer.  Should the clamped walk find nothing (the closest candidate lies outside the boxes that lead to
//     it: an irregular triangle the flattener's classification missed) the ray is walked again over its whole range.
//   best is NaN: a ray that may enter an irregular instance; walked over its whole range straight away.
// Either way the record written is NaiveSahBvhCpu::traverse's.
__device__ __forceinline__ void rewalk_in_reference_order(const SceneDev &S, const rdn_launch &L, const rdn_ray *rays, rdn_hit *hits,
                                                          uint64_t ri, float best, uint32_t *fallbacks) {
  const float4 r0 = __ldg(reinterpret_cast<const float4 *>(rays + ri));
  const float4 r1 = __ldg(reinterpret_cast<const float4 *>(rays + ri) + 1);
  WalkResult res;
  WalkCounters ctr;
  bool clamped = best == best;
  float near_walk = r0.w, far_init = r1.w;
  if (clamped) {
    const float slack = 3.0f * TIE_EPS * fabsf(best);
    near_walk = fmaxf(r0.w, best - slack);
    far_init = fminf(r1.w, best + slack);
  }
#pragma unroll 1
  for (;;) {
    reference_walk<false>(S, L, xyz(r0), xyz(r1), r0.w, r1.w, near_walk, far_init, res, ctr);
    if (res.slot != RDN_INVALID_ID || !clamped) break;
    clamped = false; near_walk = r0.w; far_init = r1.w;
    atomicAdd(fallbacks, 1u);
  }
  store_walk_result(S, hits + ri, res, r1.w);
}

// The queue drained by a separate kernel (RDN_ORDERED_VARIANT=9, the first version; kept for A/B runs): grid-stride over the
// device-side queue, no host round trip between the two kernels.
__global__ void __launch_bounds__(128) k_resolve_ties(const SceneDev S, const rdn_launch L, const rdn_ray *__restrict__ rays,
                                                      rdn_hit *__restrict__ hits, const TraceScratch scratch) {
  const uint32_t count = *scratch.tie_count;  // (reset by the memset the launcher puts behind this kernel)
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(scratch.tie_total, count);
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < count; k += gridDim.x * blockDim.x) {
    const uint32_t ri = scratch.tie_queue[k];
    scratch.tie_queue[k] = RDN_INVALID_ID;  // unpublished again: launches that drain the queue in-kernel wait on this value
    rewalk_in_reference_order(S, L, rays, hits, ri, scratch.tie_best[k], scratch.tie_unresolved);
  }
}

// ================================================================================================ ordered
constexpr int TILE_CLASSES = 8;
constexpr uint64_t MID_LIST_RAYS = 600000, LONG_LIST_RAYS = 2000000;  // ray lists from here on: throughput, not the longest ray, decides the launch (launch_trace_ordered)
enum { SHARE_NEVER = 0, SHARE_ALWAYS = 1, SHARE_LATE = 2 };  // template argument SHARE of k_trace_ordered_rounds

struct OrderedParams {
  SceneDev S;
  rdn_launch L;
  const rdn_ray *rays;
  rdn_hit *hits;
  uint64_t n;        // rays
  uint64_t n_fetch;  // fetch indices (= n, or 32 * tiles when walking 8x4 pixel tiles)
  const unsigned long long *n_ptr;  // optional (ray lists): the ray count lives on the device (a compaction's output), n is its upper bound
  uint32_t tiles_x;  // 0: linear
  uint32_t width, height;
  uint32_t world_root;  // wide reference of the bound TLAS (REF_EMPTY: every ray misses)
  uint32_t irregular_start, irregular_count;  // the bound TLAS's irregular instances (S.irregular_instances), at most IRREGULAR_LIST_MAX
  // HOT: wide nodes [hot_a_base, +hot_a_count) (top of the TLAS tree) and [hot_b_base, +hot_b_count) (top of the largest geometry
  // tree) are staged in shared memory by two bulk copies (TMA) at CTA start
  uint32_t hot_a_base, hot_a_count, hot_b_base, hot_b_count;
  // SHARE: deferred subtrees are handed to idle lanes once a pass over a tile is share_after rounds old and no more than share_busy
  // lanes are busy, by lanes that have at least share_min of them
  int share_busy, share_min, share_after;
  // TOPUP (long ray lists): a warp with no more than topup_lanes busy lanes goes back to the refill point and fills the others
  int topup_lanes;
  uint32_t wait_epoch;  // launches issued before this one on the same scratch set: they must have left before this one touches it
  // HISTORY (grids; see k_build_tile_lists): tile_lists / tile_meta = the tiles of the grid by how long they took in an earlier
  // launch over it (TILE_CLASSES lists of capacity n_tiles each, longest class last) and the lists' lengths, or null: grid order;
  // tile_cost = where this launch notes the duration of its passes, or null; tile_meta_clear = the description the order kernel
  // behind this launch will fill, zeroed by the last CTA on its way out
  const uint32_t *tile_lists;
  const TileMeta *tile_meta;
  uint32_t *tile_cost;
  TileMeta *tile_meta_clear;
  uint32_t n_tiles;
  TraceScratch scratch;
};

// the SM's cycle counter (durations of passes over tiles; the emulated build counts host nanoseconds instead)
__device__ __forceinline__ uint32_t pass_clock() {
#ifdef RDN_SIMT_EMU
  return static_cast<uint32_t>(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
#else
  return static_cast<uint32_t>(clock());
#endif
}

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
#ifdef RDN_SIMT_EMU
  ::simt::preempt();  // (a spin on another thread's store must let that thread run: cooperative scheduling in the emulated build)
#endif
  return *reinterpret_cast<const volatile uint32_t *>(p);
}

// Hits of an irregular instance (or of the irregular triangles of an otherwise regular BLAS) need not lie inside their boxes,
// so what the ordered walk would prune the reference may have found.  A ray whose ORIGINAL range meets such an instance (the
// reference's own instance test, traverse_cpu.rs:80-86) — and, when only some leaves of its BLAS are irregular, the box of
// one of those leaves in object space — can reach it in the reference's walk; this is a superset of the rays that do, and
// they are handed to the reference-order walk as they are.  Out of line: regular scenes never call it and it must not
// cost the traversal loop registers.
__device__ __noinline__ bool meets_irregular_instance(const SceneDev &S, uint32_t start, uint32_t count, uint32_t cull_mask,
                                                      const rdn_ray *ray) {
  const float4 r0 = __ldg(reinterpret_cast<const float4 *>(ray)), r1 = __ldg(reinterpret_cast<const float4 *>(ray) + 1);
  const Vec3 ro = xyz(r0), rd = xyz(r1), rinv = recip3(rd);
  const float t_min = r0.w, t_max = r1.w;
  for (uint32_t k = 0; k < count; ++k) {
    const uint32_t entry = __ldg(S.irregular_instances + start + k);
    const uint32_t slot = entry & ~IRREGULAR_WHOLE_BIT;
    const float4 *tb = reinterpret_cast<const float4 *>(S.tlas_bounding + slot);
    const float4 b0 = __ldg(tb), b1 = __ldg(tb + 1);
    float tn;
    if (!slab_test(ro, rinv, t_min, t_max, xyz(b0), xyz(b1), tn) || (cull_mask & __float_as_uint(b0.w)) == 0) continue;
    if (entry & IRREGULAR_WHOLE_BIT) return true;
    const InstanceRecord *rec = S.instances + slot;
    const uint32_t blas = rec->blas;
    if (blas >= S.n_blas_meta) continue;
    const uint32_t l0 = S.blas_meta[blas].irregular_leaf_start, ln = S.blas_meta[blas].irregular_leaf_count;
    Vec3 bo, bd;
    float scaling;
    to_object_space(rec, ro, rd, bo, bd, scaling);
    const Vec3 inv_bd = recip3(bd);
    for (uint32_t j = 0; j < ln; ++j) {
      const float4 *lb = reinterpret_cast<const float4 *>(S.irregular_leaf_boxes + l0 + j);
      const float4 m0 = __ldg(lb), m1 = __ldg(lb + 1);
      if (slab_test(bo, inv_bd, t_min * scaling, t_max * scaling, xyz(m0), xyz(m1), tn)) return true;
    }
  }
  return false;
}

// Warp-aggregated append to the device-side re-walk queue (called by whatever lanes are active): one atomic per warp; the
// count is bumped first, then the distance is written, then (after a fence when the queue is drained concurrently) the
// ray index publishes the slot.
template <bool DRAIN_TIES>
__device__ __forceinline__ void enqueue_rewalk(const TraceScratch &scratch, uint32_t lane, uint64_t ri, float best) {
  const uint32_t peers = __activemask();
  const int pl = __ffs(peers) - 1;
  uint32_t qbase = 0;
  if (static_cast<int>(lane) == pl) qbase = atomicAdd(scratch.tie_count, static_cast<uint32_t>(__popc(peers)));
  qbase = __shfl_sync(peers, qbase, pl);
  const uint32_t q = qbase + __popc(peers & ((1u << lane) - 1u));
  scratch.tie_best[q] = best;
  if (DRAIN_TIES) __threadfence();  // the distance is visible before the index publishes the slot
  *reinterpret_cast<volatile uint32_t *>(scratch.tie_queue + q) = static_cast<uint32_t>(ri);
}

// Drain of the device-side tie queue by threads that have left the traversal loop (DRAIN_TIES).  Entries are
// claimed one at a time with a CAS on the cursor; a claimed slot is awaited until its ray index is published (the
// appender bumps tie_count first, writes the closest distance, fences, then publishes the index) and handed back as
// RDN_INVALID_ID so the queue is clean for the next launch.  Warps that run dry early resolve the ties found so far
// while the stragglers finish; the last CTA sweeps whatever was appended after that.
__device__ __forceinline__ void drain_tie_queue(const OrderedParams &P) {
  const uint32_t lane = threadIdx.x & 31u;
  for (;;) {
    // lane 0 claims up to 8 entries for the warp (so a burst of ties spreads over many warps); the claimed lanes re-walk together
    uint32_t base = 0, take = 0;
    if (lane == 0) {
      for (;;) {
        const uint32_t cur = ld_volatile_u32(P.scratch.tie_cursor);
        const uint32_t cnt = ld_volatile_u32(P.scratch.tie_count);
        if (cur >= cnt) break;
        const uint32_t want = cnt - cur < 8u ? cnt - cur : 8u;
        if (atomicCAS(P.scratch.tie_cursor, cur, cur + want) == cur) { base = cur; take = want; break; }
      }
    }
    base = __shfl_sync(FULL_MASK, base, 0);
    take = __shfl_sync(FULL_MASK, take, 0);
    if (take == 0) break;
    if (lane < take) {
      const uint32_t q = base + lane;
      uint32_t ri;
      while ((ri = ld_volatile_u32(P.scratch.tie_queue + q)) == RDN_INVALID_ID) {}
      __threadfence();
      const float best = __ldcg(P.scratch.tie_best + q);
      P.scratch.tie_queue[q] = RDN_INVALID_ID;
      rewalk_in_reference_order(P.S, P.L, P.rays, P.hits, ri, best, P.scratch.tie_unresolved);
    }
    __syncwarp();
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Warp-synchronous rounds: every round the alive lanes (1) descend up to K inner nodes each, (2) re-converge
// (__syncwarp) and handle their leaf / instance / bookkeeping item TOGETHER — on Volta+ lanes do not re-converge at a
// loop exit by themselves, and a triangle test executed by 3 lanes costs the warp as much as one executed by 32 —
// (3) vote: when no lane holds a ray any more the warp goes back to the refill point.
// Refill culls rays against the TLAS root box on the spot (the reference's first test), so rays that miss the scene
// never occupy a traversal lane.
// Template switches.  K: node steps per round (3; 2 was the round-1 default).  DRAIN_TIES: near-tie queue drained inside the
// kernel (else by k_resolve_ties).  IRREGULAR: the bound TLAS lists irregular instances (checked at refill); a separate
// instantiation so that regular scenes pay nothing for the out-of-line test (the call makes ptxas save two dozen registers around
// the whole refill block).  LD256: 256-bit loads / stores for nodes, triangles, instance boxes, rays and hit records.  INST_LOOP:
// the instances of a TLAS leaf whose box the ray misses are skipped in a loop instead of costing a round each, and a BLAS of one
// geometry is entered without a geometry-iterator round (+2..4 % with K = 3, profiles/kbench_r2a_*.log).  SHARE: the lanes of a
// warp share the work of its long rays (see the vote in ordered_rounds.inc) — used for launches that are ray LISTS (bounce / shadow
// waves), whose duration is the latency of their longest rays: +24 % on config 3 (passes older than 30 rounds, at most 16 busy
// lanes, donors with two or more deferred subtrees).  SHARE_ALWAYS runs the sharing loop from the first round; SHARE_LATE (the
// default of lists under 2 M rays) runs the plain loop and hands a pass that has grown old over to the sharing loop.  Either way the kernel is 5-6 %
// slower than the plain one before anything is shared, which grids never win back: they keep the plain loop
// (profiles/kbench_r2h_*.log, kbench_r2u_*.log, kbench_r2v_*.log, kbench_r2x_late_sharing.log).  ANYHIT: candidates of non-opaque
// geometry go through the any-hit stage.  HISTORY (device-resident grids): the tiles are taken in the order of the lists
// k_build_tile_lists made from the pass durations an earlier launch over the grid noted — long tiles first — and the launch notes its
// own (+7..9 % on serialised launches of configs 1 / 2, profiles/kbench_r3h_tile_history.log).
// Experiments kept for A/B runs.  HOT: the top levels of the TLAS tree and of the largest geometry tree (breadth-first blocks of
// HOT_TOP_NODES wide nodes, 8 KB each) are copied into shared memory with cp.async.bulk (TMA, completion on an mbarrier) when the
// CTA starts, and node fetches that fall into either block read shared memory instead of L1 (measured 9-13 % slower).  WIDE4: the
// walk runs over the 128 B four-box nodes (layout.h Wide4Node: the grandchildren of a reference node, exact boxes): half the steps
// for the same box tests, the hit children entered nearest first and the others deferred farthest first (measured neutral).
// Measured, rejected and removed (logs under profiles/): whole-unit / per-SM work distribution, child prefetch,
// speculative traversal with a postponed leaf, any-hit pre-classification, topping a thinned-out tile up with new rays, the first
// 8-32 stack entries per thread in shared memory (3-10 % slower than the L1-cached local stack), rows of tiles taken from the
// middle of the frame outwards (+2..5 % on configs 1 / 2, -11 % on config 4).
template <int K, int MINB, bool DRAIN_TIES, bool IRREGULAR, bool LD256, bool HOT, bool WIDE4, bool INST_LOOP = false, int SHARE = 0, bool ANYHIT = false, bool HISTORY = false, bool TOPUP = false>
__global__ void __launch_bounds__(ORDERED_BLOCK, MINB) k_trace_ordered_rounds(const __grid_constant__ OrderedParams P) {
  RDN_COST(COST_PROLOGUE);
  const SceneDev &S = P.S;
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t stack[STACK_MAX];
  int sp = 0;
  // SHARE (lanes of a warp share the work of its long rays, see the vote): entries [lo, sp) of the stack are the deferred subtrees
  // of the space the lane is in (world, or the instance it entered) and may be handed to idle lanes, lowest = largest first;
  // `home` is the lane that owns the ray this lane works on; `helpers` (owner only) the lanes still out with pieces of its ray
  int lo = 0;  // bits 0-7: the floor of the segment the lane is in; bits 8-15: the world segment's floor while the lane is inside an instance
  uint32_t home = lane, helpers = 0;  // home: bits 0-4 the owner's lane, bits 5.. the age of the pass over the tile in rounds (a separate
                                      // loop counter ends up spilled, and a local access per round stalls the round)
  // Programmatic dependent launch: the NEXT ordered launch on this stream may start filling SM slots as soon as CTAs of this one
  // leave, i.e. while the last long rays of this launch are still being walked (a no-op when launched without the attribute).
  // The next launch reads nothing this one writes (its own rays, the read-only scene, the other scratch set).
  pdl_launch_dependents();
  // Gate: with tails overlapping, the launch after next can reach the device while the launch that last used THIS scratch
  // set still has a straggler CTA (late CTAs of the launch in between find the ray list dry and leave at once).  Every
  // earlier launch is fully resident by then (a dependent launch starts only after all CTAs of its predecessor have
  // started), so waiting here cannot deadlock.
  if (threadIdx.x == 0) {
    // (bounded: about ten seconds.  A launch that gives up says so in gate_timeout and goes on — its results may be wrong, which the
    // host reports as an error, but the device does not hang)
    uint32_t spins = 0;
    while (ld_volatile_u32(P.scratch.epoch_done) != P.wait_epoch) {
      __nanosleep(200);
      if (++spins > 40000000u) { atomicAdd(P.scratch.gate_timeout, 1u); break; }
    }
    __threadfence();
  }
  __syncthreads();

  __shared__ uint32_t s_tile_count[TILE_CLASSES];
  if constexpr (HISTORY) {
    // (written by an order kernel that completed before this launch was allowed to start: launches that record are not overlapped)
    if (P.tile_lists && threadIdx.x < TILE_CLASSES) s_tile_count[threadIdx.x] = P.tile_meta->count[threadIdx.x];
    __syncthreads();
  }
  __shared__ __align__(128) float4 s_hot[HOT ? 2 * HOT_TOP_NODES * 4 : 4];
  __shared__ __align__(8) unsigned long long s_hot_bar;
  __shared__ int s_age[SHARE == SHARE_LATE ? ORDERED_BLOCK : 1];  // SHARE_LATE: the age of the pass over the tile in rounds, one word per thread
#ifndef RDN_SIMT_EMU
  if constexpr (HOT) {
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&s_hot_bar));
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      const uint32_t bytes_a = P.hot_a_count * 64u, bytes_b = P.hot_b_count * 64u;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_a + bytes_b) : "memory");
      if (bytes_a)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         static_cast<uint32_t>(__cvta_generic_to_shared(s_hot))),
                     "l"(S.wide_nodes + P.hot_a_base), "r"(bytes_a), "r"(bar)
                     : "memory");
      if (bytes_b)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         static_cast<uint32_t>(__cvta_generic_to_shared(s_hot + HOT_TOP_NODES * 4))),
                     "l"(S.wide_nodes + P.hot_b_base), "r"(bytes_b), "r"(bar)
                     : "memory");
    }
    __syncthreads();
    uint32_t done = 0;
    while (!done) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar) : "memory");
    }
  }
#endif

  // the world pseudo-root (TLAS root box + its reference) is launch-uniform: keep it in registers
  Vec3 root_min = {0, 0, 0}, root_max = {0, 0, 0};
  uint32_t world_entry = REF_EMPTY;
  if (P.world_root != REF_EMPTY) {
    const float4 *np = WIDE4 ? reinterpret_cast<const float4 *>(S.wide4_nodes + P.world_root) : reinterpret_cast<const float4 *>(S.wide_nodes + P.world_root);
    const float4 q0 = __ldg(np), q1 = __ldg(np + 1);
    root_min = xyz(q0); root_max = xyz(q1);
    world_entry = __float_as_uint(q0.w);
  }

  bool alive = false;
  uint64_t ri = 0;
  Vec3 o = {0, 0, 0}, d = {0, 0, 1}, inv = {0, 0, 0};
  float t_near_world = 0.f, far0 = 0.f;
  float scaling = 1.f, near_s = 0.f, far_s = 0.f;
  float bound = 0.f;
  float best = 0.f, second = 0.f, best_u = 0.f, best_v = 0.f;
  uint32_t best_slot = RDN_INVALID_ID, best_inst = RDN_INVALID_ID, best_back = 0;
  uint32_t cur = REF_DONE, cur_inst = 0, cur_flags = 0, cull_bits = 0, geom_end = 0;
  bool in_object = false;
  bool warp_exhausted = false;
#ifdef RDN_DEBUG_STEPS
  // per-thread totals, reduced once per warp at kernel exit so the counters do not perturb the timeline
  unsigned long long dbg_steps = 0, dbg_tris = 0, dbg_pushes = 0, dbg_ray_steps = 0, dbg_max = 0, dbg_rays = 0, dbg_long = 0;
#endif
#if defined(RDN_DEBUG_STEPS) || defined(RDN_DEBUG_TIMELINE)
  const unsigned long long dbg_t_in = globaltimer_ns();
  if (lane == 0) atomicMin(P.scratch.counters + 6, dbg_t_in);  // first warp in
#endif

  // a full stack drops the entry and raises a sticky flag, reported once when the ray ends: no atomic (and so no branch around one)
  // inside the traversal loop
  bool stack_overflowed = false;
#define RDN_PUSH(v) do { if (sp < STACK_MAX) stack[sp++] = (v); else stack_overflowed = true; } while (0)
#define RDN_POP() (sp > 0 ? stack[--sp] : REF_DONE)

  // a finished ray: the record of its closest candidate (or the miss record), near-ties queued for the reference-order re-walk
  auto finish_ray = [&]() {
    RDN_COST(COST_FINISH);
#ifdef RDN_DEBUG_STEPS
    dbg_max = dbg_ray_steps > dbg_max ? dbg_ray_steps : dbg_max;
    dbg_long += dbg_ray_steps > 200 ? 1 : 0;
    ++dbg_rays;
    dbg_ray_steps = 0;
#endif
    if (stack_overflowed) { atomicAdd(P.scratch.stack_overflow, 1u); stack_overflowed = false; }
    rdn_hit *dst = P.hits + ri;
    if (best_slot != RDN_INVALID_ID) {
      // the ordered result is stored first; a near-tie ray is also queued (warp-aggregated append) for the exact
      // reference-order re-walk, which overwrites the record
      const SlotInfo si = S.slot_info[best_slot];
      store_hit_as<LD256>(dst, best, best_u, best_v, si.primitive_id, si.geometry_idx, best_inst,
                S.instances[best_inst].instance_custom_index,
                best_back ? RDN_HIT_KIND_BACK_FACING_TRIANGLE : RDN_HIT_KIND_FRONT_FACING_TRIANGLE);
      if (second <= best + TIE_EPS * fabsf(best)) {
        RDN_COST(COST_TIE);
        __threadfence();  // the ordered record lands before whoever drains the queue writes the exact one
        enqueue_rewalk<DRAIN_TIES>(P.scratch, lane, ri, best);
      }
    } else {
      store_hit_as<LD256>(dst, far0, 0.f, 0.f, RDN_INVALID_ID, RDN_INVALID_ID, RDN_INVALID_ID, RDN_INVALID_ID, 0);
    }
    alive = false;
  };

  for (;;) {
    // ---------------- warp-converged refill (a few rounds, so scene-missing rays are retired here)
#pragma unroll 1
    for (int attempt = 0; attempt < 4; ++attempt) {
      const uint32_t want = __ballot_sync(FULL_MASK, !alive);
      if (!want || warp_exhausted) break;
      RDN_COST(COST_REFILL);
      const int cnt = __popc(want);
      const int leader = __ffs(want) - 1;
      unsigned long long base = 0;
      if (static_cast<int>(lane) == leader) base = atomicAdd(P.scratch.work_counter, static_cast<unsigned long long>(cnt));
      base = __shfl_sync(FULL_MASK, base, leader);
      // (a wave whose size a previous kernel left on the device: read at every refill rather than kept in registers)
      const uint64_t n_fetch = P.n_ptr ? (__ldg(P.n_ptr) < P.n_fetch ? __ldg(P.n_ptr) : P.n_fetch) : P.n_fetch;
      if (base + cnt >= n_fetch) {
#if defined(RDN_DEBUG_STEPS) || defined(RDN_DEBUG_TIMELINE)
        if (!warp_exhausted && static_cast<int>(lane) == leader) atomicMin(P.scratch.counters + 7, globaltimer_ns());  // ray list ran dry
#endif
        warp_exhausted = true;
      }
      if (!alive) {
        const uint64_t f = base + __popc(want & ((1u << lane) - 1u));
        bool valid = f < n_fetch;
        uint64_t idx = f;
        if (valid && P.tiles_x) {
          uint32_t tile = static_cast<uint32_t>(f >> 5);  // launches hold < 2^31 rays: 32-bit tile arithmetic
          if constexpr (HISTORY) {
            if (P.tile_lists) {  // slot -> (class, position): the longest class is taken first
              int q = TILE_CLASSES - 1;
#pragma unroll
              for (; q > 0; --q) {
                const uint32_t len = s_tile_count[q];
                if (tile < len) break;
                tile -= len;
              }
              tile = __ldg(P.tile_lists + static_cast<uint64_t>(q) * P.n_tiles + tile);
            }
          }
          const uint32_t in_tile = static_cast<uint32_t>(f) & 31u;
          const uint32_t ty = tile / P.tiles_x, tx = tile - ty * P.tiles_x;
          const uint32_t x = tx * 8u + (in_tile & 7u), y = ty * 4u + (in_tile >> 3);
          valid = x < P.width && y < P.height;
          idx = static_cast<uint64_t>(y) * P.width + x;
        }
        if (valid) {
          RDN_COST(COST_RAY_LOAD);
          float4 r0, r1;
          load_pair<LD256>(P.rays + idx, r0, r1);
          const Vec3 ro = xyz(r0), rd = xyz(r1), rinv = recip3(rd);
          float tn;
          const bool enters = world_entry != REF_EMPTY && slab_test(ro, rinv, r0.w, r1.w, root_min, root_max, tn);
          const bool suspect = IRREGULAR && enters &&
                               meets_irregular_instance(S, P.irregular_start, P.irregular_count, P.L.cull_mask, P.rays + idx);
          if (suspect) {
            enqueue_rewalk<DRAIN_TIES>(P.scratch, lane, idx, __int_as_float(0x7FC00000));  // NaN: walk the whole range
          } else if (enters) {
            ri = idx; o = ro; d = rd; inv = rinv;
            t_near_world = r0.w; far0 = r1.w;
            scaling = 1.f; near_s = t_near_world; bound = far0; far_s = far0;
            best = INFINITY; second = INFINITY; best_slot = RDN_INVALID_ID; best_inst = RDN_INVALID_ID;
            in_object = false; sp = 0;
            if constexpr (SHARE == SHARE_ALWAYS) { lo = 0; home = lane; helpers = 0; }
            cur = world_entry;
            alive = true;
          } else {
            store_hit_as<LD256>(P.hits + idx, r1.w, 0.f, 0.f, RDN_INVALID_ID, RDN_INVALID_ID, RDN_INVALID_ID, RDN_INVALID_ID, 0);
          }
        }
      }
    }
    RDN_COST(COST_OUTER);
    const uint32_t amask = __ballot_sync(FULL_MASK, alive);
    if (amask == 0) {
      if (warp_exhausted) break;
      continue;
    }
#ifdef RDN_DEBUG_TIMELINE
    const unsigned long long dbg_pass_t0 = globaltimer_ns();
    unsigned long long dbg_pass_rounds = 0, dbg_pass_busy = 0;
#endif

    [[maybe_unused]] uint32_t pass_t0 = 0;
    if constexpr (HISTORY) pass_t0 = pass_clock();
    if constexpr (SHARE == SHARE_LATE) {
      // Two loops: the plain one for the first share_after rounds of a pass (what almost every tile needs), then — for the few passes
      // that have grown old with a handful of long rays left — the loop in which idle lanes take over deferred subtrees.  The plain
      // loop pays one shared-memory word per round for it instead of the 6 % the sharing loop costs before anything is shared.
      bool late = false;
      s_age[threadIdx.x] = 0;
      if (alive) {
        constexpr bool SH = false, LATE_FIRST = true;
        const uint32_t rmask = amask;
#include "ordered_rounds.inc"
        if (!late && cur == REF_DONE) finish_ray();
      }
      if (__ballot_sync(FULL_MASK, late) != 0) {
        // every lane joins: the busy ones find the floor of the stack segment they are in (inside an instance: what lies above the
        // exit marker; the marker already popped, or dropped by a full stack: nothing to hand out), the others are idle helpers
        home = lane | (1023u << 5);
        helpers = 0;
        lo = 0;
        if (alive && cur != REF_DONE && in_object) {
          int k = sp;
          while (k > 0 && stack[k - 1] != REF_EXIT_INSTANCE) --k;
          lo = k > 0 ? k : sp + 1;
        }
        if (!alive) { cur = REF_DONE; sp = 0; }
        constexpr bool SH = true, LATE_FIRST = false;
        const uint32_t rmask = FULL_MASK;
#include "ordered_rounds.inc"
        if (alive) finish_ray();
      }
    } else {
      // (SHARE_ALWAYS: every lane takes part in the rounds — a lane without a ray idles at cur == REF_DONE until it is handed a subtree)
      constexpr bool SH = SHARE == SHARE_ALWAYS, LATE_FIRST = false;
      [[maybe_unused]] bool late = false;
      const uint32_t rmask = SH ? FULL_MASK : amask;
      if constexpr (SH) home &= 31u;  // a new pass over a tile: age 0
      if (SH || alive) {
#include "ordered_rounds.inc"
        if (SH ? alive : cur == REF_DONE) finish_ray();
      }
    }
    if constexpr (HISTORY) {
      // how long the pass took, noted at the tile of its first ray (a pass is one tile, or the ends of two): what the next launch
      // over this grid sorts its tiles by
      if (P.tile_cost && static_cast<int>(lane) == __ffs(amask) - 1) {
        const uint32_t dt = pass_clock() - pass_t0;
        const uint32_t y = static_cast<uint32_t>(ri / P.width), x = static_cast<uint32_t>(ri - static_cast<uint64_t>(y) * P.width);
        atomicMax(P.tile_cost + ((y >> 2) * P.tiles_x + (x >> 3)), dt);
      }
    }
#ifdef RDN_DEBUG_TIMELINE
    {  // one pass over a tile: duration histogram (8 us buckets), rounds per bucket, the longest pass, passes ending after the list ran dry
      const unsigned long long now = globaltimer_ns(), dt = now - dbg_pass_t0;
      unsigned long long rounds = dbg_pass_rounds;
      for (int off = 16; off > 0; off >>= 1) { const unsigned long long r = __shfl_xor_sync(FULL_MASK, rounds, off); rounds = r > rounds ? r : rounds; }
      if (lane == 0) {
        const unsigned b = dt / 8000ull < 31ull ? static_cast<unsigned>(dt / 8000ull) : 31u;
        atomicAdd(P.scratch.counters + 12 + b, 1ull);
        atomicAdd(P.scratch.counters + 12 + 32 + b, rounds);
        atomicMax(P.scratch.counters + 12 + 64, (dt << 20) | (rounds & 0xFFFFFull));
        atomicAdd(P.scratch.counters + 12 + 66 + b, dbg_pass_busy);  // (lane 0 takes part in every ballot: its sum is the warp's)
      }
    }
#endif
    __syncwarp();
  }
#undef RDN_PUSH
#undef RDN_POP
#ifdef RDN_DEBUG_STEPS
  {
    unsigned long long vals[5] = {dbg_steps, dbg_rays, dbg_tris, dbg_pushes, dbg_long};
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      unsigned long long x = vals[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(FULL_MASK, x, off);
      if (lane == 0 && x) atomicAdd(P.scratch.counters + 1 + i, x);  // steps, rays entered, triangle tests, pushes, rays > 200 steps
    }
    atomicMax(P.scratch.counters + 0, dbg_max);                           // longest ray, inner-node steps
  }
#endif
#if defined(RDN_DEBUG_STEPS) || defined(RDN_DEBUG_TIMELINE)
  if (lane == 0) {
    const unsigned long long t_out = globaltimer_ns();
    atomicMax(P.scratch.counters + 8, t_out);               // last warp out
    atomicAdd(P.scratch.counters + 9, t_out - dbg_t_in);    // sum of the warps' busy time (before the tie drain / exit barrier)
  }
#endif
  RDN_COST(COST_EPILOGUE);
  if (DRAIN_TIES) drain_tie_queue(P);
  // the last CTA to leave sweeps the ties appended after everyone else looked, then re-arms the counters for the next
  // launch on this scratch (no memset node between launches)
  __shared__ uint32_t s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(P.scratch.blocks_done, 1u) == gridDim.x - 1u ? 1u : 0u;
  }
  __syncthreads();
  if (s_last) {
    // ... and this launch must not COMPLETE before its predecessor: whatever follows in the stream (a copy of the previous
    // launch's hits, a kernel reading them) is ordered behind this grid only
    pdl_wait();
    if (DRAIN_TIES) {
      __threadfence();
      drain_tie_queue(P);
      __syncthreads();
    }
    if constexpr (HISTORY) {
      if (P.tile_meta_clear && threadIdx.x < sizeof(TileMeta) / 4) reinterpret_cast<uint32_t *>(P.tile_meta_clear)[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) {
      *P.scratch.work_counter = 0ull;
      *P.scratch.blocks_done = 0u;
      if (DRAIN_TIES) {
        atomicAdd(P.scratch.tie_total, ld_volatile_u32(P.scratch.tie_count));
        *P.scratch.tie_count = 0u;
        *P.scratch.tie_cursor = 0u;
      }
      __threadfence();
      *reinterpret_cast<volatile uint32_t *>(P.scratch.epoch_done) = P.wait_epoch + 1u;  // the set is free for the next launch on it
    }
  }
}

// Tiles of a grid by how long they took.  A launch ends when its last long pass ends, and a pass over a tile that straddles a
// silhouette takes as long as the rest of the launch: such tiles have to start first.  Behind a launch that noted its pass durations
// (tile_cost), this kernel files every tile under one of up to TILE_CLASSES classes — longest pass more than 4x / 3x / 2x / 1.5x the
// mean pass of the launch before, or none of that (nine tiles in ten) — and the next launch over the grid takes the classes longest
// first.  One thread per tile, 1024 consecutive tiles per CTA: ranks inside the CTA by ballot, one atomic per CTA and class for its
// range in the class's list, so every list is in grid order up to the order in which the CTAs arrive, and neighbouring warps of the
// next launch still walk neighbouring tiles.  Also sums up the durations for the kernel behind the next launch, and clears them.
struct TileThresholds { uint32_t quarters[TILE_CLASSES - 1]; };  // class boundaries in quarters of the mean, ascending (unused: 0xFFFFFFFF)
__device__ __forceinline__ int tile_class(uint32_t cost, unsigned long long mean, const TileThresholds &th) {
  const unsigned long long c = 4ull * cost;
  int k = 0;
#pragma unroll
  for (int q = 0; q < TILE_CLASSES - 1; ++q) k += c > th.quarters[q] * mean ? 1 : 0;
  return k;
}
__global__ void __launch_bounds__(1024) k_build_tile_lists(uint32_t *__restrict__ cost, uint32_t *__restrict__ lists, TileMeta *__restrict__ meta,
                                                           const TileMeta *__restrict__ previous, uint32_t n_tiles, const TileThresholds th) {
  __shared__ uint32_t s_warp[TILE_CLASSES][32];
  __shared__ uint32_t s_base[TILE_CLASSES];
  __shared__ unsigned long long s_sum[32];
  __shared__ uint32_t s_ran[32];
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t t = blockIdx.x * 1024u + tid;
  const bool inside = t < n_tiles;
  const uint32_t c = inside ? cost[t] : 0u;
  // (no launch before, or none of its tiles ran a pass: everything is class 0)
  const unsigned long long mean = previous->ran ? previous->sum / previous->ran : 0xFFFFFFFFull;
  const int k = inside ? tile_class(c, mean, th) : -1;
  uint32_t rank = 0;
#pragma unroll
  for (int q = 0; q < TILE_CLASSES; ++q) {
    const uint32_t b = __ballot_sync(FULL_MASK, k == q);
    if (k == q) rank = __popc(b & ((1u << lane) - 1u));
    if (lane == 0) s_warp[q][warp] = __popc(b);
  }
  unsigned long long sum = c;
  uint32_t ran = c != 0 ? 1u : 0u;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) { sum += __shfl_xor_sync(FULL_MASK, sum, off); ran += __shfl_xor_sync(FULL_MASK, ran, off); }
  if (lane == 0) { s_sum[warp] = sum; s_ran[warp] = ran; }
  __syncthreads();
  if (warp < TILE_CLASSES) {  // warp q: the CTA's total of class q -> its range in the list (one atomic), the warps' offsets in place
    const uint32_t mine = s_warp[warp][lane];
    uint32_t incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t up = __shfl_up_sync(FULL_MASK, incl, off);
      if (lane >= static_cast<uint32_t>(off)) incl += up;
    }
    s_warp[warp][lane] = incl - mine;
    if (lane == 31) s_base[warp] = incl ? atomicAdd(&meta->count[warp], incl) : 0u;
  } else if (warp == TILE_CLASSES) {
    unsigned long long a = s_sum[lane];
    uint32_t r = s_ran[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(FULL_MASK, a, off); r += __shfl_xor_sync(FULL_MASK, r, off); }
    if (lane == 0 && r) { atomicAdd(&meta->sum, a); atomicAdd(&meta->ran, static_cast<unsigned long long>(r)); }
  }
  __syncthreads();
  if (inside) {
    lists[static_cast<uint64_t>(k) * n_tiles + s_base[k] + s_warp[k][warp] + rank] = t;
    cost[t] = 0u;
  }
}

}  // namespace

void launch_build_tile_lists(uint32_t *d_cost, uint32_t *d_lists, TileMeta *d_meta, const TileMeta *d_previous, uint32_t n_tiles, cudaStream_t stream) {
  // longest pass more than 1.5x / 2x / 3x / 4x the mean pass (RDN_TILE_THRESHOLDS: quarters of the mean, ascending — measurement knob)
  static const TileThresholds th = []() {
    TileThresholds t;
    for (uint32_t &q : t.quarters) q = 0xFFFFFFFFu;
    const uint32_t standard[4] = {6, 8, 12, 16};
    for (int i = 0; i < 4; ++i) t.quarters[i] = standard[i];
    if (const char *e = getenv("RDN_TILE_THRESHOLDS")) {
      for (uint32_t &q : t.quarters) q = 0xFFFFFFFFu;
      int i = 0;
      for (const char *p = e; *p && i < TILE_CLASSES - 1; ++i) {
        t.quarters[i] = static_cast<uint32_t>(strtoul(p, const_cast<char **>(&p), 10));
        while (*p == ',' || *p == ' ') ++p;
      }
    }
    return t;
  }();
  if (n_tiles) k_build_tile_lists<<<(n_tiles + 1023u) / 1024u, 1024, 0, stream>>>(d_cost, d_lists, d_meta, d_previous, n_tiles, th);
}

// ------------------------------------------------------------------------------------------------ launchers
void launch_trace_reference(const SceneDev &scene, const rdn_launch &launch, const rdn_ray *d_rays, uint64_t n, rdn_hit *d_hits,
                            const TraceScratch &scratch, bool count_visits, int sm_count, cudaStream_t stream, const unsigned long long *d_n) {
  if (n == 0) return;
  uint32_t tiles_x = 0, width = 0, height = 0;
  uint64_t n_fetch = n;
  static const bool tiles_enabled = []() { const char *e = getenv("RDN_REF_TILES"); return !e || atoi(e) != 0; }();
  if (tiles_enabled && !d_n && launch.grid_width != 0 && n % launch.grid_width == 0) {
    width = launch.grid_width;
    height = static_cast<uint32_t>(n / launch.grid_width);
    tiles_x = (width + 7u) / 8u;
    n_fetch = static_cast<uint64_t>(tiles_x) * ((height + 3u) / 4u) * 32u;
  }
  const int block = 128;
  uint64_t blocks = (n_fetch + block - 1) / block;
  const uint64_t cap = static_cast<uint64_t>(sm_count) * 1024;
  if (blocks > cap) blocks = cap;
  if (count_visits)
    k_trace_reference<true><<<static_cast<unsigned>(blocks), block, 0, stream>>>(scene, launch, d_rays, n, d_hits, scratch, tiles_x, width, height, n_fetch, d_n);
  else
    k_trace_reference<false><<<static_cast<unsigned>(blocks), block, 0, stream>>>(scene, launch, d_rays, n, d_hits, scratch, tiles_x, width, height, n_fetch, d_n);
}

void launch_resolve_ties(const SceneDev &scene, const rdn_launch &launch, const rdn_ray *d_rays, rdn_hit *d_hits,
                         const TraceScratch &scratch, int sm_count, cudaStream_t stream) {
  k_resolve_ties<<<static_cast<unsigned>(sm_count) * 16u, 128, 0, stream>>>(scene, launch, d_rays, d_hits, scratch);
}

static int ordered_variant() {  // (read at every launch: the parity tests walk the variants inside one process)
  const char *e = getenv("RDN_ORDERED_VARIANT");
  return e ? atoi(e) : 0;
}
int ordered_tie_mode() { return ordered_variant() == 9 ? 0 : 3; }

bool any_hit_can_end_search(const rdn_launch &launch, const rdn_anyhit_program *programs, uint32_t n_programs, const SbtHitGroup *groups,
                            uint32_t n_groups) {
  if (launch.any_hit == RDN_ANYHIT_NONE) return false;
  auto ends = [&](uint32_t k) { return k < n_programs && ((programs[k].behavior | programs[k].otherwise) & RDN_ANYHIT_BEHAVIOR_END_SEARCH) != 0; };
  if (launch.any_hit != RDN_ANYHIT_FROM_SBT) return ends(launch.any_hit - 1u);
  for (uint32_t g = 0; g < n_groups; ++g)
    if (ends(groups[g].any_hit)) return true;
  return false;
}

cudaError_t launch_trace_ordered(const SceneDev &scene, const rdn_launch &launch, const TlasRoot &tlas, const rdn_ray *d_rays, uint64_t n,
                                 rdn_hit *d_hits, const TraceScratch &scratch, int sm_count, cudaStream_t stream, bool allow_overlap,
                                 uint32_t wait_epoch, bool *ties_resolved_in_kernel, const unsigned long long *d_n,
                                 TileHistory *history) {
  *ties_resolved_in_kernel = true;
  if (n == 0) return cudaSuccess;
  OrderedParams P;
  P.S = scene; P.L = launch; P.rays = d_rays; P.hits = d_hits; P.n = n; P.scratch = scratch;
  P.tile_lists = nullptr; P.tile_meta = nullptr; P.tile_cost = nullptr; P.tile_meta_clear = nullptr; P.n_tiles = 0;
  if (history) history->used = false;
  P.tiles_x = 0; P.width = 0; P.height = 0; P.n_fetch = n; P.n_ptr = d_n;
  if (!d_n && launch.grid_width != 0 && n % launch.grid_width == 0) {
    P.width = launch.grid_width;
    P.height = static_cast<uint32_t>(n / launch.grid_width);
    P.tiles_x = (P.width + 7u) / 8u;
    P.n_fetch = static_cast<uint64_t>(P.tiles_x) * ((P.height + 3u) / 4u) * 32u;
  }
  const int variant_ = ordered_variant();
  const bool irregular_tlas = tlas.irregular_count != 0 && tlas.irregular_count != IRREGULAR_ROUTE_ALL;
  // the four-box walk needs the four-box view (emitted only when the scene was built under the same variant) and a regular TLAS
  const bool use_wide4 = (variant_ == 60 || variant_ == 61) && !irregular_tlas && tlas.wide4_root != REF_EMPTY;
  P.world_root = use_wide4 ? tlas.wide4_root : tlas.wide_root;
  P.hot_a_base = tlas.wide_root == REF_EMPTY ? 0u : tlas.wide_root;
  P.hot_a_count = tlas.wide_root == REF_EMPTY ? 0u : tlas.hot_count;
  P.hot_b_base = tlas.hot_geometry_base;
  P.hot_b_count = tlas.hot_geometry_count;
  P.irregular_start = tlas.irregular_start;
  P.irregular_count = tlas.irregular_count == IRREGULAR_ROUTE_ALL ? 0u : tlas.irregular_count;  // (the caller routes those elsewhere)
  P.wait_epoch = wait_epoch;
  {  // (RDN_SHARE=busy,min: experimentation knob)
    P.share_busy = 16; P.share_min = 2; P.share_after = 30;  // (sweeps: profiles/kbench_r2u_*.log, kbench_r2v_*.log)
    if (const char *e = getenv("RDN_SHARE")) sscanf(e, "%d,%d,%d", &P.share_busy, &P.share_min, &P.share_after);
    P.topup_lanes = 4;
    if (const char *e = getenv("RDN_TOPUP")) P.topup_lanes = atoi(e);
  }

  // RDN_ORDERED_VARIANT: experimentation knob
  const int variant = ordered_variant();
  using KernelFn = void (*)(const OrderedParams);
  KernelFn fn;
  bool inline_ties = true;
  // template arguments: <K, MINB, DRAIN_TIES, IRREGULAR, LD256, HOT, WIDE4, INST_LOOP, SHARE>
  const bool any_hit = launch.any_hit != RDN_ANYHIT_NONE;  // (launches whose any-hit stage can END_SEARCH never get here)
  const KernelFn plain = any_hit ? k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_NEVER, true>
                                 : k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_NEVER>;
  const KernelFn sharing = any_hit ? k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_ALWAYS, true>
                                   : k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_ALWAYS>;
  const KernelFn late = any_hit ? k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_LATE, true>
                                : k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_LATE>;
  const KernelFn topup = any_hit ? late : k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_NEVER, false, false, true>;
  static const uint32_t topup_grid_instances = []() { const char *e = getenv("RDN_TOPUP_GRID_INSTANCES"); return e ? static_cast<uint32_t>(strtoul(e, nullptr, 10)) : 256u; }();
  const bool grid_topup = topup_grid_instances != 0 && scene.n_instances >= topup_grid_instances;
  if (P.tiles_x != 0 && grid_topup && !getenv("RDN_TOPUP")) P.topup_lanes = 8;
  const KernelFn late_topup = any_hit ? late : k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_LATE, false, false, true>;
  const bool mid_list = P.tiles_x == 0 && n >= MID_LIST_RAYS && n < LONG_LIST_RAYS;
  if (mid_list && !getenv("RDN_TOPUP")) P.topup_lanes = 8;
  switch (any_hit ? 0 : variant) {  // (the A/B instantiations exist without the any-hit stage only)
    case 2: fn = k_trace_ordered_rounds<2, 8, true, false, true, false, false>; break;     // the round-1 default: K = 2, one round per missed instance
    case 9: fn = k_trace_ordered_rounds<3, 8, false, false, true, false, false, true>; inline_ties = false; break;  // queue drained by k_resolve_ties
    case 30: fn = k_trace_ordered_rounds<3, 8, true, false, false, false, false, true>; break;  // 128-bit loads / stores
#ifndef RDN_SIMT_EMU
    case 40: fn = k_trace_ordered_rounds<3, 8, true, false, true, true, false, true>; break;    // top levels staged in shared memory (TMA)
#endif
    case 60: fn = k_trace_ordered_rounds<2, 8, true, false, true, false, true>; break;    // four-box nodes, K = 2 steps per round
    case 61: fn = k_trace_ordered_rounds<1, 8, true, false, true, false, true>; break;    // ... one step per round
    case 100: fn = plain; break;     // never share
    case 110: fn = sharing; break;   // the sharing loop from the first round
    case 120: fn = late; break;      // the plain loop, then the sharing loop for passes that have grown old
    case 140: fn = late_topup; break;  // top-up while the list lasts, then late sharing
    case 130: fn = topup; break;     // the plain loop, thinned-out warps topped up with new rays (whatever the launch: grids lose their tiles)
    // ray lists share (from the 30th round of a pass on) while they are short enough for their longest rays to decide the launch:
    // +25 % at 0.4 M rays, +7 % at 0.8 M, even at 2 M, -4.5 % at 4 M, -7 % at 8 M (profiles/kbench_r3p_list_sizes.log); grids do not
    // ... and lists long enough for throughput to decide top their thinned-out warps up with new rays (no tile to keep together):
    // +13 % at 8 M rays and +5 % on the configs[4] frame with a threshold of 4 busy lanes; 8 lanes: +16 % on the first but -11 % on
    // the second, whose bounce wave is full of short rays (profiles/kbench_r3q_list_topup.log)
    // ... and so do grids over scenes of many instances, whose tiles fall apart anyway (rays of a tile enter different instances and
    // end at very different times: 11 of 32 lanes busy on config 4): 3,050 -> 3,480-3,510 Mrays/s serialised, 3,800 -> 4,550-4,580
    // back to back with a threshold of 4-8 lanes; the single-instance configs[1] loses 9.5 % to it (profiles/kbench_r3r_grid_topup.log).
    // The instance count is a stand-in for "tiles diverge", RDN_TOPUP_GRID_INSTANCES moves it (0: never).
    // Lists in between do both — top up while the list lasts, share work once it has run dry: 0.8 M rays +4 %, 2 M +13-15 % over
    // sharing alone (0.4 M: -7 %, which is where the lower bound comes from; profiles/kbench_r3v_mid_lists.log)
    default: fn = P.tiles_x != 0 ? (grid_topup ? topup : plain) : (n < MID_LIST_RAYS ? late : (n < LONG_LIST_RAYS ? late_topup : topup)); break;
  }
  if (history && fn == plain && !any_hit && P.tiles_x != 0 && P.irregular_count == 0 &&
      history->n_tiles == P.tiles_x * ((P.height + 3u) / 4u)) {  // a grid with a tile history (capi.cu)
    fn = k_trace_ordered_rounds<RDN_ORDERED_K, RDN_ORDERED_MINB, true, false, true, false, false, true, SHARE_NEVER, false, true>;
    P.tile_lists = history->lists; P.tile_meta = history->meta; P.tile_cost = history->cost; P.tile_meta_clear = history->meta_clear;
    P.n_tiles = history->n_tiles;
    history->used = true;
  }
  if ((variant == 60 || variant == 61) && !use_wide4) fn = plain;
  if (P.irregular_count != 0) {  // (the experimentation variants exist for regular scenes only)
    // Rays handed over at refill can be a large part of the launch, and the in-kernel drain claims entries through one CAS
    // cursor (fine for a handful of ties, 55 ms for 400 K entries): the queue of an irregular launch is walked by
    // k_resolve_ties, one thread per entry, right behind this kernel.
    inline_ties = false;
    fn = k_trace_ordered_rounds<3, 8, false, true, true, false, false, true, SHARE_NEVER>;
  }
  int blocks_per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, fn, ORDERED_BLOCK, 0);
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  uint64_t blocks = static_cast<uint64_t>(sm_count) * blocks_per_sm;
  const uint64_t needed = (P.n_fetch + ORDERED_BLOCK - 1) / ORDERED_BLOCK;
  if (blocks > needed) blocks = needed;
  // Tail overlap is only requested for grids that fill the GPU: such a grid needs every CTA slot, so launch k+2 cannot start
  // before launch k has left entirely, and two alternating scratch sets are enough (capi.cu).
  static const bool pdl_enabled = []() { const char *e = getenv("RDN_PDL"); return !e || atoi(e) != 0; }();
  const bool full_grid = blocks == static_cast<uint64_t>(sm_count) * blocks_per_sm;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(blocks));
  cfg.blockDim = dim3(ORDERED_BLOCK);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  if (allow_overlap && pdl_enabled && full_grid && inline_ties) { cfg.attrs = attr; cfg.numAttrs = 1; }
  *ties_resolved_in_kernel = inline_ties;
  return cudaLaunchKernelEx(&cfg, fn, P);
}

}  // namespace rdn
