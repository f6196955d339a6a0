// accel.cpp — BLAS/TLAS assembly and flattening.  Compiled with -ffp-contract=off.
#include "accel.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "../../include/rdn_rt.h"

namespace rdn {

// build_device.cu: 0 = built, 1 = not supported on the device, < 0 = error
int build_bvh_sah_device(const Box3 *boxes, uint64_t n, uint32_t n_buckets, const TreeBuildOption &option, int device, FlattenBVH &out,
                         std::string &err);

// ------------------------------------------------------------------------------------------------ matrices
// three-term cofactor row: x*(p*q - r*s) + y*(...) + z*(...), the shape of every entry of mat4.rs:82-100
static inline float cof3(float x, float p0, float q0, float r0, float s0, float y, float p1, float q1, float r1, float s1,
                         float z, float p2, float q2, float r2, float s2) {
  return x * (p0 * q0 - r0 * s0) + y * (p1 * q1 - r1 * s1) + z * (p2 * q2 - r2 * s2);
}

static float mat4_det(const Mat4 &m) {
  // 24-term expansion in the order of mat4.rs:40-70
  return m.a1 * m.b2 * m.c3 * m.d4 - m.a1 * m.b2 * m.c4 * m.d3 + m.a1 * m.b3 * m.c4 * m.d2 - m.a1 * m.b3 * m.c2 * m.d4 +
         m.a1 * m.b4 * m.c2 * m.d3 - m.a1 * m.b4 * m.c3 * m.d2 - m.a2 * m.b3 * m.c4 * m.d1 + m.a2 * m.b3 * m.c1 * m.d4 -
         m.a2 * m.b4 * m.c1 * m.d3 + m.a2 * m.b4 * m.c3 * m.d1 - m.a2 * m.b1 * m.c3 * m.d4 + m.a2 * m.b1 * m.c4 * m.d3 +
         m.a3 * m.b4 * m.c1 * m.d2 - m.a3 * m.b4 * m.c2 * m.d1 + m.a3 * m.b1 * m.c2 * m.d4 - m.a3 * m.b1 * m.c4 * m.d2 +
         m.a3 * m.b2 * m.c4 * m.d1 - m.a3 * m.b2 * m.c1 * m.d4 - m.a4 * m.b1 * m.c2 * m.d3 + m.a4 * m.b1 * m.c3 * m.d2 -
         m.a4 * m.b2 * m.c3 * m.d1 + m.a4 * m.b2 * m.c1 * m.d3 - m.a4 * m.b3 * m.c1 * m.d2 + m.a4 * m.b3 * m.c2 * m.d1;
}

Mat4 mat4_inverse_or_identity(const Mat4 &m) {
  const float det = mat4_det(m);
  if (det == 0.0f) return Mat4{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  const float p = 1.0f / det;  // inv_det
  const float q = -p;          // `-inv_det * (...)`
  Mat4 r;
  r.a1 = p * cof3(m.b2, m.c3, m.d4, m.c4, m.d3, m.b3, m.c4, m.d2, m.c2, m.d4, m.b4, m.c2, m.d3, m.c3, m.d2);
  r.a2 = q * cof3(m.a2, m.c3, m.d4, m.c4, m.d3, m.a3, m.c4, m.d2, m.c2, m.d4, m.a4, m.c2, m.d3, m.c3, m.d2);
  r.a3 = p * cof3(m.a2, m.b3, m.d4, m.b4, m.d3, m.a3, m.b4, m.d2, m.b2, m.d4, m.a4, m.b2, m.d3, m.b3, m.d2);
  r.a4 = q * cof3(m.a2, m.b3, m.c4, m.b4, m.c3, m.a3, m.b4, m.c2, m.b2, m.c4, m.a4, m.b2, m.c3, m.b3, m.c2);
  r.b1 = q * cof3(m.b1, m.c3, m.d4, m.c4, m.d3, m.b3, m.c4, m.d1, m.c1, m.d4, m.b4, m.c1, m.d3, m.c3, m.d1);
  r.b2 = p * cof3(m.a1, m.c3, m.d4, m.c4, m.d3, m.a3, m.c4, m.d1, m.c1, m.d4, m.a4, m.c1, m.d3, m.c3, m.d1);
  r.b3 = q * cof3(m.a1, m.b3, m.d4, m.b4, m.d3, m.a3, m.b4, m.d1, m.b1, m.d4, m.a4, m.b1, m.d3, m.b3, m.d1);
  r.b4 = p * cof3(m.a1, m.b3, m.c4, m.b4, m.c3, m.a3, m.b4, m.c1, m.b1, m.c4, m.a4, m.b1, m.c3, m.b3, m.c1);
  r.c1 = p * cof3(m.b1, m.c2, m.d4, m.c4, m.d2, m.b2, m.c4, m.d1, m.c1, m.d4, m.b4, m.c1, m.d2, m.c2, m.d1);
  r.c2 = q * cof3(m.a1, m.c2, m.d4, m.c4, m.d2, m.a2, m.c4, m.d1, m.c1, m.d4, m.a4, m.c1, m.d2, m.c2, m.d1);
  r.c3 = p * cof3(m.a1, m.b2, m.d4, m.b4, m.d2, m.a2, m.b4, m.d1, m.b1, m.d4, m.a4, m.b1, m.d2, m.b2, m.d1);
  r.c4 = q * cof3(m.a1, m.b2, m.c4, m.b4, m.c2, m.a2, m.b4, m.c1, m.b1, m.c4, m.a4, m.b1, m.c2, m.b2, m.c1);
  r.d1 = q * cof3(m.b1, m.c2, m.d3, m.c3, m.d2, m.b2, m.c3, m.d1, m.c1, m.d3, m.b3, m.c1, m.d2, m.c2, m.d1);
  r.d2 = p * cof3(m.a1, m.c2, m.d3, m.c3, m.d2, m.a2, m.c3, m.d1, m.c1, m.d3, m.a3, m.c1, m.d2, m.c2, m.d1);
  r.d3 = q * cof3(m.a1, m.b2, m.d3, m.b3, m.d2, m.a2, m.b3, m.d1, m.b1, m.d3, m.a3, m.b1, m.d2, m.b2, m.d1);
  r.d4 = p * cof3(m.a1, m.b2, m.c3, m.b3, m.c2, m.a2, m.b3, m.c1, m.b1, m.c3, m.a3, m.b1, m.c2, m.b2, m.c1);
  return r;
}

float mat4_upper3_det(const Mat4 &m) {
  // Mat4::to_mat3().det(), mat3.rs:37-42
  const float t11 = m.c3 * m.b2 - m.b3 * m.c2;
  const float t12 = m.b3 * m.c1 - m.c3 * m.b1;
  const float t13 = m.c2 * m.b1 - m.b2 * m.c1;
  return m.a1 * t11 + m.a2 * t12 + m.a3 * t13;
}

static inline Vec3 transform_point(const Mat4 &m, float x, float y, float z) {
  // Mat4 * Vec3: expand with w = 1, divide by the resulting w (mat4.rs:140-168)
  const float rx = x * m.a1 + y * m.b1 + z * m.c1 + 1.0f * m.d1;
  const float ry = x * m.a2 + y * m.b2 + z * m.c2 + 1.0f * m.d2;
  const float rz = x * m.a3 + y * m.b3 + z * m.c3 + 1.0f * m.d3;
  const float rw = x * m.a4 + y * m.b4 + z * m.c4 + 1.0f * m.d4;
  return Vec3{rx / rw, ry / rw, rz / rw};
}

Box3 box_apply_matrix(const Box3 &b, const Mat4 &m) {
  // box3.rs:22-38: untouched when empty, else the bound of the 8 transformed corners (000,001,...,111)
  if ((b.max.x < b.min.x) || (b.max.y < b.min.y) || (b.max.z < b.min.z)) return b;
  Box3 r = box_empty();
  for (int corner = 0; corner < 8; ++corner) {
    const float x = (corner & 4) ? b.max.x : b.min.x;
    const float y = (corner & 2) ? b.max.y : b.min.y;
    const float z = (corner & 1) ? b.max.z : b.min.z;
    expand(r, transform_point(m, x, y, z));
  }
  return r;
}

// ------------------------------------------------------------------------------------------------ sources
uint32_t NaiveSahBvhSource::create_blas(std::vector<GeometrySource> source) {
  ++blas_epoch_;
  blas_data_.push_back(Blas{true, std::move(source)});
  return static_cast<uint32_t>(blas_data_.size() - 1);
}
uint32_t NaiveSahBvhSource::create_tlas(std::vector<InstanceSource> source) {
  tlas_data_.push_back(Tlas{true, std::move(source)});
  return static_cast<uint32_t>(tlas_data_.size() - 1);
}
bool NaiveSahBvhSource::delete_blas(uint32_t h) {
  if (h >= blas_data_.size()) return false;
  ++blas_epoch_;
  blas_data_[h] = Blas{};
  return true;
}
bool NaiveSahBvhSource::update_tlas(uint32_t h, std::vector<InstanceSource> source) {
  if (h >= tlas_data_.size() || !tlas_data_[h].alive) return false;
  tlas_data_[h].instances = std::move(source);
  return true;
}
bool NaiveSahBvhSource::delete_tlas(uint32_t h) {
  if (h >= tlas_data_.size()) return false;
  tlas_data_[h] = Tlas{};
  return true;
}

// ------------------------------------------------------------------------------------------------ flatten
static DeviceBVHNode to_device_node(const FlattenBVHNode &n, uint32_t hit, uint32_t miss, uint32_t next_offset,
                                    uint32_t primitive_offset) {
  // flatten_bvh_to_gpu_node, mod.rs:338-367
  DeviceBVHNode d;
  std::memset(&d, 0, sizeof(d));
  d.aabb_min[0] = n.bounding.min.x; d.aabb_min[1] = n.bounding.min.y; d.aabb_min[2] = n.bounding.min.z;
  d.aabb_max[0] = n.bounding.max.x; d.aabb_max[1] = n.bounding.max.y; d.aabb_max[2] = n.bounding.max.z;
  d.hit_next = hit == INVALID_NEXT ? INVALID_NEXT : hit + next_offset;
  d.miss_next = miss == INVALID_NEXT ? INVALID_NEXT : miss + next_offset;
  d.content_range[0] = static_cast<uint32_t>(n.primitive_start) + primitive_offset;
  d.content_range[1] = static_cast<uint32_t>(n.primitive_end) + primitive_offset;
  return d;
}

static void set_child(WideNode &w, int which, const Box3 *box, uint32_t ref) {
  const float nan = NAN;  // a NaN box fails every comparison of the slab test: never entered
  float *mn = which == 0 ? w.c0_min : w.c1_min;
  float *mx = which == 0 ? w.c0_max : w.c1_max;
  mn[0] = box ? box->min.x : nan; mn[1] = box ? box->min.y : nan; mn[2] = box ? box->min.z : nan;
  mx[0] = box ? box->max.x : nan; mx[1] = box ? box->max.y : nan; mx[2] = box ? box->max.z : nan;
  (which == 0 ? w.ref0 : w.ref1) = ref;
}

uint32_t emit_wide_nodes(const BigVector<FlattenBVHNode> &nodes, uint64_t slot_offset, BigVector<WideNode> &out,
                         bool &capacity_error, const TlasBounding *item_bounds) {
  if (nodes.empty() || nodes[0].primitive_end == nodes[0].primitive_start) return REF_EMPTY;
  // Wide index of every inner reference node, after the pseudo root: the top of the tree breadth first (pseudo root + up to
  // HOT_TOP_NODES - 1 inner nodes form one contiguous block: what every ray touches, and what the kernel's staging variant
  // copies into shared memory), everything below in the reference's pre-order (a parent next to its left subtree).
  const uint64_t base = out.size();
  std::vector<uint32_t> wide_of(nodes.size(), 0);
  uint64_t n_inner = 0;
  {
    std::vector<uint8_t> in_top(nodes.size(), 0);
    std::vector<size_t> queue{0};
    for (size_t head = 0; head < queue.size() && n_inner + 1 < HOT_TOP_NODES; ++head) {
      const size_t i = queue[head];
      if (!nodes[i].has_child) continue;
      in_top[i] = 1;
      wide_of[i] = static_cast<uint32_t>(base + 1 + n_inner++);
      queue.push_back(nodes[i].left_child_offset());
      queue.push_back(nodes[i].right_child_offset());
    }
    for (size_t i = 0; i < nodes.size(); ++i)
      if (nodes[i].has_child && !in_top[i]) wide_of[i] = static_cast<uint32_t>(base + 1 + n_inner++);
  }
  // Leaves that need nodes of their own — a fine TLAS leaf becomes a small subtree, a leaf longer than REF_LEAF_MAX_COUNT a chain —
  // get them behind the inner nodes, in the order the leaves are met (root, then every inner node's left and right child): the
  // ranges are laid out first so that the nodes can be written by all threads.
  auto own_nodes = [&](const FlattenBVHNode &leaf) -> uint64_t {
    uint64_t count = leaf.primitive_end - leaf.primitive_start;
    if (count == 0 || slot_offset + leaf.primitive_start + count > REF_LEAF_START_MASK) return 0;
    if (item_bounds && count > 1) return count - 1;
    uint64_t chain = 0;
    while (count > REF_LEAF_MAX_COUNT) { ++chain; count -= REF_LEAF_MAX_COUNT; }
    return chain;
  };
  std::vector<uint64_t> own_base(nodes.size(), 0);
  uint64_t own_total = 0;
  auto lay_out = [&](size_t idx) { if (!nodes[idx].has_child) { own_base[idx] = base + 1 + n_inner + own_total; own_total += own_nodes(nodes[idx]); } };
  lay_out(0);
  for (size_t i = 0; i < nodes.size(); ++i)
    if (nodes[i].has_child) { lay_out(nodes[i].left_child_offset()); lay_out(nodes[i].right_child_offset()); }
  out.resize(base + 1 + n_inner + own_total);
  std::atomic<bool> over_capacity{false};

  // a leaf reference
  auto leaf_ref = [&](size_t leaf_idx) -> uint32_t {
    const FlattenBVHNode &leaf = nodes[leaf_idx];
    uint64_t start = slot_offset + leaf.primitive_start;
    uint64_t count = leaf.primitive_end - leaf.primitive_start;
    if (count == 0) return REF_EMPTY;
    if (start + count > REF_LEAF_START_MASK) { over_capacity = true; return REF_EMPTY; }
    auto enc = [](uint64_t s, uint64_t c) { return REF_LEAF_BIT | (static_cast<uint32_t>(c - 1) << REF_LEAF_COUNT_SHIFT) | static_cast<uint32_t>(s); };
    uint64_t cursor = own_base[leaf_idx];  // next node of this leaf's own range
    if (item_bounds && count > 1) {
      // a binary tree over contiguous halves of the leaf's slots, single slots at the bottom; a node's boxes are the unions of
      // its halves' boxes (min / max are exact: the same boxes as the unions over the slots), handed up with the reference
      struct Built { uint32_t ref; Box3 box; };
      auto slot_box = [&](uint64_t s) {
        const TlasBounding &t = item_bounds[s - slot_offset];
        return Box3{Vec3{t.world_min[0], t.world_min[1], t.world_min[2]}, Vec3{t.world_max[0], t.world_max[1], t.world_max[2]}};
      };
      auto subtree_impl = [&](auto &&self, uint64_t s, uint64_t c) -> Built {
        if (c == 1) return Built{enc(s, 1), slot_box(s)};
        const uint64_t cl = (c + 1) / 2;
        const uint64_t idx = cursor++;
        const Built l = self(self, s, cl), r = self(self, s + cl, c - cl);
        WideNode w;
        std::memset(&w, 0, sizeof(w));
        set_child(w, 0, &l.box, l.ref);
        set_child(w, 1, &r.box, r.ref);
        out[idx] = w;
        Box3 b = l.box;
        expand(b, r.box);  // (a NaN operand loses: the box of an instance nothing can hit does not swallow its neighbours')
        return Built{static_cast<uint32_t>(idx), b};
      };
      return subtree_impl(subtree_impl, start, count).ref;
    }
    if (count <= REF_LEAF_MAX_COUNT) return enc(start, count);
    // chain: node k = {first 16 slots, rest}
    const uint32_t head = static_cast<uint32_t>(cursor);
    while (count > REF_LEAF_MAX_COUNT) {
      WideNode w;
      std::memset(&w, 0, sizeof(w));
      const uint64_t rest = count - REF_LEAF_MAX_COUNT;
      const uint32_t next = rest > REF_LEAF_MAX_COUNT ? static_cast<uint32_t>(cursor + 1) : enc(start + REF_LEAF_MAX_COUNT, rest);
      set_child(w, 0, &leaf.bounding, enc(start, REF_LEAF_MAX_COUNT));
      set_child(w, 1, &leaf.bounding, next);
      out[cursor++] = w;
      start += REF_LEAF_MAX_COUNT;
      count = rest;
    }
    return head;
  };
  auto child_ref = [&](size_t idx) -> uint32_t { return nodes[idx].has_child ? wide_of[idx] : leaf_ref(idx); };

  // pseudo root: the reference tests the root's own box before anything else
  {
    WideNode w;
    std::memset(&w, 0, sizeof(w));
    const uint32_t r = child_ref(0);
    set_child(w, 0, &nodes[0].bounding, r);
    set_child(w, 1, nullptr, REF_EMPTY);
    out[base] = w;
  }
  parallel_for(nodes.size(), 1024, [&](uint64_t i0, uint64_t i1) {
    for (size_t i = i0; i < i1; ++i) {
      if (!nodes[i].has_child) continue;
      const size_t l = nodes[i].left_child_offset(), r = nodes[i].right_child_offset();
      WideNode w;
      std::memset(&w, 0, sizeof(w));
      const uint32_t lr = child_ref(l), rr = child_ref(r);
      set_child(w, 0, &nodes[l].bounding, lr);
      set_child(w, 1, &nodes[r].bounding, rr);
      out[wide_of[i]] = w;
    }
  });
  if (over_capacity) capacity_error = true;
  if (out.size() >= REF_SPECIAL) capacity_error = true;
  return static_cast<uint32_t>(base);
}

uint32_t emit_wide4_nodes(const BigVector<FlattenBVHNode> &nodes, uint64_t slot_offset, BigVector<Wide4Node> &out,
                          bool &capacity_error) {
  if (nodes.empty() || nodes[0].primitive_end == nodes[0].primitive_start) return REF_EMPTY;
  const uint64_t base = out.size();
  auto set = [](Wide4Node &w, int k, const Box3 *box, uint32_t ref) {
    const float nan = NAN;
    Wide4Node::Child &c = w.child[k];
    c.bmin[0] = box ? box->min.x : nan; c.bmin[1] = box ? box->min.y : nan; c.bmin[2] = box ? box->min.z : nan;
    c.bmax[0] = box ? box->max.x : nan; c.bmax[1] = box ? box->max.y : nan; c.bmax[2] = box ? box->max.z : nan;
    c.ref = ref; c.pad = 0;
  };
  auto blank = [&](Wide4Node &w) { for (int k = 0; k < 4; ++k) set(w, k, nullptr, REF_EMPTY); };
  // which inner nodes own a 4-wide node: the root, and every inner grandchild (or inner child of a leaf-sibling... no: an inner
  // CHILD is always absorbed into its parent's node; only its children surface) — found by a walk from the root
  std::vector<uint32_t> wide_of(nodes.size(), 0);
  std::vector<size_t> owners;
  {
    std::vector<size_t> stack;
    if (nodes[0].has_child) stack.push_back(0);
    while (!stack.empty()) {
      const size_t i = stack.back();
      stack.pop_back();
      owners.push_back(i);
      for (size_t c : {static_cast<size_t>(nodes[i].left_child_offset()), static_cast<size_t>(nodes[i].right_child_offset())}) {
        if (!nodes[c].has_child) continue;
        for (size_t g : {static_cast<size_t>(nodes[c].left_child_offset()), static_cast<size_t>(nodes[c].right_child_offset())})
          if (nodes[g].has_child) stack.push_back(g);
      }
    }
    std::sort(owners.begin(), owners.end());  // pre-order of the reference tree
    for (size_t k = 0; k < owners.size(); ++k) wide_of[owners[k]] = static_cast<uint32_t>(base + 1 + k);
  }
  out.resize(base + 1 + owners.size());
  auto leaf_ref = [&](const FlattenBVHNode &leaf) -> uint32_t {
    uint64_t start = slot_offset + leaf.primitive_start;
    uint64_t count = leaf.primitive_end - leaf.primitive_start;
    if (count == 0) return REF_EMPTY;
    if (start + count > REF_LEAF_START_MASK) { capacity_error = true; return REF_EMPTY; }
    auto enc = [](uint64_t s, uint64_t c) { return REF_LEAF_BIT | (static_cast<uint32_t>(c - 1) << REF_LEAF_COUNT_SHIFT) | static_cast<uint32_t>(s); };
    if (count <= REF_LEAF_MAX_COUNT) return enc(start, count);
    // chain of nodes that repeat the leaf's box: up to three 16-slot groups per node, the rest behind the fourth reference
    const uint32_t head = static_cast<uint32_t>(out.size());
    while (count > 0) {
      Wide4Node w;
      blank(w);
      int k = 0;
      for (; k < 3 && count > 0; ++k) {
        const uint64_t take = std::min<uint64_t>(count, REF_LEAF_MAX_COUNT);
        set(w, k, &leaf.bounding, enc(start, take));
        start += take; count -= take;
      }
      if (count > 0) {
        if (count <= REF_LEAF_MAX_COUNT) { set(w, 3, &leaf.bounding, enc(start, count)); start += count; count = 0; }
        else set(w, 3, &leaf.bounding, static_cast<uint32_t>(out.size() + 1));
      }
      out.push_back(w);
    }
    return head;
  };
  auto ref_of = [&](size_t idx) -> uint32_t { return nodes[idx].has_child ? wide_of[idx] : leaf_ref(nodes[idx]); };
  {
    Wide4Node w;
    blank(w);
    set(w, 0, &nodes[0].bounding, ref_of(0));  // pseudo root: the root's own box is tested first, as the reference does
    out[base] = w;
  }
  for (size_t i : owners) {
    Wide4Node w;
    blank(w);
    int k = 0;
    for (size_t c : {static_cast<size_t>(nodes[i].left_child_offset()), static_cast<size_t>(nodes[i].right_child_offset())}) {
      if (!nodes[c].has_child) { const uint32_t r = ref_of(c); set(w, k++, &nodes[c].bounding, r); continue; }
      for (size_t g : {static_cast<size_t>(nodes[c].left_child_offset()), static_cast<size_t>(nodes[c].right_child_offset())}) {
        const uint32_t r = ref_of(g);
        set(w, k++, &nodes[g].bounding, r);
      }
    }
    out[wide_of[i]] = w;
  }
  if (out.size() >= REF_SPECIAL) capacity_error = true;
  return static_cast<uint32_t>(base);
}

static TriRecord make_tri_record(Vec3 v0, Vec3 v1, Vec3 v2) {
  // the ray-independent part of intersect_ray_triangle_cpu (geometry/mod.rs:115-148), same operation order
  const Vec3 e1 = v1 - v0;
  const Vec3 e2 = v2 - v0;
  const Vec3 normal = normalize(cross(e1, e2));
  const float uu = dot(e1, e1);
  const float uv = dot(e1, e2);
  const float vv = dot(e2, e2);
  const float inverse_d = 1.0f / (uv * uv - uu * vv);
  TriRecord t;
  t.n[0] = normal.x; t.n[1] = normal.y; t.n[2] = normal.z; t.inv_d = inverse_d;
  t.v0[0] = v0.x; t.v0[1] = v0.y; t.v0[2] = v0.z; t.uu = uu;
  t.e1[0] = e1.x; t.e1[1] = e1.y; t.e1[2] = e1.z; t.uv = uv;
  t.e2[0] = e2.x; t.e2[1] = e2.y; t.e2[2] = e2.z; t.vv = vv;
  return t;
}

// ------------------------------------------------------------------------------------------------ regularity
// The ordered kernel visits near children first and prunes against its own closest hit, so it agrees with the reference's
// pre-order walk only where every accepted hit lies inside the boxes that lead to it (up to rounding, which the TIE_EPS
// slack of the kernel absorbs).  The reference's result is traversal-order dependent where that fails:
//   * needle triangles: uv*uv - uu*vv cancels to noise, the barycentric test passes for points far outside the triangle
//     (and its leaf box); the reference then finds or prunes them depending on what it met first;
//   * instances whose world box does not bound their hits: singular transform (inverse_or_identity gives the identity while
//     the box uses the singular matrix), non-finite or projective matrices, inverses that do not round-trip, and the
//     reference's own blas_box indexing (one entry per GEOMETRY, looked up by BLAS handle, mod.rs:239,273), which hands an
//     instance the box of another BLAS's geometry once an earlier BLAS has more or fewer than one geometry.
// Such triangles / instances are "irregular".  Rays whose ORIGINAL range meets the world box of an irregular instance — a
// superset of the rays whose reference walk can enter it — are walked in the reference's order (unclamped) instead.
// sin^2 of the angle between the edges below which a triangle counts as a needle.  uv*uv - uu*vv carries a relative error of
// about 2e-7 / sin^2, which is also how far beyond its far edge (in units of its long edge) the barycentric test still accepts:
// 2 % at this threshold, shrinking fast above it (thin but honest triangles, e.g. next to the poles of a finely tessellated
// UV sphere at sin^2 ~ 1e-4, stay regular); below it the determinant is mostly noise.
constexpr double NEEDLE_SIN2 = 1e-5;
constexpr float ROUND_TRIP_TOLERANCE = 1e-3f;  // |Minv (M c) - c| allowed, relative to the box diagonal

static bool finite3(const float *v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

static bool triangle_is_irregular(const TriRecord &t) {
  // zero normal (exactly degenerate, e.g. the pole triangles of a UV sphere): b = 0, a = -0, t = NaN, which the range asserts
  // reject in every traversal order — harmless
  if (t.n[0] == 0.0f && t.n[1] == 0.0f && t.n[2] == 0.0f) return false;
  if (!finite3(t.n) || !finite3(t.v0) || !finite3(t.e1) || !finite3(t.e2) || !std::isfinite(t.inv_d)) return true;
  const double area2 = static_cast<double>(t.uu) * t.vv - static_cast<double>(t.uv) * t.uv;  // |e1 x e2|^2
  return !(area2 >= NEEDLE_SIN2 * static_cast<double>(t.uu) * t.vv);
}

static bool box_is_empty(const Box3 &b) { return (b.max.x < b.min.x) || (b.max.y < b.min.y) || (b.max.z < b.min.z); }
static bool box_contains(const Box3 &outer, const Box3 &inner) {
  return outer.min.x <= inner.min.x && outer.min.y <= inner.min.y && outer.min.z <= inner.min.z && outer.max.x >= inner.max.x &&
         outer.max.y >= inner.max.y && outer.max.z >= inner.max.z;
}
static float box_diagonal(const Box3 &b) { return length(b.max - b.min); }

// `used_box`: the object-space box the reference transforms into the instance's world box; `true_box`: the bound of the
// triangles the instance really holds
static bool instance_is_irregular(const Mat4 &m, const Mat4 &inv, const Box3 &used_box, const Box3 &true_box) {
  if (box_is_empty(true_box)) return false;  // nothing to hit
  if (box_is_empty(used_box) || !box_contains(used_box, true_box)) return true;
  const float *mp = &m.a1, *ip = &inv.a1;
  for (int i = 0; i < 16; ++i)
    if (!std::isfinite(mp[i]) || !std::isfinite(ip[i])) return true;
  if (mat4_det(m) == 0.0f) return true;                                                    // inverse_or_identity fell back
  if (m.a4 != 0.0f || m.b4 != 0.0f || m.c4 != 0.0f || m.d4 != 1.0f) return true;          // the ray transform is affine only
  const float object_diag = box_diagonal(used_box);
  Box3 world = box_empty();
  Vec3 corners[8], images[8];
  for (int corner = 0; corner < 8; ++corner) {
    corners[corner] = Vec3{(corner & 4) ? used_box.max.x : used_box.min.x, (corner & 2) ? used_box.max.y : used_box.min.y,
                           (corner & 1) ? used_box.max.z : used_box.min.z};
    images[corner] = transform_point(m, corners[corner].x, corners[corner].y, corners[corner].z);
    expand(world, images[corner]);
  }
  const float world_diag = box_diagonal(world);
  for (int corner = 0; corner < 8; ++corner) {
    const Vec3 back = transform_point(inv, images[corner].x, images[corner].y, images[corner].z);
    const Vec3 again = transform_point(m, back.x, back.y, back.z);
    if (!(length(back - corners[corner]) <= ROUND_TRIP_TOLERANCE * object_diag)) return true;
    if (!(length(again - images[corner]) <= ROUND_TRIP_TOLERANCE * world_diag)) return true;
  }
  return false;
}

int NaiveSahBvhSource::build(const std::vector<uint32_t> &tlas_binding, FlatScene &out, std::string &err, FlatScene *previous, bool *reused) const {
  out = FlatScene{};
  if (reused) *reused = false;
  bool capacity_error = false;
  using Clock = std::chrono::steady_clock;
  const auto t_begin = Clock::now();
  double bvh_ms = 0.0;
  const bool timing = getenv("RDN_BUILD_TIMING") != nullptr;
  // the four-box view is only emitted for the kernel experiment that walks it (RDN_ORDERED_VARIANT=60/61): +45 % blob otherwise unused
  const char *variant_env = getenv("RDN_ORDERED_VARIANT");
  const bool want_wide4 = variant_env && (atoi(variant_env) == 60 || atoi(variant_env) == 61);
  // TLAS leaves of several instances (the reference bins up to ten) become small subtrees of the wide view: a data-only change of
  // what the ordered kernel walks (issue model, config 4: -19 % warp instructions).  RDN_FINE_TLAS=0 restores the multi-slot leaves.
  const char *fine_env = getenv("RDN_FINE_TLAS");
  const bool fine_tlas = !fine_env || atoi(fine_env) != 0;
  // TLAS-only rebuild: the BLAS arrays of the previous flattened scene are still right when no BLAS was created or deleted since
  const bool reuse_blas = previous && cache_.blas_epoch == blas_epoch_ && cache_.want_wide4 == want_wide4 && cache_.fine_tlas == fine_tlas &&
                          previous->wide_nodes.size() >= cache_.wide_nodes_n && previous->blas_meta.size() == blas_data_.size();
  if (reuse_blas) {
    out = std::move(*previous);
    out.tlas_root.clear(); out.tlas_bvh_forest.clear(); out.tlas_bounding.clear(); out.instances.clear(); out.irregular_instances.clear();
    out.wide_nodes.resize(cache_.wide_nodes_n);
    out.wide4_nodes.resize(cache_.wide4_nodes_n);
    out.stats = cache_.stats;
    if (reused) *reused = true;
  }
  out.tlas_binding = tlas_binding;
  double ms_boxes = 0, ms_records = 0, ms_wide = 0, ms_threaded = 0, ms_leaves = 0;
  auto since = [](Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); };
  auto timed_build = [&](const Box3 *boxes, uint64_t n, BVHBuildStrategy &strategy, const TreeBuildOption &option) {
    const auto t0 = Clock::now();
    FlattenBVH bvh;
    bool on_device = false;
    if (build_device >= 0 && n >= device_build_min) {
      if (const SAH *sah = dynamic_cast<const SAH *>(&strategy)) {
        std::string device_err;
        on_device = build_bvh_sah_device(boxes, n, sah->bucket_count(), option, build_device, bvh, device_err) == 0;
        if (on_device) out.stats.device_built_trees++;
      }
    }
    if (!on_device) bvh = FlattenBVH::build(boxes, n, strategy, option);  // (also reports the reference's build errors)
    bvh_ms += std::chrono::duration<double, std::milli>(Clock::now() - t0).count();
    if (bvh.stats.build_threads > out.stats.build_threads) out.stats.build_threads = bvh.stats.build_threads;
    return bvh;
  };

  // ---- build_blas (mod.rs:122-260).  NOTE blas_box gets one entry PER GEOMETRY of a live BLAS but one per
  // deleted BLAS, and build_tlas indexes it by BLAS handle (mod.rs:239,273) — reproduced as is.
  std::vector<OptBox> blas_box;
  std::vector<Box3> blas_true_box;       // per BLAS HANDLE: bound of all its triangle geometries
  std::vector<uint32_t> blas_irregular;  // per BLAS handle: BlasMeta::irregular_leaf_count
  std::vector<HotBlock> blas_hot;        // per BLAS handle: the top block of its largest geometry tree
  uint64_t n_indices_total = 0;  // geometry_indices.len()
  const TreeBuildOption blas_option{50, 2};

  if (reuse_blas) {
    blas_box = cache_.blas_box; blas_true_box = cache_.blas_true_box; blas_irregular = cache_.blas_irregular; blas_hot = cache_.blas_hot;
  } else {
    for (const Blas &blas : blas_data_) {
      if (!blas.alive) {
        out.blas_meta.push_back(BlasMeta{{0, 0}, 0, 0});
        blas_box.push_back(OptBox{false, box_empty()});
        blas_true_box.push_back(box_empty());
        blas_irregular.push_back(0);
        blas_hot.push_back(HotBlock{});
        continue;
      }
      HotBlock hot;
      const uint32_t tri_start = static_cast<uint32_t>(out.geometry_meta.size());
      Box3 true_box = box_empty();
      const uint32_t leaf_start = static_cast<uint32_t>(out.irregular_leaf_boxes.size());
      for (size_t g = 0; g < blas.geometries.size(); ++g) {
        const GeometrySource &src = blas.geometries[g];
        Box3 root_box = box_empty();
        if (!src.is_aabbs) {
          const uint32_t primitive_start = static_cast<uint32_t>(n_indices_total / 3);
          const uint64_t n_idx = src.has_indices ? src.indices.size() : src.positions.size();
          const uint64_t n_tri = n_idx / 3;  // as_chunks::<3>().0 drops the remainder
          auto vertex_of = [&](uint64_t tri, int k) -> uint64_t { return src.has_indices ? src.indices[3 * tri + k] : 3 * tri + k; };
          auto t_phase = Clock::now();
          std::vector<Box3> boxes(n_tri);
          std::atomic<bool> index_out_of_bounds{false};
          parallel_for(n_tri, PARALLEL_BUILD_MIN, [&](uint64_t t_begin_, uint64_t t_end_) {
            for (uint64_t t = t_begin_; t < t_end_; ++t) {
              Box3 b = box_empty();
              for (int k = 0; k < 3; ++k) {
                const uint64_t vi = vertex_of(t, k);
                if (vi >= src.positions.size()) { index_out_of_bounds = true; return; }
                expand(b, src.positions[vi]);
              }
              boxes[t] = b;
            }
          });
          if (index_out_of_bounds) { err = "triangle index out of bounds (the reference panics here)"; return RDN_ERR_BUILD; }
          ms_boxes += since(t_phase);
          SAH sah(4);
          FlattenBVH bvh = timed_build(boxes.data(), n_tri, sah, blas_option);
          if (bvh.stats.bucket_out_of_range) { err = "SAH bucket index out of range (the reference panics here)"; return RDN_ERR_BUILD; }
          out.stats.balance_fallbacks += bvh.stats.balance_fallbacks;
          out.stats.balance_fallbacks_gt10 += bvh.stats.balance_fallbacks_gt10;
          expand(root_box, bvh.nodes[0].bounding);
          if (n_tri) expand(true_box, bvh.nodes[0].bounding);
          t_phase = Clock::now();
          const auto next = compute_bvh_next(bvh.nodes);
          ms_threaded += since(t_phase);
          t_phase = Clock::now();

          // slots of this geometry start at primitive_start (indices_redirect and indices grow in lock step)
          const uint64_t slot_base = out.triangles.size();
          if (slot_base != primitive_start) { err = "internal: slot/primitive offset mismatch"; return RDN_ERR_BUILD; }
          out.prim_to_slot.resize(slot_base + n_tri, 0u);
          out.triangles.resize(slot_base + n_tri);
          out.slot_info.resize(slot_base + n_tri);
          std::vector<uint8_t> slot_irregular(n_tri, 0);
          parallel_for(n_tri, PARALLEL_BUILD_MIN, [&](uint64_t k_begin, uint64_t k_end) {
            for (uint64_t k = k_begin; k < k_end; ++k) {
              const uint64_t tri = bvh.sorted_primitive_index[k];  // indices_redirect[slot] - raw_primitive_start
              out.prim_to_slot[slot_base + tri] = static_cast<uint32_t>(slot_base + k);
              out.triangles[slot_base + k] = make_tri_record(src.positions[vertex_of(tri, 0)], src.positions[vertex_of(tri, 1)],
                                                             src.positions[vertex_of(tri, 2)]);
              out.slot_info[slot_base + k] = SlotInfo{static_cast<uint32_t>(tri), static_cast<uint32_t>(g)};
              slot_irregular[k] = triangle_is_irregular(out.triangles[slot_base + k]) ? 1 : 0;
            }
          });
          for (uint64_t k = 0; k < n_tri; ++k) out.stats.irregular_triangles += slot_irregular[k];
          ms_records += since(t_phase);
          t_phase = Clock::now();
          n_indices_total += n_tri * 3;
          // the boxes of the leaves that hold an irregular triangle: the reference can test such a triangle only after this box test
          for (const FlattenBVHNode &node : bvh.nodes) {
            if (node.has_child) continue;
            bool any = false;
            for (uint64_t k = node.primitive_start; k < node.primitive_end && !any; ++k) any = slot_irregular[k] != 0;
            if (!any) continue;
            LeafBox lb;
            std::memset(&lb, 0, sizeof(lb));
            lb.bmin[0] = node.bounding.min.x; lb.bmin[1] = node.bounding.min.y; lb.bmin[2] = node.bounding.min.z;
            lb.bmax[0] = node.bounding.max.x; lb.bmax[1] = node.bounding.max.y; lb.bmax[2] = node.bounding.max.z;
            out.irregular_leaf_boxes.push_back(lb);
          }

          ms_leaves += since(t_phase);
          t_phase = Clock::now();
          const uint32_t bvh_start = static_cast<uint32_t>(out.tri_bvh_forest.size());
          GeometryMeta gm;
          std::memset(&gm, 0, sizeof(gm));
          gm.bvh_root_idx = bvh_start;
          gm.geometry_idx = static_cast<uint32_t>(g);
          gm.primitive_start = primitive_start;
          gm.geometry_flags = src.flags;
          gm.wide_root = emit_wide_nodes(bvh.nodes, primitive_start, out.wide_nodes, capacity_error);
          gm.wide4_root = want_wide4 ? emit_wide4_nodes(bvh.nodes, primitive_start, out.wide4_nodes, capacity_error) : REF_EMPTY;
          if (gm.wide_root != REF_EMPTY && n_tri > hot.triangles) {
            const uint64_t block = out.wide_nodes.size() - gm.wide_root;
            hot = HotBlock{gm.wide_root, static_cast<uint32_t>(block < HOT_TOP_NODES ? block : HOT_TOP_NODES), n_tri};
          }
          out.geometry_meta.push_back(gm);
          ms_wide += since(t_phase);
          t_phase = Clock::now();
          out.tri_bvh_forest.resize(bvh_start + bvh.nodes.size());
          parallel_for(bvh.nodes.size(), PARALLEL_BUILD_MIN, [&](uint64_t i_begin, uint64_t i_end) {
            for (uint64_t i = i_begin; i < i_end; ++i)
              out.tri_bvh_forest[bvh_start + i] = to_device_node(bvh.nodes[i], next[i].first, next[i].second, bvh_start, primitive_start);
          });
          ms_threaded += since(t_phase);
        }
        blas_box.push_back(OptBox{true, root_box});
      }
      uint32_t leaf_count = static_cast<uint32_t>(out.irregular_leaf_boxes.size() - leaf_start);
      if (leaf_count > IRREGULAR_LEAF_MAX) {
        out.irregular_leaf_boxes.resize(leaf_start);
        leaf_count = IRREGULAR_ROUTE_ALL;
      }
      out.blas_meta.push_back(BlasMeta{{tri_start, static_cast<uint32_t>(out.geometry_meta.size())}, leaf_start, leaf_count});
      blas_true_box.push_back(true_box);
      blas_irregular.push_back(leaf_count);
      blas_hot.push_back(hot);
    }
    cache_.blas_epoch = 0;  // (filled below once the BLAS part is known to be complete)
  }

  if (!reuse_blas) {
    cache_.blas_box = blas_box; cache_.blas_true_box = blas_true_box; cache_.blas_irregular = blas_irregular; cache_.blas_hot = blas_hot;
    cache_.wide_nodes_n = out.wide_nodes.size(); cache_.wide4_nodes_n = out.wide4_nodes.size();
    cache_.want_wide4 = want_wide4; cache_.fine_tlas = fine_tlas;
    cache_.stats = out.stats;
    cache_.blas_epoch = blas_epoch_;
  }
  // ---- build_tlas per TLAS (mod.rs:262-320, 428-448)
  const TreeBuildOption tlas_option{50, 10};
  for (const Tlas &tlas : tlas_data_) {
    if (!tlas.alive) {
      out.tlas_root.push_back(TlasRoot{INVALID_NEXT, REF_EMPTY, 0, 0, 0, 0, 0, REF_EMPTY});
      continue;
    }
    HotBlock tlas_hot_geometry;
    const uint32_t bvh_start = static_cast<uint32_t>(out.tlas_bvh_forest.size());
    const uint32_t primitive_start = static_cast<uint32_t>(out.instances.size());
    std::vector<Box3> aabbs(tlas.instances.size());
    for (size_t i = 0; i < tlas.instances.size(); ++i) {
      const InstanceSource &src = tlas.instances[i];
      if (src.blas_handle >= blas_box.size() || !blas_box[src.blas_handle].some) {
        err = "instance references a deleted or unknown BLAS (the reference panics here)";
        return RDN_ERR_INVALID_HANDLE;
      }
    }
    constexpr uint64_t PARALLEL_TLAS_MIN = 2048;  // (a TLAS-only commit of ten thousand moving instances should take a millisecond or two)
    auto t_stage = Clock::now();
    double ms_t_boxes = 0, ms_t_tree = 0, ms_t_records = 0, ms_t_lists = 0, ms_t_wide = 0, ms_t_forest = 0;
    auto lap = [&](double &into) { into += since(t_stage); t_stage = Clock::now(); };
    parallel_for(tlas.instances.size(), PARALLEL_TLAS_MIN, [&](uint64_t i0, uint64_t i1) {
      for (uint64_t i = i0; i < i1; ++i) aabbs[i] = box_apply_matrix(blas_box[tlas.instances[i].blas_handle].box, tlas.instances[i].transform);
    });
    lap(ms_t_boxes);
    SAH sah(4);
    FlattenBVH bvh = timed_build(aabbs.data(), aabbs.size(), sah, tlas_option);
    lap(ms_t_tree);
    if (bvh.stats.bucket_out_of_range) { err = "SAH bucket index out of range (the reference panics here)"; return RDN_ERR_BUILD; }
    out.stats.balance_fallbacks += bvh.stats.balance_fallbacks;
    out.stats.balance_fallbacks_gt10 += bvh.stats.balance_fallbacks_gt10;
    const auto next = compute_bvh_next(bvh.nodes);

    // instance records and boxes in BVH order: independent per slot, written in place; what depends on the order (the irregular
    // list, the hot geometry) is gathered afterwards
    const uint32_t irregular_start = static_cast<uint32_t>(out.irregular_instances.size());
    const size_t n_slots = bvh.sorted_primitive_index.size();
    out.instances.resize(primitive_start + n_slots);
    out.tlas_bounding.resize(primitive_start + n_slots);
    std::vector<uint8_t> slot_irregular(n_slots, 0);  // 1: only the listed leaves of its BLAS, 2: the whole instance
    parallel_for(n_slots, PARALLEL_TLAS_MIN, [&](uint64_t k0, uint64_t k1) {
      for (uint64_t k = k0; k < k1; ++k) {
        const uint64_t box_idx = bvh.sorted_primitive_index[k];
        const InstanceSource &src = tlas.instances[box_idx];
        uint32_t flags = src.flags;
        if (mat4_upper3_det(src.transform) < 0.0f) flags ^= RDN_GEOMETRY_INSTANCE_TRIANGLE_FLIP_FACING;
        InstanceRecord rec;
        const Mat4 inv = mat4_inverse_or_identity(src.transform);
        std::memcpy(rec.transform_inv, &inv, sizeof(inv));
        rec.instance_custom_index = src.instance_custom_index;
        rec.sbt_offset = src.sbt_offset;
        rec.flags = flags;
        rec.blas = src.blas_handle;
        if (src.blas_handle < blas_true_box.size()) {  // else the kernels skip the instance (blas >= n_blas_meta)
          const bool whole = blas_irregular[src.blas_handle] == IRREGULAR_ROUTE_ALL ||
                             instance_is_irregular(src.transform, inv, blas_box[src.blas_handle].box, blas_true_box[src.blas_handle]);
          slot_irregular[k] = whole ? 2 : (blas_irregular[src.blas_handle] != 0 ? 1 : 0);
        }
        out.instances[primitive_start + k] = rec;
        TlasBounding tb;
        tb.world_min[0] = aabbs[box_idx].min.x; tb.world_min[1] = aabbs[box_idx].min.y; tb.world_min[2] = aabbs[box_idx].min.z;
        tb.world_max[0] = aabbs[box_idx].max.x; tb.world_max[1] = aabbs[box_idx].max.y; tb.world_max[2] = aabbs[box_idx].max.z;
        tb.mask = src.mask;
        tb.flags = flags;
        out.tlas_bounding[primitive_start + k] = tb;
      }
    });
    lap(ms_t_records);
    double irregular_area = 0.0;
    for (size_t k = 0; k < n_slots; ++k) {
      const uint64_t box_idx = bvh.sorted_primitive_index[k];
      const uint32_t blas_handle = tlas.instances[box_idx].blas_handle;
      if (blas_handle < blas_true_box.size() && blas_hot[blas_handle].triangles > tlas_hot_geometry.triangles) tlas_hot_geometry = blas_hot[blas_handle];
      if (slot_irregular[k]) {
        const bool whole = slot_irregular[k] == 2;
        out.irregular_instances.push_back(static_cast<uint32_t>(primitive_start + k) | (whole ? IRREGULAR_WHOLE_BIT : 0u));
        out.stats.irregular_instances++;
        if (whole) irregular_area += static_cast<double>(surface_area(aabbs[box_idx]));
      }
    }
    lap(ms_t_lists);
    TlasRoot root;
    root.bvh_root_idx = bvh_start;
    root.wide_root = emit_wide_nodes(bvh.nodes, primitive_start, out.wide_nodes, capacity_error,
                                     fine_tlas ? out.tlas_bounding.data() + primitive_start : nullptr);
    root.wide4_root = want_wide4 ? emit_wide4_nodes(bvh.nodes, primitive_start, out.wide4_nodes, capacity_error) : REF_EMPTY;
    root.hot_count = 0;
    if (root.wide_root != REF_EMPTY) {
      const uint64_t block = out.wide_nodes.size() - root.wide_root;
      root.hot_count = static_cast<uint32_t>(block < HOT_TOP_NODES ? block : HOT_TOP_NODES);
    }
    root.hot_geometry_base = tlas_hot_geometry.base;
    root.hot_geometry_count = tlas_hot_geometry.count;
    root.irregular_start = irregular_start;
    root.irregular_count = static_cast<uint32_t>(out.irregular_instances.size() - irregular_start);
    // a short list is checked per ray by the ordered kernel; a long one, or boxes that cover most of the TLAS anyway
    // (e.g. its only instance), send the whole TLAS to the reference-order kernel
    if (root.irregular_count != 0) {
      const double root_area = bvh.nodes.empty() ? 0.0 : static_cast<double>(surface_area(bvh.nodes[0].bounding));
      if (root.irregular_count > IRREGULAR_LIST_MAX || !(irregular_area < 0.5 * root_area)) {
        out.irregular_instances.resize(irregular_start);
        root.irregular_count = IRREGULAR_ROUTE_ALL;
        out.stats.reference_routed_tlas++;
      }
    }
    out.tlas_root.push_back(root);
    lap(ms_t_wide);
    for (size_t i = 0; i < bvh.nodes.size(); ++i)
      out.tlas_bvh_forest.push_back(to_device_node(bvh.nodes[i], next[i].first, next[i].second, bvh_start, primitive_start));
    lap(ms_t_forest);
    if (timing)
      fprintf(stderr, "[rdn flatten] TLAS of %zu instances: boxes %.2f ms, tree %.2f ms (+ next links), records %.2f ms, lists %.2f ms, wide nodes %.2f ms, "
                      "threaded nodes %.2f ms\n", tlas.instances.size(), ms_t_boxes, ms_t_tree, ms_t_records, ms_t_lists, ms_t_wide, ms_t_forest);
  }

  if (out.geometry_meta.size() > REF_GEOM_ITER_MAX) capacity_error = true;
  if (capacity_error) {
    err = "scene exceeds the 32-bit child-reference encoding (2^27 slots / 0x7F000000 nodes / 2^24 geometries)";
    return RDN_ERR_CAPACITY;
  }
  if (timing)
    fprintf(stderr, "[rdn flatten] BLAS geometries: boxes %.1f ms, triangle records %.1f ms, irregular-leaf scan %.1f ms, wide nodes %.1f ms, "
                    "threaded nodes %.1f ms\n", ms_boxes, ms_records, ms_leaves, ms_wide, ms_threaded);
  out.stats.bvh_build_ms = bvh_ms;
  out.stats.flatten_ms = std::chrono::duration<double, std::milli>(Clock::now() - t_begin).count() - bvh_ms;
  return RDN_OK;
}

// ------------------------------------------------------------------------------------------------ blob
template <typename V>
static void place(uint64_t &cursor, BlobHeader &h, std::vector<FlatScene::ArrayRef> &arrays, int id, const V &v) {
  using T = typename V::value_type;
  const uint64_t off = (cursor + BLOB_ALIGN - 1) / BLOB_ALIGN * BLOB_ALIGN;
  // every array keeps at least one zeroed element so kernels never see a null base (create_gpu_buffer, mod.rs:386-400)
  const uint64_t bytes = (v.empty() ? 1 : v.size()) * sizeof(T);
  cursor = off + bytes;
  h.offset[id] = off;
  h.count[id] = v.size();
  h.elem_size[id] = sizeof(T);
  arrays.push_back(FlatScene::ArrayRef{id, v.data(), v.size() * sizeof(T)});
}

BlobHeader FlatScene::layout(std::vector<ArrayRef> &arrays) const {
  BlobHeader h;
  std::memset(&h, 0, sizeof(h));
  h.magic = BLOB_MAGIC;
  h.version = BLOB_VERSION;
  h.header_bytes = sizeof(BlobHeader);
  arrays.clear();
  uint64_t cursor = sizeof(BlobHeader);
  place(cursor, h, arrays, ARR_TLAS_BINDING, tlas_binding);
  place(cursor, h, arrays, ARR_TLAS_ROOT, tlas_root);
  place(cursor, h, arrays, ARR_TLAS_BVH_FOREST, tlas_bvh_forest);
  place(cursor, h, arrays, ARR_TLAS_BOUNDING, tlas_bounding);
  place(cursor, h, arrays, ARR_INSTANCES, instances);
  place(cursor, h, arrays, ARR_BLAS_META, blas_meta);
  place(cursor, h, arrays, ARR_GEOMETRY_META, geometry_meta);
  place(cursor, h, arrays, ARR_TRI_BVH_FOREST, tri_bvh_forest);
  place(cursor, h, arrays, ARR_TRIANGLES, triangles);
  place(cursor, h, arrays, ARR_SLOT_INFO, slot_info);
  place(cursor, h, arrays, ARR_WIDE_NODES, wide_nodes);
  place(cursor, h, arrays, ARR_PRIM_TO_SLOT, prim_to_slot);
  place(cursor, h, arrays, ARR_IRREGULAR_INSTANCES, irregular_instances);
  place(cursor, h, arrays, ARR_IRREGULAR_LEAF_BOXES, irregular_leaf_boxes);
  place(cursor, h, arrays, ARR_WIDE4_NODES, wide4_nodes);
  h.total_bytes = (cursor + BLOB_ALIGN - 1) / BLOB_ALIGN * BLOB_ALIGN;
  return h;
}

std::vector<uint8_t> FlatScene::serialize() const {
  std::vector<ArrayRef> arrays;
  const BlobHeader h = layout(arrays);
  std::vector<uint8_t> blob(h.total_bytes, 0);
  std::memcpy(blob.data(), &h, sizeof(h));
  for (const ArrayRef &a : arrays)
    if (a.bytes) std::memcpy(blob.data() + h.offset[a.id], a.data, a.bytes);
  return blob;
}

}  // namespace rdn
