// build_device.cu — the reference's binned-SAH BVH build on the device, node for node the tree of
// FlattenBVH::new + SAH (content/space/src/bvh/mod.rs:55-79, strategy.rs:11-43,202-284) with its BalanceTree fallback (:67-86).
//
// Level-synchronous: all nodes of a depth are split together by six passes over flat arrays — the formulation
// tools/level_sync_build.py states in numpy and tests/test_level_sync_build.py holds to the recursive builder:
//   k_slot_begin   per splitting node: longest axis, bucket origin and width; bucket counters / boxes reset
//   k_bucket       per primitive position: bucket of its centre; counts and box unions per (node, bucket) by atomics
//                  (integer adds and min / max of order-preserving float keys: exact in any order)
//   (CUB scan)     exclusive prefix sums of the four one-hot bucket indicators over all positions
//   k_split        per node: all primitives in one bucket -> median split (BalanceTree), else the first strict minimum of the
//                  prefix costs area(L)*nL + area(R)*nR in f32; the two children, the next level's list
//   k_scatter      per position: new position = node start + primitives of earlier buckets + rank among the node's primitives of
//                  the same bucket (from the scan): the reference's STABLE bucket-by-bucket rewrite
//   k_fallback     per degenerate node: stable insertion sort of its (short) range by centre, child boxes
// The host then numbers the nodes in pre-order (subtree sizes bottom-up, offsets top-down).  Degenerate ranges longer than
// FALLBACK_MAX primitives (many identical centres) are left to the host builder.  f32 arithmetic in the reference's order, -fmad=false.
#include <cuda_runtime.h>

#ifndef RDN_SIMT_EMU  // (the CPU emulation build of the test-suite, tests/simt/, scans with a plain loop: CUB needs nvcc)
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#endif

#include <cmath>
#include <string>
#include <vector>

#include "bvh_builder.h"
#include "kernels.h"
#include "rdn_math.h"

namespace rdn {

namespace {

constexpr int MAXB = 4;               // SAH::new(4) everywhere in the reference's ray-tracing path (naive/mod.rs:146,285)
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint32_t FALLBACK_MAX = 64;  // longest degenerate range sorted on the device (one thread, insertion sort)

struct DevNode {  // BFS creation order; children of a node are `left` and `left + 1`
  float bmin[3]; uint32_t start;
  float bmax[3]; uint32_t end;
  uint32_t left; int32_t axis; uint32_t depth; uint32_t pad;
};
struct Slot {  // one splitting node of the current level
  uint32_t node; int32_t axis; float lo, step;
  uint32_t cnt[MAXB];
  uint32_t kmin[MAXB][3], kmax[MAXB][3];  // order-preserving keys of the bucket boxes
  uint32_t offset[MAXB];                  // primitives of the node in earlier buckets
  uint32_t left_count, left_next, right_next, degenerate;
};
struct Counts4 { uint32_t c[MAXB]; };
struct Counts4Add {
  __host__ __device__ Counts4 operator()(const Counts4 &a, const Counts4 &b) const {
    Counts4 r;
    for (int k = 0; k < MAXB; ++k) r.c[k] = a.c[k] + b.c[k];
    return r;
  }
};
struct OneHot {
  __host__ __device__ Counts4 operator()(const uint8_t &w) const {
    Counts4 r;
    for (int k = 0; k < MAXB; ++k) r.c[k] = (w == k) ? 1u : 0u;
    return r;
  }
};
struct Control { uint32_t node_count, next_count, unsupported, out_of_range; };

// exclusive prefix sums of the one-hot bucket indicators (d_temp == nullptr: only the scratch size is returned)
cudaError_t scan_one_hot(void *d_temp, size_t &temp_bytes, const uint8_t *d_which, Counts4 *d_scan, uint32_t n) {
#ifndef RDN_SIMT_EMU
  cub::TransformInputIterator<Counts4, OneHot, const uint8_t *> it(d_which, OneHot());
  return cub::DeviceScan::ExclusiveScan(d_temp, temp_bytes, it, d_scan, Counts4Add(), Counts4{}, static_cast<int>(n));
#else
  if (!d_temp) { temp_bytes = 16; return cudaSuccess; }
  Counts4 run{};
  for (uint32_t i = 0; i < n; ++i) { d_scan[i] = run; run = Counts4Add()(run, OneHot()(d_which[i])); }
  return cudaSuccess;
#endif
}

__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

__device__ __forceinline__ float area_of(const float *mn, const float *mx) {  // box3.rs:5-11
  const float w = mx[0] - mn[0], h = mx[1] - mn[1], d = mx[2] - mn[2];
  return 2.0f * (w * h + w * d + h * d);
}
__device__ __forceinline__ int longest_axis_of(const float *mn, const float *mx) {  // box3.rs:117-133
  const float x = mx[0] - mn[0], y = mx[1] - mn[1], z = mx[2] - mn[2];
  if (x > y) return x > z ? 0 : 2;
  if (y > z) return 1;
  return 2;
}
__device__ __forceinline__ float centre_of(const Box3 *boxes, uint32_t prim, int axis) {  // box3.rs:91-93: (min + max) * 0.5
  const float *b = reinterpret_cast<const float *>(boxes + prim);
  return (b[axis] + b[3 + axis]) * 0.5f;
}

__global__ void k_slot_begin(const uint32_t *__restrict__ active, uint32_t n_active, const DevNode *__restrict__ nodes, Slot *__restrict__ slots,
                             uint32_t n_buckets) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_active) return;
  const DevNode nd = nodes[active[s]];
  Slot &sl = slots[s];
  sl.node = active[s];
  sl.axis = longest_axis_of(nd.bmin, nd.bmax);
  sl.lo = nd.bmin[sl.axis];
  sl.step = (nd.bmax[sl.axis] - sl.lo) / static_cast<float>(n_buckets);
  for (int b = 0; b < MAXB; ++b) {
    sl.cnt[b] = 0;
    for (int c = 0; c < 3; ++c) { sl.kmin[b][c] = float_key(INFINITY); sl.kmax[b][c] = float_key(-INFINITY); }
  }
}

__global__ void k_bucket(const Box3 *__restrict__ boxes, const uint32_t *__restrict__ index_in, uint32_t *__restrict__ index_out,
                         const uint32_t *__restrict__ owner, uint32_t *__restrict__ owner_next, uint32_t n, Slot *__restrict__ slots,
                         uint32_t n_buckets, uint8_t *__restrict__ which, Control *__restrict__ ctl) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const uint32_t s = owner[pos];
  const uint32_t prim = index_in[pos];
  if (s == NONE) {  // this range is final: carried over
    which[pos] = 0xFF;
    index_out[pos] = prim;
    owner_next[pos] = NONE;
    return;
  }
  Slot &sl = slots[s];
  const float v = floorf((centre_of(boxes, prim, sl.axis) - sl.lo) / sl.step);
  unsigned long long w = (!(v == v) || v <= 0.0f) ? 0ull : (v >= 18446744073709551616.0f ? 0xFFFFFFFFFFFFFFFFull : static_cast<unsigned long long>(v));
  if (w == n_buckets) w -= 1;
  if (w >= n_buckets) { ctl->out_of_range = 1; w = n_buckets - 1; }
  which[pos] = static_cast<uint8_t>(w);
  atomicAdd(&sl.cnt[w], 1u);
  const float *b = reinterpret_cast<const float *>(boxes + prim);
  for (int c = 0; c < 3; ++c) {
    atomicMin(&sl.kmin[w][c], float_key(b[c]));
    atomicMax(&sl.kmax[w][c], float_key(b[3 + c]));
  }
}

__global__ void k_split(Slot *__restrict__ slots, uint32_t n_active, DevNode *__restrict__ nodes, uint32_t n_buckets, uint32_t max_depth,
                        uint32_t bin_size, uint32_t *__restrict__ next_active, Control *__restrict__ ctl) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_active) return;
  Slot &sl = slots[s];
  DevNode &parent = nodes[sl.node];
  const uint32_t start = parent.start, end = parent.end, count = end - start;
  float bmin[MAXB][3], bmax[MAXB][3];
  uint32_t empty = 0;
  for (uint32_t b = 0; b < n_buckets; ++b) {
    empty += sl.cnt[b] == 0;
    for (int c = 0; c < 3; ++c) { bmin[b][c] = key_float(sl.kmin[b][c]); bmax[b][c] = key_float(sl.kmax[b][c]); }
  }
  float lmin[3], lmax[3], rmin[3], rmax[3];
  uint32_t left_count;
  sl.degenerate = empty == n_buckets - 1;
  if (sl.degenerate) {
    if (count > FALLBACK_MAX) ctl->unsupported = 1;
    left_count = (end + start) / 2 - start;  // median_partition_at_axis: middle = (end + begin) / 2
    for (int c = 0; c < 3; ++c) { lmin[c] = rmin[c] = INFINITY; lmax[c] = rmax[c] = -INFINITY; }  // filled by k_fallback
  } else {
    // group_of(from, to): union of the bucket boxes and their count; cost of the n_buckets - 1 prefix partitions, first strict minimum
    uint32_t best = 0;
    float best_cost = INFINITY;
    bool have = false;
    for (uint32_t i = 0; i + 1 < n_buckets; ++i) {
      float amin[3] = {INFINITY, INFINITY, INFINITY}, amax[3] = {-INFINITY, -INFINITY, -INFINITY};
      float cmin[3] = {INFINITY, INFINITY, INFINITY}, cmax[3] = {-INFINITY, -INFINITY, -INFINITY};
      uint32_t ln = 0, rn = 0;
      for (uint32_t k = 0; k <= i; ++k) { ln += sl.cnt[k]; for (int c = 0; c < 3; ++c) { amin[c] = fminf(amin[c], bmin[k][c]); amax[c] = fmaxf(amax[c], bmax[k][c]); } }
      for (uint32_t k = i + 1; k < n_buckets; ++k) { rn += sl.cnt[k]; for (int c = 0; c < 3; ++c) { cmin[c] = fminf(cmin[c], bmin[k][c]); cmax[c] = fmaxf(cmax[c], bmax[k][c]); } }
      const float cost = area_of(amin, amax) * static_cast<float>(ln) + area_of(cmin, cmax) * static_cast<float>(rn);
      if (cost < best_cost) { best_cost = cost; best = i; have = true; }
    }
    if (!have) best = 0;  // no cost below +inf: the reference keeps its initial (bucket 0 | the rest)
    left_count = 0;
    for (int c = 0; c < 3; ++c) { lmin[c] = rmin[c] = INFINITY; lmax[c] = rmax[c] = -INFINITY; }
    for (uint32_t k = 0; k < n_buckets; ++k) {
      const bool left = k <= best;
      if (left) left_count += sl.cnt[k];
      for (int c = 0; c < 3; ++c) {
        if (left) { lmin[c] = fminf(lmin[c], bmin[k][c]); lmax[c] = fmaxf(lmax[c], bmax[k][c]); }
        else { rmin[c] = fminf(rmin[c], bmin[k][c]); rmax[c] = fmaxf(rmax[c], bmax[k][c]); }
      }
    }
  }
  uint32_t off = 0;
  for (uint32_t b = 0; b < n_buckets; ++b) { sl.offset[b] = off; off += sl.cnt[b]; }
  sl.left_count = left_count;
  // children: ids left, left + 1
  const uint32_t left_id = atomicAdd(&ctl->node_count, 2u);
  parent.left = left_id;
  parent.axis = sl.axis;
  const uint32_t depth = parent.depth + 1;
  for (int side = 0; side < 2; ++side) {
    DevNode child;
    for (int c = 0; c < 3; ++c) { child.bmin[c] = side ? rmin[c] : lmin[c]; child.bmax[c] = side ? rmax[c] : lmax[c]; }
    child.start = side ? start + left_count : start;
    child.end = side ? end : start + left_count;
    child.left = NONE; child.axis = 0; child.depth = depth; child.pad = 0;
    nodes[left_id + side] = child;
    uint32_t next_slot = NONE;
    if (depth < max_depth && child.end - child.start > bin_size) {  // TreeBuildOption::should_continue
      next_slot = atomicAdd(&ctl->next_count, 1u);
      next_active[next_slot] = left_id + side;
    }
    if (side) sl.right_next = next_slot; else sl.left_next = next_slot;
  }
}

__global__ void k_scatter(const uint32_t *__restrict__ index_in, uint32_t *__restrict__ index_out, const uint32_t *__restrict__ owner,
                          uint32_t *__restrict__ owner_next, uint32_t n, const Slot *__restrict__ slots, const DevNode *__restrict__ nodes,
                          const uint8_t *__restrict__ which, const Counts4 *__restrict__ scan) {
  const uint32_t pos = blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= n) return;
  const uint32_t s = owner[pos];
  if (s == NONE) return;  // carried over by k_bucket
  const Slot &sl = slots[s];
  const uint32_t start = nodes[sl.node].start;
  uint32_t np = pos;
  if (!sl.degenerate) {
    const uint32_t b = which[pos];
    np = start + sl.offset[b] + (scan[pos].c[b] - scan[start].c[b]);
  }
  index_out[np] = index_in[pos];
  owner_next[np] = np < start + sl.left_count ? sl.left_next : sl.right_next;
}

// BalanceTree::split for the (short) degenerate ranges: stable sort by centre along the axis, child boxes from the halves
__global__ void k_fallback(const Box3 *__restrict__ boxes, uint32_t *__restrict__ index_out, const Slot *__restrict__ slots, uint32_t n_active,
                           DevNode *__restrict__ nodes) {
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_active || !slots[s].degenerate) return;
  const Slot &sl = slots[s];
  const DevNode parent = nodes[sl.node];
  const uint32_t start = parent.start, end = parent.end;
  if (end - start > FALLBACK_MAX) return;  // (flagged unsupported: the host builder takes over)
  if ((end - start) / 2 != 0) {
    for (uint32_t i = start + 1; i < end; ++i) {  // stable insertion sort: an element only passes strictly greater ones
      const uint32_t p = index_out[i];
      const float key = centre_of(boxes, p, sl.axis);
      uint32_t j = i;
      while (j > start && centre_of(boxes, index_out[j - 1], sl.axis) > key) { index_out[j] = index_out[j - 1]; --j; }
      index_out[j] = p;
    }
  }
  for (int side = 0; side < 2; ++side) {
    DevNode &child = nodes[parent.left + side];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint32_t i = child.start; i < child.end; ++i) {
      const float *b = reinterpret_cast<const float *>(boxes + index_out[i]);
      for (int c = 0; c < 3; ++c) { mn[c] = fminf(mn[c], b[c]); mx[c] = fmaxf(mx[c], b[3 + c]); }
    }
    for (int c = 0; c < 3; ++c) { child.bmin[c] = mn[c]; child.bmax[c] = mx[c]; }
  }
}

__global__ void k_fill_u32(uint32_t *dst, uint32_t n, uint32_t v, bool iota) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = iota ? i : v;
}

#define BD_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) { err = std::string(#expr) + ": " + cudaGetErrorString(e__); rc = -1; goto done; } \
  } while (0)

}  // namespace

// 0 = built; 1 = not supported on the device (caller uses the host builder); -1 = CUDA error (err); -2 = bucket out of range
int build_bvh_sah_device(const Box3 *boxes, uint64_t n64, uint32_t n_buckets, const TreeBuildOption &option, int device, FlattenBVH &out,
                         std::string &err) {
  if (n_buckets < 2 || n_buckets > MAXB || n64 == 0 || n64 >= 0x7FFFFFFFull) return 1;
  const uint32_t n = static_cast<uint32_t>(n64);
  const uint32_t max_depth = static_cast<uint32_t>(std::min<uint64_t>(option.max_tree_depth, 0xFFFFFFFFull));
  const uint32_t bin_size = static_cast<uint32_t>(std::min<uint64_t>(option.bin_size, 0xFFFFFFFFull));
  int rc = 0;
  Box3 *d_boxes = nullptr;
  uint32_t *d_index[2] = {nullptr, nullptr}, *d_owner[2] = {nullptr, nullptr}, *d_active[2] = {nullptr, nullptr};
  uint8_t *d_which = nullptr;
  Counts4 *d_scan = nullptr;
  DevNode *d_nodes = nullptr;
  Slot *d_slots = nullptr;
  Control *d_ctl = nullptr;
  void *d_temp = nullptr;
  size_t temp_bytes = 0;
  const uint64_t node_cap = 2ull * n + 2;
  const uint32_t threads = 256, blocks_n = (n + threads - 1) / threads;
  std::vector<DevNode> h_nodes;
  std::vector<uint32_t> h_index;
  Control ctl{};
  DevNode root{};
  uint32_t n_active = 0, cur = 0;

  // root box on the host (a union: exact in any order)
  Box3 rb = box_empty();
  for (uint32_t i = 0; i < n; ++i) expand(rb, boxes[i]);
  root.bmin[0] = rb.min.x; root.bmin[1] = rb.min.y; root.bmin[2] = rb.min.z;
  root.bmax[0] = rb.max.x; root.bmax[1] = rb.max.y; root.bmax[2] = rb.max.z;
  root.start = 0; root.end = n; root.left = NONE; root.axis = 0; root.depth = 0; root.pad = 0;
  const bool root_splits = 0 < max_depth && n > bin_size;

  BD_CUDA(cudaSetDevice(device));
  BD_CUDA(cudaMalloc(&d_boxes, static_cast<size_t>(n) * sizeof(Box3)));
  for (int k = 0; k < 2; ++k) {
    BD_CUDA(cudaMalloc(&d_index[k], static_cast<size_t>(n) * 4));
    BD_CUDA(cudaMalloc(&d_owner[k], static_cast<size_t>(n) * 4));
    BD_CUDA(cudaMalloc(&d_active[k], static_cast<size_t>(n) * 4));
  }
  BD_CUDA(cudaMalloc(&d_which, n));
  BD_CUDA(cudaMalloc(&d_scan, static_cast<size_t>(n) * sizeof(Counts4)));
  BD_CUDA(cudaMalloc(&d_nodes, node_cap * sizeof(DevNode)));
  BD_CUDA(cudaMalloc(&d_slots, static_cast<size_t>(n) * sizeof(Slot)));
  BD_CUDA(cudaMalloc(&d_ctl, sizeof(Control)));
  BD_CUDA(scan_one_hot(nullptr, temp_bytes, d_which, d_scan, n));
  BD_CUDA(cudaMalloc(&d_temp, temp_bytes ? temp_bytes : 16));
  BD_CUDA(cudaMemcpy(d_boxes, boxes, static_cast<size_t>(n) * sizeof(Box3), cudaMemcpyHostToDevice));
  BD_CUDA(cudaMemcpy(d_nodes, &root, sizeof(root), cudaMemcpyHostToDevice));
  k_fill_u32<<<blocks_n, threads>>>(d_index[0], n, 0, true);
  k_fill_u32<<<blocks_n, threads>>>(d_owner[0], n, root_splits ? 0u : NONE, false);
  ctl.node_count = 1;
  if (root_splits) {
    const uint32_t zero = 0;
    BD_CUDA(cudaMemcpy(d_active[0], &zero, 4, cudaMemcpyHostToDevice));
    n_active = 1;
  }
  while (n_active) {
    ctl.next_count = 0;
    BD_CUDA(cudaMemcpy(d_ctl, &ctl, sizeof(ctl), cudaMemcpyHostToDevice));
    const uint32_t blocks_a = (n_active + threads - 1) / threads;
    k_slot_begin<<<blocks_a, threads>>>(d_active[cur], n_active, d_nodes, d_slots, n_buckets);
    k_bucket<<<blocks_n, threads>>>(d_boxes, d_index[cur], d_index[cur ^ 1], d_owner[cur], d_owner[cur ^ 1], n, d_slots, n_buckets, d_which, d_ctl);
    BD_CUDA(scan_one_hot(d_temp, temp_bytes, d_which, d_scan, n));
    k_split<<<blocks_a, threads>>>(d_slots, n_active, d_nodes, n_buckets, max_depth, bin_size, d_active[cur ^ 1], d_ctl);
    k_scatter<<<blocks_n, threads>>>(d_index[cur], d_index[cur ^ 1], d_owner[cur], d_owner[cur ^ 1], n, d_slots, d_nodes, d_which, d_scan);
    k_fallback<<<blocks_a, threads>>>(d_boxes, d_index[cur ^ 1], d_slots, n_active, d_nodes);
    BD_CUDA(cudaGetLastError());
    BD_CUDA(cudaMemcpy(&ctl, d_ctl, sizeof(ctl), cudaMemcpyDeviceToHost));
    if (ctl.out_of_range) { rc = -2; goto done; }
    if (ctl.unsupported) { rc = 1; goto done; }
    if (ctl.node_count > node_cap) { err = "internal: node capacity exceeded"; rc = -1; goto done; }
    n_active = ctl.next_count;
    cur ^= 1;
  }
  h_nodes.resize(ctl.node_count);
  h_index.resize(n);
  BD_CUDA(cudaMemcpy(h_nodes.data(), d_nodes, h_nodes.size() * sizeof(DevNode), cudaMemcpyDeviceToHost));
  BD_CUDA(cudaMemcpy(h_index.data(), d_index[cur], static_cast<size_t>(n) * 4, cudaMemcpyDeviceToHost));
  {
    // pre-order numbering: subtree sizes bottom-up (children are created after their parents), offsets top-down
    const size_t m = h_nodes.size();
    std::vector<uint64_t> size(m, 1), pre(m, 0);
    for (size_t i = m; i-- > 0;)
      if (h_nodes[i].left != NONE) size[i] += size[h_nodes[i].left] + size[h_nodes[i].left + 1];
    for (size_t i = 0; i < m; ++i)
      if (h_nodes[i].left != NONE) {
        pre[h_nodes[i].left] = pre[i] + 1;
        pre[h_nodes[i].left + 1] = pre[i] + 1 + size[h_nodes[i].left];
      }
    out = FlattenBVH{};
    out.nodes.resize(m);
    for (size_t i = 0; i < m; ++i) {
      const DevNode &d = h_nodes[i];
      FlattenBVHNode nd;
      std::memset(&nd, 0, sizeof(nd));
      nd.bounding = Box3{Vec3{d.bmin[0], d.bmin[1], d.bmin[2]}, Vec3{d.bmax[0], d.bmax[1], d.bmax[2]}};
      nd.primitive_start = d.start; nd.primitive_end = d.end; nd.self_index = pre[i];
      if (d.left != NONE) { nd.has_child = 1; nd.split_axis = d.axis; nd.left_count = size[d.left]; }
      out.nodes[pre[i]] = nd;
    }
    out.sorted_primitive_index.assign(h_index.begin(), h_index.end());
    out.stats.build_threads = 0;  // built on the device
  }
done:
  for (void *p : {static_cast<void *>(d_boxes), static_cast<void *>(d_index[0]), static_cast<void *>(d_index[1]), static_cast<void *>(d_owner[0]),
                  static_cast<void *>(d_owner[1]), static_cast<void *>(d_active[0]), static_cast<void *>(d_active[1]), static_cast<void *>(d_which),
                  static_cast<void *>(d_scan), static_cast<void *>(d_nodes), static_cast<void *>(d_slots), static_cast<void *>(d_ctl), d_temp})
    if (p) cudaFree(p);
  return rc;
}

}  // namespace rdn
