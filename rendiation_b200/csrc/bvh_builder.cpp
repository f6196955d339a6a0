// bvh_builder.cpp — see bvh_builder.h.  Compiled with -ffp-contract=off.
#include "bvh_builder.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace rdn {

int longest_axis(const Box3 &b) {
  // the exact `>` cascade of math/geometry/src/dimension3/box3.rs:117-133
  const float x_length = b.max.x - b.min.x;
  const float y_length = b.max.y - b.min.y;
  const float z_length = b.max.z - b.min.z;
  if (x_length > y_length) return x_length > z_length ? 0 : 2;
  if (y_length > z_length) return 1;
  return 2;
}

float surface_area(const Box3 &b) {
  // LebesgueMeasurable<2> for Box3, box3.rs:5-11
  const float w = b.max.x - b.min.x, h = b.max.y - b.min.y, d = b.max.z - b.min.z;
  return 2.0f * (w * h + w * d + h * d);
}

Vec3 box_center(const Box3 &b) { return (b.min + b.max) * 0.5f; }

static inline float component(const Vec3 &v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

static Box3 bounding_of_range(const std::vector<BuildPrimitive> &src, const std::vector<uint64_t> &index, uint64_t begin,
                              uint64_t end) {
  Box3 r = box_empty();
  for (uint64_t i = begin; i < end; ++i) expand(r, src[index[i]].bounding);
  return r;
}

SplitResult BalanceTree::split(const FlattenBVHNode &parent, const std::vector<BuildPrimitive> &src,
                               std::vector<uint64_t> &index, BuildStats &) {
  SplitResult r;
  r.axis = longest_axis(parent.bounding);
  const uint64_t begin = parent.primitive_start, end = parent.primitive_end;
  const uint64_t middle = (end + begin) / 2;
  if ((end - begin) / 2 != 0) {
    // median_partition_at_axis (apply.rs:19-49).  Rust's select_nth_unstable_by leaves an unspecified
    // permutation; a stable sort by centre is a valid one and equals std's insertion-sort path (<= 10 items).
    const int axis = r.axis;
    std::stable_sort(index.begin() + begin, index.begin() + end, [&](uint64_t a, uint64_t b) {
      return component(src[a].center, axis) < component(src[b].center, axis);
    });
  }
  r.left_start = begin; r.left_end = middle; r.right_start = middle; r.right_end = end;
  r.left_box = bounding_of_range(src, index, begin, middle);
  r.right_box = bounding_of_range(src, index, middle, end);
  return r;
}

SAH::SAH(uint32_t n) : pre_partition_(std::max<uint32_t>(n, 2u)) {}

// Rust `as usize` on f32: saturating, NaN -> 0
static inline uint64_t saturating_usize(float v) {
  if (!(v == v) || v <= 0.0f) return 0;
  if (v >= 18446744073709551616.0f) return UINT64_MAX;
  return static_cast<uint64_t>(v);
}

SplitResult SAH::split(const FlattenBVHNode &parent, const std::vector<BuildPrimitive> &src, std::vector<uint64_t> &index,
                       BuildStats &stats) {
  const uint64_t begin = parent.primitive_start, end = parent.primitive_end;
  const size_t n_part = pre_partition_.size();
  for (auto &p : pre_partition_) { p.primitive_bucket.clear(); p.bounding = box_empty(); }
  std::vector<uint64_t> &counts = counts_;  // member scratch: a split per inner node must not allocate
  counts.assign(n_part, 0);

  // step 1: bucket every primitive by its centre along the longest axis of the NODE box
  const int axis = longest_axis(parent.bounding);
  const float range_start = component(parent.bounding.min, axis);
  const float range_end = component(parent.bounding.max, axis);
  const float step = (range_end - range_start) / static_cast<float>(n_part);
  auto bucket_of = [&](uint64_t prim, bool &out_of_range) -> size_t {
    const float axis_value = component(src[prim].center, axis);
    uint64_t which = saturating_usize(floorf((axis_value - range_start) / step));
    if (which == n_part) which -= 1;
    if (which >= n_part) { out_of_range = true; which = n_part - 1; }
    return static_cast<size_t>(which);
  };

  // Large ranges (the top levels of a big tree, split on the calling thread before the subtrees fan out) are bucketed and
  // rewritten by all threads: chunk t counts its primitives per bucket, the rewrite offsets are the prefix sums over
  // (bucket, chunk), so the result is the sequential bucket-by-bucket stable order; box unions are exact in any order.
  // (thread count only looked up for ranges that qualify: hardware_concurrency() is a system call and there is a split per inner node)
  const unsigned threads = (parallel_split_ && end - begin >= PARALLEL_SPLIT_MIN && n_part <= 255) ? build_thread_count() : 1u;
  const bool parallel = threads > 1;
  std::vector<uint8_t> which_of;
  std::vector<std::vector<uint64_t>> chunk_counts;
  uint64_t chunk = 0;
  if (parallel) {
    const uint64_t n = end - begin;
    chunk = (n + threads - 1) / threads;
    which_of.resize(n);
    chunk_counts.assign(threads, std::vector<uint64_t>(n_part, 0));
    std::vector<std::vector<Box3>> chunk_boxes(threads, std::vector<Box3>(n_part, box_empty()));
    std::atomic<bool> out_of_range{false};
    parallel_for(n, 0, [&](uint64_t b0, uint64_t b1) {
      const size_t t = static_cast<size_t>(b0 / chunk);
      bool oor = false;
      std::vector<uint64_t> my_counts(n_part, 0);  // thread-private while counting: the shared rows sit in one cache line
      std::vector<Box3> my_boxes(n_part, box_empty());
      for (uint64_t i = b0; i < b1; ++i) {
        const uint64_t prim = index[begin + i];
        const size_t w = bucket_of(prim, oor);
        which_of[i] = static_cast<uint8_t>(w);
        my_counts[w]++;
        expand(my_boxes[w], src[prim].bounding);
      }
      chunk_counts[t] = my_counts;
      chunk_boxes[t] = my_boxes;
      if (oor) out_of_range = true;
    });
    if (out_of_range) stats.bucket_out_of_range = true;
    for (size_t t = 0; t < threads; ++t)
      for (size_t k = 0; k < n_part; ++k) { counts[k] += chunk_counts[t][k]; expand(pre_partition_[k].bounding, chunk_boxes[t][k]); }
  } else {
    for (uint64_t i = begin; i < end; ++i) {
      const uint64_t prim = index[i];
      const size_t which = bucket_of(prim, stats.bucket_out_of_range);
      expand(pre_partition_[which].bounding, src[prim].bounding);
      pre_partition_[which].primitive_bucket.push_back(prim);
      counts[which]++;
    }
  }

  size_t empty_buckets = 0;
  for (size_t k = 0; k < n_part; ++k) empty_buckets += counts[k] == 0;
  if (empty_buckets == n_part - 1) {
    stats.balance_fallbacks++;
    if (end - begin > 10) stats.balance_fallbacks_gt10++;
    BalanceTree fallback;
    return fallback.split(parent, src, index, stats);
  }

  // step 2: cost of each of the n_part-1 prefix partitions; first strict minimum wins
  struct Group { Box3 box; uint64_t count; };
  auto group_of = [&](size_t from, size_t to) {
    Group g{box_empty(), 0};
    for (size_t k = from; k < to; ++k) { expand(g.box, pre_partition_[k].bounding); g.count += counts[k]; }
    return g;
  };
  Group best_left = group_of(0, 1), best_right = group_of(1, n_part);
  float best_cost = INFINITY;
  for (size_t i = 0; i + 1 < n_part; ++i) {
    const Group l = group_of(0, i + 1), r = group_of(i + 1, n_part);
    const float cost = surface_area(l.box) * static_cast<float>(l.count) + surface_area(r.box) * static_cast<float>(r.count);
    if (cost < best_cost) { best_cost = cost; best_left = l; best_right = r; }
  }

  // step 3: rewrite the index range bucket by bucket (stable)
  if (parallel) {
    const uint64_t n = end - begin;
    std::vector<uint64_t> old(index.begin() + begin, index.begin() + end);
    std::vector<std::vector<uint64_t>> offset(threads, std::vector<uint64_t>(n_part, 0));
    uint64_t ptr = begin;
    for (size_t k = 0; k < n_part; ++k)
      for (size_t t = 0; t < threads; ++t) { offset[t][k] = ptr; ptr += chunk_counts[t][k]; }
    parallel_for(n, 0, [&](uint64_t b0, uint64_t b1) {
      std::vector<uint64_t> off = offset[static_cast<size_t>(b0 / chunk)];  // private copy (see above)
      for (uint64_t i = b0; i < b1; ++i) index[off[which_of[i]]++] = old[i];
    });
  } else {
    uint64_t ptr = begin;
    for (const auto &p : pre_partition_)
      for (uint64_t prim : p.primitive_bucket) index[ptr++] = prim;
  }

  SplitResult r;
  r.axis = axis;
  r.left_box = best_left.box; r.left_start = begin; r.left_end = begin + best_left.count;
  r.right_box = best_right.box; r.right_start = begin + best_left.count; r.right_end = end;
  return r;
}

unsigned build_thread_count() {
  unsigned n = std::thread::hardware_concurrency();
  if (n == 0) n = 1;
  if (const char *e = getenv("RDN_BUILD_THREADS")) {
    const int cap = atoi(e);
    if (cap >= 1 && static_cast<unsigned>(cap) < n) n = static_cast<unsigned>(cap);
  }
  return n;
}

void parallel_for(uint64_t n, uint64_t min_parallel, const std::function<void(uint64_t, uint64_t)> &fn) {
  const unsigned threads = build_thread_count();
  if (threads <= 1 || n < min_parallel) { fn(0, n); return; }
  const uint64_t chunk = (n + threads - 1) / threads;
  std::vector<std::thread> pool;
  for (unsigned t = 1; t < threads; ++t) {
    const uint64_t begin = std::min<uint64_t>(n, t * chunk), end = std::min<uint64_t>(n, begin + chunk);
    if (begin < end) pool.emplace_back([&fn, begin, end]() { fn(begin, end); });
  }
  fn(0, std::min<uint64_t>(n, chunk));
  for (auto &th : pool) th.join();
}

namespace {

// Pre-order construction of the subtree over [start, end) without recursion: descend left, park the right sibling; when a
// leaf is reached the most recent parked sibling is emitted next, which fixes its parent's left_count.  Node indices are local
// to `nodes` (the subtree root is node 0); left_count is position independent.
void build_subtree(const Box3 &box, uint64_t start, uint64_t end, uint64_t depth0, BVHBuildStrategy &strategy, const TreeBuildOption &option,
                   const std::vector<BuildPrimitive> &primitives, std::vector<uint64_t> &index, std::vector<FlattenBVHNode> &nodes,
                   BuildStats &stats) {
  auto make_node = [&](const Box3 &b, uint64_t s, uint64_t e) {
    FlattenBVHNode nd;
    std::memset(&nd, 0, sizeof(nd));
    nd.bounding = b; nd.primitive_start = s; nd.primitive_end = e; nd.self_index = nodes.size();
    nodes.push_back(nd);
  };
  make_node(box, start, end);
  struct Parked { uint64_t parent; Box3 box; uint64_t start, end; uint64_t depth; int32_t axis; };
  std::vector<Parked> parked;
  uint64_t cur = 0, depth = depth0;
  for (;;) {
    const FlattenBVHNode node = nodes[cur];
    if (option.should_continue(node.primitive_end - node.primitive_start, depth)) {
      const SplitResult s = strategy.split(node, primitives, index, stats);
      parked.push_back(Parked{cur, s.right_box, s.right_start, s.right_end, depth + 1, s.axis});
      make_node(s.left_box, s.left_start, s.left_end);
      cur = nodes.size() - 1;
      depth += 1;
      continue;
    }
    if (parked.empty()) break;
    const Parked p = parked.back();
    parked.pop_back();
    FlattenBVHNode &parent = nodes[p.parent];
    parent.has_child = 1;
    parent.split_axis = p.axis;
    parent.left_count = nodes.size() - (p.parent + 1);
    make_node(p.box, p.start, p.end);
    cur = nodes.size() - 1;
    depth = p.depth;
  }
}

void merge_stats(BuildStats &into, const BuildStats &from) {
  into.balance_fallbacks += from.balance_fallbacks;
  into.balance_fallbacks_gt10 += from.balance_fallbacks_gt10;
  into.bucket_out_of_range = into.bucket_out_of_range || from.bucket_out_of_range;
}

}  // namespace

FlattenBVH FlattenBVH::build(const Box3 *boxes, uint64_t n, BVHBuildStrategy &strategy, const TreeBuildOption &option, unsigned n_threads) {
  const auto t_begin = std::chrono::steady_clock::now();
  FlattenBVH out;
  std::vector<BuildPrimitive> primitives(n);
  out.sorted_primitive_index.resize(n);
  parallel_for(n, PARALLEL_BUILD_MIN, [&](uint64_t begin, uint64_t end) {
    for (uint64_t i = begin; i < end; ++i) {
      primitives[i].bounding = boxes[i];
      primitives[i].center = box_center(boxes[i]);
      out.sorted_primitive_index[i] = i;
    }
  });
  const Box3 root_box = bounding_of_range(primitives, out.sorted_primitive_index, 0, n);
  if (n_threads == 0) n_threads = build_thread_count();
  if (n_threads <= 1 || n < PARALLEL_BUILD_MIN) {
    build_subtree(root_box, 0, n, 0, strategy, option, primitives, out.sorted_primitive_index, out.nodes, out.stats);
    return out;
  }

  // ---- top of the tree on this thread: split the largest open subtree until there are enough of them
  struct Top { Box3 box; uint64_t start, end, depth; int left = -1, right = -1; int32_t axis = 0; int task = -1; };
  std::vector<Top> top;
  top.push_back(Top{root_box, 0, n, 0});
  std::vector<int> open{0};
  const size_t want_open = static_cast<size_t>(n_threads) * 8;
  const uint64_t min_task = 2048;
  while (open.size() < want_open) {
    size_t pick = open.size();
    uint64_t largest = min_task;
    for (size_t k = 0; k < open.size(); ++k) {
      const Top &t = top[open[k]];
      const uint64_t count = t.end - t.start;
      if (count > largest && option.should_continue(count, t.depth)) { largest = count; pick = k; }
    }
    if (pick == open.size()) break;
    const int ti = open[pick];
    FlattenBVHNode node;
    std::memset(&node, 0, sizeof(node));
    node.bounding = top[ti].box; node.primitive_start = top[ti].start; node.primitive_end = top[ti].end;
    const SplitResult s = strategy.split(node, primitives, out.sorted_primitive_index, out.stats);
    const uint64_t depth = top[ti].depth + 1;
    top[ti].axis = s.axis;
    top[ti].left = static_cast<int>(top.size());
    top.push_back(Top{s.left_box, s.left_start, s.left_end, depth});
    top[ti].right = static_cast<int>(top.size());
    top.push_back(Top{s.right_box, s.right_start, s.right_end, depth});
    open[pick] = top[ti].left;
    open.push_back(top[ti].right);
  }

  const bool timing = getenv("RDN_BUILD_TIMING") != nullptr;
  const auto t_top = std::chrono::steady_clock::now();
  // ---- the open subtrees on worker threads, largest first
  struct Task { int top; std::vector<FlattenBVHNode> nodes; BuildStats stats; };
  std::vector<Task> tasks(open.size());
  for (size_t k = 0; k < open.size(); ++k) { tasks[k].top = open[k]; top[open[k]].task = static_cast<int>(k); }
  std::vector<size_t> order(tasks.size());
  for (size_t k = 0; k < order.size(); ++k) order[k] = k;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    const uint64_t ca = top[tasks[a].top].end - top[tasks[a].top].start, cb = top[tasks[b].top].end - top[tasks[b].top].start;
    return ca != cb ? ca > cb : a < b;
  });
  std::atomic<size_t> next{0};
  auto worker = [&]() {
    std::unique_ptr<BVHBuildStrategy> mine = strategy.clone();
    for (;;) {
      const size_t k = next.fetch_add(1);
      if (k >= order.size()) break;
      Task &task = tasks[order[k]];
      const Top &t = top[task.top];
      build_subtree(t.box, t.start, t.end, t.depth, *mine, option, primitives, out.sorted_primitive_index, task.nodes, task.stats);
    }
  };
  const unsigned workers = static_cast<unsigned>(std::min<size_t>(n_threads, tasks.size()));
  std::vector<std::thread> pool;
  for (unsigned w = 1; w < workers; ++w) pool.emplace_back(worker);
  worker();
  for (auto &th : pool) th.join();
  out.stats.build_threads = workers;
  const auto t_workers = std::chrono::steady_clock::now();

  // ---- splice in pre-order
  uint64_t total = 0;
  for (const Task &task : tasks) { total += task.nodes.size(); merge_stats(out.stats, task.stats); }
  out.nodes.reserve(total + top.size());
  struct Frame { int top; uint64_t node; int stage; };
  std::vector<Frame> stack{Frame{0, 0, 0}};
  while (!stack.empty()) {
    Frame &f = stack.back();
    const Top &t = top[f.top];
    if (t.task >= 0) {
      const uint64_t base = out.nodes.size();
      for (FlattenBVHNode nd : tasks[t.task].nodes) { nd.self_index += base; out.nodes.push_back(nd); }
      stack.pop_back();
      continue;
    }
    if (f.stage == 0) {
      FlattenBVHNode nd;
      std::memset(&nd, 0, sizeof(nd));
      nd.bounding = t.box; nd.primitive_start = t.start; nd.primitive_end = t.end; nd.self_index = out.nodes.size();
      nd.has_child = 1; nd.split_axis = t.axis;
      f.node = out.nodes.size();
      out.nodes.push_back(nd);
      f.stage = 1;
      const int left = t.left;
      stack.push_back(Frame{left, 0, 0});
    } else if (f.stage == 1) {
      out.nodes[f.node].left_count = out.nodes.size() - (f.node + 1);
      f.stage = 2;
      const int right = t.right;
      stack.push_back(Frame{right, 0, 0});
    } else {
      stack.pop_back();
    }
  }
  if (timing) {
    const auto t_end = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[rdn build] %llu primitives: setup+top %.1f ms (%zu open subtrees), workers %.1f ms on %u threads, splice %.1f ms\n",
            static_cast<unsigned long long>(n), ms(t_begin, t_top), tasks.size(), ms(t_top, t_workers), workers, ms(t_workers, t_end));
  }
  return out;
}

std::vector<std::pair<uint32_t, uint32_t>> compute_bvh_next(const std::vector<FlattenBVHNode> &nodes) {
  std::vector<std::pair<uint32_t, uint32_t>> result;
  result.reserve(nodes.size());
  std::vector<uint32_t> pending_right;
  for (const FlattenBVHNode &node : nodes) {
    if (!pending_right.empty() && pending_right.back() == static_cast<uint32_t>(node.self_index)) pending_right.pop_back();
    const uint32_t miss = pending_right.empty() ? INVALID_NEXT : pending_right.back();
    if (node.has_child) {
      pending_right.push_back(static_cast<uint32_t>(node.right_child_offset()));
      result.emplace_back(static_cast<uint32_t>(node.left_child_offset()), miss);
    } else {
      result.emplace_back(miss, miss);
    }
  }
  return result;
}

}  // namespace rdn
