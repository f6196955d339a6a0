// bvh_builder.cpp — see bvh_builder.h.  Compiled with -ffp-contract=off.
#include "bvh_builder.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>

#include <unistd.h>
#if defined(__linux__)
#include <sched.h>
#endif

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

static Box3 bounding_of_range(const BigVector<BuildPrimitive> &src, const BigVector<uint64_t> &index, uint64_t begin,
                              uint64_t end) {
  Box3 r = box_empty();
  for (uint64_t i = begin; i < end; ++i) expand(r, src[index[i]].bounding);
  return r;
}

SplitResult BalanceTree::split(const FlattenBVHNode &parent, const BigVector<BuildPrimitive> &src,
                               BigVector<uint64_t> &index, BuildStats &) {
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

SplitResult SAH::split(const FlattenBVHNode &parent, const BigVector<BuildPrimitive> &src, BigVector<uint64_t> &index,
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
    // (the reference floors before the cast; the saturating cast truncates towards zero and sends everything negative to 0, so the
    // floor changes nothing and the libm call is left out)
    uint64_t which = saturating_usize((axis_value - range_start) / step);
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
  // (member scratch: a split per inner node must not allocate)
  std::vector<uint8_t> &which_of = which_of_;
  std::vector<uint64_t> &chunk_counts = chunk_counts_;  // [thread][bucket]
  std::vector<Box3> &chunk_boxes = chunk_boxes_;
  const uint64_t n = end - begin;
  uint64_t chunk = 0;
  if (parallel) {
    chunk = (n + threads - 1) / threads;
    which_of.resize(n);
    chunk_counts.assign(static_cast<size_t>(threads) * n_part, 0);
    chunk_boxes.assign(static_cast<size_t>(threads) * n_part, box_empty());
    std::atomic<bool> out_of_range{false};
    parallel_for(n, 0, [&](uint64_t b0, uint64_t b1) {
      const size_t t = static_cast<size_t>(b0 / chunk);
      bool oor = false;
      uint64_t my_counts[256] = {0};  // thread-private while counting: the shared rows sit in one cache line
      Box3 my_boxes[256];
      for (size_t k = 0; k < n_part; ++k) my_boxes[k] = box_empty();
      for (uint64_t i = b0; i < b1; ++i) {
        const uint64_t prim = index[begin + i];
        const size_t w = bucket_of(prim, oor);
        which_of[i] = static_cast<uint8_t>(w);
        my_counts[w]++;
        expand(my_boxes[w], src[prim].bounding);
      }
      for (size_t k = 0; k < n_part; ++k) { chunk_counts[t * n_part + k] = my_counts[k]; chunk_boxes[t * n_part + k] = my_boxes[k]; }
      if (oor) out_of_range = true;
    });
    if (out_of_range) stats.bucket_out_of_range = true;
    for (size_t t = 0; t < threads; ++t)
      for (size_t k = 0; k < n_part; ++k) { counts[k] += chunk_counts[t * n_part + k]; expand(pre_partition_[k].bounding, chunk_boxes[t * n_part + k]); }
  } else if (n_part <= 255) {
    // one pass that notes the bucket of every primitive; the range is rewritten by a counting scatter below
    which_of.resize(n);
    for (uint64_t i = 0; i < n; ++i) {
      const uint64_t prim = index[begin + i];
      const size_t which = bucket_of(prim, stats.bucket_out_of_range);
      which_of[i] = static_cast<uint8_t>(which);
      expand(pre_partition_[which].bounding, src[prim].bounding);
      counts[which]++;
    }
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
    std::vector<uint64_t> &old = old_;
    old.assign(index.begin() + begin, index.begin() + end);
    std::vector<uint64_t> &offset = offset_;  // [thread][bucket]
    offset.assign(static_cast<size_t>(threads) * n_part, 0);
    uint64_t ptr = begin;
    for (size_t k = 0; k < n_part; ++k)
      for (size_t t = 0; t < threads; ++t) { offset[t * n_part + k] = ptr; ptr += chunk_counts[t * n_part + k]; }
    parallel_for(n, 0, [&](uint64_t b0, uint64_t b1) {
      uint64_t off[256];  // private copy (see above)
      const size_t t = static_cast<size_t>(b0 / chunk);
      for (size_t k = 0; k < n_part; ++k) off[k] = offset[t * n_part + k];
      for (uint64_t i = b0; i < b1; ++i) index[off[which_of[i]]++] = old[i];
    });
  } else if (n_part <= 255) {
    std::vector<uint64_t> &old = old_;
    old.assign(index.begin() + begin, index.begin() + end);
    uint64_t off[256];
    uint64_t ptr = begin;
    for (size_t k = 0; k < n_part; ++k) { off[k] = ptr; ptr += counts[k]; }
    for (uint64_t i = 0; i < n; ++i) index[off[which_of[i]]++] = old[i];
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

// CPUs this process may reasonably use: those its affinity mask allows, and — one process per GPU under torchrun, several of them
// bound to the same socket — no more than its share of the machine (LOCAL_WORLD_SIZE).  Eight ranks that each start a thread per
// CPU of the box run slower than eight ranks with eight threads each (the pageable host-buffer path at 8 GPUs: 604 against 914
// Mrays/s for the driver's own staging, profiles/bench_r3m_n8.json).  RDN_POOL_THREADS overrides.
unsigned usable_cpu_count() {
  unsigned hw = std::thread::hardware_concurrency();
  if (hw == 0) hw = 1;
  unsigned allowed = hw;
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) allowed = static_cast<unsigned>(CPU_COUNT(&set));
#endif
  unsigned ranks = 1;
  if (const char *e = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(e); if (v > 1) ranks = static_cast<unsigned>(v); }
  unsigned n = std::min(allowed, (hw + ranks - 1) / ranks);
  if (const char *e = getenv("RDN_POOL_THREADS")) { const int v = atoi(e); if (v >= 1) n = static_cast<unsigned>(v); }
  return n ? n : 1u;
}

unsigned build_thread_count() {
  unsigned n = usable_cpu_count();
  if (const char *e = getenv("RDN_BUILD_THREADS")) {
    const int cap = atoi(e);
    if (cap >= 1 && static_cast<unsigned>(cap) < n) n = static_cast<unsigned>(cap);
  }
  return n;
}

// ------------------------------------------------------------------------------------------------ worker pool
// A commit of a few thousand moving instances is a handful of parallel sections of some tens of microseconds each; starting and
// joining sixteen threads costs more than that per section.  The sections run on a process-wide pool instead.  A section is a
// list of chunks claimed one at a time through one atomic word (no lock on the way in: sixteen threads queueing on a mutex cost
// more than the section) by the calling thread and by whatever workers are awake: the
// caller never waits for a worker to wake up, only for the chunks that were claimed to finish — workers that were asleep (they
// spin for a moment after a section, then sleep on a condition variable) join when they get there, or find nothing left.
// Sections of different callers (concurrent commits of different scenes) run on the same workers one after another; a section
// requested from inside a worker runs on the caller alone.  After a fork() the child gets a fresh pool.
namespace {

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#elif defined(__aarch64__)
  asm volatile("yield" ::: "memory");
#endif
}

thread_local bool t_pool_worker = false;

class WorkerPool {
 public:
  explicit WorkerPool(unsigned workers) {
    for (unsigned i = 0; i < workers; ++i) threads_.emplace_back([this]() { worker_main(); });
  }
  ~WorkerPool() {
    stop_.store(true);
    { std::lock_guard<std::mutex> g(sleep_mutex_); sleep_cv_.notify_all(); }
    for (auto &t : threads_) t.join();
  }

  // a hint that sections are about to follow: sleeping workers go back to spinning (waking one takes longer than a short section)
  void wake() {
    wake_epoch_.fetch_add(1);
    if (sleepers_.load() != 0) { std::lock_guard<std::mutex> g(sleep_mutex_); sleep_cv_.notify_all(); }
  }

  void run(unsigned n, const std::function<void(unsigned)> &fn) {
    std::lock_guard<std::mutex> one_section(run_mutex_);
    struct InSection {  // (a section opened from inside a chunk the caller itself runs stays on the caller: see the_pool)
      InSection() { t_pool_worker = true; }
      ~InSection() { t_pool_worker = false; }
    } in_section;
    const uint64_t g = (ticket_.load(std::memory_order_relaxed) >> 32) + 1;
    fn_.store(&fn, std::memory_order_relaxed);
    n_.store(n, std::memory_order_relaxed);
    done_.store(0, std::memory_order_relaxed);
    ticket_.store(g << 32);                   // publishes the section (sequentially consistent with the sleepers' count, see worker_main)
    if (sleepers_.load() != 0) { std::lock_guard<std::mutex> lk(sleep_mutex_); sleep_cv_.notify_all(); }
    work(g);
    for (unsigned spins = 0; done_.load(std::memory_order_acquire) != n; ++spins) {
      if (spins < 4096) cpu_relax(); else std::this_thread::yield();
    }
  }

 private:
  // Claims chunks of section g until none is left.  The ticket holds (section, next chunk): a claim is a compare-exchange on it, so a
  // worker that read the section's description late — the caller may have moved on to the next section meanwhile — fails the
  // exchange and throws what it read away; a chunk that was claimed is always run (the caller waits for done_ before `fn` dies).
  void work(uint64_t g) {
    for (;;) {
      uint64_t t = ticket_.load(std::memory_order_acquire);
      if ((t >> 32) != g) return;
      const std::function<void(unsigned)> *fn = fn_.load(std::memory_order_relaxed);
      const unsigned n = n_.load(std::memory_order_relaxed);
      const unsigned c = static_cast<unsigned>(t & 0xFFFFFFFFu);
      if (c >= n) {
        if (ticket_.load(std::memory_order_acquire) == t) return;  // (n belongs to this section: nothing left)
        continue;
      }
      if (!ticket_.compare_exchange_weak(t, t + 1, std::memory_order_acq_rel)) continue;
      (*fn)(c);
      done_.fetch_add(1, std::memory_order_release);
    }
  }

  void worker_main() {
    t_pool_worker = true;
    uint64_t seen = 0;
    for (;;) {
      auto t_idle = std::chrono::steady_clock::now();
      for (unsigned spins = 0; (ticket_.load(std::memory_order_acquire) >> 32) == seen && !stop_.load(std::memory_order_relaxed); ++spins) {
        if ((spins & 255u) != 255u || std::chrono::steady_clock::now() - t_idle < std::chrono::microseconds(300)) { cpu_relax(); continue; }
        std::unique_lock<std::mutex> lk(sleep_mutex_);
        const uint64_t woken = wake_epoch_.load();
        sleepers_.fetch_add(1);
        sleep_cv_.wait(lk, [&]() { return (ticket_.load() >> 32) != seen || wake_epoch_.load() != woken || stop_.load(); });
        sleepers_.fetch_sub(1);
        t_idle = std::chrono::steady_clock::now();  // (woken by a hint: spin for another while)
      }
      if (stop_.load()) return;
      seen = ticket_.load(std::memory_order_acquire) >> 32;
      work(seen);
    }
  }

  std::vector<std::thread> threads_;
  std::mutex run_mutex_, sleep_mutex_;
  std::condition_variable sleep_cv_;
  std::atomic<uint64_t> ticket_{0};   // (section number << 32) | next chunk of it
  std::atomic<const std::function<void(unsigned)> *> fn_{nullptr};
  std::atomic<unsigned> n_{0}, done_{0};
  std::atomic<uint64_t> wake_epoch_{0};
  std::atomic<unsigned> sleepers_{0};
  std::atomic<bool> stop_{false};
};

struct PoolHolder {
  std::mutex mutex;
  WorkerPool *pool = nullptr;
  pid_t pid = 0;
  ~PoolHolder() { if (pool && pid == getpid()) delete pool; }  // (a forked child must not join threads it never had)
};
PoolHolder g_pool;

WorkerPool *the_pool() {
  if (t_pool_worker) return nullptr;
  static const bool disabled = []() { const char *e = getenv("RDN_BUILD_POOL"); return e && atoi(e) == 0; }();
  if (disabled) return nullptr;
  std::lock_guard<std::mutex> g(g_pool.mutex);
  const pid_t me = getpid();
  if (!g_pool.pool || g_pool.pid != me) {  // first use, or first use in a forked child (the parent's pool object is leaked there)
    const unsigned cpus = usable_cpu_count();
    if (cpus <= 1) return nullptr;
    g_pool.pool = new WorkerPool(cpus - 1);
    g_pool.pid = me;
  }
  return g_pool.pool;
}

bool run_on_pool(unsigned n, const std::function<void(unsigned)> &fn) {
  WorkerPool *pool = the_pool();
  if (!pool) return false;
  pool->run(n, fn);
  return true;
}

}  // namespace

void warm_worker_pool() {
  if (WorkerPool *pool = the_pool()) pool->wake();
}

void run_parallel(unsigned n, const std::function<void(unsigned)> &fn) {
  if (n <= 1) { if (n == 1) fn(0); return; }
  if (run_on_pool(n, fn)) return;
  for (unsigned t = 0; t < n; ++t) fn(t);  // (inside a worker, a single-core host, RDN_BUILD_POOL=0: the caller alone)
}

void parallel_for(uint64_t n, uint64_t min_parallel, const std::function<void(uint64_t, uint64_t)> &fn) {
  const unsigned threads = build_thread_count();
  if (threads <= 1 || n < min_parallel) { fn(0, n); return; }
  const uint64_t chunk = (n + threads - 1) / threads;
  const unsigned used = static_cast<unsigned>((n + chunk - 1) / chunk);  // chunks that hold anything
  run_parallel(used, [&](unsigned t) {
    const uint64_t begin = std::min<uint64_t>(n, t * chunk), end = std::min<uint64_t>(n, begin + chunk);
    if (begin < end) fn(begin, end);
  });
}

namespace {

// Pre-order construction of the subtree over [start, end) without recursion: descend left, park the right sibling; when a
// leaf is reached the most recent parked sibling is emitted next, which fixes its parent's left_count.  Node indices are local
// to `nodes` (the subtree root is node 0); left_count is position independent.
void build_subtree(const Box3 &box, uint64_t start, uint64_t end, uint64_t depth0, BVHBuildStrategy &strategy, const TreeBuildOption &option,
                   const BigVector<BuildPrimitive> &primitives, BigVector<uint64_t> &index, BigVector<FlattenBVHNode> &nodes,
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
  BigVector<BuildPrimitive> primitives(n);
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

  const auto t_setup = std::chrono::steady_clock::now();
  // ---- top of the tree on this thread: split the largest open subtree until there are enough of them
  struct Top { Box3 box; uint64_t start, end, depth; int left = -1, right = -1; int32_t axis = 0; int task = -1; };
  std::vector<Top> top;
  top.push_back(Top{root_box, 0, n, 0});
  std::vector<int> open{0};
  const size_t want_open = static_cast<size_t>(n_threads) * 8;
  // (subtrees below this size are not split further on the calling thread; RDN_BUILD_MIN_TASK: measurement knob)
  static const uint64_t min_task = []() { const char *e = getenv("RDN_BUILD_MIN_TASK"); return e && atoll(e) > 0 ? static_cast<uint64_t>(atoll(e)) : 512ull; }();
  // Rounds: every open subtree that is still worth splitting is split once per round.  Several small ones are split side by side,
  // one thread each with its own copy of the strategy (disjoint index ranges); a few large ones one after another, each by all
  // threads (SAH::split).  Which subtrees end up as tasks does not change the tree: a split depends on its own range only.
  while (open.size() < want_open) {
    std::vector<size_t> cand;
    for (size_t k = 0; k < open.size(); ++k) {
      const Top &t = top[open[k]];
      const uint64_t count = t.end - t.start;
      if (count > min_task && option.should_continue(count, t.depth)) cand.push_back(k);
    }
    if (cand.empty()) break;
    std::sort(cand.begin(), cand.end(), [&](size_t a, size_t b) {
      const uint64_t ca = top[open[a]].end - top[open[a]].start, cb = top[open[b]].end - top[open[b]].start;
      return ca != cb ? ca > cb : a < b;
    });
    if (cand.size() > want_open - open.size()) cand.resize(want_open - open.size());
    const uint64_t largest = top[open[cand[0]]].end - top[open[cand[0]]].start;
    std::vector<SplitResult> results(cand.size());
    std::vector<BuildStats> split_stats(cand.size());
    auto split_one = [&](BVHBuildStrategy &with, size_t k) {
      const Top &t = top[open[cand[k]]];
      FlattenBVHNode node;
      std::memset(&node, 0, sizeof(node));
      node.bounding = t.box; node.primitive_start = t.start; node.primitive_end = t.end;
      const auto t_split = std::chrono::steady_clock::now();
      results[k] = with.split(node, primitives, out.sorted_primitive_index, split_stats[k]);
      if (getenv("RDN_BUILD_TIMING_SPLITS"))
        fprintf(stderr, "[rdn build] top split of %llu primitives: %.1f us\n", static_cast<unsigned long long>(t.end - t.start),
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_split).count());
    };
    if (cand.size() > 1 && (cand.size() >= n_threads || largest < 8192)) {  // (enough of them to occupy every thread, or small)
      run_parallel(static_cast<unsigned>(cand.size()), [&](unsigned k) {
        std::unique_ptr<BVHBuildStrategy> mine = strategy.clone();
        split_one(*mine, k);
      });
    } else {
      for (size_t k = 0; k < cand.size(); ++k) split_one(strategy, k);
    }
    for (size_t k = 0; k < cand.size(); ++k) {
      merge_stats(out.stats, split_stats[k]);
      const int ti = open[cand[k]];
      const SplitResult &s = results[k];
      const uint64_t depth = top[ti].depth + 1;
      top[ti].axis = s.axis;
      top[ti].left = static_cast<int>(top.size());
      top.push_back(Top{s.left_box, s.left_start, s.left_end, depth});
      top[ti].right = static_cast<int>(top.size());
      top.push_back(Top{s.right_box, s.right_start, s.right_end, depth});
      open[cand[k]] = top[ti].left;
      open.push_back(top[ti].right);
    }
  }

  const bool timing = getenv("RDN_BUILD_TIMING") != nullptr;
  const auto t_top = std::chrono::steady_clock::now();
  // ---- the open subtrees on worker threads, largest first
  struct Task { int top; BigVector<FlattenBVHNode> nodes; BuildStats stats; };
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
  run_parallel(workers, [&](unsigned) { worker(); });
  out.stats.build_threads = workers;
  const auto t_workers = std::chrono::steady_clock::now();

  // ---- splice in pre-order: the top nodes and every task's position are laid out by one walk over the (small) top of the tree,
  // then the tasks' nodes are copied to their places by all threads
  uint64_t total = 0;
  for (const Task &task : tasks) { total += task.nodes.size(); merge_stats(out.stats, task.stats); }
  out.nodes.resize(total + top.size() - tasks.size());
  std::vector<uint64_t> task_base(tasks.size(), 0);
  {
    uint64_t cursor = 0;
    struct Frame { int top; uint64_t node; int stage; };
    std::vector<Frame> stack{Frame{0, 0, 0}};
    while (!stack.empty()) {
      Frame &f = stack.back();
      const Top &t = top[f.top];
      if (t.task >= 0) {
        task_base[t.task] = cursor;
        cursor += tasks[t.task].nodes.size();
        stack.pop_back();
        continue;
      }
      if (f.stage == 0) {
        FlattenBVHNode nd;
        std::memset(&nd, 0, sizeof(nd));
        nd.bounding = t.box; nd.primitive_start = t.start; nd.primitive_end = t.end; nd.self_index = cursor;
        nd.has_child = 1; nd.split_axis = t.axis;
        f.node = cursor;
        out.nodes[cursor++] = nd;
        f.stage = 1;
        const int left = t.left;
        stack.push_back(Frame{left, 0, 0});
      } else if (f.stage == 1) {
        out.nodes[f.node].left_count = cursor - (f.node + 1);
        f.stage = 2;
        const int right = t.right;
        stack.push_back(Frame{right, 0, 0});
      } else {
        stack.pop_back();
      }
    }
  }
  {
    std::atomic<size_t> next_copy{0};
    run_parallel(workers, [&](unsigned) {
      for (size_t k; (k = next_copy.fetch_add(1)) < tasks.size();) {
        const uint64_t base = task_base[k];
        const BigVector<FlattenBVHNode> &from = tasks[k].nodes;
        for (size_t i = 0; i < from.size(); ++i) { FlattenBVHNode nd = from[i]; nd.self_index += base; out.nodes[base + i] = nd; }
      }
    });
  }
  if (timing) {
    const auto t_end = std::chrono::steady_clock::now();
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    fprintf(stderr, "[rdn build] %llu primitives: setup %.2f ms, top %.2f ms (%zu open subtrees), workers %.2f ms on %u threads, splice %.2f ms\n",
            static_cast<unsigned long long>(n), ms(t_begin, t_setup), ms(t_setup, t_top), tasks.size(), ms(t_top, t_workers), workers, ms(t_workers, t_end));
  }
  return out;
}

std::vector<std::pair<uint32_t, uint32_t>> compute_bvh_next(const BigVector<FlattenBVHNode> &nodes) {
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
