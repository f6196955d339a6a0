// bvh_builder.h — host-side FlattenBVH builder (the flattener's input).
//
// Mirrors the reference's space-query surface, same names and argument meaning:
//   FlattenBVH::new / FlattenBVHNode      content/space/src/bvh/mod.rs:26-79, bvh/node.rs:5-53
//   BVHBuildStrategy / BalanceTree / SAH  content/space/src/bvh/strategy.rs:11-284
//   TreeBuildOption                       content/space/src/utils.rs:20-37
//   compute_bvh_next                      shader/ray-tracing/.../geometry/naive/mod.rs:612-632
// The tree must be the reference's tree node for node (pre-order, left = self+1,
// right = self+left_count+1): visit order decides which of two equal-distance hits is reported.
#pragma once
#include <cstdint>
#include <functional>
#include <memory>
#include <utility>
#include <vector>

#include "big_vector.h"
#include "rdn_math.h"

namespace rdn {

struct TreeBuildOption {
  uint64_t max_tree_depth = 10;
  uint64_t bin_size = 50;
  bool should_continue(uint64_t item_count, uint64_t depth) const { return depth < max_tree_depth && item_count > bin_size; }
};

// layout-compatible with rdn_flat_bvh_node (include/rdn_rt.h)
struct FlattenBVHNode {
  Box3 bounding;
  uint64_t primitive_start, primitive_end;  // range into sorted_primitive_index
  uint64_t self_index;
  uint64_t left_count;                      // nodes in the left subtree, valid iff has_child
  int32_t has_child;
  int32_t split_axis;                       // 0 X, 1 Y, 2 Z
  bool is_leaf() const { return !has_child; }
  uint64_t left_child_offset() const { return self_index + 1; }
  uint64_t right_child_offset() const { return self_index + left_count + 1; }
};

struct BuildPrimitive {
  Box3 bounding;
  Vec3 center;
};

struct SplitResult {
  Box3 left_box, right_box;
  uint64_t left_start, left_end, right_start, right_end;
  int32_t axis;
};

struct BuildStats {
  uint64_t balance_fallbacks = 0;       // SAH -> BalanceTree fallbacks (all primitives in one bucket)
  uint64_t balance_fallbacks_gt10 = 0;  // ... over more than 10 primitives (Rust's select_nth order is unspecified there)
  bool bucket_out_of_range = false;     // the reference would have panicked
  // regularity classification of the flattener (accel.cpp): what the ordered kernel may not prune around
  uint64_t irregular_triangles = 0;     // needle / non-finite triangle records (their test can pass far outside their leaf box)
  uint64_t irregular_instances = 0;     // singular / non-finite / ill-conditioned transforms, or instances of a BLAS with irregular triangles
  uint64_t reference_routed_tlas = 0;   // TLASes whose rays all take the reference-order kernel
  // wall-clock of the last build (ms): FlattenBVH::build calls, the rest of build_blas/build_tlas (boxes, triangle records,
  // wide nodes, threaded layout), blob serialisation + upload
  double bvh_build_ms = 0, flatten_ms = 0, upload_ms = 0;
  uint32_t build_threads = 1;           // worker threads FlattenBVH::build used for its largest tree
  uint64_t device_built_trees = 0;      // geometry trees that came from the device SAH builder (build_device.cu)
};

class BVHBuildStrategy {
 public:
  virtual ~BVHBuildStrategy() = default;
  virtual SplitResult split(const FlattenBVHNode &parent, const BigVector<BuildPrimitive> &build_source,
                            BigVector<uint64_t> &index_source, BuildStats &stats) = 0;
  // a strategy object holds scratch (SAH's buckets): every worker thread of the parallel build splits with its own copy
  virtual std::unique_ptr<BVHBuildStrategy> clone() const = 0;
};

class BalanceTree : public BVHBuildStrategy {
 public:
  SplitResult split(const FlattenBVHNode &parent, const BigVector<BuildPrimitive> &build_source,
                    BigVector<uint64_t> &index_source, BuildStats &stats) override;
  std::unique_ptr<BVHBuildStrategy> clone() const override { return std::make_unique<BalanceTree>(); }
};

class SAH : public BVHBuildStrategy {
 public:
  explicit SAH(uint32_t pre_partition_check_count);
  uint32_t bucket_count() const { return static_cast<uint32_t>(pre_partition_.size()); }
  SplitResult split(const FlattenBVHNode &parent, const BigVector<BuildPrimitive> &build_source,
                    BigVector<uint64_t> &index_source, BuildStats &stats) override;
  // (a worker's copy splits on its own thread only: the workers already occupy the cores)
  std::unique_ptr<BVHBuildStrategy> clone() const override {
    auto copy = std::make_unique<SAH>(static_cast<uint32_t>(pre_partition_.size()));
    copy->parallel_split_ = false;
    return copy;
  }

 private:
  bool parallel_split_ = true;
  struct Bucket {
    Box3 bounding;
    std::vector<uint64_t> primitive_bucket;
  };
  std::vector<Bucket> pre_partition_;
  std::vector<uint64_t> counts_;
  // scratch of one split (a split per inner node must not allocate): bucket of every primitive of the range, the range before the
  // rewrite, per-thread bucket counts / boxes / rewrite offsets of the all-threads split
  std::vector<uint8_t> which_of_;
  std::vector<uint64_t> old_, chunk_counts_, offset_;
  std::vector<Box3> chunk_boxes_;
};

struct FlattenBVH {
  BigVector<FlattenBVHNode> nodes;
  BigVector<uint64_t> sorted_primitive_index;
  BuildStats stats;

  // The reference builds on one thread (`// todo par`, naive/mod.rs:147).  A split only reads and rewrites its own index
  // range, so disjoint subtrees are independent: above PARALLEL_BUILD_MIN primitives the top of the tree is split on the
  // calling thread until there are enough open subtrees, the subtrees are built by worker threads (each with its own clone
  // of the strategy) and spliced back in pre-order — node for node the tree of the sequential build.
  // n_threads: 0 = hardware concurrency capped by the RDN_BUILD_THREADS environment variable, 1 = sequential.
  static FlattenBVH build(const Box3 *boxes, uint64_t n, BVHBuildStrategy &strategy, const TreeBuildOption &option, unsigned n_threads = 0);
};
constexpr uint64_t PARALLEL_BUILD_MIN = 1u << 12;
constexpr uint64_t PARALLEL_SPLIT_MIN = 1u << 11;  // a single SAH split over at least this many primitives uses all threads too

// CPUs this process may use: its affinity mask, capped by its share of the machine under torchrun (LOCAL_WORLD_SIZE); RDN_POOL_THREADS overrides
unsigned usable_cpu_count();
// worker threads for host-side loops: min(usable_cpu_count(), RDN_BUILD_THREADS), at least 1
unsigned build_thread_count();
// fn(begin, end) over [0, n) in contiguous chunks, one per thread (sequential below `min_parallel` items)
void parallel_for(uint64_t n, uint64_t min_parallel, const std::function<void(uint64_t, uint64_t)> &fn);
// fn(0) on the calling thread and fn(1) .. fn(n - 1) on the process-wide worker pool (bvh_builder.cpp; own threads when the pool is taken)
void run_parallel(unsigned n, const std::function<void(unsigned)> &fn);
// hint that parallel sections are about to follow (a commit): sleeping pool workers start spinning again
void warm_worker_pool();

constexpr uint32_t INVALID_NEXT = 0xFFFFFFFFu;
// (hit_next, miss_next) per node for the stackless threaded walk
std::vector<std::pair<uint32_t, uint32_t>> compute_bvh_next(const BigVector<FlattenBVHNode> &nodes);

// box3.rs helpers the builder and the TLAS assembly need
int longest_axis(const Box3 &b);
float surface_area(const Box3 &b);
Vec3 box_center(const Box3 &b);

}  // namespace rdn
