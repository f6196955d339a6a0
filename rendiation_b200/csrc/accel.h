// accel.h — host side of the software acceleration-structure system (build + flatten).
//
// Mirrors NaiveSahBVHSystem / NaiveSahBvhSource (shader/ray-tracing/src/backend/wavefront_compute/geometry/
// naive/mod.rs:73-119,122-493): create/delete BLAS & TLAS only store sources and invalidate; the actual build
// is lazy (mod.rs:521-536) and always from scratch (mod.rs:121 `todo incremental change`).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bvh_builder.h"
#include "layout.h"

namespace rdn {

struct GeometrySource {  // BottomLevelAccelerationStructureBuildSource (api/backend.rs:104-118)
  std::vector<Vec3> positions;
  std::vector<uint32_t> indices;
  bool has_indices = false;
  bool is_aabbs = false;  // accepted, ignored by the naive builder (mod.rs:201-237)
  uint32_t flags = 0;
};

struct InstanceSource {  // TopLevelAccelerationStructureSourceInstance (api/backend.rs:160-167)
  Mat4 transform;
  uint32_t instance_custom_index, mask, sbt_offset, flags, blas_handle;
};

// The flattened scene on the host, one vector per blob array.
struct FlatScene {
  std::vector<uint32_t> tlas_binding;
  std::vector<TlasRoot> tlas_root;
  std::vector<DeviceBVHNode> tlas_bvh_forest;
  std::vector<TlasBounding> tlas_bounding;
  std::vector<InstanceRecord> instances;
  std::vector<BlasMeta> blas_meta;
  std::vector<GeometryMeta> geometry_meta;
  BigVector<DeviceBVHNode> tri_bvh_forest;
  BigVector<TriRecord> triangles;
  BigVector<SlotInfo> slot_info;
  BigVector<WideNode> wide_nodes;
  std::vector<uint32_t> prim_to_slot;
  std::vector<uint32_t> irregular_instances;
  std::vector<LeafBox> irregular_leaf_boxes;
  BigVector<Wide4Node> wide4_nodes;
  BuildStats stats;

  // pack into one contiguous, BLOB_ALIGN-aligned byte image starting with a BlobHeader
  std::vector<uint8_t> serialize() const;
  // the same image without building it: where every array goes (header filled in, total_bytes included) and where it comes from
  struct ArrayRef { int id; const void *data; uint64_t bytes; };
  BlobHeader layout(std::vector<ArrayRef> &arrays) const;
};

class NaiveSahBvhSource {
 public:
  uint32_t create_blas(std::vector<GeometrySource> source);
  uint32_t create_tlas(std::vector<InstanceSource> source);
  bool delete_blas(uint32_t handle);
  bool delete_tlas(uint32_t handle);

  // the reference has no such call (mod.rs:121 `todo incremental change`): replace the instances of a live TLAS in place.  When only
  // TLASes changed since the last build, the next build keeps every BLAS array of the previous FlatScene and redoes build_tlas alone.
  bool update_tlas(uint32_t handle, std::vector<InstanceSource> source);

  // NaiveSahBvhSource::build: returns 0 or a negative rdn_status; `err` gets a message on failure.
  // `previous`: the FlatScene of the previous successful build, consumed when it can be reused (same BLAS set, same flattener
  // switches): `out` then starts from it and only the TLAS part is rebuilt; *reused says whether that happened.
  int build(const std::vector<uint32_t> &tlas_binding, FlatScene &out, std::string &err, FlatScene *previous = nullptr, bool *reused = nullptr) const;
  // first wide node / four-box node of the TLAS part (what a reusing build rewrote from there on)
  uint64_t tlas_part_wide_start() const { return cache_.wide_nodes_n; }
  uint64_t tlas_part_wide4_start() const { return cache_.wide4_nodes_n; }

  // Opt-in (RDN_COMMIT_DEVICE_BUILD=1, set by the C ABI for scenes that own a CUDA device): geometry trees of at least
  // `device_build_min` triangles are built by the device SAH builder (build_device.cu: the reference's tree node for node) and by
  // the host builder when the device declines (more than four buckets, a long degenerate range).  -1: host builder only.
  int build_device = -1;
  uint64_t device_build_min = 1u << 16;

 private:
  struct Blas { bool alive = false; std::vector<GeometrySource> geometries; };
  struct Tlas { bool alive = false; std::vector<InstanceSource> instances; };
  std::vector<Blas> blas_data_;
  std::vector<Tlas> tlas_data_;
  uint64_t blas_epoch_ = 1;   // bumped by every BLAS mutation
  // what build_blas leaves behind for build_tlas (per BLAS handle), kept between builds
  struct OptBox { bool some; Box3 box; };
  struct HotBlock { uint32_t base = 0, count = 0; uint64_t triangles = 0; };
  struct BlasPartCache {
    uint64_t blas_epoch = 0;  // 0: nothing cached
    bool want_wide4 = false, fine_tlas = true;
    std::vector<OptBox> blas_box;
    std::vector<Box3> blas_true_box;
    std::vector<uint32_t> blas_irregular;
    std::vector<HotBlock> blas_hot;
    uint64_t wide_nodes_n = 0, wide4_nodes_n = 0;
    BuildStats stats;
  };
  mutable BlasPartCache cache_;
};

// Mat4 helpers (math/algebra/src/mat/mat4.rs:40-104, mat3.rs:37-42) used by the TLAS assembly
Mat4 mat4_inverse_or_identity(const Mat4 &m);
float mat4_upper3_det(const Mat4 &m);
Box3 box_apply_matrix(const Box3 &b, const Mat4 &m);

// Emit the wide (two-boxes-per-node) view of one reference tree into `out`; leaf references address
// `slot_offset + primitive_range`.  Returns the pseudo-root reference (REF_EMPTY for an empty tree) or
// sets `capacity_error`.
// `item_bounds` (optional; TLAS only): the exact box of every slot of this tree, slot order.  A leaf of more than one slot is then
// emitted as a small binary tree over contiguous halves of its slots (boxes = unions of the slots' boxes, leaves = single slots)
// instead of one multi-slot leaf reference: the ordered kernel finds the instances a ray can enter with node steps rather than
// one instance-box test per slot.  Visibility is unchanged — nested boxes keep the slab test monotone, and entering an instance
// still takes the reference's test of its own box.
uint32_t emit_wide_nodes(const BigVector<FlattenBVHNode> &nodes, uint64_t slot_offset, BigVector<WideNode> &out,
                         bool &capacity_error, const TlasBounding *item_bounds = nullptr);
// the same for the 4-wide view: one node per inner reference node at even depth below the root (its inner children are absorbed)
uint32_t emit_wide4_nodes(const BigVector<FlattenBVHNode> &nodes, uint64_t slot_offset, BigVector<Wide4Node> &out,
                          bool &capacity_error);

}  // namespace rdn
