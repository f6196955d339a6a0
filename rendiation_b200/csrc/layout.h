// layout.h — the flattened scene as it lives in HBM: ONE contiguous blob of 128-byte-aligned arrays
// (so replication to another GPU is a single NVLink copy / NCCL broadcast), described by BlobHeader.
//
// Two views of the same reference-built trees are stored:
//   * the reference's threaded layout (DeviceBVHNode 48 B, naive/mod.rs:43-53; TlasBounding 32 B, :33-41) —
//     walked by the reference-order kernel, which reproduces NaiveSahBvhCpu::traverse visit for visit;
//   * a wide layout for the ordered kernel: one 64 B node per INNER reference node holding both child boxes
//     (exactly the reference's f32 boxes, so every box decision is the reference's) and two child references.
// Triangles are pre-gathered in BVH-sorted slot order (64 B each: one contiguous fetch instead of the
// reference's redirect -> indices -> vertices chain, traverse_gpu.rs:526-533) together with the
// ray-independent terms of intersect_ray_triangle (geometry/mod.rs:105-155) evaluated in the reference's order.
#pragma once
#include <cstdint>

namespace rdn {

// ---- reference threaded node, byte-identical to DeviceBVHNode ----
struct DeviceBVHNode {
  float aabb_min[3];
  uint32_t hit_next;
  float aabb_max[3];
  uint32_t miss_next;
  uint32_t content_range[2];
  uint32_t tail[2];
};
static_assert(sizeof(DeviceBVHNode) == 48, "DeviceBVHNode");

struct TlasBounding {
  float world_min[3];
  uint32_t mask;
  float world_max[3];
  uint32_t flags;
};
static_assert(sizeof(TlasBounding) == 32, "TlasBounding");

// world->object matrix + the four u32 of TopLevelAccelerationStructureSourceDeviceInstance (naive/mod.rs:22-32);
// object->world is kept host side only (the traversal never reads it)
struct InstanceRecord {
  float transform_inv[16];  // column-major a1..d4
  uint32_t instance_custom_index, sbt_offset, flags, blas;
};
static_assert(sizeof(InstanceRecord) == 80, "InstanceRecord");

struct BlasMeta {
  uint32_t tri_root_range[2];     // into GeometryMeta[]
  uint32_t irregular_leaf_start;  // [start, start + count) in irregular_leaf_boxes: object-space boxes of the BVH leaves that hold an
  uint32_t irregular_leaf_count;  // irregular triangle; IRREGULAR_ROUTE_ALL: too many — every instance of the BLAS is irregular
};
static_assert(sizeof(BlasMeta) == 16, "BlasMeta");

struct LeafBox {  // 32 B
  float bmin[3]; uint32_t pad0;
  float bmax[3]; uint32_t pad1;
};
static_assert(sizeof(LeafBox) == 32, "LeafBox");

struct GeometryMeta {
  uint32_t bvh_root_idx;    // root in tri_bvh_forest (reference layout)
  uint32_t geometry_idx;
  uint32_t primitive_start;
  uint32_t geometry_flags;
  uint32_t wide_root;       // child reference of the pseudo-root in wide_nodes, or REF_EMPTY
  uint32_t wide4_root;      // ... in wide4_nodes
  uint32_t pad[2];
};
static_assert(sizeof(GeometryMeta) == 32, "GeometryMeta");

struct TlasRoot {
  uint32_t bvh_root_idx;     // root in tlas_bvh_forest or INVALID_NEXT for a deleted TLAS
  uint32_t wide_root;        // REF_EMPTY when deleted / empty
  uint32_t irregular_start;  // [start, start + count) in irregular_instances: instance slots whose hits need not lie inside
  uint32_t irregular_count;  // their boxes (see accel.cpp "regularity"); IRREGULAR_ROUTE_ALL: walk every ray in reference order
  uint32_t hot_count;        // wide nodes [wide_root, wide_root + hot_count): the top of the TLAS tree (<= HOT_TOP_NODES)
  uint32_t hot_geometry_base, hot_geometry_count;  // the same block of the largest geometry tree among the instances' BLASes
  uint32_t wide4_root;       // pseudo root of the TLAS tree in wide4_nodes (REF_EMPTY when deleted / empty)
};
static_assert(sizeof(TlasRoot) == 32, "TlasRoot");
constexpr uint32_t IRREGULAR_ROUTE_ALL = 0xFFFFFFFFu;
constexpr uint32_t IRREGULAR_LIST_MAX = 8;       // more irregular instances than this in one TLAS: the whole TLAS is walked in reference order
constexpr uint32_t IRREGULAR_LEAF_MAX = 16;      // more irregular leaves than this in one BLAS: its instances are irregular as a whole
constexpr uint32_t IRREGULAR_WHOLE_BIT = 1u << 31;  // irregular_instances entry: the whole instance (else only its BLAS's listed leaves)

// 64 B, sector 0 = plane test operands, sector 1 = barycentric operands
struct TriRecord {
  float n[3];  float inv_d;   // normalize(e1 x e2), 1/(uv*uv - uu*vv)
  float v0[3]; float uu;
  float e1[3]; float uv;
  float e2[3]; float vv;
};
static_assert(sizeof(TriRecord) == 64, "TriRecord");

struct SlotInfo {
  uint32_t primitive_id;  // indices_redirect[slot] - primitive_start (original triangle index of the geometry)
  uint32_t geometry_idx;
};

// 64 B: q0 = {c0.min, ref0}, q1 = {c0.max, ref1}, q2 = {c1.min, 0}, q3 = {c1.max, 0}
struct WideNode {
  float c0_min[3]; uint32_t ref0;
  float c0_max[3]; uint32_t ref1;
  float c1_min[3]; uint32_t pad0;
  float c1_max[3]; uint32_t pad1;
};
static_assert(sizeof(WideNode) == 64, "WideNode");

// 128 B: up to four children — the grandchildren of an inner reference node, a child that is a leaf kept as it is — child k in
// floats [8k, 8k+8): {min.xyz, ref} {max.xyz, 0}; an unused slot has a NaN box and REF_EMPTY.  Inner references index wide4_nodes.
struct Wide4Node {
  struct Child { float bmin[3]; uint32_t ref; float bmax[3]; uint32_t pad; } child[4];
};
static_assert(sizeof(Wide4Node) == 128, "Wide4Node");

// ---- child reference encoding (u32) ----
//   inner  : index into wide_nodes (< REF_SPECIAL)
//   leaf   : bit 31 | (count-1) << 27 | start     count 1..16, start < 2^27
//            (object space: triangle slots; world space: instance slots)
//   special: 0x7F000000 | payload  — REF_EMPTY, REF_EXIT_INSTANCE, geometry iterator (payload = GeometryMeta index)
constexpr uint32_t REF_LEAF_BIT = 0x80000000u;
constexpr uint32_t REF_LEAF_COUNT_SHIFT = 27;
constexpr uint32_t REF_LEAF_START_MASK = (1u << 27) - 1u;
#ifndef RDN_REF_LEAF_MAX_COUNT
#define RDN_REF_LEAF_MAX_COUNT 16  // (the emulated test build of tests/simt lowers it so that ordinary scenes walk leaf chains)
#endif
constexpr uint32_t REF_LEAF_MAX_COUNT = RDN_REF_LEAF_MAX_COUNT;
constexpr uint32_t REF_SPECIAL = 0x7F000000u;
constexpr uint32_t REF_DONE = 0x7FFFFFFDu;           // traversal stack exhausted (kernel-internal)
constexpr uint32_t REF_EMPTY = 0x7FFFFFFEu;
constexpr uint32_t REF_EXIT_INSTANCE = 0x7FFFFFFFu;
constexpr uint32_t REF_GEOM_ITER_MAX = 0x00FFFFFCu;
// the first HOT_TOP_NODES wide nodes of every tree (pseudo root first) are its top levels in breadth-first order
constexpr uint32_t HOT_TOP_NODES = 128;

enum ArrayId : int {
  ARR_TLAS_BINDING = 0,   // u32
  ARR_TLAS_ROOT,          // TlasRoot
  ARR_TLAS_BVH_FOREST,    // DeviceBVHNode
  ARR_TLAS_BOUNDING,      // TlasBounding
  ARR_INSTANCES,          // InstanceRecord
  ARR_BLAS_META,          // BlasMeta
  ARR_GEOMETRY_META,      // GeometryMeta
  ARR_TRI_BVH_FOREST,     // DeviceBVHNode
  ARR_TRIANGLES,          // TriRecord
  ARR_SLOT_INFO,          // SlotInfo
  ARR_WIDE_NODES,         // WideNode
  ARR_PRIM_TO_SLOT,       // u32: (primitive_start + original triangle index) -> slot; inverse of the reference's indices_redirect
  ARR_IRREGULAR_INSTANCES,  // u32: instance slots (| IRREGULAR_WHOLE_BIT), grouped per TLAS (TlasRoot::irregular_start / _count)
  ARR_IRREGULAR_LEAF_BOXES, // LeafBox, grouped per BLAS (BlasMeta::irregular_leaf_start / _count)
  ARR_WIDE4_NODES,          // Wide4Node: the 4-wide view of the same trees (experiment: RDN_ORDERED_VARIANT=60)
  ARR_COUNT
};

constexpr uint64_t BLOB_MAGIC = 0x52444E5F424C4F42ull;  // "RDN_BLOB"
constexpr uint32_t BLOB_VERSION = 5;
constexpr uint64_t BLOB_ALIGN = 128;

struct BlobHeader {
  uint64_t magic;
  uint32_t version;
  uint32_t header_bytes;
  uint64_t total_bytes;
  uint64_t offset[ARR_COUNT];  // byte offset from the blob base
  uint64_t count[ARR_COUNT];   // element count
  uint64_t elem_size[ARR_COUNT];
};

}  // namespace rdn
