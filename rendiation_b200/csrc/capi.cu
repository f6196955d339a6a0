// capi.cu — the extern "C" boundary (include/rdn_rt.h): scene object, lazy commit, replication, trace drivers.
//
// Host-side mirror of NaiveSahBVHSystem (shader/ray-tracing/src/backend/wavefront_compute/geometry/naive/mod.rs:
// 495-610): mutations store + invalidate under a write lock, the first use after a mutation rebuilds everything
// (get_or_build_gpu_data, mod.rs:521-536).  Errors are status codes + a thread-local message, never an abort.
#include <cuda_runtime.h>

#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <deque>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <vector>

#include "../../include/rdn_rt.h"
#include "accel.h"
#include "kernels.h"

using namespace rdn;

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
  g_last_error = msg;
  return code;
}

#define RDN_CUDA(expr)                                                                                      \
  do {                                                                                                      \
    cudaError_t e__ = (expr);                                                                               \
    if (e__ != cudaSuccess) return fail(RDN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// rays per pipelined H2D/trace/D2H chunk of the host-buffer path (8 MiB in, 8 MiB out): small enough that the un-overlapped
// head (first H2D) and tail (last D2H) of the pipeline are a few percent of a 2 M-ray frame, large enough for PCIe efficiency
// and to fill the 148 SMs (8 Ki warps per chunk)
constexpr uint64_t HOST_CHUNK_RAYS = 1u << 18;
constexpr uint64_t MAX_LAUNCH_RAYS = 1ull << 31;
constexpr int N_SLOTS = 4;

struct Scratch {
  void *base = nullptr;        // small block: work_counter u64 | tie_count | tie_unresolved | stack_overflow | blocks_done | tie_cursor | tie_total | counters[10]
  uint32_t *tie_queue = nullptr;
  float *tie_best = nullptr;
  uint64_t capacity = 0;       // rays
  mutable uint32_t ordered_launches = 0;  // ordered kernels issued on this set so far (their completion count lives in epoch_done)
  TraceScratch view() const {
    TraceScratch s;
    char *b = static_cast<char *>(base);
    s.work_counter = reinterpret_cast<unsigned long long *>(b);
    s.tie_count = reinterpret_cast<uint32_t *>(b + 8);
    s.tie_unresolved = reinterpret_cast<uint32_t *>(b + 12);
    s.stack_overflow = reinterpret_cast<uint32_t *>(b + 16);
    s.blocks_done = reinterpret_cast<uint32_t *>(b + 20);
    s.tie_cursor = reinterpret_cast<uint32_t *>(b + 24);
    s.tie_total = reinterpret_cast<uint32_t *>(b + 28);
    s.epoch_done = reinterpret_cast<uint32_t *>(b + 112);
    s.gate_timeout = reinterpret_cast<uint32_t *>(b + 116);
    s.counters = reinterpret_cast<unsigned long long *>(b + 32);
    s.tie_queue = tie_queue;
    s.tie_best = tie_best;
    return s;
  }
};
// 32 B of cursors/flags + 10 u64 counters (6 visit counters, 4 debug) + epoch_done at 112
constexpr size_t SCRATCH_BASE_BYTES = 1024;  // (bytes 128.. : tile histogram of the RDN_DEBUG_TIMELINE build, counters[16..])

struct Slot {  // one pipeline lane of the host-buffer path
  cudaStream_t stream = nullptr;
  rdn_ray *d_rays = nullptr;
  rdn_hit *d_hits = nullptr;
  uint64_t capacity = 0;
  Scratch scratch;
  uint32_t *h_flags = nullptr;  // pinned: tie_count | tie_unresolved | stack_overflow read back after the last chunk
  // page-locked staging of one chunk for callers whose buffers are ordinary (pageable) memory, and the event behind the chunk
  rdn_ray *h_rays = nullptr;
  rdn_hit *h_hits = nullptr;
  uint64_t staging_capacity = 0;
  cudaEvent_t chunk_done = nullptr;
  bool chunk_in_flight = false;
};

enum KernelKind : int { KERNEL_ORDERED = 0, KERNEL_TIES = 1, KERNEL_REFERENCE = 2, KERNEL_KIND_COUNT = 3 };
struct TimedLaunch {
  int kind;
  cudaEvent_t begin, end;
};

struct DeviceCtx {
  int device = 0;
  int sm_count = 148;
  void *d_blob = nullptr;
  uint64_t blob_bytes = 0;
  SceneDev dev{};
  Slot slots[N_SLOTS];
  Scratch ext_scratch[2];         // for caller-stream (device-resident) traces; launches alternate so that consecutive ordered
                                  // launches on one stream may overlap their tails (programmatic dependent launch, traverse.cu)
  int ext_next = 0;
  cudaEvent_t ext_done = nullptr; // orders device-resident traces issued on DIFFERENT streams (they share the scratch sets)
  cudaStream_t ext_last_stream = nullptr;
  bool ext_pending = false;
  // the previous device-resident trace, as far as tail overlap is concerned: valid while nothing else of this library went into
  // that stream since, and only if it was ONE ordered launch that resolves its ties itself
  struct { bool valid = false; uintptr_t rays_lo = 0, rays_hi = 0, hits_lo = 0, hits_hi = 0; } ext_prev;
  bool last_enqueue_overlappable = false;  // set by enqueue_trace
  // tile history of grid launches (device-resident traces): how long each 8x4 tile took in the last launch over a grid of this size,
  // and the order built from it for the next one (traverse.cu k_build_tile_order)
  uint32_t *d_tile_cost = nullptr, *d_tile_lists = nullptr;  // lists: two sets of TILE_CLASSES lists (one read by launches, one being built)
  TileMeta *d_tile_meta = nullptr;                           // [2]
  uint64_t tile_cap = 0;
  struct { uint32_t width = 0, height = 0; int current = 0; bool built = false; uint32_t builds = 0, since_build = 0; } tile_hist;
  unsigned long long *d_compact_status = nullptr;
  uint64_t compact_status_cap = 0;
  float *d_ao_payload = nullptr;   // AO accumulation: per-pixel payload, 1.0f between samples
  uint64_t ao_payload_cap = 0;
  uint8_t *d_keep = nullptr;       // bounce step: keep flags + identity indices feeding the compaction
  uint32_t *d_iota = nullptr;
  uint64_t bounce_cap = 0;
  // wavefront executor (rdn_rt_trace_ray): buffers of one launch grid, kept between calls
  struct WaveScratch {
    uint64_t cap = 0;        // rays
    uint32_t buckets = 0;    // task lists (closest-hit + miss shaders)
    uint32_t rows = 0;       // rounds + 1
    rdn_ray *rays[2] = {nullptr, nullptr}, *next = nullptr;
    uint32_t *launch[2] = {nullptr, nullptr}, *iota = nullptr, *idx = nullptr, *task = nullptr, *segments = nullptr;
    rdn_hit *hits = nullptr;
    uint8_t *spawn = nullptr, *keep = nullptr;
    uint64_t *counters = nullptr;          // [0] the grid size, [1] [2] wave sizes (alternating), [3 ..] task-list sizes
    uint64_t *rows_dev = nullptr;          // per round: wave, spawned, task-list sizes
    const uint64_t **ptrs = nullptr;       // device array of counter addresses for one row
    unsigned long long *status = nullptr;
  } wave;
  rdn_anyhit_program *d_anyhit = nullptr;  // device copy of the scene's any-hit programs
  uint32_t n_anyhit = 0;
  bool anyhit_stale = true;
  bool timing = false;                 // rdn_rt_kernel_timing_begin .. _end: events around every traversal kernel
  std::vector<TimedLaunch> timed;
};

}  // namespace

struct rdn_rt_scene {
  std::shared_mutex lock;
  std::mutex launch_lock;          // serialises use of the per-device scratch / slots
  std::mutex wave_lock;            // serialises rdn_rt_trace_ray calls (their wave buffers)
  NaiveSahBvhSource source;
  std::vector<uint32_t> tlas_binding;
  bool dirty = true;               // invalidate(): cpu_data = gpu_data = None
  bool adopted = false;            // blob came from another rank; local sources are not authoritative
  FlatScene flat;                  // host copy (empty when adopted)
  std::vector<uint8_t> host_blob;  // kept only by host-only scenes (n_devices == 0)
  BlobHeader blob_header{};        // layout of the blob the devices hold (valid when flat is)
  uint64_t tlas_only_commits = 0;  // commits that kept every BLAS array and patched the TLAS arrays of the device blobs in place
  std::atomic<uint64_t> kernels_enqueued{0};  // traversal-path kernels put into streams so far (rdn_build_stats)
  void *patch_staging = nullptr;   // page-locked staging buffer of those patches
  uint64_t patch_staging_cap = 0;
  std::vector<uint32_t> h_tlas_binding;  // host copies used to resolve the wide root of a launch
  std::vector<TlasRoot> h_tlas_root;
  std::vector<DeviceCtx> devices;
  // any-hit stage: the programs of the pipeline (rdn_rt_set_any_hit_programs) and the executor's current table (rdn_rt_bind_sbt)
  std::vector<rdn_anyhit_program> anyhit_programs;
  struct rdn_sbt *bound_sbt = nullptr;
};

struct rdn_sbt {
  rdn_rt_scene *scene = nullptr;
  uint32_t ray_stride = 0, max_geometry = 0, max_tlas_offset = 0;
  std::mutex lock;
  std::vector<SbtHitGroup> hit_groups;   // ray_ty_idx + geometry_idx * ray_stride + tlas_offset (sbt.rs:20-38)
  std::vector<uint32_t> miss;            // ray_type_count entries
  uint32_t ray_gen = RDN_SBT_NO_SHADER;
  struct PerDevice {
    SbtHitGroup *d_hit_groups = nullptr;
    uint32_t *d_miss = nullptr;
    bool stale = true;
    uint8_t *d_keep = nullptr;           // grouping scratch, `cap` rays
    uint32_t *d_iota = nullptr, *d_segment = nullptr;
    uint64_t *d_count = nullptr;
    unsigned long long *d_status = nullptr;
    uint64_t cap = 0;
  };
  std::vector<PerDevice> per_device;
  std::vector<int> device_ids;           // CUDA device of each entry (kept here: the table may outlive the scene object)
};

namespace {
int sbt_upload(rdn_sbt *t, int device_index) {
  rdn_sbt::PerDevice &pd = t->per_device[device_index];
  if (!pd.stale) return RDN_OK;
  if (!pd.d_hit_groups) RDN_CUDA(cudaMalloc(&pd.d_hit_groups, std::max<size_t>(t->hit_groups.size(), 1) * sizeof(SbtHitGroup)));
  if (!pd.d_miss) RDN_CUDA(cudaMalloc(&pd.d_miss, std::max<size_t>(t->miss.size(), 1) * sizeof(uint32_t)));
  // (synchronous copies: a table is configured once per pipeline, not per launch)
  RDN_CUDA(cudaMemcpy(pd.d_hit_groups, t->hit_groups.data(), t->hit_groups.size() * sizeof(SbtHitGroup), cudaMemcpyHostToDevice));
  RDN_CUDA(cudaMemcpy(pd.d_miss, t->miss.data(), t->miss.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  pd.stale = false;
  return RDN_OK;
}
void sbt_touch(rdn_sbt *t) { for (auto &pd : t->per_device) pd.stale = true; }
}  // namespace

namespace {

int ensure_scratch(Scratch &s, uint64_t n) {
  bool fresh = false;
  if (!s.base) {
    RDN_CUDA(cudaMalloc(&s.base, SCRATCH_BASE_BYTES));
    RDN_CUDA(cudaMemset(s.base, 0, SCRATCH_BASE_BYTES));
    fresh = true;
  }
  if (s.capacity < n) {
    if (s.tie_queue) cudaFree(s.tie_queue);
    if (s.tie_best) cudaFree(s.tie_best);
    s.tie_queue = nullptr; s.tie_best = nullptr; s.capacity = 0;
    RDN_CUDA(cudaMalloc(&s.tie_queue, std::max<uint64_t>(n, 1) * sizeof(uint32_t)));
    RDN_CUDA(cudaMalloc(&s.tie_best, std::max<uint64_t>(n, 1) * sizeof(float)));
    RDN_CUDA(cudaMemset(s.tie_queue, 0xFF, std::max<uint64_t>(n, 1) * sizeof(uint32_t)));  // RDN_INVALID_ID = slot not published
    s.capacity = n;
    fresh = true;
  }
  // the memsets above run on the legacy default stream, asynchronously to the host; the consumers are non-blocking slot streams and
  // caller streams, which do not synchronise with it: make the initial state visible before anything is launched on it
  if (fresh) RDN_CUDA(cudaDeviceSynchronize());
  return RDN_OK;
}

void free_scratch(Scratch &s) {
  if (s.base) cudaFree(s.base);
  if (s.tie_queue) cudaFree(s.tie_queue);
  if (s.tie_best) cudaFree(s.tie_best);
  s = Scratch{};
}

void bind_blob(DeviceCtx &dc, const BlobHeader &h) {
  const char *b = static_cast<const char *>(dc.d_blob);
  SceneDev &d = dc.dev;
  d.tlas_binding = reinterpret_cast<const uint32_t *>(b + h.offset[ARR_TLAS_BINDING]);
  d.tlas_root = reinterpret_cast<const TlasRoot *>(b + h.offset[ARR_TLAS_ROOT]);
  d.tlas_bvh_forest = reinterpret_cast<const DeviceBVHNode *>(b + h.offset[ARR_TLAS_BVH_FOREST]);
  d.tlas_bounding = reinterpret_cast<const TlasBounding *>(b + h.offset[ARR_TLAS_BOUNDING]);
  d.instances = reinterpret_cast<const InstanceRecord *>(b + h.offset[ARR_INSTANCES]);
  d.blas_meta = reinterpret_cast<const BlasMeta *>(b + h.offset[ARR_BLAS_META]);
  d.geometry_meta = reinterpret_cast<const GeometryMeta *>(b + h.offset[ARR_GEOMETRY_META]);
  d.tri_bvh_forest = reinterpret_cast<const DeviceBVHNode *>(b + h.offset[ARR_TRI_BVH_FOREST]);
  d.triangles = reinterpret_cast<const TriRecord *>(b + h.offset[ARR_TRIANGLES]);
  d.slot_info = reinterpret_cast<const SlotInfo *>(b + h.offset[ARR_SLOT_INFO]);
  d.wide_nodes = reinterpret_cast<const WideNode *>(b + h.offset[ARR_WIDE_NODES]);
  d.prim_to_slot = reinterpret_cast<const uint32_t *>(b + h.offset[ARR_PRIM_TO_SLOT]);
  d.irregular_instances = reinterpret_cast<const uint32_t *>(b + h.offset[ARR_IRREGULAR_INSTANCES]);
  d.irregular_leaf_boxes = reinterpret_cast<const LeafBox *>(b + h.offset[ARR_IRREGULAR_LEAF_BOXES]);
  d.wide4_nodes = reinterpret_cast<const Wide4Node *>(b + h.offset[ARR_WIDE4_NODES]);
  d.n_tlas_binding = static_cast<uint32_t>(h.count[ARR_TLAS_BINDING]);
  d.n_tlas_root = static_cast<uint32_t>(h.count[ARR_TLAS_ROOT]);
  d.n_blas_meta = static_cast<uint32_t>(h.count[ARR_BLAS_META]);
  d.n_instances = static_cast<uint32_t>(h.count[ARR_INSTANCES]);
}

bool header_ok(const BlobHeader &h, uint64_t bytes) {
  if (h.magic != BLOB_MAGIC || h.version != BLOB_VERSION || h.header_bytes != sizeof(BlobHeader) || h.total_bytes != bytes) return false;
  for (int i = 0; i < ARR_COUNT; ++i)
    if (h.offset[i] % 16 != 0 || h.offset[i] + std::max<uint64_t>(h.count[i], 1) * h.elem_size[i] > bytes) return false;
  return true;
}

void parallel_copy(void *dst, const void *src, uint64_t bytes);  // (below)

// build -> flatten -> upload to device 0 -> replicate to the other devices (peer copy: NVLink when available)
int commit_locked(rdn_rt_scene *s) {
  if (!s->dirty) return RDN_OK;
  if (s->adopted && ((!s->devices.empty() && s->devices[0].d_blob) || (s->devices.empty() && !s->host_blob.empty()))) { s->dirty = false; return RDN_OK; }
  std::string err;
  FlatScene flat;
  const char *device_build_env = getenv("RDN_COMMIT_DEVICE_BUILD");  // (read at every commit: a test flips it between scenes)
  const bool device_build = device_build_env && atoi(device_build_env) != 0;
  s->source.build_device = (device_build && !s->devices.empty()) ? s->devices[0].device : -1;
  if (const char *e = getenv("RDN_COMMIT_DEVICE_BUILD_MIN")) s->source.device_build_min = strtoull(e, nullptr, 10);
  // A commit after TLAS-only changes (rdn_rt_tlas_update, TLAS create / delete, bind) keeps the BLAS arrays of the previous flattened
  // scene and redoes build_tlas alone; when no array changed its length the device blobs are patched in place — a few hundred
  // kilobytes instead of the whole scene.  (The reference rebuilds and re-uploads everything on any change, naive/mod.rs:121,546-549.)
  bool reused = false;
  const bool have_previous = !s->adopted && !s->flat.blas_meta.empty();
  const int rc = s->source.build(s->tlas_binding, flat, err, have_previous ? &s->flat : nullptr, &reused);
  if (rc != RDN_OK) { s->flat = FlatScene{}; return fail(rc, err); }
  const auto t_upload = std::chrono::steady_clock::now();
  if (reused && !s->devices.empty()) {
    const BlobHeader &h = s->blob_header;
    const uint64_t wide0 = s->source.tlas_part_wide_start(), wide4_0 = s->source.tlas_part_wide4_start();
    auto same = [&](int id, uint64_t count) { return h.count[id] == count; };
    const bool same_layout = h.magic == BLOB_MAGIC && same(ARR_TLAS_BINDING, flat.tlas_binding.size()) && same(ARR_TLAS_ROOT, flat.tlas_root.size()) &&
                             same(ARR_TLAS_BVH_FOREST, flat.tlas_bvh_forest.size()) && same(ARR_TLAS_BOUNDING, flat.tlas_bounding.size()) &&
                             same(ARR_INSTANCES, flat.instances.size()) && same(ARR_WIDE_NODES, flat.wide_nodes.size()) &&
                             same(ARR_IRREGULAR_INSTANCES, flat.irregular_instances.size()) && same(ARR_WIDE4_NODES, flat.wide4_nodes.size());
    if (same_layout) {
      // the changed ranges are gathered into one page-locked staging buffer (pageable sources would be staged by the driver one
      // copy at a time, each with its own synchronisation) and go out as asynchronous copies behind one another
      struct Range { int id; const char *src; uint64_t first_byte, bytes, staged_at; };
      std::vector<Range> ranges;
      uint64_t staged_bytes = 0;
      auto add = [&](int id, const void *src, uint64_t first, uint64_t count, uint64_t elem) {
        if (count == 0) return;
        ranges.push_back(Range{id, static_cast<const char *>(src) + first * elem, first * elem, count * elem, staged_bytes});
        staged_bytes += (count * elem + 255) / 256 * 256;
      };
      add(ARR_TLAS_BINDING, flat.tlas_binding.data(), 0, flat.tlas_binding.size(), 4);
      add(ARR_TLAS_ROOT, flat.tlas_root.data(), 0, flat.tlas_root.size(), sizeof(TlasRoot));
      add(ARR_TLAS_BVH_FOREST, flat.tlas_bvh_forest.data(), 0, flat.tlas_bvh_forest.size(), sizeof(DeviceBVHNode));
      add(ARR_TLAS_BOUNDING, flat.tlas_bounding.data(), 0, flat.tlas_bounding.size(), sizeof(TlasBounding));
      add(ARR_INSTANCES, flat.instances.data(), 0, flat.instances.size(), sizeof(InstanceRecord));
      add(ARR_WIDE_NODES, flat.wide_nodes.data(), wide0, flat.wide_nodes.size() - wide0, sizeof(WideNode));
      add(ARR_IRREGULAR_INSTANCES, flat.irregular_instances.data(), 0, flat.irregular_instances.size(), 4);
      add(ARR_WIDE4_NODES, flat.wide4_nodes.data(), wide4_0, flat.wide4_nodes.size() - wide4_0, sizeof(Wide4Node));
      RDN_CUDA(cudaSetDevice(s->devices[0].device));
      if (s->patch_staging_cap < staged_bytes) {
        if (s->patch_staging) cudaFreeHost(s->patch_staging);
        s->patch_staging = nullptr; s->patch_staging_cap = 0;
        RDN_CUDA(cudaHostAlloc(&s->patch_staging, staged_bytes + staged_bytes / 4, cudaHostAllocPortable));
        s->patch_staging_cap = staged_bytes + staged_bytes / 4;
      }
      char *staging = static_cast<char *>(s->patch_staging);
      run_parallel(static_cast<unsigned>(ranges.size()), [&](unsigned k) { std::memcpy(staging + ranges[k].staged_at, ranges[k].src, ranges[k].bytes); });
      for (DeviceCtx &dc : s->devices) {
        RDN_CUDA(cudaSetDevice(dc.device));
        RDN_CUDA(cudaDeviceSynchronize());  // nothing may still be walking the arrays that change
        char *base = static_cast<char *>(dc.d_blob);
        for (const Range &r : ranges)
          RDN_CUDA(cudaMemcpyAsync(base + h.offset[r.id] + r.first_byte, staging + r.staged_at, r.bytes, cudaMemcpyHostToDevice, cudaStreamPerThread));
      }
      for (DeviceCtx &dc : s->devices) {
        RDN_CUDA(cudaSetDevice(dc.device));
        RDN_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
      }
      flat.stats.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_upload).count();
      s->h_tlas_binding = flat.tlas_binding;
      s->h_tlas_root = flat.tlas_root;
      s->flat = std::move(flat);
      s->tlas_only_commits++;
      s->dirty = false;
      return RDN_OK;
    }
  }
  // The image is never built on the host for a scene with devices: every array goes to its place in the device blob through two
  // page-locked staging buffers, filled by the worker pool (streaming stores) while the previous piece is on the link — a 176 MB
  // scene in a few milliseconds instead of 95 (one pageable vector, one synchronous copy).
  std::vector<FlatScene::ArrayRef> arrays;
  const BlobHeader h = flat.layout(arrays);
  std::vector<uint8_t> blob;
  if (s->devices.empty()) blob = flat.serialize();
  for (size_t i = 0; i < s->devices.size(); ++i) {
    DeviceCtx &dc = s->devices[i];
    RDN_CUDA(cudaSetDevice(dc.device));
    if (dc.d_blob) { cudaFree(dc.d_blob); dc.d_blob = nullptr; }
    RDN_CUDA(cudaMalloc(&dc.d_blob, h.total_bytes));
    dc.blob_bytes = h.total_bytes;
    if (i == 0) {
      constexpr uint64_t PIECE = 16ull << 20;
      if (s->patch_staging_cap < 2 * PIECE) {
        if (s->patch_staging) cudaFreeHost(s->patch_staging);
        s->patch_staging = nullptr; s->patch_staging_cap = 0;
        RDN_CUDA(cudaHostAlloc(&s->patch_staging, 2 * PIECE, cudaHostAllocPortable));
        s->patch_staging_cap = 2 * PIECE;
      }
      struct Events {  // destroyed on every return path
        cudaEvent_t e[2] = {nullptr, nullptr};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
      } ev;
      for (cudaEvent_t &x : ev.e) RDN_CUDA(cudaEventCreateWithFlags(&x, cudaEventDisableTiming));
      char *base = static_cast<char *>(dc.d_blob);
      RDN_CUDA(cudaMemsetAsync(base, 0, h.total_bytes, cudaStreamPerThread));  // (padding and the one zeroed element of empty arrays)
      RDN_CUDA(cudaMemcpyAsync(base, &h, sizeof(h), cudaMemcpyHostToDevice, cudaStreamPerThread));  // (pageable, small: staged by the driver before it returns)
      uint64_t piece = 0;
      for (const FlatScene::ArrayRef &a : arrays) {
        for (uint64_t done = 0; done < a.bytes; done += PIECE, ++piece) {
          const uint64_t m = std::min(PIECE, a.bytes - done);
          char *stage = static_cast<char *>(s->patch_staging) + (piece & 1) * PIECE;
          if (piece >= 2) RDN_CUDA(cudaEventSynchronize(ev.e[piece & 1]));  // the copy that last read this half has left it
          parallel_copy(stage, static_cast<const char *>(a.data) + done, m);
          RDN_CUDA(cudaMemcpyAsync(base + h.offset[a.id] + done, stage, m, cudaMemcpyHostToDevice, cudaStreamPerThread));
          RDN_CUDA(cudaEventRecord(ev.e[piece & 1], cudaStreamPerThread));
        }
      }
      RDN_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    } else {
      RDN_CUDA(cudaMemcpyPeer(dc.d_blob, dc.device, s->devices[0].d_blob, s->devices[0].device, h.total_bytes));
    }
    bind_blob(dc, h);
  }
  flat.stats.upload_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_upload).count();
  s->h_tlas_binding = flat.tlas_binding;
  s->h_tlas_root = flat.tlas_root;
  s->flat = std::move(flat);
  s->blob_header = h;
  if (s->devices.empty()) s->host_blob = std::move(blob);
  s->dirty = false;
  return RDN_OK;
}

int ensure_committed(rdn_rt_scene *s) {
  {
    std::shared_lock<std::shared_mutex> rd(s->lock);
    if (!s->dirty) return RDN_OK;
  }
  std::unique_lock<std::shared_mutex> wr(s->lock);
  if (s->dirty) warm_worker_pool();
  return commit_locked(s);
}

TlasRoot resolve_tlas(const rdn_rt_scene *s, uint32_t tlas_idx) {
  const TlasRoot none{INVALID_NEXT, REF_EMPTY, 0, 0};
  if (tlas_idx >= s->h_tlas_binding.size()) return none;
  const uint32_t handle = s->h_tlas_binding[tlas_idx];
  if (handle >= s->h_tlas_root.size()) return none;
  return s->h_tlas_root[handle];
}

// CUDA events around one kernel launch while kernel timing is enabled (recorded on the launching stream)
struct ScopedKernelTimer {
  DeviceCtx &dc;
  cudaStream_t stream;
  TimedLaunch t{};
  bool on;
  ScopedKernelTimer(DeviceCtx &d, int kind, cudaStream_t s) : dc(d), stream(s), on(d.timing) {
    if (!on) return;
    t.kind = kind;
    if (cudaEventCreate(&t.begin) != cudaSuccess || cudaEventCreate(&t.end) != cudaSuccess) { on = false; return; }
    cudaEventRecord(t.begin, stream);
  }
  ~ScopedKernelTimer() {
    if (!on) return;
    cudaEventRecord(t.end, stream);
    dc.timed.push_back(t);
  }
};

void free_wave_scratch(DeviceCtx &dc) {
  auto &w = dc.wave;
  void *all[] = {w.rays[0], w.rays[1], w.next, w.launch[0], w.launch[1], w.iota, w.idx, w.task, w.segments, w.hits, w.spawn, w.keep, w.counters,
                 w.rows_dev, const_cast<uint64_t **>(w.ptrs), w.status};
  for (void *p : all) if (p) cudaFree(p);
  w = DeviceCtx::WaveScratch{};
}

int ensure_wave_scratch(DeviceCtx &dc, uint64_t n, uint32_t buckets, uint32_t rows) {
  auto &w = dc.wave;
  if (w.cap >= n && w.buckets >= buckets && w.rows >= rows) return RDN_OK;
  const uint64_t cap = std::max<uint64_t>(std::max(w.cap, n), 1);
  const uint32_t nb = std::max(std::max(w.buckets, buckets), 1u), nr = std::max(std::max(w.rows, rows), 1u);
  free_wave_scratch(dc);
  for (int k = 0; k < 2; ++k) {
    RDN_CUDA(cudaMalloc(&w.rays[k], cap * sizeof(rdn_ray)));
    RDN_CUDA(cudaMalloc(&w.launch[k], cap * sizeof(uint32_t)));
  }
  RDN_CUDA(cudaMalloc(&w.next, cap * sizeof(rdn_ray)));
  RDN_CUDA(cudaMalloc(&w.hits, cap * sizeof(rdn_hit)));
  RDN_CUDA(cudaMalloc(&w.iota, cap * sizeof(uint32_t)));
  RDN_CUDA(cudaMalloc(&w.idx, cap * sizeof(uint32_t)));
  RDN_CUDA(cudaMalloc(&w.task, cap * sizeof(uint32_t)));
  RDN_CUDA(cudaMalloc(&w.segments, cap * nb * sizeof(uint32_t)));
  RDN_CUDA(cudaMalloc(&w.spawn, cap));
  RDN_CUDA(cudaMalloc(&w.keep, cap));
  RDN_CUDA(cudaMalloc(&w.counters, (3 + nb) * sizeof(uint64_t)));
  RDN_CUDA(cudaMalloc(&w.rows_dev, static_cast<size_t>(nr) * (2 + nb) * sizeof(uint64_t)));
  RDN_CUDA(cudaMalloc(&w.ptrs, (2 + nb) * sizeof(uint64_t *)));
  RDN_CUDA(cudaMalloc(&w.status, compact_status_words(cap) * sizeof(unsigned long long)));
  w.cap = cap; w.buckets = nb; w.rows = nr;
  return RDN_OK;
}

// The two safety-net words of a scratch set — traversal stack overflow, a launch that gave up waiting at its gate — read and
// cleared (the device is idle on this set when this is called), so that one bad launch is reported once and not for ever after.
int take_error_flags(const Scratch &scratch, uint32_t *out_flags = nullptr) {
  uint32_t overflow = 0, gate = 0;
  RDN_CUDA(cudaMemcpy(&overflow, static_cast<char *>(scratch.base) + 16, 4, cudaMemcpyDeviceToHost));
  RDN_CUDA(cudaMemcpy(&gate, static_cast<char *>(scratch.base) + 116, 4, cudaMemcpyDeviceToHost));
  if (overflow) RDN_CUDA(cudaMemset(static_cast<char *>(scratch.base) + 16, 0, 4));
  if (gate) RDN_CUDA(cudaMemset(static_cast<char *>(scratch.base) + 116, 0, 4));
  if (overflow || gate) RDN_CUDA(cudaDeviceSynchronize());
  if (out_flags) *out_flags |= (overflow ? RDN_ERROR_FLAG_STACK_OVERFLOW : 0u) | (gate ? RDN_ERROR_FLAG_GATE_TIMEOUT : 0u);
  if (overflow) return fail(RDN_ERR_CAPACITY, "traversal stack overflow (hit records of that launch are not reliable)");
  if (gate) return fail(RDN_ERR_CUDA, "an ordered launch gave up waiting for the previous launch on its scratch set");
  return RDN_OK;
}

// The any-hit stage of a launch: device copies of the scene's programs and, for RDN_ANYHIT_FROM_SBT, of the bound table's hit groups,
// wired into the SceneDev the kernels get; *can_end_search = the stage may stop a traversal (the answer is then order dependent).
int prepare_any_hit(rdn_rt_scene *s, DeviceCtx &dc, const rdn_launch &launch, SceneDev &dev, bool &can_end_search) {
  can_end_search = false;
  if (launch.any_hit == RDN_ANYHIT_NONE) return RDN_OK;
  if (dc.anyhit_stale) {
    if (dc.d_anyhit) { cudaFree(dc.d_anyhit); dc.d_anyhit = nullptr; }
    dc.n_anyhit = static_cast<uint32_t>(s->anyhit_programs.size());
    if (dc.n_anyhit) {
      RDN_CUDA(cudaMalloc(&dc.d_anyhit, dc.n_anyhit * sizeof(rdn_anyhit_program)));
      RDN_CUDA(cudaMemcpy(dc.d_anyhit, s->anyhit_programs.data(), dc.n_anyhit * sizeof(rdn_anyhit_program), cudaMemcpyHostToDevice));
    }
    dc.anyhit_stale = false;
  }
  dev.anyhit_programs = dc.d_anyhit; dev.n_anyhit_programs = dc.n_anyhit;
  if (launch.any_hit == RDN_ANYHIT_FROM_SBT) {
    rdn_sbt *t = s->bound_sbt;
    if (!t) return fail(RDN_ERR_INVALID_ARGUMENT, "RDN_ANYHIT_FROM_SBT without a bound table (rdn_rt_bind_sbt)");
    const int di = static_cast<int>(&dc - s->devices.data());
    std::lock_guard<std::mutex> tl(t->lock);
    const int rc = sbt_upload(t, di);
    if (rc != RDN_OK) return rc;
    dev.sbt_hit_groups = t->per_device[di].d_hit_groups;
    dev.n_sbt_hit_groups = static_cast<uint32_t>(t->hit_groups.size());
    can_end_search = any_hit_can_end_search(launch, s->anyhit_programs.data(), dc.n_anyhit, t->hit_groups.data(), dev.n_sbt_hit_groups);
  } else {
    can_end_search = any_hit_can_end_search(launch, s->anyhit_programs.data(), dc.n_anyhit, nullptr, 0);
  }
  return RDN_OK;
}

// enqueue the kernels of one trace on `stream`; returns the number of kernels launched
int enqueue_trace(rdn_rt_scene *s, DeviceCtx &dc, const Scratch &scratch, const rdn_launch &launch, const rdn_ray *d_rays, uint64_t n,
                  rdn_hit *d_hits, int mode, cudaStream_t stream, uint32_t *launches, bool count_ties = false, bool allow_overlap = false,
                  const unsigned long long *d_n = nullptr, bool tile_history = false) {
  const TraceScratch ts = scratch.view();
  // work_counter is zero here: zeroed at allocation and re-armed by the last CTA of every ordered launch.  tie_count and
  // the two error flags accumulate until read (the stats path clears tie_count first).
  const int tie_mode = ordered_tie_mode();
  const bool end_search = (launch.ray_flags & RDN_RAY_FLAG_ACCEPT_FIRST_HIT_AND_END_SEARCH) != 0;
  const TlasRoot tlas = resolve_tlas(s, launch.tlas_idx);
  // launches whose queue is walked by k_resolve_ties (the separate-kernel variant, and every TLAS that lists irregular instances)
  // start from an empty queue and leave one (in-kernel drains re-arm it themselves)
  const bool separate_resolve = tie_mode == 0 || (tlas.irregular_count != 0 && tlas.irregular_count != IRREGULAR_ROUTE_ALL);
  if (separate_resolve) RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 8, 0, 4, stream));
  if (count_ties) RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 28, 0, 4, stream));  // tie_total: re-walked rays of this launch
  if (count_ties) RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 12, 0, 4, stream));
  // a TLAS made mostly of irregular instances (accel.cpp "regularity") has a traversal-order dependent answer: reference order
  // ACCEPT_FIRST_HIT_AND_END_SEARCH (shadow / AO rays): the answer is the first candidate the reference's pre-order walk accepts,
  // order dependent by definition.  (Measured alternative, profiles/kbench_r1_anyhit.log: the ordered kernel stopping at its first
  // candidate to separate misses, the reference-order walk only for the occluded rays — slower than walking everything in
  // reference order, whose early exit is what makes these rays cheap.)
  dc.last_enqueue_overlappable = false;
  // any-hit stage: device copies of the programs and of the bound table's hit groups; a stage that can stop a traversal
  // (END_SEARCH) makes the answer order dependent, and so does one on a TLAS with irregular instances: reference order
  SceneDev dev = dc.dev;
  bool any_hit_reference_order = false;
  {
    const int rc = prepare_any_hit(s, dc, launch, dev, any_hit_reference_order);
    if (rc != RDN_OK) return rc;
    if (launch.any_hit != RDN_ANYHIT_NONE && tlas.irregular_count != 0) any_hit_reference_order = true;
  }
  if (mode == RDN_TRACE_REFERENCE_ORDER || end_search || tlas.irregular_count == IRREGULAR_ROUTE_ALL || any_hit_reference_order) {
    ScopedKernelTimer tm(dc, KERNEL_REFERENCE, stream);
    launch_trace_reference(dev, launch, d_rays, n, d_hits, ts, false, dc.sm_count, stream, d_n);
  } else {
    bool ties_done = true;
    {
      ScopedKernelTimer tm(dc, KERNEL_ORDERED, stream);
      // (a launch that failed never advances the set's epoch: the host-side count moves only when the kernel is in the stream,
      // or the next launch on the set would wait at its gate for an epoch that never comes)
      // Tile history (device-resident grid launches): the launch takes its tiles by how long they took in an earlier launch over a
      // grid of the same size — longest first, see k_build_tile_lists — and notes its own pass durations for the next.  Frames of
      // an animation, the samples of a pixel, the bench's repeated frame: what was long stays long.  A launch that may overlap its
      // predecessor's tail only reads the lists (a kernel behind every launch would separate the launches again).  The first two
      // launches over a grid run in grid order: the classes are measured by the mean of the launch before.  RDN_TILE_HISTORY=0
      // switches all of it off.
      static const bool history_enabled = []() { const char *e = getenv("RDN_TILE_HISTORY"); return !e || atoi(e) != 0; }();
      constexpr int TILE_LISTS = 8;  // (= TILE_CLASSES of traverse.cu: sizeof(TileMeta::count))
      static_assert(sizeof(TileMeta::count) / sizeof(uint32_t) == TILE_LISTS, "one list per class");
      TileHistory hist{};
      const bool overlap = allow_overlap && !dc.timing && !count_ties;
      // (launches of 64 Ki to 32 Mi rays: below, a launch is a handful of tiles per warp; above, its tail no longer matters and the
      // lists — 64 B per tile — would run into hundreds of megabytes)
      const bool with_history = tile_history && history_enabled && !d_n && launch.grid_width != 0 && n % launch.grid_width == 0 && n >= (1u << 16) &&
                                n <= (1u << 25);
      if (with_history) {
        const uint32_t w = launch.grid_width, h = static_cast<uint32_t>(n / launch.grid_width);
        const uint32_t n_tiles = ((w + 7u) / 8u) * ((h + 3u) / 4u);
        if (dc.tile_cap < n_tiles) {
          if (dc.d_tile_cost) cudaFree(dc.d_tile_cost);
          if (dc.d_tile_lists) cudaFree(dc.d_tile_lists);
          dc.d_tile_cost = dc.d_tile_lists = nullptr; dc.tile_cap = 0; dc.tile_hist.width = 0;
          RDN_CUDA(cudaMalloc(&dc.d_tile_cost, n_tiles * sizeof(uint32_t)));
          RDN_CUDA(cudaMalloc(&dc.d_tile_lists, 2ull * TILE_LISTS * n_tiles * sizeof(uint32_t)));
          if (!dc.d_tile_meta) RDN_CUDA(cudaMalloc(&dc.d_tile_meta, 2 * sizeof(TileMeta)));
          dc.tile_cap = n_tiles;
        }
        if (dc.tile_hist.width != w || dc.tile_hist.height != h) {  // another grid: start over
          RDN_CUDA(cudaMemsetAsync(dc.d_tile_cost, 0, n_tiles * sizeof(uint32_t), stream));
          RDN_CUDA(cudaMemsetAsync(dc.d_tile_meta, 0, 2 * sizeof(TileMeta), stream));
          dc.tile_hist.width = w; dc.tile_hist.height = h; dc.tile_hist.current = 0; dc.tile_hist.built = false;
          dc.tile_hist.builds = 0; dc.tile_hist.since_build = 0;
        }
        const int cur = dc.tile_hist.current;
        hist.n_tiles = n_tiles;
        if (dc.tile_hist.built) { hist.lists = dc.d_tile_lists + static_cast<uint64_t>(cur) * TILE_LISTS * n_tiles; hist.meta = dc.d_tile_meta + cur; }
        // (the lists are rebuilt behind the first three launches over a grid — the classes need the mean of the launch before — and
        // behind every fourth from then on: what is long changes slowly, and the kernel behind the launch is 2 % of a 2 M-ray launch)
        const bool due = dc.tile_hist.builds < 3 || dc.tile_hist.since_build >= 3;
        if (!overlap && due) { hist.cost = dc.d_tile_cost; hist.meta_clear = dc.d_tile_meta + (cur ^ 1); }
      }
      RDN_CUDA(launch_trace_ordered(dev, launch, tlas, d_rays, n, d_hits, ts, dc.sm_count, stream, overlap, scratch.ordered_launches, &ties_done, d_n,
                                    with_history ? &hist : nullptr));
      if (n) scratch.ordered_launches++;
      dc.last_enqueue_overlappable = ties_done && n != 0;
      if (with_history && hist.used && hist.cost) {
        // the lists for the launches to come, into the set nobody reads (its description was zeroed by the launch's last CTA)
        const int cur = dc.tile_hist.current;
        launch_build_tile_lists(dc.d_tile_cost, dc.d_tile_lists + static_cast<uint64_t>(cur ^ 1) * TILE_LISTS * hist.n_tiles, dc.d_tile_meta + (cur ^ 1),
                                dc.d_tile_meta + cur, hist.n_tiles, stream);
        dc.tile_hist.current = cur ^ 1;
        dc.tile_hist.built = true;
        dc.tile_hist.builds++;
        dc.tile_hist.since_build = 0;
        s->kernels_enqueued++;
        if (launches) *launches += 1;
      } else if (with_history && hist.used) {
        dc.tile_hist.since_build++;
      } else if (with_history && !hist.used) {
        dc.tile_hist.built = false;  // (another instantiation took the launch — any-hit stage, an A/B variant —: the lists are stale)
      }
    }
    if (!ties_done) {
      ScopedKernelTimer tm(dc, KERNEL_TIES, stream);
      launch_resolve_ties(dev, launch, d_rays, d_hits, ts, dc.sm_count, stream);
      s->kernels_enqueued++;
      if (launches) *launches += 1;
      RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 8, 0, 4, stream));  // the next launch on this set may drain in-kernel
    }
  }
  if (n) s->kernels_enqueued++;
  if (launches) *launches += 1;
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int ensure_slot_staging(Slot &slot, uint64_t rays) {
  if (!slot.chunk_done) RDN_CUDA(cudaEventCreateWithFlags(&slot.chunk_done, cudaEventDisableTiming));
  if (slot.staging_capacity < rays) {
    if (slot.h_rays) cudaFreeHost(slot.h_rays);
    if (slot.h_hits) cudaFreeHost(slot.h_hits);
    slot.h_rays = nullptr; slot.h_hits = nullptr; slot.staging_capacity = 0;
    RDN_CUDA(cudaHostAlloc(&slot.h_rays, rays * sizeof(rdn_ray), cudaHostAllocPortable));
    RDN_CUDA(cudaHostAlloc(&slot.h_hits, rays * sizeof(rdn_hit), cudaHostAllocPortable));
    slot.staging_capacity = rays;
  }
  return RDN_OK;
}

// is this host pointer ordinary pageable memory (neither allocated nor registered through CUDA)?
bool is_pageable(const void *p) {
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return attr.type == cudaMemoryTypeUnregistered;
}

// One thread's share of a staging copy.  Non-temporal stores where the ISA has them: the destination is not read again by this
// core (the DMA engine reads the staging buffer, the caller's result array is far larger than the caches), and an ordinary store
// would first fetch every destination line it is about to overwrite — a third of the copy's memory traffic.
void copy_streaming(char *dst, const char *src, uint64_t bytes) {
#if defined(__x86_64__) || defined(_M_X64)
  const uint64_t head = std::min<uint64_t>(bytes, (16u - (reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u);
  if (head) { std::memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
  const uint64_t lines = bytes / 64;
  for (uint64_t i = 0; i < lines; ++i) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src)), b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32)), d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst), a);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), d);
    src += 64; dst += 64;
  }
  _mm_sfence();
  if (bytes % 64) std::memcpy(dst, src, bytes % 64);
#else
  std::memcpy(dst, src, bytes);
#endif
}

// a staging copy by several threads of the worker pool (one thread moves 5-10 GB/s, a PCIe 5 x16 link wants 50)
void parallel_copy(void *dst, const void *src, uint64_t bytes) {
  static const unsigned max_threads = []() {
    const char *e = getenv("RDN_STAGE_THREADS");
    const int v = e ? atoi(e) : 0;
    return static_cast<unsigned>(v > 0 ? v : 16);
  }();
  static const bool streaming = []() { const char *e = getenv("RDN_STAGE_STREAMING"); return !e || atoi(e) != 0; }();
  constexpr uint64_t PIECE = 256u << 10;
  const unsigned pieces = static_cast<unsigned>(std::min<uint64_t>(std::min<unsigned>(max_threads, build_thread_count()), (bytes + PIECE - 1) / PIECE));
  auto one = [&](uint64_t b, uint64_t e) {
    if (streaming) copy_streaming(static_cast<char *>(dst) + b, static_cast<const char *>(src) + b, e - b);
    else std::memcpy(static_cast<char *>(dst) + b, static_cast<const char *>(src) + b, e - b);
  };
  if (pieces <= 1) { one(0, bytes); return; }
  const uint64_t each = ((bytes + pieces - 1) / pieces + 63) / 64 * 64;
  run_parallel(pieces, [&](unsigned k) {
    const uint64_t b = std::min<uint64_t>(bytes, k * each), e = std::min<uint64_t>(bytes, b + each);
    if (b < e) one(b, e);
  });
}

int ensure_slot(Slot &slot, uint64_t rays) {
  if (!slot.stream) RDN_CUDA(cudaStreamCreateWithFlags(&slot.stream, cudaStreamNonBlocking));
  if (slot.capacity < rays) {
    if (slot.d_rays) cudaFree(slot.d_rays);
    if (slot.d_hits) cudaFree(slot.d_hits);
    slot.d_rays = nullptr; slot.d_hits = nullptr; slot.capacity = 0;
    RDN_CUDA(cudaMalloc(&slot.d_rays, rays * sizeof(rdn_ray)));
    RDN_CUDA(cudaMalloc(&slot.d_hits, rays * sizeof(rdn_hit)));
    slot.capacity = rays;
  }
  if (!slot.h_flags) RDN_CUDA(cudaMallocHost(&slot.h_flags, 4 * sizeof(uint32_t)));
  return ensure_scratch(slot.scratch, rays);
}

}  // namespace

// ================================================================================================ C ABI
extern "C" {

const char *rdn_rt_last_error(void) { return g_last_error.c_str(); }
const char *rdn_rt_version(void) { return "rendiation_b200 0.1 (sm_100a)"; }

int rdn_rt_scene_create(int n_devices, const int *device_ids, rdn_rt_scene **out) {
  if (!out || n_devices < 0) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_create: need out and n_devices >= 0");
  int available = 0;
  if (n_devices > 0) RDN_CUDA(cudaGetDeviceCount(&available));  // n_devices == 0: host-only scene (build/flatten inspection, no tracing)
  auto *s = new rdn_rt_scene();
  s->devices.resize(n_devices);
  for (int i = 0; i < n_devices; ++i) {
    const int dev = device_ids ? device_ids[i] : i;
    if (dev < 0 || dev >= available) { delete s; return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_create: no such CUDA device"); }
    s->devices[i].device = dev;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) s->devices[i].sm_count = prop.multiProcessorCount;
  }
  // let device 0 push the blob to its peers over NVLink
  for (int i = 1; i < n_devices; ++i) {
    int can = 0;
    if (cudaDeviceCanAccessPeer(&can, s->devices[i].device, s->devices[0].device) == cudaSuccess && can) {
      cudaSetDevice(s->devices[i].device);
      cudaDeviceEnablePeerAccess(s->devices[0].device, 0);
      cudaGetLastError();
    }
  }
  *out = s;
  return RDN_OK;
}

void rdn_rt_scene_destroy(rdn_rt_scene *s) {
  if (!s) return;
  if (s->bound_sbt) s->bound_sbt->scene = nullptr;  // (a table may outlive its scene; it must not point back at it then)
  for (DeviceCtx &dc : s->devices) {
    cudaSetDevice(dc.device);
    cudaDeviceSynchronize();
    if (dc.d_blob) cudaFree(dc.d_blob);
    for (Slot &slot : dc.slots) {
      if (slot.d_rays) cudaFree(slot.d_rays);
      if (slot.d_hits) cudaFree(slot.d_hits);
      free_scratch(slot.scratch);
      if (slot.h_flags) cudaFreeHost(slot.h_flags);
      if (slot.h_rays) cudaFreeHost(slot.h_rays);
      if (slot.h_hits) cudaFreeHost(slot.h_hits);
      if (slot.chunk_done) cudaEventDestroy(slot.chunk_done);
      if (slot.stream) cudaStreamDestroy(slot.stream);
    }
    free_scratch(dc.ext_scratch[0]);
    free_scratch(dc.ext_scratch[1]);
    for (TimedLaunch &t : dc.timed) { cudaEventDestroy(t.begin); cudaEventDestroy(t.end); }
    if (dc.ext_done) cudaEventDestroy(dc.ext_done);
    if (dc.d_compact_status) cudaFree(dc.d_compact_status);
    if (dc.d_tile_cost) cudaFree(dc.d_tile_cost);
    if (dc.d_tile_lists) cudaFree(dc.d_tile_lists);
    if (dc.d_tile_meta) cudaFree(dc.d_tile_meta);
    if (dc.d_anyhit) cudaFree(dc.d_anyhit);
    free_wave_scratch(dc);
    if (dc.d_ao_payload) cudaFree(dc.d_ao_payload);
    if (dc.d_keep) cudaFree(dc.d_keep);
    if (dc.d_iota) cudaFree(dc.d_iota);
  }
  if (s->patch_staging) cudaFreeHost(s->patch_staging);
  delete s;
}

int rdn_rt_blas_create(rdn_rt_scene *s, const rdn_blas_geometry *geoms, uint32_t n, uint32_t *out_handle) {
  if (!s || !out_handle || (n && !geoms)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_blas_create: null argument");
  std::vector<GeometrySource> src(n);
  for (uint32_t g = 0; g < n; ++g) {
    const rdn_blas_geometry &in = geoms[g];
    src[g].flags = in.flags;
    src[g].is_aabbs = in.kind != 0;
    if (src[g].is_aabbs) continue;
    if (in.n_positions && !in.positions) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_blas_create: null positions");
    src[g].positions.resize(in.n_positions);
    if (in.n_positions) std::memcpy(src[g].positions.data(), in.positions, in.n_positions * sizeof(Vec3));
    src[g].has_indices = in.indices != nullptr;
    if (in.indices) src[g].indices.assign(in.indices, in.indices + in.n_indices);
  }
  std::unique_lock<std::shared_mutex> wr(s->lock);
  s->dirty = true; s->adopted = false;
  *out_handle = s->source.create_blas(std::move(src));
  return RDN_OK;
}

int rdn_rt_blas_destroy(rdn_rt_scene *s, uint32_t handle) {
  if (!s) return fail(RDN_ERR_INVALID_ARGUMENT, "null scene");
  std::unique_lock<std::shared_mutex> wr(s->lock);
  s->dirty = true; s->adopted = false;
  return s->source.delete_blas(handle) ? RDN_OK : fail(RDN_ERR_INVALID_HANDLE, "rdn_rt_blas_destroy: unknown handle");
}

int rdn_rt_tlas_create(rdn_rt_scene *s, const rdn_instance *inst, uint32_t n, uint32_t *out_handle) {
  if (!s || !out_handle || (n && !inst)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_tlas_create: null argument");
  std::vector<InstanceSource> src(n);
  for (uint32_t i = 0; i < n; ++i) {
    std::memcpy(&src[i].transform, inst[i].transform, sizeof(Mat4));
    src[i].instance_custom_index = inst[i].instance_custom_index;
    src[i].mask = inst[i].mask;
    src[i].sbt_offset = inst[i].instance_shader_binding_table_record_offset;
    src[i].flags = inst[i].flags;
    src[i].blas_handle = inst[i].blas_handle;
  }
  std::unique_lock<std::shared_mutex> wr(s->lock);
  s->dirty = true; s->adopted = false;
  *out_handle = s->source.create_tlas(std::move(src));
  return RDN_OK;
}

int rdn_rt_tlas_update(rdn_rt_scene *s, uint32_t handle, const rdn_instance *inst, uint32_t n) {
  if (!s || (n && !inst)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_tlas_update: null argument");
  warm_worker_pool();  // (the commit that follows is a handful of short parallel sections)
  std::vector<InstanceSource> src(n);
  for (uint32_t i = 0; i < n; ++i) {
    std::memcpy(&src[i].transform, inst[i].transform, sizeof(Mat4));
    src[i].instance_custom_index = inst[i].instance_custom_index;
    src[i].mask = inst[i].mask;
    src[i].sbt_offset = inst[i].instance_shader_binding_table_record_offset;
    src[i].flags = inst[i].flags;
    src[i].blas_handle = inst[i].blas_handle;
  }
  std::unique_lock<std::shared_mutex> wr(s->lock);
  if (s->adopted) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_tlas_update: this scene adopted another rank's blob and has no sources");
  if (!s->source.update_tlas(handle, std::move(src))) return fail(RDN_ERR_INVALID_HANDLE, "rdn_rt_tlas_update: unknown or deleted TLAS");
  s->dirty = true;
  return RDN_OK;
}

int rdn_rt_tlas_destroy(rdn_rt_scene *s, uint32_t handle) {
  if (!s) return fail(RDN_ERR_INVALID_ARGUMENT, "null scene");
  std::unique_lock<std::shared_mutex> wr(s->lock);
  s->dirty = true; s->adopted = false;
  return s->source.delete_tlas(handle) ? RDN_OK : fail(RDN_ERR_INVALID_HANDLE, "rdn_rt_tlas_destroy: unknown handle");
}

int rdn_rt_bind_tlas(rdn_rt_scene *s, const uint32_t *handles, uint32_t n) {
  if (!s || (n && !handles)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_bind_tlas: null argument");
  std::vector<uint32_t> nb(handles, handles + n);
  std::unique_lock<std::shared_mutex> wr(s->lock);
  if (nb != s->tlas_binding) {  // set_binding, mod.rs:539-545
    s->tlas_binding = std::move(nb);
    s->dirty = true; s->adopted = false;
  }
  return RDN_OK;
}

uint32_t rdn_rt_bind_tlas_max_len(const rdn_rt_scene *) { return 0xFFFFFFFFu; }  // mod.rs:569-571

int rdn_rt_commit(rdn_rt_scene *s) {
  if (!s) return fail(RDN_ERR_INVALID_ARGUMENT, "null scene");
  std::unique_lock<std::shared_mutex> wr(s->lock);
  if (s->dirty) warm_worker_pool();
  return commit_locked(s);
}

static int trace_device_impl(rdn_rt_scene *s, int device_index, const rdn_launch *launch, const rdn_ray *d_rays, uint64_t n,
                             rdn_hit *d_hits, void *cuda_stream, int mode, rdn_trace_stats *stats, const unsigned long long *d_n) {
  if (!s || !launch || (n && (!d_rays || !d_hits))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_closest_device: null argument");
  if (d_n && n > MAX_LAUNCH_RAYS) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_closest_device_n: n_max above 2^31 rays");
  if ((reinterpret_cast<uintptr_t>(d_rays) | reinterpret_cast<uintptr_t>(d_hits)) & 31u)
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_closest_device: ray and hit arrays must be 32-byte aligned (one 256-bit access per record)");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index (host-only scene?)");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  std::lock_guard<std::mutex> lg(s->launch_lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (stats) std::memset(stats, 0, sizeof(*stats));
  // Tail overlap with the previous launch (programmatic dependent launch) is the caller's request AND the library's own check:
  // the previous thing this library put into the same stream was a single ordered launch, and this launch neither reads what that
  // one writes nor writes what it reads or writes.  (What the caller itself enqueued in between the library cannot see: that is
  // the contract of RDN_TRACE_OVERLAP_PREVIOUS, include/rdn_rt.h.)
  const bool want_overlap = (mode & RDN_TRACE_OVERLAP_PREVIOUS) != 0;
  mode &= ~RDN_TRACE_OVERLAP_PREVIOUS;
  const uintptr_t rays_lo = reinterpret_cast<uintptr_t>(d_rays), rays_hi = rays_lo + n * sizeof(rdn_ray);
  const uintptr_t hits_lo = reinterpret_cast<uintptr_t>(d_hits), hits_hi = hits_lo + n * sizeof(rdn_hit);
  auto overlaps = [](uintptr_t a0, uintptr_t a1, uintptr_t b0, uintptr_t b1) { return a0 < b1 && b0 < a1; };
  const bool overlap_ok = want_overlap && !stats && dc.ext_prev.valid && dc.ext_last_stream == stream &&
                          !overlaps(rays_lo, rays_hi, dc.ext_prev.hits_lo, dc.ext_prev.hits_hi) &&
                          !overlaps(hits_lo, hits_hi, dc.ext_prev.hits_lo, dc.ext_prev.hits_hi) &&
                          !overlaps(hits_lo, hits_hi, dc.ext_prev.rays_lo, dc.ext_prev.rays_hi);
  dc.ext_prev.valid = false;
  // The scratch sets are shared by every caller stream: a call on another stream than the previous one first waits for that
  // stream's work (event recorded there now).  Calls on the same stream need nothing — and get nothing between their kernels.
  if (!dc.ext_done) RDN_CUDA(cudaEventCreateWithFlags(&dc.ext_done, cudaEventDisableTiming));
  if (dc.ext_pending && dc.ext_last_stream != stream) {
    if (cudaEventRecord(dc.ext_done, dc.ext_last_stream) == cudaSuccess) RDN_CUDA(cudaStreamWaitEvent(stream, dc.ext_done, 0));
    else { cudaGetLastError(); RDN_CUDA(cudaDeviceSynchronize()); }  // the previous stream no longer exists
  }
  dc.ext_last_stream = stream;
  dc.ext_pending = true;
  const uint64_t chunk = std::min<uint64_t>(std::max<uint64_t>(n, 1), MAX_LAUNCH_RAYS);
  rc = ensure_scratch(dc.ext_scratch[0], chunk);
  if (rc == RDN_OK) rc = ensure_scratch(dc.ext_scratch[1], chunk);
  if (rc != RDN_OK) return rc;

  struct EventPair {  // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } events;
  cudaEvent_t &e0 = events.a, &e1 = events.b;
  if (stats) { RDN_CUDA(cudaEventCreate(&e0)); RDN_CUDA(cudaEventCreate(&e1)); }
  uint64_t ties = 0, fallbacks = 0;
  float ms_total = 0.f;
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = std::min(chunk, n - off);
    rdn_launch l = *launch;
    if (off != 0 || m != n) l.grid_width = (l.grid_width && m % l.grid_width == 0 && off % l.grid_width == 0) ? l.grid_width : 0;
    Scratch &scratch = dc.ext_scratch[dc.ext_next];
    dc.ext_next ^= 1;
#if defined(RDN_DEBUG_STEPS) || defined(RDN_DEBUG_TIMELINE)
    if (stats) {
      RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 32, 0, 6 * 8, stream));
      RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 32 + 6 * 8, 0xFF, 2 * 8, stream));
      RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 32 + 8 * 8, 0, 16, stream));
      RDN_CUDA(cudaMemsetAsync(static_cast<char *>(scratch.base) + 128, 0, SCRATCH_BASE_BYTES - 128, stream));
    }
#endif
    if (stats) RDN_CUDA(cudaEventRecord(e0, stream));
    rc = enqueue_trace(s, dc, scratch, l, d_rays + off, m, d_hits + off, mode, stream, stats ? &stats->kernel_launches : nullptr, stats != nullptr,
                       overlap_ok && off == 0, d_n, /*tile_history=*/m == n);
    if (rc != RDN_OK) return rc;
    if (!stats && m == n && dc.last_enqueue_overlappable) {
      dc.ext_prev.valid = true;
      dc.ext_prev.rays_lo = rays_lo; dc.ext_prev.rays_hi = rays_hi; dc.ext_prev.hits_lo = hits_lo; dc.ext_prev.hits_hi = hits_hi;
    }
    if (stats) {
      RDN_CUDA(cudaEventRecord(e1, stream));
      RDN_CUDA(cudaEventSynchronize(e1));
      float ms = 0.f;
      RDN_CUDA(cudaEventElapsedTime(&ms, e0, e1));
      ms_total += ms;
      uint32_t small[6] = {0, 0, 0, 0, 0, 0};  // tie_count | tie_unresolved | stack_overflow | blocks_done | tie_cursor | tie_total
      RDN_CUDA(cudaMemcpy(small, static_cast<char *>(scratch.base) + 8, sizeof(small), cudaMemcpyDeviceToHost));
      ties += small[5];
      fallbacks += small[1];
      rc = take_error_flags(scratch);  // (reported once: the flags are cleared)
      if (rc != RDN_OK) return rc;
#if defined(RDN_DEBUG_STEPS) || defined(RDN_DEBUG_TIMELINE)
      unsigned long long c[10];
      RDN_CUDA(cudaMemcpy(c, static_cast<char *>(scratch.base) + 32, sizeof(c), cudaMemcpyDeviceToHost));
      const double r = double(c[2] ? c[2] : 1);
      const double span_us = (double(c[8]) - double(c[6])) * 1e-3;
      const double warps = double((m + 31) / 32 < 148ull * 32 ? (m + 31) / 32 : 148ull * 32);
      fprintf(stderr, "[dbg steps] max=%llu steps=%llu rays_entered=%llu tri_tests=%llu pushes=%llu rays>200steps=%llu | per entered ray: %.2f steps "
              "%.2f tris %.2f pushes | timeline: list dry after %.1f us, last warp out after %.1f us (tail %.1f us), mean warp busy %.1f us = %.0f%% of the span\n",
              c[0], c[1], c[2], c[3], c[4], c[5], double(c[1]) / r, double(c[3]) / r, double(c[4]) / r,
              (double(c[7]) - double(c[6])) * 1e-3, span_us, (double(c[8]) - double(c[7])) * 1e-3, double(c[9]) * 1e-3 / warps,
              100.0 * double(c[9]) * 1e-3 / warps / (span_us > 0 ? span_us : 1));
      {  // tile passes by duration (8 us buckets): count, mean rounds; the longest pass
        unsigned long long h[100];
        RDN_CUDA(cudaMemcpy(h, static_cast<char *>(scratch.base) + 128, sizeof(h), cudaMemcpyDeviceToHost));
        fprintf(stderr, "[dbg tiles] passes by duration (8 us buckets) count/mean_rounds/mean_busy_lanes_at_round_start:");
        for (int b = 0; b < 32; ++b)
          if (h[b]) fprintf(stderr, " [%d-%d us] %llu/%.1f/%.1f", b * 8, b * 8 + 8, h[b], double(h[32 + b]) / double(h[b]), double(h[66 + b]) / double(h[32 + b] ? h[32 + b] : 1));
        fprintf(stderr, " | longest pass %.1f us, %llu rounds\n", double(h[64] >> 20) * 1e-3, h[64] & 0xFFFFFull);
      }
#endif
    }
  }
  if (stats) {
    stats->rays = n; stats->tie_rays = ties; stats->kernel_ms = ms_total; stats->whole_range_rewalks = fallbacks;
  }
  return RDN_OK;
}

int rdn_rt_trace_closest_device(rdn_rt_scene *s, int device_index, const rdn_launch *launch, const rdn_ray *d_rays, uint64_t n,
                                rdn_hit *d_hits, void *cuda_stream, int mode, rdn_trace_stats *stats) {
  return trace_device_impl(s, device_index, launch, d_rays, n, d_hits, cuda_stream, mode, stats, nullptr);
}

int rdn_rt_trace_closest_device_n(rdn_rt_scene *s, int device_index, const rdn_launch *launch, const rdn_ray *d_rays, const uint64_t *d_n,
                                  uint64_t n_max, rdn_hit *d_hits, void *cuda_stream, int mode) {
  if (!d_n) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_closest_device_n: null count");
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "u64");
  return trace_device_impl(s, device_index, launch, d_rays, n_max, d_hits, cuda_stream, mode, nullptr, reinterpret_cast<const unsigned long long *>(d_n));
}

int rdn_rt_trace_closest(rdn_rt_scene *s, const rdn_launch *launch, const rdn_ray *rays, uint64_t n, rdn_hit *out_hits) {
  if (!s || !launch || (n && (!rays || !out_hits))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_closest: null argument");
  if (s->devices.empty()) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_closest: host-only scene has no CUDA device (there is no CPU fallback)");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  if (n == 0) return RDN_OK;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  std::lock_guard<std::mutex> lg(s->launch_lock);

  // chunk = whole rows of the launch grid (multiple of 4 rows) when a grid hint is given, so every chunk keeps the hint
  static const uint64_t chunk_rays = []() {  // RDN_HOST_CHUNK_RAYS: experimentation knob
    const char *e = getenv("RDN_HOST_CHUNK_RAYS");
    const long long v = e ? atoll(e) : 0;
    return v > 0 ? static_cast<uint64_t>(v) : HOST_CHUNK_RAYS;
  }();
  uint64_t chunk = chunk_rays;
  const uint32_t gw = launch->grid_width;
  const bool grid = gw != 0 && n % gw == 0;
  if (grid) {
    uint64_t rows = std::max<uint64_t>(4, (chunk_rays / gw) / 4 * 4);
    chunk = rows * gw;
  }
  const uint64_t n_chunks = (n + chunk - 1) / chunk;
  const size_t n_dev = s->devices.size();

  // Ordinary (pageable) caller memory — what the reference's own callers hand over — is not copied from directly: the driver would
  // stage it through its own small buffer, synchronously, at a fifth of the link's rate.  Each chunk goes through page-locked
  // staging buffers of its slot instead, filled / emptied by several host threads while earlier chunks are on the link and in the
  // kernels.  Buffers from rdn_rt_host_alloc / rdn_rt_host_register are copied from directly.
  const bool stage_in = is_pageable(rays), stage_out = is_pageable(out_hits);
  // tiles (chunks) are dealt round-robin to devices; each device pipelines its chunks over N_SLOTS streams:
  // H2D(k+1) || kernels(k) || D2H(k-1)
  for (size_t di = 0; di < n_dev; ++di) {
    DeviceCtx &dc = s->devices[di];
    RDN_CUDA(cudaSetDevice(dc.device));
    for (Slot &slot : dc.slots) {
      rc = ensure_slot(slot, std::min<uint64_t>(chunk, n));
      if (rc != RDN_OK) return rc;
      if (stage_in || stage_out) {
        rc = ensure_slot_staging(slot, std::min<uint64_t>(chunk, n));
        if (rc != RDN_OK) return rc;
      }
      slot.chunk_in_flight = false;
    }
  }
  struct Pending { Slot *slot; int device; uint64_t off, m; };
  std::deque<Pending> pending;  // staged chunks in issue order: results still to be copied out of the slot's staging buffer
  auto retire_front = [&]() -> cudaError_t {
    const Pending p = pending.front();
    pending.pop_front();
    cudaError_t e = cudaSetDevice(p.device);
    if (e == cudaSuccess) e = cudaEventSynchronize(p.slot->chunk_done);
    if (e != cudaSuccess) return e;
    if (stage_out) parallel_copy(out_hits + p.off, p.slot->h_hits, p.m * sizeof(rdn_hit));
    p.slot->chunk_in_flight = false;
    return cudaSuccess;
  };
  std::vector<uint64_t> per_dev_seq(n_dev, 0);
  for (uint64_t c = 0; c < n_chunks; ++c) {
    const size_t di = c % n_dev;
    DeviceCtx &dc = s->devices[di];
    Slot &slot = dc.slots[per_dev_seq[di]++ % N_SLOTS];
    const uint64_t off = c * chunk, m = std::min(chunk, n - off);
    rdn_launch l = *launch;
    l.grid_width = grid ? gw : 0;
    const rdn_ray *src = rays + off;
    rdn_hit *dst = out_hits + off;
    if (stage_in || stage_out) {
      // the slot's previous chunk must have left its staging buffers (chunks retire in issue order)
      while (slot.chunk_in_flight) RDN_CUDA(retire_front());
      if (stage_in) { parallel_copy(slot.h_rays, src, m * sizeof(rdn_ray)); src = slot.h_rays; }
      if (stage_out) dst = slot.h_hits;
    }
    RDN_CUDA(cudaSetDevice(dc.device));
    RDN_CUDA(cudaMemcpyAsync(slot.d_rays, src, m * sizeof(rdn_ray), cudaMemcpyHostToDevice, slot.stream));
    rc = enqueue_trace(s, dc, slot.scratch, l, slot.d_rays, m, slot.d_hits, RDN_TRACE_AUTO, slot.stream, nullptr);
    if (rc != RDN_OK) return rc;
    RDN_CUDA(cudaMemcpyAsync(dst, slot.d_hits, m * sizeof(rdn_hit), cudaMemcpyDeviceToHost, slot.stream));
    if (stage_in || stage_out) {
      RDN_CUDA(cudaEventRecord(slot.chunk_done, slot.stream));
      slot.chunk_in_flight = true;
      pending.push_back(Pending{&slot, dc.device, off, m});
    }
  }
  while (!pending.empty()) RDN_CUDA(retire_front());
  // error flags of every used slot come back through pinned memory behind the slot's last chunk: one wait per stream
  for (size_t di = 0; di < n_dev; ++di) {
    DeviceCtx &dc = s->devices[di];
    RDN_CUDA(cudaSetDevice(dc.device));
    const uint64_t used = std::min<uint64_t>(per_dev_seq[di], N_SLOTS);
    for (uint64_t k = 0; k < used; ++k) {
      Slot &slot = dc.slots[k];
      RDN_CUDA(cudaMemcpyAsync(slot.h_flags, static_cast<char *>(slot.scratch.base) + 8, 3 * sizeof(uint32_t), cudaMemcpyDeviceToHost, slot.stream));
    }
  }
  int host_rc = RDN_OK;
  for (size_t di = 0; di < n_dev; ++di) {
    DeviceCtx &dc = s->devices[di];
    RDN_CUDA(cudaSetDevice(dc.device));
    const uint64_t used = std::min<uint64_t>(per_dev_seq[di], N_SLOTS);
    for (uint64_t k = 0; k < used; ++k) {
      Slot &slot = dc.slots[k];
      RDN_CUDA(cudaStreamSynchronize(slot.stream));
      if (slot.h_flags[2]) { take_error_flags(slot.scratch); host_rc = RDN_ERR_CAPACITY; }
    }
  }
  return host_rc;
}

// Page-locked host memory for the host-buffer entry points.  Transfers from pageable memory (a plain Vec, a numpy array) are staged by
// the driver: they neither overlap with the kernels nor reach PCIe speed.  The library does NOT page-lock caller memory behind the
// caller's back: a registration outlives a free() it cannot see, and the stale entry then fails every later copy that touches a new
// allocation at the same addresses (tried: a cache keyed by pointer and size broke an unrelated cudaMemcpy two tests later).  The
// owner of the memory decides: allocate it here, or register what it owns for as long as it lives.
int rdn_rt_host_alloc(uint64_t bytes, void **out) {
  if (!out) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_host_alloc: null argument");
  *out = nullptr;
  RDN_CUDA(cudaHostAlloc(out, std::max<uint64_t>(bytes, 1), cudaHostAllocPortable));
  return RDN_OK;
}
void rdn_rt_host_free(void *ptr) {
  if (ptr) cudaFreeHost(ptr);
}
int rdn_rt_host_register(void *ptr, uint64_t bytes) {
  if (!ptr || !bytes) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_host_register: null argument");
  RDN_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
  return RDN_OK;
}
int rdn_rt_host_unregister(void *ptr) {
  if (!ptr) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_host_unregister: null argument");
  RDN_CUDA(cudaHostUnregister(ptr));
  return RDN_OK;
}

int rdn_rt_set_any_hit_programs(rdn_rt_scene *s, const rdn_anyhit_program *programs, uint32_t n) {
  if (!s || (n && !programs)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_set_any_hit_programs: null argument");
  for (uint32_t k = 0; k < n; ++k)
    if (programs[k].kind > RDN_ANYHIT_MIN_DISTANCE) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_set_any_hit_programs: unknown program kind");
  std::unique_lock<std::shared_mutex> wr(s->lock);
  std::lock_guard<std::mutex> lg(s->launch_lock);
  s->anyhit_programs.assign(programs, programs + n);
  for (DeviceCtx &dc : s->devices) dc.anyhit_stale = true;
  return RDN_OK;
}

int rdn_rt_bind_sbt(rdn_rt_scene *s, rdn_sbt *sbt) {
  if (!s) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_bind_sbt: null scene");
  if (sbt && sbt->scene != s) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_bind_sbt: the table belongs to another scene");
  std::unique_lock<std::shared_mutex> wr(s->lock);
  s->bound_sbt = sbt;
  return RDN_OK;
}

int rdn_rt_poll_errors(rdn_rt_scene *s, int device_index, void *cuda_stream, uint32_t *out_flags) {
  if (!s) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_poll_errors: null scene");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index (host-only scene?)");
  if (out_flags) *out_flags = 0;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  std::lock_guard<std::mutex> lg(s->launch_lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  RDN_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(cuda_stream)));
  int rc = RDN_OK;
  for (Scratch &scratch : dc.ext_scratch) {
    if (!scratch.base) continue;
    const int r = take_error_flags(scratch, out_flags);
    if (r != RDN_OK && rc == RDN_OK) rc = r;
  }
  return rc;
}

int rdn_rt_trace_counted(rdn_rt_scene *s, const rdn_launch *launch, const rdn_ray *rays, uint64_t n, rdn_hit *out_hits,
                         rdn_counters *out_counters) {
  if (!s || !launch || (n && (!rays || !out_hits))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_counted: null argument");
  if (s->devices.empty()) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_counted: host-only scene has no CUDA device (there is no CPU fallback)");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  std::lock_guard<std::mutex> lg(s->launch_lock);
  DeviceCtx &dc = s->devices[0];
  RDN_CUDA(cudaSetDevice(dc.device));
  if (out_counters) std::memset(out_counters, 0, sizeof(*out_counters));
  const uint64_t chunk = 1u << 22;
  Slot &slot = dc.slots[0];
  rc = ensure_slot(slot, std::min<uint64_t>(chunk, std::max<uint64_t>(n, 1)));
  if (rc != RDN_OK) return rc;
  for (uint64_t off = 0; off < n; off += chunk) {
    const uint64_t m = std::min(chunk, n - off);
    RDN_CUDA(cudaMemcpyAsync(slot.d_rays, rays + off, m * sizeof(rdn_ray), cudaMemcpyHostToDevice, slot.stream));
    RDN_CUDA(cudaMemsetAsync(static_cast<char *>(slot.scratch.base) + 32, 0, 10 * 8, slot.stream));  // the visit counters only
    SceneDev dev = dc.dev;
    bool unused = false;
    rc = prepare_any_hit(s, dc, *launch, dev, unused);
    if (rc != RDN_OK) return rc;
    launch_trace_reference(dev, *launch, slot.d_rays, m, slot.d_hits, slot.scratch.view(), true, dc.sm_count, slot.stream);
    RDN_CUDA(cudaGetLastError());
    RDN_CUDA(cudaMemcpyAsync(out_hits + off, slot.d_hits, m * sizeof(rdn_hit), cudaMemcpyDeviceToHost, slot.stream));
    unsigned long long c[6];
    RDN_CUDA(cudaMemcpyAsync(c, static_cast<char *>(slot.scratch.base) + 32, sizeof(c), cudaMemcpyDeviceToHost, slot.stream));
    RDN_CUDA(cudaStreamSynchronize(slot.stream));
    if (out_counters) {
      out_counters->bvh_visit += c[0]; out_counters->bvh_hit += c[1]; out_counters->tri_visit += c[2];
      out_counters->tri_hit += c[3]; out_counters->inst_visit += c[4]; out_counters->ref_abort += c[5];
    }
  }
  return RDN_OK;
}

// ------------------------------------------------------------------------------------------------ ray generation / bounce (f1)
int rdn_rt_gen_pinhole_rays_device(rdn_rt_scene *s, int device_index, const rdn_pinhole *p, rdn_ray *d_rays, void *cuda_stream) {
  if (!s || !p || !d_rays) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_pinhole_rays_device: null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  if (p->width == 0 || p->height == 0 || p->rect_x + static_cast<uint64_t>(p->rect_w) > p->width || p->rect_y + static_cast<uint64_t>(p->rect_h) > p->height)
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_pinhole_rays_device: rectangle outside the launch");
  RDN_CUDA(cudaSetDevice(s->devices[device_index].device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
  launch_gen_pinhole_rays(*p, d_rays, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int rdn_rt_gen_pinhole_rays_batch_device(rdn_rt_scene *s, int device_index, const rdn_pinhole *params, uint32_t n_params, rdn_ray *d_rays,
                                         void *cuda_stream) {
  if (!s || (n_params && (!params || !d_rays))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_pinhole_rays_batch_device: null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  if (n_params == 0) return RDN_OK;
  std::vector<uint64_t> offsets(n_params);
  uint64_t total = 0, biggest = 0;
  for (uint32_t k = 0; k < n_params; ++k) {
    const rdn_pinhole &p = params[k];
    if (p.width == 0 || p.height == 0 || p.rect_x + static_cast<uint64_t>(p.rect_w) > p.width || p.rect_y + static_cast<uint64_t>(p.rect_h) > p.height)
      return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_pinhole_rays_batch_device: rectangle outside the launch");
    offsets[k] = total;
    const uint64_t m = static_cast<uint64_t>(p.rect_w) * p.rect_h;
    total += m;
    biggest = std::max(biggest, m);
  }
  RDN_CUDA(cudaSetDevice(s->devices[device_index].device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  // descriptors travel through a stream-ordered allocation, so calls on different streams never share them
  rdn_pinhole *d_params = nullptr;
  uint64_t *d_offsets = nullptr;
  RDN_CUDA(cudaMallocAsync(&d_params, n_params * sizeof(rdn_pinhole), stream));
  RDN_CUDA(cudaMallocAsync(&d_offsets, n_params * sizeof(uint64_t), stream));
  RDN_CUDA(cudaMemcpyAsync(d_params, params, n_params * sizeof(rdn_pinhole), cudaMemcpyHostToDevice, stream));
  RDN_CUDA(cudaMemcpyAsync(d_offsets, offsets.data(), n_params * sizeof(uint64_t), cudaMemcpyHostToDevice, stream));
  for (uint32_t first = 0; first < n_params; first += 65535u) {
    const uint32_t m = std::min<uint32_t>(65535u, n_params - first);
    launch_gen_pinhole_rays_batch(d_params + first, d_offsets + first, m, biggest, d_rays, stream);
  }
  RDN_CUDA(cudaGetLastError());
  RDN_CUDA(cudaFreeAsync(d_params, stream));
  RDN_CUDA(cudaFreeAsync(d_offsets, stream));
  return RDN_OK;
}

int rdn_rt_gen_camera_rays_device(rdn_rt_scene *s, int device_index, const rdn_camera *p, rdn_ray *d_rays, void *cuda_stream) {
  if (!s || !p || !d_rays) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_camera_rays_device: null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  if (p->width == 0 || p->height == 0 || p->rect_x + static_cast<uint64_t>(p->rect_w) > p->width || p->rect_y + static_cast<uint64_t>(p->rect_h) > p->height)
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_camera_rays_device: rectangle outside the launch");
  RDN_CUDA(cudaSetDevice(s->devices[device_index].device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
  launch_gen_camera_rays(*p, d_rays, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int rdn_rt_gen_bounce_rays_device(rdn_rt_scene *s, int device_index, const rdn_bounce *p, const rdn_ray *d_rays_in, const rdn_hit *d_hits,
                                  uint64_t n, rdn_ray *d_rays_out, uint32_t *d_src_index, uint64_t *d_out_n, void *cuda_stream) {
  if (!s || !p || !d_out_n || (n && (!d_rays_in || !d_hits || !d_rays_out || !d_src_index)))
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_bounce_rays_device: null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  if (p->mode > 2 || (p->mode == 1 && p->max_sample == 0)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_gen_bounce_rays_device: bad mode");
  if (n > MAX_LAUNCH_RAYS) return fail(RDN_ERR_CAPACITY, "rdn_rt_gen_bounce_rays_device: more than 2^31 rays in one call");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  {
    std::shared_lock<std::shared_mutex> rd(s->lock);
    std::lock_guard<std::mutex> lg(s->launch_lock);
    DeviceCtx &dc = s->devices[device_index];
    RDN_CUDA(cudaSetDevice(dc.device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
    if (dc.bounce_cap < n) {
      if (dc.d_keep) cudaFree(dc.d_keep);
      if (dc.d_iota) cudaFree(dc.d_iota);
      dc.d_keep = nullptr; dc.d_iota = nullptr; dc.bounce_cap = 0;
      RDN_CUDA(cudaMalloc(&dc.d_keep, std::max<uint64_t>(n, 1)));
      RDN_CUDA(cudaMalloc(&dc.d_iota, std::max<uint64_t>(n, 1) * sizeof(uint32_t)));
      dc.bounce_cap = n;
    }
    launch_mark_hits(d_hits, n, dc.d_keep, dc.d_iota, stream);
  }
  DeviceCtx &dc = s->devices[device_index];
  rc = rdn_rt_compact_u32_device(s, device_index, dc.d_iota, dc.d_keep, n, d_src_index, d_out_n, cuda_stream);
  if (rc != RDN_OK) return rc;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  launch_gen_bounce_rays(dc.dev, *p, d_rays_in, d_hits, d_src_index, d_out_n, n, d_rays_out, stream);
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int rdn_rt_kernel_timing_begin(rdn_rt_scene *s, int device_index) {
  if (!s || device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_kernel_timing_begin: bad argument");
  std::lock_guard<std::mutex> lg(s->launch_lock);
  DeviceCtx &dc = s->devices[device_index];
  for (TimedLaunch &t : dc.timed) { cudaEventDestroy(t.begin); cudaEventDestroy(t.end); }
  dc.timed.clear();
  dc.timing = true;
  return RDN_OK;
}

int rdn_rt_kernel_timing_end(rdn_rt_scene *s, int device_index, rdn_kernel_times *out) {
  if (!s || !out || device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_kernel_timing_end: bad argument");
  std::lock_guard<std::mutex> lg(s->launch_lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  std::memset(out, 0, sizeof(*out));
  dc.timing = false;
  for (TimedLaunch &t : dc.timed) {
    RDN_CUDA(cudaEventSynchronize(t.end));
    float ms = 0.f;
    RDN_CUDA(cudaEventElapsedTime(&ms, t.begin, t.end));
    if (t.kind == KERNEL_ORDERED) { out->ordered_launches++; out->ordered_ms += ms; }
    else if (t.kind == KERNEL_TIES) { out->tie_launches++; out->tie_ms += ms; }
    else { out->reference_launches++; out->reference_ms += ms; }
    cudaEventDestroy(t.begin); cudaEventDestroy(t.end);
  }
  dc.timed.clear();
  return RDN_OK;
}

int rdn_rt_compact_u32_device(rdn_rt_scene *s, int device_index, const uint32_t *d_in, const uint8_t *d_keep, uint64_t n,
                              uint32_t *d_out, uint64_t *d_out_n, void *cuda_stream) {
  if (!s || !d_out_n || (n && (!d_in || !d_keep || !d_out))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_compact_u32_device: null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  std::lock_guard<std::mutex> lg(s->launch_lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
  const uint64_t words = compact_status_words(n);
  if (dc.compact_status_cap < words) {
    if (dc.d_compact_status) cudaFree(dc.d_compact_status);
    dc.d_compact_status = nullptr; dc.compact_status_cap = 0;
    RDN_CUDA(cudaMalloc(&dc.d_compact_status, words * sizeof(unsigned long long)));
    dc.compact_status_cap = words;
  }
  launch_compact_u32(d_in, d_keep, n, d_out, d_out_n, dc.d_compact_status, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int rdn_rt_compact_u32(rdn_rt_scene *s, const uint32_t *in, const uint8_t *keep, uint64_t n, uint32_t *out, uint64_t *out_n) {
  if (!s || !out_n || (n && (!in || !keep || !out))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_compact_u32: null argument");
  if (s->devices.empty()) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_compact_u32: host-only scene has no CUDA device (there is no CPU fallback)");
  DeviceCtx &dc = s->devices[0];
  RDN_CUDA(cudaSetDevice(dc.device));
  uint32_t *d_in = nullptr, *d_out = nullptr;
  uint8_t *d_keep = nullptr;
  uint64_t *d_n = nullptr;
  const uint64_t m = std::max<uint64_t>(n, 1);
  RDN_CUDA(cudaMalloc(&d_in, m * 4)); RDN_CUDA(cudaMalloc(&d_out, m * 4)); RDN_CUDA(cudaMalloc(&d_keep, m)); RDN_CUDA(cudaMalloc(&d_n, 8));
  RDN_CUDA(cudaMemcpy(d_in, in, n * 4, cudaMemcpyHostToDevice));
  RDN_CUDA(cudaMemcpy(d_keep, keep, n, cudaMemcpyHostToDevice));
  int rc = rdn_rt_compact_u32_device(s, 0, d_in, d_keep, n, d_out, d_n, nullptr);
  if (rc == RDN_OK) {
    RDN_CUDA(cudaDeviceSynchronize());
    RDN_CUDA(cudaMemcpy(out, d_out, n * 4, cudaMemcpyDeviceToHost));
    RDN_CUDA(cudaMemcpy(out_n, d_n, 8, cudaMemcpyDeviceToHost));
  }
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_keep); cudaFree(d_n);
  return rc;
}

int rdn_rt_scene_blob(rdn_rt_scene *s, int device_index, void **out_ptr, uint64_t *out_bytes) {
  if (!s || !out_ptr || !out_bytes) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_blob: null argument");
  const bool host_only = s->devices.empty() && device_index == -1;
  if (!host_only && (device_index < 0 || device_index >= static_cast<int>(s->devices.size()))) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  if (host_only) {  // host-only scene (n_devices == 0): the blob lives in host memory
    *out_ptr = s->host_blob.data();
    *out_bytes = s->host_blob.size();
    return RDN_OK;
  }
  *out_ptr = s->devices[device_index].d_blob;
  *out_bytes = s->devices[device_index].blob_bytes;
  return RDN_OK;
}

int rdn_rt_scene_adopt_blob(rdn_rt_scene *s, int device_index, const void *d_blob, uint64_t bytes) {
  if (!s || !d_blob || bytes < sizeof(BlobHeader)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_adopt_blob: null/short blob");
  if (s->devices.empty() && device_index == -1) {  // host-only scene adopting a host blob (replication logic without a GPU)
    BlobHeader hh;
    std::memcpy(&hh, d_blob, sizeof(hh));
    if (!header_ok(hh, bytes)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_adopt_blob: bad header");
    std::unique_lock<std::shared_mutex> wr(s->lock);
    s->host_blob.assign(static_cast<const uint8_t *>(d_blob), static_cast<const uint8_t *>(d_blob) + bytes);
    const uint8_t *b = s->host_blob.data();
    s->h_tlas_binding.assign(reinterpret_cast<const uint32_t *>(b + hh.offset[ARR_TLAS_BINDING]),
                             reinterpret_cast<const uint32_t *>(b + hh.offset[ARR_TLAS_BINDING]) + hh.count[ARR_TLAS_BINDING]);
    s->h_tlas_root.assign(reinterpret_cast<const TlasRoot *>(b + hh.offset[ARR_TLAS_ROOT]),
                          reinterpret_cast<const TlasRoot *>(b + hh.offset[ARR_TLAS_ROOT]) + hh.count[ARR_TLAS_ROOT]);
    s->flat = FlatScene{};
    s->adopted = true;
    s->dirty = false;
    return RDN_OK;
  }
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  std::unique_lock<std::shared_mutex> wr(s->lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  BlobHeader h;
  RDN_CUDA(cudaMemcpy(&h, d_blob, sizeof(h), cudaMemcpyDeviceToHost));
  if (!header_ok(h, bytes)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_adopt_blob: bad header");
  if (dc.d_blob) { cudaFree(dc.d_blob); dc.d_blob = nullptr; }
  RDN_CUDA(cudaMalloc(&dc.d_blob, bytes));
  RDN_CUDA(cudaMemcpy(dc.d_blob, d_blob, bytes, cudaMemcpyDeviceToDevice));
  dc.blob_bytes = bytes;
  bind_blob(dc, h);
  s->h_tlas_binding.assign(h.count[ARR_TLAS_BINDING], 0);
  s->h_tlas_root.assign(h.count[ARR_TLAS_ROOT], TlasRoot{INVALID_NEXT, REF_EMPTY, 0, 0, 0, 0, 0, REF_EMPTY});
  if (!s->h_tlas_binding.empty())
    RDN_CUDA(cudaMemcpy(s->h_tlas_binding.data(), dc.dev.tlas_binding, s->h_tlas_binding.size() * 4, cudaMemcpyDeviceToHost));
  if (!s->h_tlas_root.empty())
    RDN_CUDA(cudaMemcpy(s->h_tlas_root.data(), dc.dev.tlas_root, s->h_tlas_root.size() * sizeof(TlasRoot), cudaMemcpyDeviceToHost));
  s->flat = FlatScene{};
  s->adopted = true;
  s->dirty = false;
  return RDN_OK;
}

int rdn_rt_scene_array(rdn_rt_scene *s, int array_id, void *out, uint64_t capacity, uint64_t *out_bytes) {
  if (!s || !out_bytes || array_id < 0 || array_id >= ARR_COUNT) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_array: bad argument");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  BlobHeader h;
  if (s->devices.empty()) {
    std::memcpy(&h, s->host_blob.data(), sizeof(h));
  } else {
    RDN_CUDA(cudaSetDevice(s->devices[0].device));
    RDN_CUDA(cudaMemcpy(&h, s->devices[0].d_blob, sizeof(h), cudaMemcpyDeviceToHost));
  }
  const uint64_t bytes = h.count[array_id] * h.elem_size[array_id];
  *out_bytes = bytes;
  if (!out) return RDN_OK;
  if (capacity < bytes) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_array: buffer too small");
  if (bytes) {
    if (s->devices.empty()) std::memcpy(out, s->host_blob.data() + h.offset[array_id], bytes);
    else RDN_CUDA(cudaMemcpy(out, static_cast<const char *>(s->devices[0].d_blob) + h.offset[array_id], bytes, cudaMemcpyDeviceToHost));
  }
  return RDN_OK;
}

int rdn_rt_ao_accumulate_device(rdn_rt_scene *s, int device_index, const rdn_hit *d_secondary_hits, const uint32_t *d_src_index,
                                const uint64_t *d_n_secondary, uint64_t n_pixels, uint32_t sample_count, uint32_t max_sample,
                                float *d_ao_buffer, void *cuda_stream) {
  if (!s || !d_n_secondary || (n_pixels && (!d_secondary_hits || !d_src_index || !d_ao_buffer)))
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_ao_accumulate_device: null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  std::lock_guard<std::mutex> lg(s->launch_lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  if (dc.ao_payload_cap < n_pixels) {
    if (dc.d_ao_payload) cudaFree(dc.d_ao_payload);
    dc.d_ao_payload = nullptr; dc.ao_payload_cap = 0;
    RDN_CUDA(cudaMalloc(&dc.d_ao_payload, std::max<uint64_t>(n_pixels, 1) * sizeof(float)));
    dc.ao_payload_cap = n_pixels;
    launch_fill_f32(dc.d_ao_payload, n_pixels, 1.0f, stream);
  }
  launch_ao_accumulate(d_secondary_hits, d_src_index, d_n_secondary, n_pixels, sample_count, max_sample, dc.d_ao_payload, d_ao_buffer, stream);
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

// ------------------------------------------------------------------------------------------------ mesh picking (f2)
struct rdn_pick_mesh {
  int device = -1;
  int sm_count = 148;
  PickMeshDev dev{};
  float *d_positions = nullptr;
  uint32_t *d_indices = nullptr;
  // per-call scratch, grown on demand
  rdn_ray *d_rays = nullptr;
  rdn_mesh_hit *d_hits = nullptr;
  unsigned long long *d_best_key = nullptr;
  uint32_t *d_first_hit = nullptr;
  uint64_t ray_cap = 0;
  uint8_t *d_keep = nullptr;
  uint32_t *d_iota = nullptr, *d_index = nullptr;
  rdn_mesh_hit *d_records = nullptr, *d_gathered = nullptr;
  uint64_t *d_n_kept = nullptr;
  unsigned long long *d_compact_status = nullptr;
  uint64_t gathered_cap = 0;
  std::mutex lock;
};

static void free_pick_mesh(rdn_pick_mesh *m) {
  void *ptrs[] = {m->d_positions, m->d_indices, m->d_rays, m->d_hits, m->d_best_key, m->d_first_hit, m->d_keep, m->d_iota, m->d_index,
                  m->d_records, m->d_gathered, m->d_n_kept, m->d_compact_status};
  if (m->device >= 0) cudaSetDevice(m->device);
  for (void *p : ptrs)
    if (p) cudaFree(p);
}

int rdn_pick_mesh_create(const rdn_mesh_view *mesh, uint32_t topology, int device, rdn_pick_mesh **out) {
  if (!mesh || !out || (!mesh->positions && mesh->n_positions) || topology > RDN_TOPOLOGY_TRIANGLE_STRIP)
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_pick_mesh_create: bad argument");
  const bool indexed = mesh->indices != nullptr;
  if (indexed)
    for (uint64_t k = 0; k < mesh->n_indices; ++k)
      if (mesh->indices[k] >= mesh->n_positions) return fail(RDN_ERR_BUILD, "rdn_pick_mesh_create: vertex index out of bounds");
  static const uint64_t stride_of[5] = {1, 2, 2, 3, 3}, step_of[5] = {1, 2, 1, 3, 1};
  const uint64_t count = indexed ? mesh->n_indices : mesh->n_positions;
  const uint64_t n_prims = count + step_of[topology] < stride_of[topology] ? 0 : (count + step_of[topology] - stride_of[topology]) / step_of[topology];
  if (n_prims >= 0xFFFFFFFFull) return fail(RDN_ERR_CAPACITY, "rdn_pick_mesh_create: more than 2^32 - 1 primitives");
  int available = 0;
  RDN_CUDA(cudaGetDeviceCount(&available));
  if (device < 0 || device >= available) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_pick_mesh_create: no such CUDA device (there is no CPU fallback)");
  auto *m = new rdn_pick_mesh();
  m->device = device;
  auto bail = [&](cudaError_t e, const char *what) { free_pick_mesh(m); delete m; return fail(RDN_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); };
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return bail(e, "cudaSetDevice");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) m->sm_count = prop.multiProcessorCount;
  e = cudaMalloc(&m->d_positions, std::max<uint64_t>(mesh->n_positions, 1) * 3 * sizeof(float));
  if (e != cudaSuccess) return bail(e, "cudaMalloc");
  e = cudaMemcpy(m->d_positions, mesh->positions, mesh->n_positions * 3 * sizeof(float), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return bail(e, "cudaMemcpy");
  if (indexed) {
    e = cudaMalloc(&m->d_indices, std::max<uint64_t>(mesh->n_indices, 1) * sizeof(uint32_t));
    if (e != cudaSuccess) return bail(e, "cudaMalloc");
    e = cudaMemcpy(m->d_indices, mesh->indices, mesh->n_indices * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return bail(e, "cudaMemcpy");
  }
  m->dev = PickMeshDev{m->d_positions, m->d_indices, n_prims, topology};
  *out = m;
  return RDN_OK;
}

void rdn_pick_mesh_destroy(rdn_pick_mesh *m) {
  if (!m) return;
  free_pick_mesh(m);
  delete m;
}

int rdn_pick_mesh_primitive_count(const rdn_pick_mesh *m, uint64_t *out_count) {
  if (!m || !out_count) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_pick_mesh_primitive_count: null argument");
  *out_count = m->dev.n_prims;
  return RDN_OK;
}

int rdn_pick_mesh_nearest(rdn_pick_mesh *m, const rdn_pick_config *config, const rdn_ray *rays, uint64_t n, rdn_mesh_hit *out) {
  if (!m || !config || (n && (!rays || !out))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_pick_mesh_nearest: null argument");
  if (config->triangle_face > RDN_FACE_DOUBLE) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_pick_mesh_nearest: bad triangle_face");
  if (n == 0) return RDN_OK;
  std::lock_guard<std::mutex> lg(m->lock);
  RDN_CUDA(cudaSetDevice(m->device));
  if (m->ray_cap < n) {
    for (void *p : {static_cast<void *>(m->d_rays), static_cast<void *>(m->d_hits), static_cast<void *>(m->d_best_key), static_cast<void *>(m->d_first_hit)})
      if (p) cudaFree(p);
    m->d_rays = nullptr; m->d_hits = nullptr; m->d_best_key = nullptr; m->d_first_hit = nullptr; m->ray_cap = 0;
    RDN_CUDA(cudaMalloc(&m->d_rays, n * sizeof(rdn_ray)));
    RDN_CUDA(cudaMalloc(&m->d_hits, n * sizeof(rdn_mesh_hit)));
    RDN_CUDA(cudaMalloc(&m->d_best_key, n * sizeof(unsigned long long)));
    RDN_CUDA(cudaMalloc(&m->d_first_hit, n * sizeof(uint32_t)));
    RDN_CUDA(cudaMemset(m->d_best_key, 0xFF, n * sizeof(unsigned long long)));
    RDN_CUDA(cudaMemset(m->d_first_hit, 0xFF, n * sizeof(uint32_t)));
    m->ray_cap = n;
  }
  RDN_CUDA(cudaMemcpy(m->d_rays, rays, n * sizeof(rdn_ray), cudaMemcpyHostToDevice));
  launch_pick_nearest(m->dev, m->d_rays, n, config->tolerance_local, config->triangle_face, m->d_best_key, m->d_first_hit, m->d_hits, m->sm_count, nullptr);
  RDN_CUDA(cudaGetLastError());
  RDN_CUDA(cudaMemcpy(out, m->d_hits, n * sizeof(rdn_mesh_hit), cudaMemcpyDeviceToHost));
  return RDN_OK;
}

int rdn_pick_mesh_all(rdn_pick_mesh *m, const rdn_pick_config *config, const rdn_ray *ray, rdn_mesh_hit *out, uint64_t capacity, uint64_t *out_total) {
  if (!m || !config || !ray || !out_total || (capacity && !out)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_pick_mesh_all: null argument");
  if (config->triangle_face > RDN_FACE_DOUBLE) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_pick_mesh_all: bad triangle_face");
  *out_total = 0;
  const uint64_t np = m->dev.n_prims;
  if (np == 0) return RDN_OK;
  std::lock_guard<std::mutex> lg(m->lock);
  RDN_CUDA(cudaSetDevice(m->device));
  if (!m->d_keep) {
    RDN_CUDA(cudaMalloc(&m->d_keep, np));
    RDN_CUDA(cudaMalloc(&m->d_iota, np * sizeof(uint32_t)));
    RDN_CUDA(cudaMalloc(&m->d_index, np * sizeof(uint32_t)));
    RDN_CUDA(cudaMalloc(&m->d_records, np * sizeof(rdn_mesh_hit)));
    RDN_CUDA(cudaMalloc(&m->d_n_kept, sizeof(uint64_t)));
    RDN_CUDA(cudaMalloc(&m->d_compact_status, compact_status_words(np) * sizeof(unsigned long long)));
  }
  if (m->ray_cap < 1) {
    RDN_CUDA(cudaMalloc(&m->d_rays, sizeof(rdn_ray)));
    RDN_CUDA(cudaMalloc(&m->d_hits, sizeof(rdn_mesh_hit)));
    RDN_CUDA(cudaMalloc(&m->d_best_key, sizeof(unsigned long long)));
    RDN_CUDA(cudaMalloc(&m->d_first_hit, sizeof(uint32_t)));
    RDN_CUDA(cudaMemset(m->d_best_key, 0xFF, sizeof(unsigned long long)));
    RDN_CUDA(cudaMemset(m->d_first_hit, 0xFF, sizeof(uint32_t)));
    m->ray_cap = 1;
  }
  RDN_CUDA(cudaMemcpy(m->d_rays, ray, sizeof(rdn_ray), cudaMemcpyHostToDevice));
  launch_pick_all_mark(m->dev, m->d_rays, config->tolerance_local, config->triangle_face, m->d_keep, m->d_iota, m->d_records, m->sm_count, nullptr);
  launch_compact_u32(m->d_iota, m->d_keep, np, m->d_index, m->d_n_kept, m->d_compact_status, nullptr);
  RDN_CUDA(cudaGetLastError());
  uint64_t total = 0;
  RDN_CUDA(cudaMemcpy(&total, m->d_n_kept, sizeof(total), cudaMemcpyDeviceToHost));
  *out_total = total;
  const uint64_t take = std::min(total, capacity);
  if (take) {
    if (m->gathered_cap < take) {
      if (m->d_gathered) cudaFree(m->d_gathered);
      m->d_gathered = nullptr; m->gathered_cap = 0;
      RDN_CUDA(cudaMalloc(&m->d_gathered, take * sizeof(rdn_mesh_hit)));
      m->gathered_cap = take;
    }
    launch_pick_all_gather(m->d_index, m->d_n_kept, take, m->d_records, m->d_gathered, m->sm_count, nullptr);
    RDN_CUDA(cudaGetLastError());
    RDN_CUDA(cudaMemcpy(out, m->d_gathered, take * sizeof(rdn_mesh_hit), cudaMemcpyDeviceToHost));
  }
  return RDN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// f4: shader binding table + dispatch (sbt.cu)
int rdn_sbt_create(rdn_rt_scene *scene, uint32_t max_geometry_count_in_blas, uint32_t max_tlas_offset, uint32_t ray_type_count, rdn_sbt **out) {
  if (!scene || !out) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_create: null argument");
  const uint64_t groups = static_cast<uint64_t>(max_geometry_count_in_blas) * max_tlas_offset * ray_type_count;
  if (groups > (1ull << 28)) return fail(RDN_ERR_CAPACITY, "rdn_sbt_create: more than 2^28 hit groups");
  rdn_sbt *t = new rdn_sbt;
  t->scene = scene;
  t->ray_stride = ray_type_count;
  t->max_geometry = max_geometry_count_in_blas;
  t->max_tlas_offset = max_tlas_offset;
  t->hit_groups.assign(groups, SbtHitGroup{RDN_SBT_NO_SHADER, RDN_SBT_NO_SHADER, RDN_SBT_NO_SHADER});
  t->miss.assign(ray_type_count, RDN_SBT_NO_SHADER);
  t->per_device.resize(scene->devices.size());
  for (const DeviceCtx &dc : scene->devices) t->device_ids.push_back(dc.device);
  *out = t;
  return RDN_OK;
}

void rdn_sbt_destroy(rdn_sbt *t) {
  if (!t) return;
  if (t->scene && t->scene->bound_sbt == t) t->scene->bound_sbt = nullptr;
  for (size_t i = 0; i < t->per_device.size(); ++i) {
    rdn_sbt::PerDevice &pd = t->per_device[i];
    cudaSetDevice(t->device_ids[i]);
    cudaFree(pd.d_hit_groups); cudaFree(pd.d_miss); cudaFree(pd.d_keep); cudaFree(pd.d_iota); cudaFree(pd.d_segment);
    cudaFree(pd.d_count); cudaFree(pd.d_status);
  }
  delete t;
}

int rdn_sbt_config_ray_generation(rdn_sbt *t, uint32_t shader) {
  if (!t) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_config_ray_generation: null table");
  std::lock_guard<std::mutex> lg(t->lock);
  t->ray_gen = shader;
  return RDN_OK;
}

int rdn_sbt_ray_generation(const rdn_sbt *t, uint32_t *out_shader) {
  if (!t || !out_shader) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_ray_generation: null argument");
  *out_shader = t->ray_gen;
  return RDN_OK;
}

int rdn_sbt_config_hit_group(rdn_sbt *t, uint32_t geometry_idx, uint32_t tlas_offset, uint32_t ray_ty_idx, uint32_t closest_hit,
                             uint32_t any_hit, uint32_t intersection) {
  if (!t) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_config_hit_group: null table");
  std::lock_guard<std::mutex> lg(t->lock);
  // the reference's set_value(...).unwrap() panics outside the allocated range
  const uint64_t idx = static_cast<uint64_t>(ray_ty_idx) + static_cast<uint64_t>(geometry_idx) * t->ray_stride + tlas_offset;
  if (idx >= t->hit_groups.size()) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_config_hit_group: record outside the table");
  if ((closest_hit != RDN_SBT_NO_SHADER && (closest_hit & RDN_TASK_MISS_BIT)))
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_config_hit_group: shader handles are below 2^31");
  t->hit_groups[idx] = SbtHitGroup{closest_hit, any_hit, intersection};
  sbt_touch(t);
  return RDN_OK;
}

int rdn_sbt_config_missing(rdn_sbt *t, uint32_t ray_ty_idx, uint32_t shader) {
  if (!t) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_config_missing: null table");
  std::lock_guard<std::mutex> lg(t->lock);
  if (ray_ty_idx >= t->miss.size()) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_config_missing: ray type outside the table");
  if (shader != RDN_SBT_NO_SHADER && (shader & RDN_TASK_MISS_BIT)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_sbt_config_missing: shader handles are below 2^31");
  t->miss[ray_ty_idx] = shader;
  sbt_touch(t);
  return RDN_OK;
}

int rdn_rt_sbt_dispatch_device(rdn_rt_scene *s, int device_index, rdn_sbt *t, const rdn_sbt_ray_config *config, const rdn_hit *d_hits,
                               uint64_t n, uint32_t *d_task, void *cuda_stream) {
  if (!s || !t || !config || (n && (!d_hits || !d_task))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_sbt_dispatch_device: null argument");
  if (t->scene != s) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_sbt_dispatch_device: the table belongs to another scene");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  if (reinterpret_cast<uintptr_t>(d_hits) & 15u) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_sbt_dispatch_device: hit array must be 16-byte aligned");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  std::lock_guard<std::mutex> lg(t->lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
  rc = sbt_upload(t, device_index);
  if (rc != RDN_OK) return rc;
  const rdn_sbt::PerDevice &pd = t->per_device[device_index];
  launch_sbt_dispatch(dc.dev, pd.d_hit_groups, static_cast<uint32_t>(t->hit_groups.size()), pd.d_miss, static_cast<uint32_t>(t->miss.size()),
                      *config, d_hits, n, d_task, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int rdn_rt_sbt_group_device(rdn_rt_scene *s, int device_index, rdn_sbt *t, const uint32_t *d_task, uint64_t n, uint32_t n_closest_shaders,
                            uint32_t n_miss_shaders, uint32_t *d_queue, uint64_t *d_offsets, void *cuda_stream) {
  if (!s || !t || !d_offsets || (n && (!d_task || !d_queue))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_sbt_group_device: null argument");
  if (t->scene != s) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_sbt_group_device: the table belongs to another scene");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  if (n > MAX_LAUNCH_RAYS) return fail(RDN_ERR_CAPACITY, "rdn_rt_sbt_group_device: more than 2^31 rays in one call");
  if (static_cast<uint64_t>(n_closest_shaders) + n_miss_shaders > 4096) return fail(RDN_ERR_CAPACITY, "rdn_rt_sbt_group_device: more than 4096 shaders");
  std::lock_guard<std::mutex> lg(t->lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  s->devices[device_index].ext_prev.valid = false;  // (something of ours now sits between two traces on that stream)
  rdn_sbt::PerDevice &pd = t->per_device[device_index];
  if (pd.cap < n || !pd.d_count) {
    cudaFree(pd.d_keep); cudaFree(pd.d_iota); cudaFree(pd.d_segment); cudaFree(pd.d_count); cudaFree(pd.d_status);
    pd.d_keep = nullptr; pd.d_iota = nullptr; pd.d_segment = nullptr; pd.d_count = nullptr; pd.d_status = nullptr; pd.cap = 0;
    const uint64_t m = std::max<uint64_t>(n, 1);
    RDN_CUDA(cudaMalloc(&pd.d_keep, m));
    RDN_CUDA(cudaMalloc(&pd.d_iota, m * sizeof(uint32_t)));
    RDN_CUDA(cudaMalloc(&pd.d_segment, m * sizeof(uint32_t)));
    RDN_CUDA(cudaMalloc(&pd.d_count, sizeof(uint64_t)));
    RDN_CUDA(cudaMalloc(&pd.d_status, compact_status_words(m) * sizeof(unsigned long long)));
    pd.cap = m;
  }
  launch_sbt_group(d_task, n, n_closest_shaders, n_miss_shaders, pd.d_keep, pd.d_iota, pd.d_segment, pd.d_count, pd.d_status, d_queue, d_offsets,
                   static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// a20 / f4: the wavefront executor in one call (include/rdn_rt.h "a20 / f4"; device glue in wavefront.cu)
static int trace_device_impl(rdn_rt_scene *s, int device_index, const rdn_launch *launch, const rdn_ray *d_rays, uint64_t n,
                             rdn_hit *d_hits, void *cuda_stream, int mode, rdn_trace_stats *stats, const unsigned long long *d_n);

int rdn_rt_trace_ray(rdn_rt_scene *s, int device_index, rdn_sbt *t, const rdn_trace_ray_desc *d, void *cuda_stream) {
  if (!s || !t || !d || !d->ray_generation || !d->round_launch || d->n_round_launch == 0)
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_ray: null argument");
  if (t->scene != s) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_ray: the table belongs to another scene");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index (host-only scene?)");
  if ((d->n_closest_hit && !d->closest_hit) || (d->n_miss && !d->miss) || d->n_closest_hit + static_cast<uint64_t>(d->n_miss) > 64)
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_ray: bad shader lists (at most 64 closest-hit + miss shaders)");
  const uint64_t n0 = static_cast<uint64_t>(d->width) * d->height;
  if (n0 == 0) return RDN_OK;
  if (n0 > MAX_LAUNCH_RAYS) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_ray: more than 2^31 launch indices");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  std::lock_guard<std::mutex> wl(s->wave_lock);
  DeviceCtx &dc = s->devices[device_index];
  RDN_CUDA(cudaSetDevice(dc.device));
  cudaStream_t stream = static_cast<cudaStream_t>(cuda_stream);
  const uint32_t rounds = d->execution_round_hint, buckets = d->n_closest_hit + d->n_miss;
  rc = ensure_wave_scratch(dc, n0, buckets, rounds + 1);
  if (rc != RDN_OK) return rc;
  auto &w = dc.wave;
  {
    std::lock_guard<std::mutex> tl(t->lock);
    rc = sbt_upload(t, device_index);
    if (rc != RDN_OK) return rc;
  }
  rdn_sbt *previous_sbt;
  {  // the executor's current table for the duration of the call (RDN_ANYHIT_FROM_SBT rounds read it)
    std::unique_lock<std::shared_mutex> wr(s->lock);
    previous_sbt = s->bound_sbt;
    s->bound_sbt = t;
  }
  struct Restore { rdn_rt_scene *s; rdn_sbt *prev; ~Restore() { std::unique_lock<std::shared_mutex> wr(s->lock); s->bound_sbt = prev; } } restore{s, previous_sbt};
  const rdn_sbt::PerDevice &pd = t->per_device[device_index];
  const uint32_t row_width = 2 + w.buckets;
  uint64_t *grid_size = w.counters, *wave_size[2] = {w.counters + 1, w.counters + 2}, *list_size = w.counters + 3;
  launch_store_u64(grid_size, n0, stream);

  rdn_wave wave{};
  wave.width = d->width; wave.height = d->height; wave.d_payload = d->d_payload;
  wave.d_next_rays = w.next; wave.d_spawn = w.spawn; wave.max_tasks = n0;
  auto record_row = [&](uint32_t round, const uint64_t *wave_n, const uint64_t *spawned, bool with_lists) -> int {
    std::vector<const uint64_t *> ptrs(row_width);
    ptrs[0] = wave_n; ptrs[1] = spawned;
    for (uint32_t b = 0; b < w.buckets; ++b) ptrs[2 + b] = (with_lists && b < buckets) ? list_size + b : nullptr;
    // (nullptr entries: the row keeps the zero it was cleared to)
    std::vector<const uint64_t *> live;
    std::vector<uint32_t> col;
    for (uint32_t c = 0; c < row_width; ++c) if (ptrs[c]) { live.push_back(ptrs[c]); col.push_back(c); }
    for (size_t k = 0; k < live.size(); ++k)
      RDN_CUDA(cudaMemcpyAsync(w.rows_dev + static_cast<size_t>(round) * row_width + col[k], live[k], sizeof(uint64_t), cudaMemcpyDeviceToDevice, stream));
    return RDN_OK;
  };
  RDN_CUDA(cudaMemsetAsync(w.rows_dev, 0, static_cast<size_t>(rounds + 1) * row_width * sizeof(uint64_t), stream));

  // ---- round 0: ray generation over the launch grid
  RDN_CUDA(cudaMemsetAsync(w.spawn, 0, n0, stream));
  wave.round = 0; wave.shader = t->ray_gen; wave.d_tasks = nullptr; wave.d_task_count = grid_size;
  wave.d_rays = nullptr; wave.d_hits = nullptr; wave.d_launch_index = nullptr;
  rc = d->ray_generation(d->ray_generation_user, &wave, cuda_stream);
  if (rc != 0) return fail(rc < 0 ? rc : RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_ray: the ray generation stage failed");
  int cur = 0;
  launch_wave_clip_spawn(w.spawn, grid_size, n0, w.iota, stream);
  launch_compact_u32(w.iota, w.spawn, n0, w.idx, wave_size[cur], w.status, stream);
  launch_wave_gather(w.next, nullptr, w.idx, wave_size[cur], n0, w.rays[cur], w.launch[cur], stream);
  rc = record_row(0, grid_size, wave_size[cur], false);
  if (rc != RDN_OK) return rc;

  // ---- rounds: trace, pick the stage of every ray, run the stages over their task lists, compact what they spawned
  for (uint32_t r = 1; r <= rounds; ++r) {
    rdn_launch L = d->round_launch[std::min(r - 1, d->n_round_launch - 1)];
    L.grid_width = 0;
    rc = trace_device_impl(s, device_index, &L, w.rays[cur], n0, w.hits, cuda_stream, RDN_TRACE_AUTO, nullptr,
                           reinterpret_cast<const unsigned long long *>(wave_size[cur]));
    if (rc != RDN_OK) return rc;
    dc.ext_prev.valid = false;
    rdn_sbt_ray_config cfg{};
    cfg.ray_flags = L.ray_flags; cfg.sbt_ray_offset = L.sbt_ray_offset; cfg.sbt_ray_stride = L.sbt_ray_stride; cfg.miss_index = L.miss_index;
    launch_sbt_dispatch_n(dc.dev, pd.d_hit_groups, static_cast<uint32_t>(t->hit_groups.size()), pd.d_miss, static_cast<uint32_t>(t->miss.size()), cfg,
                          w.hits, wave_size[cur], n0, w.task, stream);
    for (uint32_t b = 0; b < buckets; ++b) {
      const uint32_t code = b < d->n_closest_hit ? b : ((b - d->n_closest_hit) | RDN_TASK_MISS_BIT);
      launch_wave_mark(w.task, wave_size[cur], n0, code, w.keep, w.iota, stream);
      launch_compact_u32(w.iota, w.keep, n0, w.segments + static_cast<size_t>(b) * w.cap, list_size + b, w.status, stream);
    }
    RDN_CUDA(cudaMemsetAsync(w.spawn, 0, n0, stream));
    wave.round = r; wave.d_rays = w.rays[cur]; wave.d_hits = w.hits; wave.d_launch_index = w.launch[cur];
    for (uint32_t b = 0; b < buckets; ++b) {
      const bool closest = b < d->n_closest_hit;
      const uint32_t k = closest ? b : b - d->n_closest_hit;
      const rdn_stage_fn fn = closest ? d->closest_hit[k] : d->miss[k];
      if (!fn) continue;
      wave.shader = k; wave.d_tasks = w.segments + static_cast<size_t>(b) * w.cap; wave.d_task_count = list_size + b;
      void *user = closest ? (d->closest_hit_user ? d->closest_hit_user[k] : nullptr) : (d->miss_user ? d->miss_user[k] : nullptr);
      rc = fn(user, &wave, cuda_stream);
      if (rc != 0) return fail(rc < 0 ? rc : RDN_ERR_INVALID_ARGUMENT, "rdn_rt_trace_ray: a shader stage failed");
    }
    RDN_CUDA(cudaSetDevice(dc.device));  // (a stage may have switched devices)
    const int nxt = cur ^ 1;
    launch_wave_clip_spawn(w.spawn, wave_size[cur], n0, w.iota, stream);
    launch_compact_u32(w.iota, w.spawn, n0, w.idx, wave_size[nxt], w.status, stream);
    launch_wave_gather(w.next, w.launch[cur], w.idx, wave_size[nxt], n0, w.rays[nxt], w.launch[nxt], stream);
    rc = record_row(r, wave_size[cur], wave_size[nxt], true);
    if (rc != RDN_OK) return rc;
    cur = nxt;
  }
  RDN_CUDA(cudaGetLastError());
  if (d->counts && d->n_counts) {
    std::vector<uint64_t> rows(static_cast<size_t>(rounds + 1) * row_width);
    RDN_CUDA(cudaMemcpyAsync(rows.data(), w.rows_dev, rows.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost, stream));
    RDN_CUDA(cudaStreamSynchronize(stream));
    for (uint32_t r = 0; r <= rounds && r < d->n_counts; ++r) {
      const uint64_t *row = rows.data() + static_cast<size_t>(r) * row_width;
      rdn_wave_counts &c = d->counts[r];
      c.wave = row[0]; c.spawned = row[1]; c.closest_tasks = 0; c.miss_tasks = 0;
      for (uint32_t b = 0; b < buckets; ++b) (b < d->n_closest_hit ? c.closest_tasks : c.miss_tasks) += row[2 + b];
      c.no_task = r == 0 ? 0 : c.wave - c.closest_tasks - c.miss_tasks;
    }
  }
  return RDN_OK;
}

static int stage_device(rdn_rt_scene *s, int device_index, const rdn_wave *wave, const char *what) {
  if (!s || !wave) return fail(RDN_ERR_INVALID_ARGUMENT, std::string(what) + ": null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  RDN_CUDA(cudaSetDevice(s->devices[device_index].device));
  return RDN_OK;
}
int rdn_rt_stage_spawn_all(rdn_rt_scene *s, int device_index, const rdn_wave *wave, void *cuda_stream) {
  const int rc = stage_device(s, device_index, wave, "rdn_rt_stage_spawn_all");
  if (rc != RDN_OK) return rc;
  launch_stage_spawn_all(wave->d_spawn, wave->d_task_count, wave->max_tasks, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}
int rdn_rt_stage_bounce(rdn_rt_scene *s, int device_index, const rdn_bounce *p, const rdn_wave *wave, void *cuda_stream) {
  const int rc = stage_device(s, device_index, wave, "rdn_rt_stage_bounce");
  if (rc != RDN_OK) return rc;
  if (!p || !wave->d_rays || !wave->d_hits || !wave->d_tasks) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_stage_bounce: needs a traced wave (round >= 1)");
  if (p->mode > 2) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_stage_bounce: unknown mode");
  launch_stage_bounce_rays(s->devices[device_index].dev, *p, wave->d_rays, wave->d_hits, wave->d_tasks, wave->d_task_count, wave->max_tasks,
                           wave->d_launch_index, wave->d_next_rays, wave->d_spawn, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}
int rdn_rt_stage_store_f32(rdn_rt_scene *s, int device_index, const rdn_wave *wave, float value, float *d_dst, void *cuda_stream) {
  const int rc = stage_device(s, device_index, wave, "rdn_rt_stage_store_f32");
  if (rc != RDN_OK) return rc;
  if (!d_dst) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_stage_store_f32: null destination");
  launch_stage_store_f32(wave->d_tasks, wave->d_task_count, wave->max_tasks, wave->d_launch_index, value, d_dst, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}
int rdn_rt_ao_resolve_device(rdn_rt_scene *s, int device_index, float *d_payload, uint64_t n_pixels, uint32_t sample_count, uint32_t max_sample,
                             float *d_ao_buffer, void *cuda_stream) {
  if (!s || !d_payload || !d_ao_buffer) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_ao_resolve_device: null argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  RDN_CUDA(cudaSetDevice(s->devices[device_index].device));
  launch_ao_resolve(d_payload, n_pixels, sample_count, max_sample, d_ao_buffer, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int rdn_rt_sbt_dispatch(rdn_rt_scene *s, rdn_sbt *t, const rdn_sbt_ray_config *config, const rdn_hit *hits, uint64_t n,
                        uint32_t n_closest_shaders, uint32_t n_miss_shaders, uint32_t *task, uint32_t *queue, uint64_t *offsets) {
  if (!s || !t || !config || (n && !hits) || (queue && !offsets)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_sbt_dispatch: null argument");
  if (s->devices.empty()) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_sbt_dispatch: host-only scene has no CUDA device (there is no CPU fallback)");
  RDN_CUDA(cudaSetDevice(s->devices[0].device));
  const uint64_t m = std::max<uint64_t>(n, 1);
  const size_t n_groups = static_cast<size_t>(n_closest_shaders) + n_miss_shaders;
  rdn_hit *d_hits = nullptr;
  uint32_t *d_task = nullptr, *d_queue = nullptr;
  uint64_t *d_offsets = nullptr;
  int rc = RDN_OK;
  auto release = [&]() { cudaFree(d_hits); cudaFree(d_task); cudaFree(d_queue); cudaFree(d_offsets); };
  if (cudaMalloc(&d_hits, m * sizeof(rdn_hit)) != cudaSuccess || cudaMalloc(&d_task, m * 4) != cudaSuccess ||
      cudaMalloc(&d_queue, m * 4) != cudaSuccess || cudaMalloc(&d_offsets, (n_groups + 1) * 8) != cudaSuccess) {
    release();
    return fail(RDN_ERR_CUDA, "rdn_rt_sbt_dispatch: out of device memory");
  }
  if (cudaMemcpy(d_hits, hits, n * sizeof(rdn_hit), cudaMemcpyHostToDevice) != cudaSuccess) rc = fail(RDN_ERR_CUDA, "rdn_rt_sbt_dispatch: copy failed");
  if (rc == RDN_OK) rc = rdn_rt_sbt_dispatch_device(s, 0, t, config, d_hits, n, d_task, nullptr);
  if (rc == RDN_OK && offsets) rc = rdn_rt_sbt_group_device(s, 0, t, d_task, n, n_closest_shaders, n_miss_shaders, d_queue, d_offsets, nullptr);
  if (rc == RDN_OK && cudaDeviceSynchronize() != cudaSuccess) rc = fail(RDN_ERR_CUDA, "rdn_rt_sbt_dispatch: kernel failed");
  if (rc == RDN_OK && task && cudaMemcpy(task, d_task, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(RDN_ERR_CUDA, "rdn_rt_sbt_dispatch: copy failed");
  if (rc == RDN_OK && offsets) {
    if (cudaMemcpy(offsets, d_offsets, (n_groups + 1) * 8, cudaMemcpyDeviceToHost) != cudaSuccess) rc = fail(RDN_ERR_CUDA, "rdn_rt_sbt_dispatch: copy failed");
    if (rc == RDN_OK && queue && cudaMemcpy(queue, d_queue, offsets[n_groups] * 4, cudaMemcpyDeviceToHost) != cudaSuccess)
      rc = fail(RDN_ERR_CUDA, "rdn_rt_sbt_dispatch: copy failed");
  }
  release();
  return rc;
}

int rdn_rt_measure_l2_read_gbs(rdn_rt_scene *s, int device_index, uint64_t bytes, int passes, double *out_gbs) {
  if (!s || !out_gbs || bytes < (1u << 20) || passes < 1) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_measure_l2_read_gbs: bad argument");
  if (device_index < 0 || device_index >= static_cast<int>(s->devices.size())) return fail(RDN_ERR_INVALID_ARGUMENT, "bad device_index");
  std::lock_guard<std::mutex> lg(s->launch_lock);
  RDN_CUDA(cudaSetDevice(s->devices[device_index].device));
  if (measure_l2_read_gbs(bytes, passes, s->devices[device_index].sm_count, out_gbs) != 0) return fail(RDN_ERR_CUDA, "rdn_rt_measure_l2_read_gbs: CUDA error");
  return RDN_OK;
}

int rdn_rt_scene_build_stats(rdn_rt_scene *s, rdn_build_stats *out) {
  if (!s || !out) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_rt_scene_build_stats: null argument");
  int rc = ensure_committed(s);
  if (rc != RDN_OK) return rc;
  std::shared_lock<std::shared_mutex> rd(s->lock);
  std::memset(out, 0, sizeof(*out));
  if (!s->adopted) {
    out->balance_fallbacks = s->flat.stats.balance_fallbacks;
    out->balance_fallbacks_gt10 = s->flat.stats.balance_fallbacks_gt10;
    out->irregular_triangles = s->flat.stats.irregular_triangles;
    out->irregular_instances = s->flat.stats.irregular_instances;
    out->reference_routed_tlas = s->flat.stats.reference_routed_tlas;
    out->bvh_build_ms = s->flat.stats.bvh_build_ms;
    out->flatten_ms = s->flat.stats.flatten_ms;
    out->upload_ms = s->flat.stats.upload_ms;
    out->build_threads = s->flat.stats.build_threads;
    out->device_built_trees = s->flat.stats.device_built_trees;
    out->tlas_only_commits = s->tlas_only_commits;
  } else {
    for (const TlasRoot &t : s->h_tlas_root) {
      if (t.irregular_count == IRREGULAR_ROUTE_ALL) out->reference_routed_tlas++;
      else out->irregular_instances += t.irregular_count;
    }
  }
  out->kernels_enqueued = s->kernels_enqueued.load();
  return RDN_OK;
}

// ------------------------------------------------------------------------------------------------ space query (path A)
struct rdn_flat_bvh {
  FlattenBVH bvh;
  uint64_t depth = 0;
  // device-resident copy made by rdn_bvh_upload: nodes + triangles pre-gathered in sorted_primitive_index order
  int device = -1;
  bool built_on_device = false;
  PathANode *d_nodes = nullptr;
  PathATri *d_tris = nullptr;
  std::mutex lock;
};

static uint64_t tree_depth(const BigVector<FlattenBVHNode> &nodes) {
  // pre-order walk with an explicit stack of (node, depth)
  uint64_t best = 0;
  std::vector<std::pair<uint64_t, uint64_t>> st;
  if (!nodes.empty()) st.emplace_back(0, 1);
  while (!st.empty()) {
    auto [i, d] = st.back();
    st.pop_back();
    best = std::max(best, d);
    if (nodes[i].has_child) { st.emplace_back(nodes[i].left_child_offset(), d + 1); st.emplace_back(nodes[i].right_child_offset(), d + 1); }
  }
  return best;
}

static int build_bvh_common(const Box3 *boxes, uint64_t n, int strategy, uint32_t buckets, const rdn_tree_build_option *option,
                            rdn_flat_bvh **out) {
  TreeBuildOption opt;
  if (option) { opt.max_tree_depth = option->max_tree_depth; opt.bin_size = option->bin_size; }
  auto *r = new rdn_flat_bvh();
  if (strategy == RDN_BVH_SAH) {
    SAH sah(buckets ? buckets : 4);
    r->bvh = FlattenBVH::build(boxes, n, sah, opt);
  } else if (strategy == RDN_BVH_BALANCE_TREE) {
    BalanceTree bt;
    r->bvh = FlattenBVH::build(boxes, n, bt, opt);
  } else {
    delete r;
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_build: unknown strategy");
  }
  if (r->bvh.stats.bucket_out_of_range) { delete r; return fail(RDN_ERR_BUILD, "SAH bucket index out of range (the reference panics here)"); }
  r->depth = tree_depth(r->bvh.nodes);
  *out = r;
  return RDN_OK;
}

int rdn_bvh_build(const float *boxes6, uint64_t n, int strategy, uint32_t sah_buckets, const rdn_tree_build_option *option, rdn_flat_bvh **out) {
  if (!out || (n && !boxes6)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_build: null argument");
  static_assert(sizeof(Box3) == 24, "Box3 layout");
  return build_bvh_common(reinterpret_cast<const Box3 *>(boxes6), n, strategy, sah_buckets, option, out);
}

int rdn_bvh_build_device(const float *boxes6, uint64_t n, uint32_t sah_buckets, const rdn_tree_build_option *option, int device,
                         rdn_flat_bvh **out) {
  if (!out || (n && !boxes6)) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_build_device: null argument");
  int available = 0;
  RDN_CUDA(cudaGetDeviceCount(&available));
  if (device < 0 || device >= available) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_build_device: no such CUDA device");
  TreeBuildOption opt;
  if (option) { opt.max_tree_depth = option->max_tree_depth; opt.bin_size = option->bin_size; }
  auto *r = new rdn_flat_bvh();
  std::string err;
  const int rc = build_bvh_sah_device(reinterpret_cast<const Box3 *>(boxes6), n, sah_buckets ? sah_buckets : 4, opt, device, r->bvh, err);
  if (rc == 0) {
    r->built_on_device = true;
    r->depth = tree_depth(r->bvh.nodes);
    *out = r;
    return RDN_OK;
  }
  delete r;
  if (rc == -1) return fail(RDN_ERR_CUDA, "rdn_bvh_build_device: " + err);
  if (rc == -2) return fail(RDN_ERR_BUILD, "SAH bucket index out of range (the reference panics here)");
  return build_bvh_common(reinterpret_cast<const Box3 *>(boxes6), n, RDN_BVH_SAH, sah_buckets, option, out);  // not covered on the device
}

int rdn_bvh_built_on_device(const rdn_flat_bvh *b) { return b && b->built_on_device ? 1 : 0; }

int rdn_bvh_build_for_mesh(const rdn_mesh_view *mesh, int strategy, uint32_t sah_buckets, const rdn_tree_build_option *option, rdn_flat_bvh **out) {
  if (!out || !mesh || !mesh->positions || !mesh->indices) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_build_for_mesh: null argument");
  const uint64_t n_tri = mesh->n_indices / 3;
  std::vector<Box3> boxes(n_tri);
  for (uint64_t t = 0; t < n_tri; ++t) {
    Box3 b = box_empty();  // Triangle::to_bounding, bounding_impl.rs:3-12
    for (int k = 0; k < 3; ++k) {
      const uint64_t vi = mesh->indices[3 * t + k];
      if (vi >= mesh->n_positions) return fail(RDN_ERR_BUILD, "triangle index out of bounds");
      expand(b, Vec3{mesh->positions[3 * vi], mesh->positions[3 * vi + 1], mesh->positions[3 * vi + 2]});
    }
    boxes[t] = b;
  }
  return build_bvh_common(boxes.data(), n_tri, strategy, sah_buckets, option, out);
}

static void free_bvh_device(rdn_flat_bvh *b) {
  if (b->device >= 0) {
    cudaSetDevice(b->device);
    if (b->d_nodes) cudaFree(b->d_nodes);
    if (b->d_tris) cudaFree(b->d_tris);
  }
  b->d_nodes = nullptr; b->d_tris = nullptr; b->device = -1;
}

void rdn_bvh_destroy(rdn_flat_bvh *b) {
  if (!b) return;
  free_bvh_device(b);
  delete b;
}

int rdn_bvh_upload(rdn_flat_bvh *b, const rdn_mesh_view *mesh, int device) {
  if (!b || !mesh || !mesh->positions || !mesh->indices) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_upload: null argument");
  if (b->depth > static_cast<uint64_t>(PATHA_MAX_DEPTH)) return fail(RDN_ERR_CAPACITY, "rdn_bvh_upload: tree deeper than 128 levels");
  const auto &nodes = b->bvh.nodes;
  const auto &sorted = b->bvh.sorted_primitive_index;
  if (sorted.size() * 3 > mesh->n_indices) return fail(RDN_ERR_INVALID_ARGUMENT, "mesh smaller than the BVH it was built for");
  if (nodes.size() >= 0xFFFFFFFFull || sorted.size() >= 0xFFFFFFFFull) return fail(RDN_ERR_CAPACITY, "rdn_bvh_upload: more than 2^32 nodes / primitives");
  std::vector<PathANode> pn(nodes.size());
  for (size_t i = 0; i < nodes.size(); ++i) {
    const FlattenBVHNode &nd = nodes[i];
    PathANode &p = pn[i];
    p.bmin[0] = nd.bounding.min.x; p.bmin[1] = nd.bounding.min.y; p.bmin[2] = nd.bounding.min.z;
    p.bmax[0] = nd.bounding.max.x; p.bmax[1] = nd.bounding.max.y; p.bmax[2] = nd.bounding.max.z;
    if (nd.has_child) { p.a = static_cast<uint32_t>(nd.right_child_offset()); p.b = 0xFFFFFFFFu; }
    else { p.a = static_cast<uint32_t>(nd.primitive_start); p.b = static_cast<uint32_t>(nd.primitive_end); }
  }
  std::vector<PathATri> pt(sorted.size());
  for (size_t k = 0; k < sorted.size(); ++k) {
    const uint64_t prim = sorted[k];
    PathATri &t = pt[k];
    std::memset(&t, 0, sizeof(t));
    float *dst[3] = {t.a, t.b, t.c};
    for (int v = 0; v < 3; ++v) {
      const uint64_t vi = mesh->indices[3 * prim + v];
      if (vi >= mesh->n_positions) return fail(RDN_ERR_BUILD, "triangle index out of bounds");
      std::memcpy(dst[v], mesh->positions + 3 * vi, 3 * sizeof(float));
    }
    t.prim = static_cast<uint32_t>(prim);
  }
  std::lock_guard<std::mutex> lg(b->lock);
  free_bvh_device(b);
  RDN_CUDA(cudaSetDevice(device));
  b->device = device;
  RDN_CUDA(cudaMalloc(&b->d_nodes, std::max<size_t>(pn.size(), 1) * sizeof(PathANode)));
  RDN_CUDA(cudaMalloc(&b->d_tris, std::max<size_t>(pt.size(), 1) * sizeof(PathATri)));
  RDN_CUDA(cudaMemcpy(b->d_nodes, pn.data(), pn.size() * sizeof(PathANode), cudaMemcpyHostToDevice));
  RDN_CUDA(cudaMemcpy(b->d_tris, pt.data(), pt.size() * sizeof(PathATri), cudaMemcpyHostToDevice));
  return RDN_OK;
}

int rdn_bvh_query_nearest_device(const rdn_flat_bvh *b, const rdn_ray *d_rays, uint64_t n, uint32_t face_side, rdn_mesh_hit *d_out,
                                 void *cuda_stream) {
  if (!b || (n && (!d_rays || !d_out))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_query_nearest_device: null argument");
  if (face_side > RDN_FACE_DOUBLE) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_query_nearest_device: bad face_side");
  if (b->device < 0 || !b->d_nodes) return fail(RDN_ERR_NOT_COMMITTED, "rdn_bvh_query_nearest_device: call rdn_bvh_upload first");
  RDN_CUDA(cudaSetDevice(b->device));
  if (b->bvh.nodes.empty()) {  // empty tree: every query is OptionalNearest::none()
    RDN_CUDA(cudaMemsetAsync(d_out, 0, n * sizeof(rdn_mesh_hit), static_cast<cudaStream_t>(cuda_stream)));
    return RDN_OK;
  }
  launch_patha_nearest(b->d_nodes, b->d_tris, d_rays, n, face_side, d_out, static_cast<cudaStream_t>(cuda_stream));
  RDN_CUDA(cudaGetLastError());
  return RDN_OK;
}

int rdn_bvh_nodes(const rdn_flat_bvh *b, const rdn_flat_bvh_node **out_nodes, uint64_t *out_n) {
  if (!b || !out_nodes || !out_n) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_nodes: null argument");
  static_assert(sizeof(FlattenBVHNode) == sizeof(rdn_flat_bvh_node), "FlattenBVHNode layout");
  *out_nodes = reinterpret_cast<const rdn_flat_bvh_node *>(b->bvh.nodes.data());
  *out_n = b->bvh.nodes.size();
  return RDN_OK;
}

int rdn_bvh_sorted_primitive_index(const rdn_flat_bvh *b, const uint64_t **out_index, uint64_t *out_n) {
  if (!b || !out_index || !out_n) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_sorted_primitive_index: null argument");
  *out_index = b->bvh.sorted_primitive_index.data();
  *out_n = b->bvh.sorted_primitive_index.size();
  return RDN_OK;
}

int rdn_bvh_query_nearest(const rdn_flat_bvh *b, const rdn_mesh_view *mesh, const rdn_ray *rays, uint64_t n, uint32_t face_side,
                          int device, rdn_mesh_hit *out) {
  if (!b || !mesh || !mesh->positions || !mesh->indices || (n && (!rays || !out))) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_query_nearest: null argument");
  if (face_side > RDN_FACE_DOUBLE) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_query_nearest: bad face_side");
  // the mesh is borrowed per call (as in intersect_nearest_bvh(mesh, ray, bvh, conf)): gather + upload it, then run the resident query
  int rc = rdn_bvh_upload(const_cast<rdn_flat_bvh *>(b), mesh, device);
  if (rc != RDN_OK) return rc;
  rdn_ray *d_rays = nullptr; rdn_mesh_hit *d_out = nullptr;
  auto cleanup = [&]() { cudaFree(d_rays); cudaFree(d_out); };
#define RDN_CUDA_C(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { cleanup(); return fail(RDN_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); } } while (0)
  RDN_CUDA_C(cudaMalloc(&d_rays, std::max<uint64_t>(n, 1) * sizeof(rdn_ray)));
  RDN_CUDA_C(cudaMalloc(&d_out, std::max<uint64_t>(n, 1) * sizeof(rdn_mesh_hit)));
  RDN_CUDA_C(cudaMemcpy(d_rays, rays, n * sizeof(rdn_ray), cudaMemcpyHostToDevice));
  rc = rdn_bvh_query_nearest_device(b, d_rays, n, face_side, d_out, nullptr);
  if (rc != RDN_OK) { cleanup(); return rc; }
  RDN_CUDA_C(cudaDeviceSynchronize());
  RDN_CUDA_C(cudaMemcpy(out, d_out, n * sizeof(rdn_mesh_hit), cudaMemcpyDeviceToHost));
  cleanup();
  return RDN_OK;
}

int rdn_bvh_query_list(const rdn_flat_bvh *b, const rdn_mesh_view *mesh, const rdn_ray *rays, uint64_t n, uint32_t face_side, int device,
                       uint64_t *out_offsets, rdn_mesh_hit *out_hits, uint64_t capacity, uint64_t *out_total) {
  if (!b || !mesh || !mesh->positions || !mesh->indices || !out_offsets || !out_total || (n && !rays))
    return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_query_list: null argument");
  if (face_side > RDN_FACE_DOUBLE) return fail(RDN_ERR_INVALID_ARGUMENT, "rdn_bvh_query_list: bad face_side");
  if (b->bvh.nodes.empty()) {  // an empty tree has no hits (and its device arrays hold nothing to walk): all offsets zero
    for (uint64_t i = 0; i <= n; ++i) out_offsets[i] = 0;
    *out_total = 0;
    return RDN_OK;
  }
  int rc = rdn_bvh_upload(const_cast<rdn_flat_bvh *>(b), mesh, device);
  if (rc != RDN_OK) return rc;
  rdn_ray *d_rays = nullptr; uint32_t *d_counts = nullptr; uint64_t *d_offsets = nullptr; rdn_mesh_hit *d_out = nullptr;
  auto cleanup = [&]() { cudaFree(d_rays); cudaFree(d_counts); cudaFree(d_offsets); cudaFree(d_out); };
  RDN_CUDA_C(cudaMalloc(&d_rays, std::max<uint64_t>(n, 1) * sizeof(rdn_ray)));
  RDN_CUDA_C(cudaMalloc(&d_counts, std::max<uint64_t>(n, 1) * sizeof(uint32_t)));
  RDN_CUDA_C(cudaMemcpy(d_rays, rays, n * sizeof(rdn_ray), cudaMemcpyHostToDevice));
  launch_patha_list(b->d_nodes, b->d_tris, d_rays, n, face_side, d_counts, nullptr, nullptr, nullptr);
  RDN_CUDA_C(cudaGetLastError());
  std::vector<uint32_t> counts(n);
  RDN_CUDA_C(cudaMemcpy(counts.data(), d_counts, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  uint64_t total = 0;  // exclusive scan on the host: the call is host-buffer bound anyway
  for (uint64_t i = 0; i < n; ++i) { out_offsets[i] = total; total += counts[i]; }
  out_offsets[n] = total;
  *out_total = total;
  if (!out_hits || capacity < total || total == 0) { cleanup(); return RDN_OK; }  // size query, or nothing to write
  RDN_CUDA_C(cudaMalloc(&d_offsets, (n + 1) * sizeof(uint64_t)));
  RDN_CUDA_C(cudaMalloc(&d_out, total * sizeof(rdn_mesh_hit)));
  RDN_CUDA_C(cudaMemcpy(d_offsets, out_offsets, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice));
  launch_patha_list(b->d_nodes, b->d_tris, d_rays, n, face_side, nullptr, d_offsets, d_out, nullptr);
  RDN_CUDA_C(cudaGetLastError());
  RDN_CUDA_C(cudaDeviceSynchronize());
  RDN_CUDA_C(cudaMemcpy(out_hits, d_out, total * sizeof(rdn_mesh_hit), cudaMemcpyDeviceToHost));
  cleanup();
  return RDN_OK;
}

}  // extern "C"
