/* ORACLE — TEST INFRASTRUCTURE ONLY.  FlattenBVH builder restatement.
 *
 * Follows /root/reference:
 *   content/space/src/bvh/mod.rs:55-79        FlattenBVH::new
 *   content/space/src/bvh/node.rs:30-53       pre-order child offsets (left = self+1, right = self+left_count+1)
 *   content/space/src/bvh/strategy.rs:11-43   BVHBuildStrategy::build (recursive, pre-order push)
 *   content/space/src/bvh/strategy.rs:67-86   BalanceTree::split
 *   content/space/src/bvh/strategy.rs:202-284 SAH::split
 *   content/space/src/bvh/apply.rs:19-49      median_partition_at_axis (select_nth_unstable_by)
 *   content/space/src/utils.rs:20-66          TreeBuildOption, BuildPrimitive, bounding_from_build_source
 *   shader/ray-tracing/.../geometry/naive/mod.rs:612-632  compute_bvh_next
 *
 * One deliberate approximation: Rust's select_nth_unstable_by leaves an implementation-defined
 * permutation.  For slices of <= 10 elements std uses a stable insertion sort, which the stable
 * sort below reproduces; larger fallbacks are counted in balance_fallbacks_gt10 so parity tests
 * can assert they never happen on the benchmark scenes.
 */
#include <stdlib.h>
#include <string.h>
#include "oracle.h"

typedef struct {
  const obox *boxes;   /* BuildPrimitive.bounding */
  ov3 *centers;        /* BuildPrimitive.center   */
  orc_bvh *out;
  int strategy;
  uint32_t n_buckets;
  uint64_t max_tree_depth, bin_size;
  /* scratch */
  uint32_t *bucket_of;   /* per position in the current range */
  uint64_t *tmp_index;
  float *tmp_key;
} build_ctx;

static void push_node(orc_bvh *b, obox bounding, uint64_t start, uint64_t end) {
  if (b->n_nodes == b->cap_nodes) {
    b->cap_nodes = b->cap_nodes ? b->cap_nodes * 2 : 64;
    b->nodes = (orc_bvh_node *)realloc(b->nodes, b->cap_nodes * sizeof(orc_bvh_node));
  }
  orc_bvh_node *n = &b->nodes[b->n_nodes];
  n->bounding = bounding;
  n->start = start;
  n->end = end;
  n->self_index = b->n_nodes;
  n->left_count = 0;
  n->has_child = 0;
  n->split_axis = 0;
  b->n_nodes++;
}

static obox bounding_from_build_source(const build_ctx *c, uint64_t start, uint64_t end) {
  obox r = obox_empty();
  for (uint64_t i = start; i < end; i++) obox_expand_box(&r, c->boxes[c->out->sorted_primitive_index[i]]);
  return r;
}

static inline float axis_of(ov3 v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

/* stable merge sort of idx[0..n) by key (partial_cmp(..).unwrap_or(Less) has no NaNs on finite meshes) */
static void stable_sort_by_key(uint64_t *idx, float *key, uint64_t n, uint64_t *tmp_i, float *tmp_k) {
  if (n <= 16) {
    for (uint64_t i = 1; i < n; i++) {
      uint64_t vi = idx[i]; float vk = key[i];
      uint64_t j = i;
      while (j > 0 && vk < key[j - 1]) { idx[j] = idx[j - 1]; key[j] = key[j - 1]; j--; }
      idx[j] = vi; key[j] = vk;
    }
    return;
  }
  uint64_t h = n / 2;
  stable_sort_by_key(idx, key, h, tmp_i, tmp_k);
  stable_sort_by_key(idx + h, key + h, n - h, tmp_i, tmp_k);
  uint64_t a = 0, b = h, o = 0;
  while (a < h && b < n) {
    if (key[b] < key[a]) { tmp_i[o] = idx[b]; tmp_k[o] = key[b]; b++; }
    else { tmp_i[o] = idx[a]; tmp_k[o] = key[a]; a++; }
    o++;
  }
  while (a < h) { tmp_i[o] = idx[a]; tmp_k[o] = key[a]; a++; o++; }
  while (b < n) { tmp_i[o] = idx[b]; tmp_k[o] = key[b]; b++; o++; }
  memcpy(idx, tmp_i, n * sizeof(uint64_t));
  memcpy(key, tmp_k, n * sizeof(float));
}

typedef struct { obox lbox; uint64_t lstart, lend; int axis; obox rbox; uint64_t rstart, rend; } split_result;

static split_result balance_split(build_ctx *c, const orc_bvh_node *parent) {
  split_result r;
  uint64_t *index = c->out->sorted_primitive_index;
  int axis = obox_longest_axis(parent->bounding);
  uint64_t start = parent->start, end = parent->end;
  uint64_t middle = (end + start) / 2;
  uint64_t range_middle = (end - start) / 2;
  if (range_middle != 0) {
    uint64_t n = end - start;
    for (uint64_t i = 0; i < n; i++) c->tmp_key[i] = axis_of(c->centers[index[start + i]], axis);
    stable_sort_by_key(index + start, c->tmp_key, n, c->tmp_index, c->tmp_key + n);
  }
  r.axis = axis;
  r.lstart = start; r.lend = middle; r.rstart = middle; r.rend = end;
  r.lbox = bounding_from_build_source(c, start, middle);
  r.rbox = bounding_from_build_source(c, middle, end);
  return r;
}

/* Rust `f32 as usize`: saturating, NaN -> 0 */
static inline uint64_t f32_as_usize(float v) {
  if (!(v == v)) return 0;
  if (v <= 0.0f) return 0;
  if (v >= 18446744073709551616.0f) return UINT64_MAX;
  return (uint64_t)v;
}

static split_result sah_split(build_ctx *c, const orc_bvh_node *parent) {
  orc_bvh *b = c->out;
  uint64_t *index = b->sorted_primitive_index;
  const uint32_t nb = c->n_buckets;
  uint64_t start = parent->start, end = parent->end;

  obox bucket_box[64];
  uint64_t bucket_cnt[64];
  for (uint32_t k = 0; k < nb; k++) { bucket_box[k] = obox_empty(); bucket_cnt[k] = 0; }

  int axis = obox_longest_axis(parent->bounding);
  float axis_start = axis_of(parent->bounding.min, axis);
  float axis_end = axis_of(parent->bounding.max, axis);
  float step = (axis_end - axis_start) / (float)nb;

  for (uint64_t i = start; i < end; i++) {
    uint64_t prim = index[i];
    float axis_value = axis_of(c->centers[prim], axis);
    uint64_t which = f32_as_usize(floorf((axis_value - axis_start) / step));
    if (which == nb) which -= 1;
    if (which >= nb) { b->error = 1; which = nb - 1; } /* the reference would panic: index out of bounds */
    obox_expand_box(&bucket_box[which], c->boxes[prim]);
    bucket_cnt[which]++;
    c->bucket_of[i - start] = (uint32_t)which;
  }

  uint32_t empty = 0;
  for (uint32_t k = 0; k < nb; k++) empty += (bucket_cnt[k] == 0);
  if (empty == nb - 1) {
    b->balance_fallbacks++;
    if (end - start > 10) b->balance_fallbacks_gt10++;
    return balance_split(c, parent);
  }

  /* step 2: evaluate the nb-1 prefix splits, first strict minimum wins */
  uint32_t best = 0;
  float best_cost = INFINITY;
  obox best_l = obox_empty(), best_r = obox_empty();
  uint64_t best_nl = 0;
  for (uint32_t i = 0; i + 1 < nb; i++) {
    obox l = obox_empty(), r = obox_empty();
    uint64_t nl = 0, nr = 0;
    for (uint32_t k = 0; k <= i; k++) { obox_expand_box(&l, bucket_box[k]); nl += bucket_cnt[k]; }
    for (uint32_t k = i + 1; k < nb; k++) { obox_expand_box(&r, bucket_box[k]); nr += bucket_cnt[k]; }
    float cost = obox_surface_area(l) * (float)nl + obox_surface_area(r) * (float)nr;
    if (i == 0) { best_l = l; best_r = r; best_nl = nl; } /* partition_decision[0] is the default pick */
    if (cost < best_cost) { best_cost = cost; best = i; best_l = l; best_r = r; best_nl = nl; }
  }
  (void)best;

  /* step 3: stable bucket-by-bucket rewrite of the index range */
  uint64_t offs[64];
  uint64_t acc = 0;
  for (uint32_t k = 0; k < nb; k++) { offs[k] = acc; acc += bucket_cnt[k]; }
  uint64_t n = end - start;
  for (uint64_t i = 0; i < n; i++) c->tmp_index[offs[c->bucket_of[i]]++] = index[start + i];
  memcpy(index + start, c->tmp_index, n * sizeof(uint64_t));

  split_result r;
  r.axis = axis;
  r.lbox = best_l; r.lstart = start; r.lend = start + best_nl;
  r.rbox = best_r; r.rstart = start + best_nl; r.rend = end;
  return r;
}

static uint64_t build_rec(build_ctx *c, uint64_t depth) {
  orc_bvh *b = c->out;
  uint64_t node_index = b->n_nodes - 1;
  orc_bvh_node node = b->nodes[node_index];
  uint64_t count = node.end - node.start;
  if (!(depth < c->max_tree_depth && count > c->bin_size)) return 1;

  split_result s = (c->strategy == ORC_STRATEGY_SAH) ? sah_split(c, &node) : balance_split(c, &node);

  push_node(b, s.lbox, s.lstart, s.lend);
  uint64_t left_count = build_rec(c, depth + 1);
  push_node(b, s.rbox, s.rstart, s.rend);
  uint64_t right_count = build_rec(c, depth + 1);

  b->nodes[node_index].has_child = 1;
  b->nodes[node_index].left_count = left_count;
  b->nodes[node_index].split_axis = s.axis;
  return left_count + right_count + 1;
}

orc_bvh *orc_bvh_build(const obox *boxes, uint64_t n, int strategy, uint32_t sah_buckets,
                       uint64_t max_tree_depth, uint64_t bin_size) {
  orc_bvh *b = (orc_bvh *)calloc(1, sizeof(orc_bvh));
  if (sah_buckets < 2) sah_buckets = 2;
  if (sah_buckets > 64) sah_buckets = 64;
  b->n_prims = n;
  b->sorted_primitive_index = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
  build_ctx c;
  memset(&c, 0, sizeof(c));
  c.boxes = boxes;
  c.centers = (ov3 *)malloc((n ? n : 1) * sizeof(ov3));
  c.out = b;
  c.strategy = strategy;
  c.n_buckets = sah_buckets;
  c.max_tree_depth = max_tree_depth;
  c.bin_size = bin_size;
  c.bucket_of = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
  c.tmp_index = (uint64_t *)malloc((n ? n : 1) * sizeof(uint64_t));
  c.tmp_key = (float *)malloc((n ? n : 1) * 2 * sizeof(float));
  for (uint64_t i = 0; i < n; i++) {
    b->sorted_primitive_index[i] = i;
    c.centers[i] = obox_center(boxes[i]);
  }
  obox root = bounding_from_build_source(&c, 0, n);
  push_node(b, root, 0, n);
  build_rec(&c, 0);
  free(c.centers); free(c.bucket_of); free(c.tmp_index); free(c.tmp_key);
  return b;
}

void orc_bvh_free(orc_bvh *b) {
  if (!b) return;
  free(b->nodes);
  free(b->sorted_primitive_index);
  free(b);
}

void orc_bvh_compute_next(const orc_bvh *b, uint32_t *out) {
  uint32_t *stack = (uint32_t *)malloc((b->n_nodes + 1) * sizeof(uint32_t));
  uint64_t sp = 0;
  for (uint64_t i = 0; i < b->n_nodes; i++) {
    const orc_bvh_node *n = &b->nodes[i];
    if (sp > 0 && stack[sp - 1] == (uint32_t)n->self_index) sp--;
    uint32_t miss = sp > 0 ? stack[sp - 1] : ORC_INVALID_NEXT;
    uint32_t hit;
    if (n->has_child) {
      hit = (uint32_t)(n->self_index + 1);
      stack[sp++] = (uint32_t)(n->self_index + n->left_count + 1);
    } else {
      hit = miss;
    }
    out[2 * i] = hit;
    out[2 * i + 1] = miss;
  }
  free(stack);
}
