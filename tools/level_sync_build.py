"""Design prototype for SURVEY.md §8f row f3 (GPU BVH build): the reference's recursive SAH builder
(content/space/src/bvh/strategy.rs:11-43,202-284; BalanceTree fallback :67-86, apply.rs:19-49) restated as LEVEL-SYNCHRONOUS
data-parallel passes over flat arrays — the shape a CUDA build takes (one kernel per pass, all nodes of a level at once) — and
checked node for node against the sequential build (tests/test_level_sync_build.py).  numpy stands in for the kernels:

  per level, for the nodes that still split (TreeBuildOption::should_continue):
    1. per node      axis = longest axis of the node box, range start, bucket width            (tiny kernel, one thread per node)
    2. per primitive bucket = saturating floor((centre - start) / width), clamped to the last  (one thread per primitive position)
    3. per (node, bucket)  primitive count and box union                                        (segmented reduction; min / max are exact in any order)
    4. per node      all primitives in one bucket -> median split by a stable sort of the centres (the reference's BalanceTree fallback);
                     else the n_bucket-1 prefix costs area(L)*nL + area(R)*nR in f32, first strict minimum   (one thread per node)
    5. per primitive new position = node start + primitives of the node in earlier buckets + rank among the node's primitives of
                     the same bucket at earlier positions                                        (segmented exclusive scan per bucket: STABLE, as the reference's bucket-by-bucket rewrite)
    6. per node      two children (box, range); the next level's list
  afterwards: subtree sizes bottom-up, pre-order index top-down (left = parent + 1, right = parent + 1 + size(left)): the FlattenBVH node array.

Every f32 expression is evaluated in the reference's order (np.float32 scalars / arrays, no fused operations)."""
from __future__ import annotations

import numpy as np

f32 = np.float32


def _longest_axis(bmin, bmax):
    """box3.rs:117-133: the exact `>` cascade (vectorised over nodes)"""
    l = (bmax - bmin).astype(f32)
    x, y, z = l[:, 0], l[:, 1], l[:, 2]
    return np.where(x > y, np.where(x > z, 0, 2), np.where(y > z, 1, 2))


def _area(bmin, bmax):
    """box3.rs:5-11, f32 in the written order; an empty union (+inf, -inf) gives NaN / -inf like the reference's"""
    with np.errstate(all="ignore"):
        w = (bmax[..., 0] - bmin[..., 0]).astype(f32); h = (bmax[..., 1] - bmin[..., 1]).astype(f32); d = (bmax[..., 2] - bmin[..., 2]).astype(f32)
        return (f32(2.0) * (((w * h).astype(f32) + (w * d).astype(f32)).astype(f32) + (h * d).astype(f32)).astype(f32)).astype(f32)


def build(boxes: np.ndarray, n_buckets: int = 4, max_tree_depth: int = 50, bin_size: int = 2):
    """boxes[n, 6] (min xyz, max xyz) -> (nodes in pre-order as a dict of arrays, sorted_primitive_index)"""
    boxes = np.ascontiguousarray(boxes, f32).reshape(-1, 6)
    n = boxes.shape[0]
    pmin, pmax = boxes[:, :3], boxes[:, 3:]
    centre = ((pmin + pmax).astype(f32) * f32(0.5)).astype(f32)
    index = np.arange(n, dtype=np.int64)
    INF = f32(np.inf)
    # BFS node records
    nb_min = [pmin.min(0) if n else np.full(3, INF, f32)]
    nb_max = [pmax.max(0) if n else np.full(3, -INF, f32)]
    n_start, n_end, n_depth, n_left, n_right, n_axis = [0], [n], [0], [-1], [-1], [0]
    active = [0] if (0 < max_tree_depth and n > bin_size) else []
    while active:
        a = np.array(active)
        A = a.size
        starts, ends = np.array([n_start[i] for i in a]), np.array([n_end[i] for i in a])
        bmin, bmax = np.stack([nb_min[i] for i in a]), np.stack([nb_max[i] for i in a])
        # 1. per node
        axis = _longest_axis(bmin, bmax)
        lo = bmin[np.arange(A), axis]; hi = bmax[np.arange(A), axis]
        with np.errstate(all="ignore"):
            step = ((hi - lo).astype(f32) / f32(n_buckets)).astype(f32)
        # 2. per primitive position (segment id by position)
        counts_per_node = ends - starts
        seg = np.repeat(np.arange(A), counts_per_node)
        pos = np.concatenate([np.arange(s, e) for s, e in zip(starts, ends)])
        prim = index[pos]
        with np.errstate(all="ignore"):
            q = np.floor((((centre[prim, axis[seg]] - lo[seg]).astype(f32)) / step[seg]).astype(f32))
        which = np.where(np.isnan(q) | (q <= 0), 0, np.minimum(q, f32(2 ** 62))).astype(np.int64)   # Rust `as usize`: saturating, NaN -> 0
        which = np.where(which == n_buckets, n_buckets - 1, which)
        assert (which < n_buckets).all(), "bucket index out of range (the reference panics here)"
        # 3. per (node, bucket): counts and boxes
        key = seg * n_buckets + which
        cnt = np.bincount(key, minlength=A * n_buckets).reshape(A, n_buckets)
        kmin = np.full((A * n_buckets, 3), INF, f32); kmax = np.full((A * n_buckets, 3), -INF, f32)
        np.minimum.at(kmin, key, pmin[prim]); np.maximum.at(kmax, key, pmax[prim])
        kmin = kmin.reshape(A, n_buckets, 3); kmax = kmax.reshape(A, n_buckets, 3)
        # 4. per node: fallback or best prefix split
        degenerate = (cnt == 0).sum(1) == n_buckets - 1
        pre_min = np.minimum.accumulate(kmin, axis=1); pre_max = np.maximum.accumulate(kmax, axis=1)                  # union of buckets 0..i
        suf_min = np.minimum.accumulate(kmin[:, ::-1], axis=1)[:, ::-1]; suf_max = np.maximum.accumulate(kmax[:, ::-1], axis=1)[:, ::-1]
        pre_cnt = np.cumsum(cnt, 1)
        with np.errstate(all="ignore"):
            cost = ((_area(pre_min[:, :-1], pre_max[:, :-1]) * pre_cnt[:, :-1].astype(f32)).astype(f32) +
                    (_area(suf_min[:, 1:], suf_max[:, 1:]) * (counts_per_node[:, None] - pre_cnt[:, :-1]).astype(f32)).astype(f32)).astype(f32)
        # first strict minimum below +inf (NaN never wins); when none does, the reference keeps its initial (bucket 0 | rest)
        best = np.zeros(A, np.int64); best_cost = np.full(A, INF, f32)
        for i in range(n_buckets - 1):
            better = cost[:, i] < best_cost
            best = np.where(better, i, best); best_cost = np.where(better, cost[:, i], best_cost)
        left_count = pre_cnt[np.arange(A), best]
        lmin, lmax = pre_min[np.arange(A), best], pre_max[np.arange(A), best]
        rmin, rmax = suf_min[np.arange(A), best + 1], suf_max[np.arange(A), best + 1]
        # 5. stable rewrite: node start + earlier buckets + rank within (node, bucket)
        bucket_offset = np.cumsum(cnt, 1) - cnt
        order = np.argsort(key, kind="stable")                     # = segmented exclusive scan per bucket
        rank = np.empty_like(order); 
        first_of_key = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=A * n_buckets))[:-1]])
        rank[order] = np.arange(order.size) - first_of_key[key[order]]
        new_pos = starts[seg] + bucket_offset[seg, which] + rank
        keep = ~degenerate[seg]
        new_index = index.copy()
        new_index[new_pos[keep]] = prim[keep]
        index = new_index
        # the reference's BalanceTree fallback for degenerate nodes: stable sort of the range by centre, split at the middle
        for j in np.nonzero(degenerate)[0]:
            s, e = starts[j], ends[j]
            if (e - s) // 2 != 0:
                rng_prims = index[s:e]
                index[s:e] = rng_prims[np.argsort(centre[rng_prims, axis[j]], kind="stable")]
            mid = (s + e) // 2
            left_count[j] = mid - s
            lp, rp = index[s:mid], index[mid:e]
            lmin[j] = pmin[lp].min(0) if lp.size else INF; lmax[j] = pmax[lp].max(0) if lp.size else -INF
            rmin[j] = pmin[rp].min(0) if rp.size else INF; rmax[j] = pmax[rp].max(0) if rp.size else -INF
        # 6. children and the next level
        nxt = []
        for j, i in enumerate(a):
            d = n_depth[i] + 1
            for (cmin, cmax, s, e, side) in ((lmin[j], lmax[j], starts[j], starts[j] + left_count[j], "l"), (rmin[j], rmax[j], starts[j] + left_count[j], ends[j], "r")):
                cid = len(n_start)
                nb_min.append(cmin.copy()); nb_max.append(cmax.copy()); n_start.append(int(s)); n_end.append(int(e)); n_depth.append(d)
                n_left.append(-1); n_right.append(-1); n_axis.append(0)
                if side == "l": n_left[i] = cid
                else: n_right[i] = cid
                if d < max_tree_depth and (e - s) > bin_size: nxt.append(cid)
            n_axis[i] = int(axis[j])
        active = nxt
    # subtree sizes bottom-up (children are created after their parents), pre-order index top-down
    m = len(n_start)
    size = np.ones(m, np.int64)
    for i in range(m - 1, -1, -1):
        if n_left[i] >= 0: size[i] += size[n_left[i]] + size[n_right[i]]
    pre = np.zeros(m, np.int64)
    for i in range(m):
        if n_left[i] >= 0:
            pre[n_left[i]] = pre[i] + 1
            pre[n_right[i]] = pre[i] + 1 + size[n_left[i]]
    out = {"bmin": np.zeros((m, 3), f32), "bmax": np.zeros((m, 3), f32), "start": np.zeros(m, np.uint64), "end": np.zeros(m, np.uint64),
           "self_index": np.arange(m, dtype=np.uint64), "left_count": np.zeros(m, np.uint64), "has_child": np.zeros(m, np.int32), "split_axis": np.zeros(m, np.int32)}
    for i in range(m):
        k = pre[i]
        out["bmin"][k], out["bmax"][k], out["start"][k], out["end"][k] = nb_min[i], nb_max[i], n_start[i], n_end[i]
        if n_left[i] >= 0:
            out["has_child"][k] = 1; out["split_axis"][k] = n_axis[i]; out["left_count"][k] = size[n_left[i]]
    return out, index.astype(np.uint64)
