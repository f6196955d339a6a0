"""Multi-GPU driver of the traversal path: one process per GPU, BVH replicated, rays sharded by tile.

The reference renders through a single wgpu queue; its only sharding notion is the 512 x 512 launch tile of
``GPUWaveFrontComputeRaytracingEncoder::trace_ray`` (shader/ray-tracing/src/backend/wavefront_compute/mod.rs:134) cut
by ``rect_split_iter`` (mod.rs:234-244), which silently drops the remainder columns when the width is not a multiple
of the split (``sub_width = w / split_count``).  Here the same tile quantum is the unit dealt to ranks, but the cut is
exact: edge tiles are simply narrower.

Rays never interact and the scene is read-only during a launch, so the data path has NO collective.  The only
exchanges are (1) one broadcast of the flattened-scene blob after ``commit`` (NCCL over NVLink between GPUs, gloo
between host-only scenes in the CPU tests) and (2) an optional gather of the hit records onto one rank.

``torch.distributed`` is plumbing here: process groups, the broadcast and the gather.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

TILE = 512  # the reference's launch tile (wavefront_compute/mod.rs:134)

Tile = Tuple[int, int, int, int]  # x0, y0, w, h


def launch_tiles(width: int, height: int, tile: int = TILE, tile_h: int | None = None) -> List[Tile]:
    """Row-major list of tiles (``tile`` wide, ``tile_h`` high — square by default) covering a ``width x height`` launch exactly
    (edge tiles are narrower; nothing dropped)."""
    tile_h = tile if tile_h is None else tile_h
    if width < 0 or height < 0 or tile <= 0 or tile_h <= 0:
        raise ValueError("launch_tiles: bad extent")
    out = []
    for y0 in range(0, height, tile_h):
        for x0 in range(0, width, tile):
            out.append((x0, y0, min(tile, width - x0), min(tile_h, height - y0)))
    return out


def shard_tiles(n_tiles: int, world: int, rank: int) -> List[int]:
    """Tiles dealt round-robin: rank r owns tiles r, r + world, r + 2 world, ..."""
    if not 0 <= rank < world:
        raise ValueError("shard_tiles: rank outside the world")
    return list(range(rank, n_tiles, world))


def tile_ray_indices(width: int, tile: Tile) -> np.ndarray:
    """Linear (row-major) launch indices of one tile's rays, in the tile's own row-major order."""
    x0, y0, w, h = tile
    return ((np.arange(y0, y0 + h, dtype=np.int64)[:, None] * width) + np.arange(x0, x0 + w, dtype=np.int64)[None, :]).reshape(-1)


class TileShard:
    """This rank's slice of a 2-D launch: which tiles it owns, how to pull its rays out of the full launch and how to
    put its hits back.  Rays of a tile are stored contiguously in the tile's row-major order, so each tile is traced
    as its own small 2-D launch (``grid_width`` = tile width keeps the kernel's 8 x 4 pixel-tile walk)."""

    def __init__(self, width: int, height: int, world: int, rank: int, tile: int = TILE, tile_h: int | None = None):
        self.width, self.height, self.world, self.rank, self.tile, self.tile_h = width, height, world, rank, tile, tile_h
        self.tiles_all = launch_tiles(width, height, tile, tile_h)
        self.tile_ids = shard_tiles(len(self.tiles_all), world, rank)
        self.tiles = [self.tiles_all[i] for i in self.tile_ids]
        counts = [w * h for (_, _, w, h) in self.tiles]
        self.offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.n_rays = int(self.offsets[-1])

    def indices(self) -> np.ndarray:
        """launch indices of this rank's rays, tile after tile"""
        if not self.tiles:
            return np.zeros(0, np.int64)
        return np.concatenate([tile_ray_indices(self.width, t) for t in self.tiles])

    def gather_rays(self, launch_rays: np.ndarray) -> np.ndarray:
        return np.ascontiguousarray(launch_rays.reshape(-1)[self.indices()])

    def launches(self):
        """(offset, count, grid_width) of each owned tile inside the shard buffer"""
        for k, (_, _, w, h) in enumerate(self.tiles):
            yield int(self.offsets[k]), w * h, w

    def scatter_hits(self, shard_hits: np.ndarray, launch_hits: np.ndarray) -> None:
        launch_hits.reshape(-1)[self.indices()] = shard_hits


def trace_shard(system, shard: TileShard, shard_rays: np.ndarray, **launch_kw) -> np.ndarray:
    """Trace this rank's tiles through the host-buffer C-ABI call, one 2-D launch per tile."""
    from . import api
    hits = np.empty(shard.n_rays, api.HIT_DTYPE)
    for off, cnt, gw in shard.launches():
        system.trace_closest_batch(shard_rays[off:off + cnt], grid_width=gw, out=hits[off:off + cnt], **launch_kw)
    return hits


# ----------------------------------------------------------------------------------------------- one frame, sharded
class ShardedFrame:
    """One rank's share of a multi-sample frame — BASELINE configs[4]: ``width x height x spp`` jittered pinhole rays plus one
    cosine bounce per hit — kept on the device from the camera parameters to the last hit record.

    Sharding: the frame is cut into SMALL tiles (64 x 32 pixels by default) dealt round-robin, so every rank gets an even
    sample of cheap (background) and expensive (silhouette) image regions; the 512 x 512 launch tiles of the reference
    (wavefront_compute/mod.rs:134) left 40 unequal tiles for 8 GPUs.  Small tiles cost nothing here because all tiles and all
    samples of a rank are ONE wave: its rays form a virtual image, tile width wide, with the tiles and sample planes stacked
    vertically (tile heights are multiples of 4 and widths multiples of 8, so the kernel's 8 x 4 pixel blocks never straddle two
    tiles), generated by one batched launch and traced by one launch; the bounce wave is sized on the device
    (rdn_rt_trace_closest_device_n), so a frame is six kernels and no host round trip.
    """

    def __init__(self, system, width: int, height: int, spp: int, world: int, rank: int, device, tile_w: int = 64, tile_h: int = 32,
                 tmin: float = 0.01, tmax: float = 100.0, primary_flags: int = 0x10, bounce_flags: int = 0):
        import torch

        from . import scenes as S
        if tile_w % 8 or tile_h % 4:
            raise ValueError("ShardedFrame: tile width must be a multiple of 8 and tile height a multiple of 4")
        self.system, self.spp, self.device = system, spp, device
        self.primary_flags, self.bounce_flags, self.tmin, self.tmax = primary_flags, bounce_flags, tmin, tmax
        self.shard = TileShard(width, height, world, rank, tile=tile_w, tile_h=tile_h)
        aspect = float(np.float32(width / height))
        jitters = [S.sample_2d(np.full(1, s, np.uint32))[0] if s else np.array([0.5, 0.5], np.float32) for s in range(spp)]
        # tiles of one width form one wave (only the last column of a frame can be narrower: at most two waves per rank)
        by_width = {}
        for t in self.shard.tiles:
            by_width.setdefault(t[2], []).append(t)
        self.waves = []
        for gw, tiles in by_width.items():
            rects = [t for s in range(spp) for t in tiles]
            jits = [(float(jitters[s][0]), float(jitters[s][1])) for s in range(spp) for _ in tiles]
            batch = system.pinhole_batch(width, height, rects, jits, tmin=tmin, tmax=tmax, aspect=aspect)
            n = batch[2]
            buf = lambda: torch.empty((max(n, 1), 32), dtype=torch.uint8, device=device)
            self.waves.append(dict(gw=gw, n=n, tiles=tiles, batch=batch, rays=buf(), hits=buf(), brays=buf(), bhits=buf(),
                                   src=torch.empty(max(n, 1), dtype=torch.int32, device=device), cnt=torch.zeros(1, dtype=torch.int64, device=device)))
        self.n_primary = sum(w["n"] for w in self.waves)

    def enqueue(self, stream: int) -> None:
        """all kernels of one frame on ``stream``; returns without waiting"""
        p = self.system
        for w in self.waves:
            if w["n"] == 0:
                continue
            p.gen_pinhole_rays_prebuilt_device(w["batch"], w["rays"].data_ptr(), stream=stream)
            p.trace_closest_device(w["rays"].data_ptr(), w["n"], w["hits"].data_ptr(), ray_flags=self.primary_flags, grid_width=w["gw"], stream=stream)
            p.gen_bounce_rays_device(w["rays"].data_ptr(), w["hits"].data_ptr(), w["n"], w["brays"].data_ptr(), w["src"].data_ptr(),
                                     w["cnt"].data_ptr(), mode=0, index_base=0, tmin=self.tmin, tmax=self.tmax, stream=stream)
            p.trace_closest_device_n(w["brays"].data_ptr(), w["cnt"].data_ptr(), w["n"], w["bhits"].data_ptr(), ray_flags=self.bounce_flags, stream=stream)

    def n_bounce(self) -> int:
        """bounce rays of the last frame (reads the device-side counts: call after the stream has been waited for)"""
        return int(sum(int(w["cnt"].item()) for w in self.waves if w["n"]))


# ----------------------------------------------------------------------------------------------- replication
def replicate_scene(system, src: int = 0, group=None, device=None) -> float:
    """Broadcast the flattened scene of rank ``src`` to every rank of ``group`` and adopt it there.

    GPU scenes (``system.devices`` non-empty) broadcast the device blob in place with the group's backend (NCCL: NVLink /
    NVSwitch); host-only scenes (``devices == ()``, CPU tests) broadcast the host blob (gloo).  Returns the broadcast time in
    milliseconds (CUDA events for the device path, wall clock for the host path)."""
    import time

    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    host_only = len(system.devices) == 0
    if host_only:
        nbytes = torch.zeros(1, dtype=torch.int64)
        if rank == src:
            ptr, nb = system.blob(device_index=-1)
            nbytes[0] = nb
        dist.broadcast(nbytes, src, group=group)
        nb = int(nbytes.item())
        buf = torch.empty(nb, dtype=torch.uint8)
        if rank == src:
            import ctypes
            buf.copy_(torch.from_numpy(np.ctypeslib.as_array((ctypes.c_uint8 * nb).from_address(ptr)).copy()))
        t0 = time.perf_counter()
        dist.broadcast(buf, src, group=group)
        ms = (time.perf_counter() - t0) * 1e3
        if rank != src:
            system.adopt_blob(buf.data_ptr(), nb, device_index=-1)
        return ms

    from cuda.bindings import runtime as cudart
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    nbytes = torch.zeros(1, dtype=torch.int64, device=dev)
    ptr = 0
    if rank == src:
        ptr, nb = system.blob()
        nbytes[0] = nb
    dist.broadcast(nbytes, src, group=group)
    nb = int(nbytes.item())
    buf = torch.empty(nb, dtype=torch.uint8, device=dev)
    if rank == src:
        (err,) = cudart.cudaMemcpy(buf.data_ptr(), ptr, nb, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
        if int(err) != 0:
            raise RuntimeError(f"cudaMemcpy of the scene blob failed: {err}")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dist.broadcast(buf, src, group=group)  # the one collective of the path: BVH replication
    e1.record()
    torch.cuda.synchronize()
    if rank != src:
        system.adopt_blob(buf.data_ptr(), nb)
    return float(e0.elapsed_time(e1))


def gather_launch_hits(shard: TileShard, shard_hits: np.ndarray, dst: int = 0, group=None):
    """Assemble the full launch on rank ``dst`` from every rank's shard (host records; returns None elsewhere)."""
    import torch.distributed as dist

    from . import api
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    parts: Sequence = [None] * world if rank == dst else None
    dist.gather_object(shard_hits, parts, dst=dst, group=group)
    if rank != dst:
        return None
    full = np.empty(shard.width * shard.height, api.HIT_DTYPE)
    for r, part in enumerate(parts):
        TileShard(shard.width, shard.height, world, r, tile=shard.tile).scatter_hits(part, full)
    return full


def bind_process_to_gpu_numa_node(device_index: int) -> int:
    """One process per GPU: run this process (and, by first touch, place its pinned host buffers) on the CPUs NVML reports as
    local to the GPU, so that host<->device copies of the ranks on the second socket do not cross the inter-socket link.
    Returns the number of CPUs bound to, 0 when NVML / the affinity call is unavailable (nothing changed).  Call before
    allocating pinned memory."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return 0
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0
