#!/usr/bin/env python
"""bench.py — closest-hit Mrays/s of the BVH traversal hot path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU traversal (oracle port) on host cores

Workload (config.workload): BASELINE.json configs[1] — 1920x1080 primary rays vs a 1,002,528-triangle generated
parametric-surface mesh (708x708 torus, x5 at z=-10, rotate_x(-0.5)), RAY_FLAG_CULL_BACK_FACING_TRIANGLES, range
(0.01, 100), one TLAS with one instance.  One "step" = one closest-hit pass over one frame of rays on every GPU.
For N > 1 the flattened BVH is built on rank 0, replicated with one NCCL broadcast, and every rank traces its own
(differently jittered) frame: rays are sharded by frame tile, no collective on the data path -> "weak" scaling.

value   = whole-job Mrays/s with rays/hits resident in HBM, device-timed (CUDA events), max over ranks.  Successive steps
          trace N_FRAMES different frames (ray + hit buffers 4 x 133 MB, scene 171 MB: inputs far larger than the 126 MB
          L2, no artificial flush — the BVH staying L2-warm between frames is the steady state of a renderer);
          details.value_l2_flushed is the same loop with a 256 MiB flush between steps, details.value_serialized the same
          loop without tail overlap between launches.
e2e     = same metric through the host-buffer C-ABI call (pinned host rays -> H2D -> traversal -> D2H hits); e2e_pageable =
          the same call on caller-owned arrays (what a Rust Vec is): `value` page-locked once by their owner (rdn_rt_host_register),
          `value_unregistered` left ordinary pageable memory, which the library stages chunk by chunk through its own page-locked
          buffers with several host threads.
roofline= the kernel is bound by instruction issue at ~20 of 32 active lanes; the memory level that serves it is L2.  `frac` =
          reference-defined algorithmic bytes (64 + 48*V_node + 52*V_tri + 176*V_inst per ray, V counted by
          the oracle under the reference traversal order) / serialised kernel time / the L2 read peak measured in this run;
          the bytes the kernel ACTUALLY moves (lts__t_bytes, dram__bytes of the committed ncu capture of the shipped
          instantiation) are reported against the L2 and HBM peaks beside it.
c5      = (every N) BASELINE configs[4], strong scaling: ONE 3840x2160 x 16 spp primary + 1 bounce frame cut into 64x32-pixel
          tiles dealt round-robin to the ranks, device-resident from the camera to the hit records (--config c5 makes it the
          headline of the line instead).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H = 1920, 1080
SEG = 708
RAY_FLAGS = 0x10  # RAY_FLAG_CULL_BACK_FACING_TRIANGLES
TMIN, TMAX = 0.01, 100.0
N_FRAMES = 4  # distinct ray/hit buffers cycled by successive steps
WORKLOAD = "1920x1080 primary rays vs 1,002,528-triangle generated torus mesh (BASELINE configs[1])"
C5_W, C5_H, C5_SPP = 3840, 2160, 16
C5_WORKLOAD = "3840x2160 x 16 spp primary + 1 cosine bounce rays vs the 1,002,528-triangle torus, ray tiles sharded across the GPUs, BVH replicated (BASELINE configs[4])"
# the instantiation launch_trace_ordered selects for grid launches (rendiation_b200/csrc/traverse.cu, `plain`), as ncu prints it:
# the committed capture whose numbers the roofline quotes must be of this kernel (tests/test_bench_contract.py holds the two together)
SHIPPED_ORDERED_KERNEL = "k_trace_ordered_rounds<3, 8, 1, 0, 1, 0, 0, 1, 0, 0, 1, 0>"  # (the last two arguments: tile history, no top-up)


def base_config(workload=WORKLOAD, rays=W * H):
    """the workload description both arms print (identical keys and values: the driver compares them)"""
    return {"workload": workload, "rays_per_gpu_per_step": int(rays), "triangles": int(SEG * SEG * 2), "ray_flags": RAY_FLAGS}


def shipped_kernel_capture():
    """(path, launch) of the newest committed ncu --set full summary of the shipped ordered kernel on configs[1], or (None, None)"""
    import glob
    import re
    want = re.sub(r"[()\s]|int|bool", "", SHIPPED_ORDERED_KERNEL)
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r*_k_trace_ordered_c2.json"))):
        try:
            launches = json.load(open(path)).get("launches", [])
        except (OSError, ValueError):
            continue
        for x in launches:
            if want in re.sub(r"[()\s]|int|bool", "", x.get("kernel", "")) and "dram_traffic_bytes" in x:
                best = (path, x)  # (sorted by name: later rounds / sessions win)
    return best if best else (None, None)


def scene_inputs():
    from rendiation_b200 import scenes as S
    pos, idx = S.torus_mesh(SEG, SEG, 1.0, 0.35)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    return pos, idx, m


def frame_rays(sample_index: int):
    """frame `sample_index`: pixel-centre rays for 0, (van der Corput, Sobol) sub-pixel jitter otherwise"""
    from rendiation_b200 import scenes as S
    jitter = None
    if sample_index:
        j = S.sample_2d(np.full(1, sample_index, np.uint32))
        jitter = np.repeat(j, W * H, axis=0)
    return S.pinhole_rays(W, H, TMIN, TMAX, aspect_correct=True, jitter=jitter)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled in-process
    every ~1 ms (a timed region of K sub-millisecond steps is far shorter than one `nvidia-smi -lms 200` period); the same
    fields through the nvidia-smi CLI are the fallback when pynvml is unusable."""

    def __init__(self, torch_device_index: int):
        self.rows = []  # (sm_mhz, sm_max_mhz, power_w, reasons_bitmask)
        self.stop = threading.Event()
        self.gpu = torch_device_index
        self.th = None
        self.nvml = None
        self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_device_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(torch_device_index)
            self.nvml = pynvml
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
        reasons = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        try:
            power = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
        except Exception:
            power = float("nan")
        self.rows.append((sm, self.sm_max, power, reasons))

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        if not out:
            return
        c = [x.strip() for x in out.split(",")]
        bits = 0
        for bit, v in zip((0x8, 0x40, 0x20, 0x4), c[3:7]):  # NVML bit values of the four reasons
            if v.lower().startswith("active"):
                bits |= bit
        self.rows.append((float(c[0]), float(c[1]), float(c[2]), bits))

    def _run(self):
        while not self.stop.is_set():
            try:
                self._sample_nvml() if self.nvml else self._sample_smi()
            except Exception:
                pass
            self.stop.wait(0.001 if self.nvml else 0.1)

    def __enter__(self):
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        reasons = sorted({nm for r in self.rows for bit, nm in names.items() if r[3] & bit})
        sm = [r[0] for r in self.rows]
        pw = [r[2] for r in self.rows if r[2] == r[2]]
        return {"sm_mhz": float(np.median(sm)), "sm_min_mhz": float(min(sm)), "sm_max_mhz": float(max(r[1] for r in self.rows)),
                "power_w_max": float(max(pw)) if pw else None, "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml else "nvidia-smi"}


def algorithmic_bytes_per_ray(osc, rays, n_threads):
    """SURVEY.md §8(d) bytes model from the oracle's reference-order counters on a 1/8 row-strided sample of the frame"""
    sample = rays.reshape(H, W)[::8].reshape(-1)
    _, c = osc.trace(sample, ray_flags=RAY_FLAGS, n_threads=n_threads)
    n = sample.shape[0]
    per_ray = 64 + (48 * c["bvh_visit"] + 52 * c["tri_visit"] + 176 * c["inst_visit"]) / n
    return per_ray, {k: v / n for k, v in c.items()}, n


def build_oracle_scene():
    import oracle
    from rendiation_b200 import scenes as S
    pos, idx, m = scene_inputs()
    osc = oracle.Scene()
    b = osc.create_blas([(pos, idx.reshape(-1), 1)])
    osc.bind_tlas([osc.create_tlas(S.make_instance(m, b))])
    assert osc.build() == 0
    return osc


def run_reference(args):
    """--impl reference: the reference's own CPU traversal (NaiveSahBvhCpu::traverse restated in C, oracle/) on all host
    cores; each step = one full 1920x1080 frame of the same workload (about 1.7 core-seconds of CPU work per step)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    osc = build_oracle_scene()
    if args.config == "c5":
        return run_reference_c5(args, osc, cores)
    rays = frame_rays(0)
    for _ in range(args.warmup):
        osc.trace(rays, ray_flags=RAY_FLAGS, n_threads=cores, want_counters=False)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        osc.trace(rays, ray_flags=RAY_FLAGS, n_threads=cores, want_counters=False)
    dt = time.perf_counter() - t0
    v = rays.shape[0] * args.steps / dt / 1e6
    sample = f"one full 1920x1080 frame ({rays.shape[0]} rays) per step, {args.steps} steps, {dt:.2f} s wall on {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": "closest-hit Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": base_config(), "details": {"host_threads": cores},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def c5_sample_rays(osc, cores, tile=512):
    """a bounded sample of configs[4] for the CPU: the central 512 x 512 pixel tile of the 3840 x 2160 frame at one sample per pixel
    (pixel centres) and the cosine bounce off its hits — (primary rays, bounce rays)"""
    from rendiation_b200 import scenes as S
    pos, idx, m = scene_inputs()
    x0, y0 = (C5_W - tile) // 2, (C5_H - tile) // 2
    full_row = np.arange(C5_W, dtype=np.float32)
    aspect = np.float32(C5_W / C5_H)
    i = np.arange(x0, x0 + tile, dtype=np.float32)[None, :].repeat(tile, 0).reshape(-1)
    j = np.arange(y0, y0 + tile, dtype=np.float32)[:, None].repeat(tile, 1).reshape(-1)
    f32 = np.float32
    x = ((i + f32(0.5)) / f32(C5_W) * f32(2) - f32(1)) * aspect
    y = f32(1) - (j + f32(0.5)) / f32(C5_H) * f32(2)
    d = np.stack([x, y, np.full_like(x, -1)], -1).astype(np.float32)
    d = d * (f32(1) / np.sqrt((d * d).sum(-1, dtype=np.float32)))[:, None]
    rays = np.zeros(tile * tile, S.RAY_DTYPE)
    rays["tmin"], rays["tmax"] = TMIN, TMAX
    rays["dx"], rays["dy"], rays["dz"] = d[:, 0], d[:, 1], d[:, 2]
    del full_row
    ph = osc.trace(rays, ray_flags=RAY_FLAGS, n_threads=cores, want_counters=False)
    hit = ph["instance_id"] != 0xFFFFFFFF
    normals = np.zeros((rays.shape[0], 3), np.float32)
    normals[hit] = S.geometric_normals(pos, idx, ph["primitive_id"][hit], m, d[hit])
    bounce, _ = S.bounce_rays(rays, ph, normals)
    return rays, bounce


def run_reference_c5(args, osc, cores):
    """--impl reference --config c5: the CPU traversal on a bounded sample of configs[4] per step (one 512 x 512 tile, primary + bounce)"""
    rays, bounce = c5_sample_rays(osc, cores)
    n = rays.shape[0] + bounce.shape[0]

    def step():
        osc.trace(rays, ray_flags=RAY_FLAGS, n_threads=cores, want_counters=False)
        osc.trace(bounce, ray_flags=0, n_threads=cores, want_counters=False)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt / 1e6
    sample = (f"the central 512x512 tile at 1 spp + its bounce rays ({n} rays) per step, {args.steps} steps, {dt:.2f} s wall on {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "closest-hit Mrays/s", "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": base_config(C5_WORKLOAD, C5_W * C5_H * C5_SPP // max(args.gpus, 1)), "details": {"host_threads": cores},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def measure_c5(sysm, world, rank, dev, steps, warmup=1, want_e2e=True):
    """BASELINE configs[4] on the ranks of this job: one frame per step, strong scaling.  Device time = CUDA events around the K
    frames on the launching stream, max over ranks; rays = primary + bounce of all ranks.  e2e adds, inside the timed region, the
    device->host copy of both waves' hit records into pinned memory (the input of this workload is a camera: a few hundred bytes)."""
    import torch
    import torch.distributed as dist

    from rendiation_b200 import multi_gpu
    frame = multi_gpu.ShardedFrame(sysm, C5_W, C5_H, C5_SPP, world, rank, dev, tmin=TMIN, tmax=TMAX, primary_flags=RAY_FLAGS, bounce_flags=0)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return float(t.item())

    for _ in range(max(warmup, 1)):
        frame.enqueue(stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        frame.enqueue(stream)
    e1.record()
    barrier()
    sysm.poll_errors(stream=stream)
    ms = reduce(e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None) / steps
    ms_min = -reduce(-e0.elapsed_time(e1), dist.ReduceOp.MAX if world > 1 else None) / steps
    n_bounce_local = frame.n_bounce()
    n_primary = reduce(frame.n_primary, dist.ReduceOp.SUM if world > 1 else None)
    n_bounce = reduce(n_bounce_local, dist.ReduceOp.SUM if world > 1 else None)
    rays = n_primary + n_bounce
    out = {"workload": C5_WORKLOAD, "scaling": "strong", "value": rays / (ms * 1e-3) / 1e6, "unit": "Mrays/s", "ms_per_frame": ms,
           "ms_per_frame_fastest_rank": ms_min, "steps": steps, "primary_rays": int(n_primary), "bounce_rays": int(n_bounce),
           "tiles": "%dx%d pixels, %d tiles dealt round-robin, one wave per rank (all tiles x %d spp)" % (frame.shard.tile, frame.shard.tile_h, len(frame.shard.tiles_all), C5_SPP),
           "kernels_per_frame": 6 * len([w for w in frame.waves if w["n"]]),
           "includes": "device ray generation, traversal, hit compaction, bounce generation, bounce traversal sized on the device: no host round trip"}
    if want_e2e:
        bytes_local = 0
        pins = []
        for w in frame.waves:
            if w["n"]:
                pins.append((torch.empty((w["n"], 32), dtype=torch.uint8, pin_memory=True), torch.empty((w["n"], 32), dtype=torch.uint8, pin_memory=True), w))
        e2e_steps = max(1, min(steps, 2))

        def e2e_frame():
            nb = 0
            frame.enqueue(stream)
            for h1, h2, w in pins:
                h1.copy_(w["hits"], non_blocking=True)
                k = int(w["cnt"].item())  # (the result size is part of the result: this read is the frame's own synchronisation point)
                h2[:k].copy_(w["bhits"][:k], non_blocking=True)
                nb += 32 * (w["n"] + k)
            torch.cuda.synchronize()
            return nb

        e2e_frame()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            bytes_local = e2e_frame()
        dt = reduce(time.perf_counter() - t0, dist.ReduceOp.MAX if world > 1 else None)
        out["e2e"] = {"value": rays * e2e_steps / dt / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 512, "d2h_bytes_per_step": int(bytes_local),
                      "steps": e2e_steps, "note": "camera parameters in, both waves' hit records out to pinned host memory (bytes are this rank's)"}
    del frame
    torch.cuda.empty_cache()
    return out


def measure_other_configs(sysm_c2, dev, iters=12):
    """BASELINE configs[0], [2] and [3] on this GPU (N = 1 only), device-resident like `value`: Mrays/s of back-to-back launches with
    tail overlap, of serialised launches (L2 flushed before each), the shipped kernel of each, and a 1-in-7 sample of the records
    compared with the oracle.  configs[2]'s rays are the cosine bounce off configs[1]'s hits, generated on the device."""
    import torch

    import oracle
    from rendiation_b200 import api, scenes as S
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cores = len(os.sched_getaffinity(0)) or 1
    out = {}

    def run(name, sysm, osc, d_rays, n, flags, grid, note):
        hits = [torch.full((n, 32), 0xAB, dtype=torch.uint8, device=dev) for _ in range(iters)]
        ref = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
        for _ in range(3):
            sysm.trace_closest_device(d_rays.data_ptr(), n, ref.data_ptr(), ray_flags=flags, grid_width=grid, stream=stream)
        st = sysm.trace_closest_device(d_rays.data_ptr(), n, ref.data_ptr(), ray_flags=flags, grid_width=grid, stream=stream, want_stats=True)
        ms = []
        for k in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); sysm.trace_closest_device(d_rays.data_ptr(), n, hits[k].data_ptr(), ray_flags=flags, grid_width=grid, stream=stream); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        for h in hits:
            h.fill_(0xAB)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(iters):
            sysm.trace_closest_device(d_rays.data_ptr(), n, hits[k].data_ptr(), ray_flags=flags, grid_width=grid, stream=stream, overlap_previous=True)
        e1.record()
        torch.cuda.synchronize()
        b2b = e0.elapsed_time(e1) / iters
        same = all(bool(torch.equal(h, ref)) for h in hits)
        rays_np = d_rays.cpu().numpy().view(S.RAY_DTYPE).reshape(-1)
        sel = np.arange(0, n, 7)
        t0 = time.perf_counter()
        want = osc.trace(rays_np[sel], ray_flags=flags, n_threads=cores, want_counters=False)
        t_cpu = time.perf_counter() - t0
        one = rays_np[::97]
        t0 = time.perf_counter()
        osc.trace(one, ray_flags=flags, n_threads=1, want_counters=False)
        t_cpu1 = time.perf_counter() - t0
        got = ref.cpu().numpy().view(api.HIT_DTYPE).reshape(-1)[sel]
        out[name] = {"workload": note, "rays": int(n), "value_serialized": n / float(np.mean(ms)) / 1e3, "ms_serialized": float(np.mean(ms)),
                     "value": n / b2b / 1e3, "ms": b2b, "unit": "Mrays/s", "tie_rays": st["tie_rays"],
                     "cpu_mrays": sel.size / t_cpu / 1e6, "cpu_cores": cores, "cpu_single_thread_mrays": one.shape[0] / t_cpu1 / 1e6,
                     "sample_bit_identical_to_oracle": bool(got.tobytes() == want.tobytes()), "overlapped_launches_identical": same}

    # configs[0]: 1024 x 1024 primary rays vs the 64 x 64-segment sphere (path B)
    pos, idx = S.uv_sphere_mesh(64, 64)
    m = S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5))
    s1 = api.NaiveSahBVHSystem(devices=(dev.index,))
    b = s1.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
    s1.bind_tlas([s1.create_top_level_acceleration_structure(S.make_instance(m, b.id))]); s1.commit()
    o1 = oracle.Scene(); ob = o1.create_blas([(pos, idx.reshape(-1), 1)]); o1.bind_tlas([o1.create_tlas(S.make_instance(m, ob))]); assert o1.build() == 0
    r1 = torch.from_numpy(S.pinhole_rays(1024, 1024, 0.0, 100.0).view(np.uint8).reshape(-1, 32).copy()).to(dev)
    run("c1", s1, o1, r1, 1024 * 1024, RAY_FLAGS, 1024, "1,048,576 coherent primary rays vs the 64x64-segment sphere (8,192 triangles), BASELINE configs[0], path B")
    del s1, r1
    # configs[2]: one incoherent cosine bounce ray per primary hit of configs[1]
    o2 = build_oracle_scene()
    n = W * H
    d_rays = torch.from_numpy(frame_rays(0).view(np.uint8).reshape(-1, 32).copy()).to(dev)
    d_hits = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    d_b = torch.zeros((n, 32), dtype=torch.uint8, device=dev)
    d_src = torch.zeros(n, dtype=torch.int32, device=dev); d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    sysm_c2.trace_closest_device(d_rays.data_ptr(), n, d_hits.data_ptr(), ray_flags=RAY_FLAGS, grid_width=W, stream=stream)
    sysm_c2.gen_bounce_rays_device(d_rays.data_ptr(), d_hits.data_ptr(), n, d_b.data_ptr(), d_src.data_ptr(), d_n.data_ptr(), mode=0, tmin=TMIN, tmax=TMAX, stream=stream)
    k = int(d_n.item())
    run("c3", sysm_c2, o2, d_b[:k].contiguous(), k, 0, 0, "incoherent cosine-weighted bounce rays off the hits of configs[1] (no culling), BASELINE configs[2]")
    del d_rays, d_hits, d_b
    # configs[3]: 10,000 instances of a 100,352-triangle sphere
    pos, idx = S.uv_sphere_mesh(224, 224)
    s4 = api.NaiveSahBVHSystem(devices=(dev.index,))
    inst, moved = S.instance_grid(100, 100, 0, 3.5, -200.0), S.instance_grid(100, 100, 0, 3.51, -201.0)   # (BLAS handle 0)
    t0 = time.perf_counter()
    b = s4.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
    t4 = s4.create_top_level_acceleration_structure(inst)
    s4.bind_tlas([t4]); s4.commit()
    commit_ms = (time.perf_counter() - t0) * 1e3
    o4 = oracle.Scene(); ob = o4.create_blas([(pos, idx.reshape(-1), 1)]); o4.bind_tlas([o4.create_tlas(S.instance_grid(100, 100, ob, 3.5, -200.0))]); assert o4.build() == 0
    r4 = torch.from_numpy(S.pinhole_rays(W, H, 0.0, 1000.0, aspect_correct=True).view(np.uint8).reshape(-1, 32).copy()).to(dev)
    run("c4", s4, o4, r4, n, RAY_FLAGS, W, "1920x1080 primary rays vs 10,000 transform-instanced copies of a 100,352-triangle sphere, BASELINE configs[3]")
    # every instance moves: rdn_rt_tlas_update + commit (the TLAS part alone is rebuilt, the device blob patched in place); the first
    # update also allocates the page-locked staging buffer of the patches, so the figure is the median of five
    updates = []
    for k in range(5):
        nxt = moved if k % 2 == 0 else inst
        t0 = time.perf_counter()
        s4.update_top_level_acceleration_structure(t4, nxt)
        s4.commit()
        updates.append((time.perf_counter() - t0) * 1e3)
    out["c4"]["commit_ms"] = commit_ms
    out["c4"]["tlas_only_update_ms"] = sorted(updates)[2]
    out["c4"]["tlas_only_update_ms_first"] = updates[0]
    return out


def setup_job():
    """one process per GPU: device, process group, the scene built on rank 0 and replicated with one NCCL broadcast"""
    import torch
    import torch.distributed as dist

    from rendiation_b200 import api, multi_gpu, scenes as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    numa_cpus = multi_gpu.bind_process_to_gpu_numa_node(local_rank) if world > 1 else 0  # pinned e2e buffers local to the GPU's socket
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's own banner ("NCCL version ...") must not land on stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    # ---- scene: build + flatten on rank 0, one NCCL broadcast of the blob, every other rank adopts it
    sysm = api.NaiveSahBVHSystem(devices=(local_rank,))
    t_build = 0.0
    if rank == 0:
        pos, idx, m = scene_inputs()
        t0 = time.perf_counter()
        b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
        t = sysm.create_top_level_acceleration_structure(S.make_instance(m, b.id))
        sysm.bind_tlas([t])
        sysm.commit()
        t_build = time.perf_counter() - t0
    t_repl_ms = 0.0
    if world > 1:  # the one collective of the path: BVH replication (NCCL broadcast over NVLink), then the other ranks adopt the blob
        t_repl_ms = multi_gpu.replicate_scene(sysm, src=0, device=dev)
    return world, rank, local_rank, dev, sysm, t_build, t_repl_ms, numa_cpus


def run_c5(args):
    """--config c5: BASELINE configs[4] is the line — one 3840x2160x16spp + bounce frame per step, tiles sharded over the ranks"""
    import torch
    import torch.distributed as dist

    world, rank, local_rank, dev, sysm, t_build, t_repl_ms, _ = setup_job()
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    with ClockSampler(local_rank) as clocks:
        c5 = measure_c5(sysm, world, rank, dev, steps=args.steps, warmup=max(args.warmup, 3))
    if rank == 0:
        import oracle  # noqa: F401  (checker / baseline only)
        cores = len(os.sched_getaffinity(0)) or 1
        osc = build_oracle_scene()
        rays, bounce = c5_sample_rays(osc, cores)
        t0 = time.perf_counter()
        osc.trace(rays, ray_flags=RAY_FLAGS, n_threads=cores, want_counters=False)
        osc.trace(bounce, ray_flags=0, n_threads=cores, want_counters=False)
        t_cpu = time.perf_counter() - t0
        n_cpu = rays.shape[0] + bounce.shape[0]
        print(json.dumps({
            "metric": "closest-hit Mrays/s", "value": c5["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": c5["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": base_config(C5_WORKLOAD, C5_W * C5_H * C5_SPP // world),
            "details": {k: v for k, v in c5.items() if k not in ("e2e", "value", "unit")} | {"build_s": round(t_build, 3), "blob_broadcast_ms": t_repl_ms,
                        "l2": "inputs larger than L2 (each rank streams %.1f GB of rays and hits per frame)" % (4 * 32 * C5_W * C5_H * C5_SPP / world / 1e9)},
            "e2e": c5.get("e2e"), "gpu_launches": c5["kernels_per_frame"] * args.steps,
            "roofline": None,
            "cpu_baseline": {"value": n_cpu / t_cpu / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"the central 512x512 tile at 1 spp + its bounce rays ({n_cpu} rays), {t_cpu:.2f} s wall on {cores} threads"},
            "clocks": clocks.summary()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_ours(args):
    import torch
    import torch.distributed as dist

    from rendiation_b200 import api

    world, rank, local_rank, dev, sysm, t_build, t_repl_ms, numa_cpus = setup_job()
    blob_bytes = sysm.blob()[1]

    # ---- rays: this rank's frames (tile shard of the job), resident in HBM; steps cycle through N_FRAMES buffers
    frames_np = [frame_rays(rank * N_FRAMES + j) for j in range(N_FRAMES)]
    rays_np = frames_np[0]
    n = rays_np.shape[0]
    d_rays = [torch.from_numpy(f.view(np.uint8).reshape(-1, 32)).to(dev) for f in frames_np]
    # every step of a timed loop writes its OWN hit buffer (up to 64 of them, pre-filled with 0xAB before each loop), so a ray
    # skipped by one of the overlapping launches cannot hide behind the result of an earlier step; d_ref = one synchronised
    # launch per frame, the yardstick for those buffers (and itself compared with the oracle on rank 0)
    n_hit_bufs = max(N_FRAMES, min(args.steps, 64))
    d_hits = [torch.empty((n, 32), dtype=torch.uint8, device=dev) for _ in range(n_hit_bufs)]
    d_ref = [torch.zeros((n, 32), dtype=torch.uint8, device=dev) for _ in range(N_FRAMES)]
    h_rays = torch.from_numpy(rays_np.view(np.uint8).reshape(-1, 32).copy()).pin_memory()
    h_hits = torch.zeros((n, 32), dtype=torch.uint8).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))  # a created stream (not the legacy default stream) for all device work below
    stream = torch.cuda.current_stream().cuda_stream

    def step_device(k=0, stats=False, out=None):
        j = k % N_FRAMES
        dst = out if out is not None else d_hits[k % n_hit_bufs]
        # (overlap_previous: consecutive steps are independent frames with their own ray and hit buffers and nothing else goes
        # into the stream between them — the contract of RDN_TRACE_OVERLAP_PREVIOUS)
        return sysm.trace_closest_device(d_rays[j].data_ptr(), n, dst.data_ptr(), ray_flags=RAY_FLAGS, grid_width=W,
                                         stream=stream, want_stats=stats, overlap_previous=True)

    def reset_hit_buffers():
        for h in d_hits:
            h.fill_(0xAB)

    def hit_buffers_match_reference():
        """every buffer written by the last K-step loop == the synchronised launch of its frame"""
        ok = True
        for k in range(min(args.steps, n_hit_bufs)):
            last_k = k + ((args.steps - 1 - k) // n_hit_bufs) * n_hit_bufs  # the last step that wrote buffer k
            ok = ok and bool(torch.equal(d_hits[k], d_ref[last_k % N_FRAMES]))
        return ok

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for j in range(N_FRAMES):  # the synchronised yardstick, one launch per frame
        step_device(j, out=d_ref[j])
        torch.cuda.synchronize()
    for k in range(max(args.warmup, 3)):
        step_device(k)
    st = step_device(0, stats=True)
    launches_per_step = st["kernel_launches"]
    tie_rays = st["tie_rays"]

    def timed_loop(do_flush=False, per_kernel=False):
        """K steps; returns (total device ms between one event before the first and one after the last step, per-kernel times).
        No event is recorded between steps unless per_kernel/do_flush ask for it: consecutive ordered launches on a stream overlap
        their tails (programmatic dependent launch) and a marker between two kernels would serialise them."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reset_hit_buffers()
        barrier()
        if per_kernel:
            sysm.kernel_timing_begin()  # the library brackets each of its kernels with events on the launching stream
        if do_flush:
            total = 0.0
            for k in range(args.steps):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); step_device(k); b.record()
                torch.cuda.synchronize()
                total += a.elapsed_time(b)
            barrier()
            return total, None
        e0.record()
        for k in range(args.steps):
            step_device(k)
        e1.record()
        barrier()
        kt = sysm.kernel_timing_end() if per_kernel else None
        return e0.elapsed_time(e1), kt

    # ---- timed region: exactly K steps back to back on one stream, CUDA events on that stream at both ends
    with ClockSampler(local_rank) as clocks:
        t_wall0 = time.perf_counter()
        launched_before = sysm.build_stats()["kernels_enqueued"]
        total_ms_local, _ = timed_loop()
        launched_in_timed_region = sysm.build_stats()["kernels_enqueued"] - launched_before   # counted by the library, this rank
        t_wall = time.perf_counter() - t_wall0
        timed_loop_ok = hit_buffers_match_reference()  # what the overlapped launches of the timed region wrote
        serial_ms_local, kernel_times = timed_loop(per_kernel=True)   # same K steps, each kernel bracketed by events (no overlap)
        flushed_ms_local, _ = timed_loop(do_flush=True)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms = max_over_ranks(total_ms_local)
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    ms_per_step_serial = max_over_ranks(serial_ms_local) / args.steps
    value_serial = world * n / (ms_per_step_serial * 1e-3) / 1e6
    ms_per_step_flushed = max_over_ranks(flushed_ms_local) / args.steps
    value_flushed = world * n / (ms_per_step_flushed * 1e-3) / 1e6

    # ---- e2e: host-buffer C-ABI call (pinned host memory; H2D + traversal + D2H inside the timed region)
    def step_host():
        sysm.trace_closest_host_ptr(h_rays.data_ptr(), n, h_hits.data_ptr(), ray_flags=RAY_FLAGS, grid_width=W)

    for _ in range(2):
        step_host()
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = world * n * e2e_steps / float(t_e2e.item()) / 1e6
    # the same call on host arrays that are NOT torch-pinned (plain numpy arrays — what a Rust Vec<Ray> is): first as they are
    # (pageable: the library stages every chunk through page-locked buffers of its own, copied by several host threads), then
    # page-locked by their owner through rdn_rt_host_register, which is what the Rust wrapper does for buffers that live longer
    # than a frame
    p_rays = np.ascontiguousarray(rays_np).copy()
    p_hits = np.zeros(n, api.HIT_DTYPE)

    def step_pageable():
        sysm.trace_closest_host_ptr(p_rays.ctypes.data, n, p_hits.ctypes.data, ray_flags=RAY_FLAGS, grid_width=W)

    def timed_host_steps():
        step_pageable()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_pageable()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n * e2e_steps / float(t.item()) / 1e6

    e2e_unpinned_value = timed_host_steps()
    t0 = time.perf_counter()
    with api.HostRegistration(p_rays), api.HostRegistration(p_hits):
        t_register = time.perf_counter() - t0
        e2e_pageable_value = timed_host_steps()
    same_pageable = p_hits.tobytes() == h_hits.numpy().tobytes()
    # device-resident and host-path results must be the same bits; every rank's timed loop must have matched its yardstick
    same = bool(torch.equal(h_hits.to(dev), d_ref[0]))
    ok_all = torch.tensor([1.0 if timed_loop_ok else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
    timed_loop_ok_all_ranks = bool(ok_all.item() > 0.5)

    # ---- BASELINE configs[4] on the same ranks (strong scaling of one frame), after the headline loops
    del d_hits, d_rays, flush
    torch.cuda.empty_cache()
    c5 = measure_c5(sysm, world, rank, dev, steps=args.c5_steps) if args.c5_steps > 0 else None
    other = measure_other_configs(sysm, dev) if (world == 1 and args.other_configs) else None

    if rank == 0:
        # ---- roofline + cpu baseline + parity spot check (oracle = checker / baseline only)
        cores = len(os.sched_getaffinity(0)) or 1  # (the CPUs this rank may run on: see host_affinity)
        osc = build_oracle_scene()
        bytes_per_ray, visits, n_sample = algorithmic_bytes_per_ray(osc, rays_np, cores)
        # bounded CPU sample: the N_FRAMES full frames of this rank, twice (~14 core-seconds of traversal work)
        cpu_reps = 2
        osc.trace(frames_np[0], ray_flags=RAY_FLAGS, n_threads=cores, want_counters=False)  # warm-up (page in the BVH)
        t0 = time.perf_counter()
        for _ in range(cpu_reps):
            ohits_all = [osc.trace(f, ray_flags=RAY_FLAGS, n_threads=cores, want_counters=False) for f in frames_np]
        t_cpu = time.perf_counter() - t0
        cpu_rays = cpu_reps * N_FRAMES * n
        one = rays_np.reshape(H, W)[::8].reshape(-1).copy()
        t0 = time.perf_counter()
        osc.trace(one, ray_flags=RAY_FLAGS, n_threads=1, want_counters=False)
        t_cpu1 = time.perf_counter() - t0
        # parity: the yardstick launches == the oracle on all N_FRAMES full frames (whole 32-byte records); the K buffers written
        # by the overlapped launches of the timed region == the yardstick (checked above on every rank)
        parity_bits = all(g.cpu().numpy().view(api.HIT_DTYPE).reshape(-1).tobytes() == o.tobytes() for g, o in zip(d_ref, ohits_all))

        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        # dominant kernel = k_trace_ordered_rounds: average launch duration over K timed steps from the CUDA events the library
        # records around it on the launching stream (the serialised pass: a kernel's own duration is only defined without overlap)
        kernel_ms = kernel_times["ordered_ms"] / max(kernel_times["ordered_launches"], 1)
        tie_ms = kernel_times["tie_ms"] / max(kernel_times["tie_launches"], 1)
        achieved = bytes_per_ray * n / (kernel_ms * 1e-3) / 1e9
        compulsory = (64.0 * n + blob_bytes) / (kernel_ms * 1e-3) / 1e9
        l2_peak = sysm.measure_l2_read_gbs(64 << 20, 50)  # resident-set read microbenchmark, this box, this run
        cap_path, cap = shipped_kernel_capture()
        traffic = traffic_l2 = None
        traffic_src = "no committed ncu capture of " + SHIPPED_ORDERED_KERNEL
        if cap:  # per launch, from the committed ncu --set full capture of the shipped instantiation (cold-cache replay of one launch)
            traffic, traffic_l2 = float(cap["dram_traffic_bytes"]), float(cap["l2_traffic_bytes"])
            traffic_src = os.path.relpath(cap_path, ROOT) + " (" + cap["kernel"].split("::")[-1].split("(")[0] + ")"
        per_s = 1.0 / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": "closest-hit Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": base_config(),
            "details": {"l2": "inputs larger than L2: steps cycle %d frames of rays+hits (%.0f MB) over a %.0f MB scene, no flush; "
                              "value_l2_flushed = same loop with a 256 MiB write between steps" %
                              (N_FRAMES, N_FRAMES * 64 * n / 1e6, blob_bytes / 1e6),
                        "value_l2_flushed": value_flushed, "ms_per_step_l2_flushed": ms_per_step_flushed,
                        "launch_overlap": "the K steps are issued back to back on one stream with RDN_TRACE_OVERLAP_PREVIOUS: each ordered "
                                          "launch lets the next one start filling SM slots while its own last long rays finish (programmatic "
                                          "dependent launch, distinct ray/hit buffers per step).  value_serialized = the same K steps with every "
                                          "kernel bracketed by CUDA events, which serialises them (RDN_PDL=0 gives the same)",
                        "value_serialized": value_serial, "ms_per_step_serialized": ms_per_step_serial,
                        "parallelism": f"rays sharded by frame x{world}, BVH replicated ({blob_bytes / 1e6:.0f} MB blob, "
                                       f"NCCL broadcast {t_repl_ms:.2f} ms)", "build_s": round(t_build, 3), "build_stats": sysm.build_stats(),
                        "host_affinity": (f"each rank bound to the {numa_cpus} CPUs NVML reports local to its GPU" if numa_cpus else "unchanged")},
            "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n,
                    "steps": e2e_steps, "matches_device_path": same},
            "e2e_pageable": {"value": e2e_pageable_value, "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n,
                             "steps": e2e_steps, "matches_device_path": same_pageable, "register_ms": t_register * 1e3,
                             "value_unregistered": e2e_unpinned_value,
                             "note": "host arrays owned by the caller (numpy, not torch-pinned) and page-locked by it once with "
                                     "rdn_rt_host_register (register_ms, outside the timed steps) — what the Rust wrapper does for a long-lived "
                                     "Vec; value_unregistered = the same arrays left pageable: the library stages each chunk through page-locked buffers of "
                                     "its own, filled and emptied by up to 16 host threads while earlier chunks are on the link (224 Mrays/s when the "
                                     "driver staged the copies, profiles/bench_r2w.json)"},
            "gpu_launches": int(launched_in_timed_region),
            "roofline": {"bound": "l2", "binding": "instruction issue at ~20 of 32 active lanes (see profiles/ncu_*: issue slots ~50 % busy over the "
                                                   "launch incl. its tail, DRAM < 10 % and L2 < 15 % of peak); L2 is the memory level that serves the kernel",
                         "achieved": achieved, "peak": l2_peak, "unit": "GB/s", "frac": achieved / l2_peak,
                         "peak_source": "measured in this run: uint4 reads bypassing L1 over a 64 MiB L2-resident buffer, 50 sweeps "
                                        "(rdn_rt_measure_l2_read_gbs); MEASURED_PEAKS.json holds no L2 figure",
                         "frac_with_launch_overlap": bytes_per_ray * n / (ms_per_step * 1e-3) / 1e9 / l2_peak,
                         "traffic": traffic, "traffic_l2": traffic_l2, "traffic_source": traffic_src,
                         "actual": None if traffic is None else {
                             "l2_gbs": traffic_l2 * per_s, "l2_frac": traffic_l2 * per_s / l2_peak,
                             "hbm_gbs": traffic * per_s, "hbm_frac": traffic * per_s / peak,
                             "note": "bytes the kernel actually moves per launch (ncu lts__t_sectors x 32, dram__bytes_read + _write) over the "
                                     "serialised kernel time measured here"},
                         "hbm": {"peak": peak, "peak_source": peak_src, "frac_of_reference_bytes": achieved / peak,
                                 "compulsory_gbs": compulsory},
                         "algorithmic_bytes_per_ray": bytes_per_ray, "algorithmic_bytes_per_launch": bytes_per_ray * n,
                         "visits_per_ray": visits, "bytes_model_sample_rays": n_sample,
                         "kernel": SHIPPED_ORDERED_KERNEL, "kernel_ms_avg": kernel_ms, "kernel_launches_timed": kernel_times["ordered_launches"],
                         "kernel_share_of_step": kernel_ms / ms_per_step_serial, "tie_kernel_ms_avg": tie_ms, "tie_rays_per_step": tie_rays,
                         "note": "algorithmic bytes are defined on the REFERENCE's traversal (48 B threaded nodes in pre-order, 52 B triangle "
                                 "chains); the ordered kernel visits about half the nodes through 64 B two-box nodes and most of its fetches "
                                 "hit L1 / L2, so the reference-defined figure is several times what moves"},
            "cpu_baseline": {"value": cpu_rays / t_cpu / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
                             "sample": f"{cpu_reps} x {N_FRAMES} full frames ({cpu_rays} rays), {t_cpu:.2f} s wall on {cores} threads "
                                       f"= {t_cpu * cores:.1f} core-seconds",
                             "single_thread_mrays": one.shape[0] / t_cpu1 / 1e6, "parity_bit_identical_full_frames": parity_bits,
                             "timed_loop_results_identical_to_synchronised_launches_all_ranks": timed_loop_ok_all_ranks},
            "clocks": clocks.summary(), "wall_s_timed_region": t_wall, "c5": c5, "other_configs": other,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c5"], help="c2: BASELINE configs[1] per GPU (the headline, weak scaling; "
                    "configs[4] rides along as the line's `c5` object).  c5: BASELINE configs[4] is the line (strong scaling of one frame)")
    ap.add_argument("--other-configs", type=int, default=1, help="1: also time BASELINE configs[0], [2], [3] on this GPU after the headline (N = 1 only)")
    ap.add_argument("--c5-steps", type=int, default=3, help="frames of configs[4] timed after the headline loops (0: skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.config == "c5":
        run_c5(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
