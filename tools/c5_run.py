"""BASELINE configs[4]: 3840x2160 x 16 spp primary + 1 cosine bounce vs the 1M-triangle torus, ray tiles (512 x 512) dealt round-robin
to the ranks, BVH replicated with one NCCL broadcast.  Everything between the camera parameters and the final hit records stays on
the device (ray generation, traversal, compaction, bounce generation: rendiation_b200/csrc/{raygen,traverse,compact}.cu).

    python tools/c5_run.py [--spp 16] [--check]                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_run.py

Prints one JSON line on rank 0: whole-job Mrays/s (primary + bounce rays of all ranks / max-over-ranks device time).  --check
copies the first owned tile of the last sample back and compares both waves with the oracle bit for bit.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

W, H, SEG = 3840, 2160, 708
TMIN, TMAX = 0.01, 100.0
CULL_BACK = 0x10


def main():
    import torch
    import torch.distributed as dist

    from rendiation_b200 import api, multi_gpu as mg, scenes as S

    ap = argparse.ArgumentParser()
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--tile", type=int, default=512, help="sharding tile edge (512 = the reference's launch tile; smaller balances better)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's own banner ("NCCL version ...") must not land on stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    pos, idx = S.torus_mesh(SEG, SEG, 1.0, 0.35)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    sysm = api.NaiveSahBVHSystem(devices=(local,))
    t_build = 0.0
    if rank == 0:
        t0 = time.perf_counter()
        b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
        sysm.bind_tlas([sysm.create_top_level_acceleration_structure(S.make_instance(m, b.id))])
        sysm.commit()
        t_build = time.perf_counter() - t0
    t_repl = mg.replicate_scene(sysm, src=0, device=dev) if world > 1 else 0.0

    # this rank's tiles, grouped by width so that each group is one 2-D launch (tiles stacked vertically)
    shard = mg.TileShard(W, H, world, rank, tile=args.tile)
    groups = {}
    for t in shard.tiles:
        groups.setdefault(t[2], []).append(t)
    # all spp samples of a group form ONE wave (sample after sample, tiles stacked vertically): one traversal launch, one
    # compaction + bounce generation, one wave-size read-back and one bounce launch per group instead of one of each per sample
    spp = args.spp
    bufs = {}
    for gw, tiles in groups.items():
        n1 = sum(t[2] * t[3] for t in tiles)
        n = n1 * spp
        bufs[gw] = dict(n=n, n1=n1, tiles=tiles, rays=torch.empty((n, 32), dtype=torch.uint8, device=dev), hits=torch.empty((n, 32), dtype=torch.uint8, device=dev),
                        brays=torch.empty((n, 32), dtype=torch.uint8, device=dev), bhits=torch.empty((n, 32), dtype=torch.uint8, device=dev),
                        src=torch.empty(n, dtype=torch.int32, device=dev), cnt=torch.zeros(1, dtype=torch.int64, device=dev))
    st = torch.cuda.current_stream().cuda_stream
    aspect = float(np.float32(W / H))
    jitters = [S.sample_2d(np.full(1, s, np.uint32))[0] if s else np.array([0.5, 0.5], np.float32) for s in range(spp)]

    def frame():
        n_bounce = 0
        for gw, B in bufs.items():
            rects = [t for s in range(spp) for t in B["tiles"]]
            jits = [(float(jitters[s][0]), float(jitters[s][1])) for s in range(spp) for _ in B["tiles"]]
            sysm.gen_pinhole_rays_batch_device(B["rays"].data_ptr(), W, H, rects, jits, tmin=TMIN, tmax=TMAX, aspect=aspect, stream=st)
            sysm.trace_closest_device(B["rays"].data_ptr(), B["n"], B["hits"].data_ptr(), ray_flags=CULL_BACK, grid_width=gw, stream=st)
            sysm.gen_bounce_rays_device(B["rays"].data_ptr(), B["hits"].data_ptr(), B["n"], B["brays"].data_ptr(), B["src"].data_ptr(),
                                        B["cnt"].data_ptr(), mode=0, index_base=0, tmin=TMIN, tmax=TMAX, stream=st)
        for gw, B in bufs.items():
            k = int(B["cnt"].item())  # wave size read back, as the reference does after its compaction (task_group.rs:259-277)
            sysm.trace_closest_device(B["brays"].data_ptr(), k, B["bhits"].data_ptr(), ray_flags=0, stream=st)
            B["k"] = k
            n_bounce += k
        return n_bounce

    frame()  # warm-up (allocates scratch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_bounce = frame()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    counts = torch.tensor([shard.n_rays * spp, n_bounce], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)

    check = None
    if args.check:
        import oracle
        osc = oracle.Scene()
        ob = osc.create_blas([(pos, idx.reshape(-1), 1)])
        osc.bind_tlas([osc.create_tlas(S.make_instance(m, ob))])
        assert osc.build() == 0
        gw, B = next(iter(bufs.items()))
        n1 = B["tiles"][0][2] * B["tiles"][0][3]  # first owned tile of sample 0
        cores = os.cpu_count() or 1
        r1 = B["rays"][:n1].cpu().numpy().view(S.RAY_DTYPE).reshape(-1)
        h1 = B["hits"][:n1].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
        k1 = min(B["k"], 1 << 18)
        r2 = B["brays"][:k1].cpu().numpy().view(S.RAY_DTYPE).reshape(-1)
        h2 = B["bhits"][:k1].cpu().numpy().view(api.HIT_DTYPE).reshape(-1)
        ok1 = h1.tobytes() == osc.trace(r1, ray_flags=CULL_BACK, n_threads=cores, want_counters=False).tobytes()
        ok2 = h2.tobytes() == osc.trace(r2, ray_flags=0, n_threads=cores, want_counters=False).tobytes()
        flags = torch.tensor([float(ok1), float(ok2)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        check = {"primary_tile_bit_identical_all_ranks": bool(flags[0].item()), "bounce_bit_identical_all_ranks": bool(flags[1].item()),
                 "rays_checked_per_rank": int(n1 + k1)}

    if rank == 0:
        total = float(counts.sum().item())
        print(json.dumps({"config": "BASELINE configs[4]: 3840x2160 x %d spp primary + 1 bounce, %dx%d tiles round-robin over %d GPU(s)" % (args.spp, args.tile, args.tile, world),
                          "metric": "closest-hit Mrays/s", "value": total / (float(ms.item()) * 1e-3) / 1e6, "unit": "Mrays/s", "n_gpus": world,
                          "primary_rays": int(counts[0].item()), "bounce_rays": int(counts[1].item()), "ms_total": float(ms.item()),
                          "tile": args.tile, "tiles_total": len(shard.tiles_all), "tiles_rank0": len(shard.tiles), "blob_broadcast_ms": t_repl, "build_s": round(t_build, 3),
                          "includes": "device ray generation, traversal, compaction, bounce generation, per-wave size read-back", "check": check}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
