"""TEST INFRASTRUCTURE: the N > 1 launch path on two gloo ranks with the emulated kernels standing in for two GPUs — every
rank builds the scene, traces ITS tiles of the launch through the product's host-buffer call (multi_gpu.trace_shard: one 2-D
launch per tile, emulated k_trace_ordered_rounds underneath), rank 0 assembles the frame (multi_gpu.gather_launch_hits) and
compares it with the oracle.  Run by tests/test_simt_emulation.py; exits non-zero on any mismatch."""
import os
import socket
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def rank_main(rank: int, world: int, port: int, W: int, H: int, tile: int):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ.setdefault("RDN_SIMT_THREADS", "2")
    import numpy as np
    import torch.distributed as dist

    import build_emu
    from rendiation_b200 import api, multi_gpu as mg, scenes as S
    api.LIB_PATH = build_emu.build()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pos, idx = S.torus_mesh(64, 48, 1.0, 0.35)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    sysm = api.NaiveSahBVHSystem(devices=(0,))
    b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
    sysm.bind_tlas([sysm.create_top_level_acceleration_structure(S.make_instance(m, b.id))])
    sysm.commit()
    rays = S.pinhole_rays(W, H, 0.01, 100.0)
    shard = mg.TileShard(W, H, world, rank, tile=tile)
    hits = mg.trace_shard(sysm, shard, shard.gather_rays(rays), ray_flags=api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES)
    full = mg.gather_launch_hits(shard, hits, dst=0)
    if rank == 0:
        import oracle
        oracle.build()
        osc = oracle.Scene()
        ob = osc.create_blas([(pos, idx.reshape(-1), 1)])
        osc.bind_tlas([osc.create_tlas(S.make_instance(m, ob))])
        assert osc.build() == 0
        want = osc.trace(rays, ray_flags=api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, want_counters=False)
        assert full.tobytes() == want.tobytes(), "assembled frame differs from the oracle"
        n_hit = int((full["instance_id"] != api.INVALID_ID).sum())
        assert n_hit > 1000, n_hit
        print(f"two emulated ranks: {W}x{H} rays in {len(mg.launch_tiles(W, H, tile))} tiles, {n_hit} hits, identical to the oracle")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=rank_main, args=(r, 2, port, 160, 128, 32)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    sys.exit(0 if all(p.exitcode == 0 for p in procs) else 1)
