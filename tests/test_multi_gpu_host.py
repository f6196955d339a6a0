"""Host logic of the multi-GPU driver (rendiation_b200/multi_gpu.py) on CPU: exact tile sharding, and — with two gloo ranks —
replication of the flattened scene plus assembly of a launch from per-rank tile shards.  No CUDA here: the ranks hold host-only
scenes and the oracle stands in for the tracer (test infrastructure only), so what is exercised is the sharding, broadcast,
adoption and gather code that the NCCL path shares."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from rendiation_b200 import api, multi_gpu as mg, scenes as S


@pytest.mark.parametrize("wh", [(1920, 1080), (3840, 2160), (513, 7), (1, 1), (512, 512), (1000, 1025)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_tiles_partition_the_launch_exactly(wh, world):
    W, H = wh
    tiles = mg.launch_tiles(W, H)
    assert len(tiles) == -(-W // mg.TILE) * -(-H // mg.TILE)
    seen = np.zeros(W * H, np.int32)
    owned = 0
    for r in range(world):
        sh = mg.TileShard(W, H, world, r)
        idx = sh.indices()
        assert idx.size == sh.n_rays == sum(c for _, c, _ in sh.launches())
        seen[idx] += 1
        owned += len(sh.tile_ids)
    assert owned == len(tiles)
    assert np.all(seen == 1)  # every ray exactly once: nothing dropped (the reference's rect_split_iter drops remainder columns)


def test_tile_order_is_row_major_inside_a_tile():
    sh = mg.TileShard(1030, 600, 2, 1)
    (x0, y0, w, h) = sh.tiles[0]
    idx = mg.tile_ray_indices(1030, (x0, y0, w, h))
    assert idx[0] == y0 * 1030 + x0 and idx[1] == idx[0] + 1 and idx[w] == (y0 + 1) * 1030 + x0
    with pytest.raises(ValueError):
        mg.shard_tiles(4, 2, 2)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _scene(system_or_oracle):
    pos, idx = S.torus_mesh(48, 32, 1.0, 0.35)
    m = S.mat4_mul(S.mat4_mul(S.mat4_translate(0, 0, -10), S.mat4_scale(5, 5, 5)), S.mat4_rotate_x(-0.5))
    return pos, idx, m


def _rank_main(rank: int, world: int, port: int, W: int, H: int, tile: int, errors):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        pos, idx, m = _scene(None)
        sysm = api.NaiveSahBVHSystem(devices=())  # host-only scene: flattener + blob, no CUDA
        if rank == 0:
            b = sysm.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
            sysm.bind_tlas([sysm.create_top_level_acceleration_structure(S.make_instance(m, b.id))])
            sysm.commit()
        ms = mg.replicate_scene(sysm, src=0)
        assert ms >= 0.0
        # every rank now holds the same flattened arrays as a scene built locally
        ref = api.NaiveSahBVHSystem(devices=())
        b = ref.create_bottom_level_acceleration_structure([api.BottomLevelAccelerationStructureBuildSource(pos, idx.reshape(-1))])
        ref.bind_tlas([ref.create_top_level_acceleration_structure(S.make_instance(m, b.id))])
        ref.commit()
        for aid in range(len(api.ARRAYS)):
            assert sysm.array(aid).tobytes() == ref.array(aid).tobytes(), api.ARRAYS[aid][0]
        # an adopted scene refuses to trace without a device (no CPU fallback), like any host-only scene
        with pytest.raises(api.RdnError):
            sysm.trace_closest_batch(S.pinhole_rays(4, 4))

        # tile-sharded launch: each rank resolves its own tiles (the oracle stands in for the GPU), rank 0 assembles
        rays = S.pinhole_rays(W, H, 0.01, 100.0)
        osc = oracle.Scene()
        ob = osc.create_blas([(pos, idx.reshape(-1), 1)])
        osc.bind_tlas([osc.create_tlas(S.make_instance(m, ob))])
        assert osc.build() == 0
        shard = mg.TileShard(W, H, world, rank, tile=tile)
        mine = shard.gather_rays(rays)
        hits = np.empty(shard.n_rays, api.HIT_DTYPE)
        for off, cnt, gw in shard.launches():
            assert cnt % gw == 0
            hits[off:off + cnt] = osc.trace(mine[off:off + cnt], ray_flags=api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, want_counters=False)
        full = mg.gather_launch_hits(shard, hits, dst=0)
        if rank == 0:
            want = osc.trace(rays, ray_flags=api.RAY_FLAG_CULL_BACK_FACING_TRIANGLES, want_counters=False)
            assert full.tobytes() == want.tobytes()
            assert int((full["instance_id"] != api.INVALID_ID).sum()) > 0
        else:
            assert full is None
        dist.barrier()
        dist.destroy_process_group()
    except BaseException as e:  # noqa: BLE001 - surfaced in the parent
        import traceback
        errors.put(f"rank {rank}: {e!r}\n{traceback.format_exc()}")
        raise


@pytest.mark.timeout(300)
def test_two_ranks_replicate_and_assemble_over_gloo():
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, 96, 80, 32, errors)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    assert not msgs, "\n".join(msgs)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
