"""oracle/sbt.py known answers: the record index of config_hit_group and the hit-group formula of the dispatch are two halves
of the same addressing (sbt.rs:20-38 vs api/ctx.rs:53-55) — a record configured for (geometry, tlas_offset, ray type) must be
the one a hit on that geometry of an instance with that record offset selects for that ray type."""
import numpy as np

from oracle import sbt as osbt
from rendiation_b200 import api


def _hits(rows):
    h = np.zeros(len(rows), api.HIT_DTYPE)
    for k, (geom, inst) in enumerate(rows):
        h[k]["geometry_id"] = geom if inst is not None else 0xFFFFFFFF
        h[k]["instance_id"] = 0xFFFFFFFF if inst is None else inst
    return h


def test_addressing_round_trip_and_outcomes():
    ray_types = 2
    t = osbt.ShaderBindingTable(4, 3, ray_types)          # 24 records
    t.config_hit_group(2, 1, 0, closest_hit=5)             # index 0 + 2*2 + 1 = 5
    t.config_hit_group(2, 1, 1, closest_hit=6)             # index 6
    t.config_hit_group(0, 0, 0, closest_hit=None, any_hit=9)
    t.config_missing(1, 4)
    sbt_offset = np.array([1, 0, 100], np.uint32)          # per TLAS slot
    hits = _hits([(2, 0), (2, 1), (0, 1), (0, None), (3, 2)])
    # ray type 0: stride = ray type count, offset = ray type
    task = osbt.dispatch(t, hits, sbt_offset, sbt_ray_offset=0, sbt_ray_stride=ray_types, miss_index=0)
    assert task.tolist() == [5, osbt.TASK_NONE, osbt.TASK_NONE, osbt.TASK_NONE, osbt.TASK_NONE]   # slot 1 has offset 0 -> record 4 (empty)
    task = osbt.dispatch(t, hits, sbt_offset, sbt_ray_offset=1, sbt_ray_stride=ray_types, miss_index=1)
    assert task.tolist() == [6, 5, osbt.TASK_NONE, 4 | osbt.TASK_MISS_BIT, osbt.TASK_NONE]        # record 1+4+0 = 5 for slot 1
    skip = osbt.dispatch(t, hits, sbt_offset, ray_flags=osbt.RAY_FLAG_SKIP_CLOSEST_HIT_SHADER, sbt_ray_offset=1, sbt_ray_stride=2, miss_index=1)
    assert skip.tolist() == [osbt.TASK_NONE] * 3 + [4 | osbt.TASK_MISS_BIT, osbt.TASK_NONE]
    queue, offsets = osbt.group(np.array([1, 0, osbt.TASK_NONE, 1, 0 | osbt.TASK_MISS_BIT, 1, 7], np.uint32), 2, 1)
    assert offsets.tolist() == [0, 1, 4, 5] and queue.tolist() == [1, 0, 3, 5, 4]


def test_constants_match_the_header():
    assert (api.SBT_NO_SHADER, api.TASK_NONE, api.TASK_MISS_BIT, api.RAY_FLAG_SKIP_CLOSEST_HIT_SHADER) == \
        (osbt.NO_SHADER, osbt.TASK_NONE, osbt.TASK_MISS_BIT, osbt.RAY_FLAG_SKIP_CLOSEST_HIT_SHADER)
