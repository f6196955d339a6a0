"""Concurrent host<->device bandwidth of all ranks of one box (experiment harness): every rank copies 66 MB pinned buffers to and
from its own GPU at the same time — the ceiling of the host-buffer `e2e` figure at N GPUs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_probe_multi.py"""
import json
import os
import time

import torch
import torch.distributed as dist

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
n = 66355200
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True); h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


out = {"n_gpus": world, "bytes_each_way": n}
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    t = float(dt.item())
    out[name] = {"ms_slowest_rank": t * 1e3, "gbs_per_direction_per_gpu": n / t / 1e9, "gbs_aggregate_per_direction": world * n / t / 1e9,
                 "mrays_per_s_ceiling": world * (n / 32) / t / 1e6}
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
