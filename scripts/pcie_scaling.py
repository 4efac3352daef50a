#!/usr/bin/env python
"""Host-fabric ceiling of the end-to-end (host-buffer) path at N ranks on one node.

Launched like the bench (torchrun --nproc-per-node N, or plain python for N = 1).  Every rank, bound to the CPUs next to its GPU,
moves 1 GiB blocks between pinned host memory and its GPU -- H2D alone, D2H alone, both at once -- with all ranks running
together (barrier before, max over ranks of the elapsed time), and copies 1 GiB host-to-host (memcpy) the same way.  Rank 0 prints one
JSON line with the per-rank and aggregate GB/s: the aggregate is what the e2e pipeline of bench.py can move per second on this box at
N GPUs, whatever the kernels do."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bind_to_gpu_numa_node  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
lr = int(os.environ.get("LOCAL_RANK", "0"))
cpus = bind_to_gpu_numa_node(lr)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
N = 1 << 30
h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
h_in.fill_(1)
d_in = torch.empty(N, dtype=torch.uint8, device=dev)
d_out = torch.zeros(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=4):
    fn(1)
    sync_all()
    t0 = time.perf_counter()
    fn(reps)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return reps * N / float(t.item()) / 1e9          # GB/s per rank, at the pace of the slowest rank


def copies(h2d, d2h):
    def fn(reps):
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
    return fn


a = np.frombuffer(h_in.numpy(), dtype=np.uint8)
b = np.frombuffer(h_out.numpy(), dtype=np.uint8)


def memcpy(reps):
    for _ in range(reps):
        np.copyto(b, a)


r = dict(n_gpus=world, numa_cpus_per_rank=(len(cpus) if cpus else None), block_bytes=N)
r["h2d_GBs_per_rank"] = timed(copies(True, False))
r["d2h_GBs_per_rank"] = timed(copies(False, True))
r["bidir_each_GBs_per_rank"] = timed(copies(True, True))
r["host_memcpy_GBs_per_rank"] = timed(memcpy, 2)
for k in ("h2d", "d2h", "bidir_each", "host_memcpy"):
    r[f"{k}_GBs_aggregate"] = world * r[f"{k}_GBs_per_rank"]
# what the bench's e2e step needs per rank: 6.04 GB up + 5.37 GB down; the floor of its time on this fabric
r["e2e_floor_ms_bidir"] = max(6.04 / r["bidir_each_GBs_per_rank"], 5.37 / r["bidir_each_GBs_per_rank"]) * 1e3
r["e2e_floor_iters_per_s"] = world * 1e3 / r["e2e_floor_ms_bidir"]
if rank == 0:
    print(json.dumps(r))
if world > 1:
    dist.destroy_process_group()
