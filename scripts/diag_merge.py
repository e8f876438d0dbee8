#!/usr/bin/env python
"""Diagnostic (torchrun): time the bond-key merge alone on N ranks."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgraphs_b200.mdhbond.sharding import merge_keys_device
rank = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
k = torch.arange(1507, dtype=torch.int64, device=dev) * 7 + rank
for _ in range(5):
    merge_keys_device(k)
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(20):
    out = merge_keys_device(k)
e1.record(); torch.cuda.synchronize()
print("rank", rank, "merge ms (events)", e0.elapsed_time(e1) / 20, "wall", 1e3 * (time.perf_counter() - t0) / 20, "keys", out.numel())
# pieces
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    buf = torch.full((2049,), -1, dtype=torch.int64, device=dev); buf[0] = 1507; buf[1:1508] = k
torch.cuda.synchronize(); print("rank", rank, "pack ms", 1e3 * (time.perf_counter() - t0) / 20)
g = torch.empty((dist.get_world_size(), 2049), dtype=torch.int64, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    dist.all_gather_into_tensor(g, buf)
torch.cuda.synchronize(); print("rank", rank, "all_gather ms", 1e3 * (time.perf_counter() - t0) / 20)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    c = g[:, 0].tolist()
torch.cuda.synchronize(); print("rank", rank, "tolist ms", 1e3 * (time.perf_counter() - t0) / 20)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    u = torch.unique(g[:, 1:].reshape(-1))
torch.cuda.synchronize(); print("rank", rank, "unique ms", 1e3 * (time.perf_counter() - t0) / 20)
dist.destroy_process_group()
