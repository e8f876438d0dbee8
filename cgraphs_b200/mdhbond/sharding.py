"""Frame sharding across ranks and the bond-key merge.

Frames are independent (reference hbond.py:103,238 keep no cross-frame state except OR-ing into the
per-key vectors), so rank g processes one contiguous block of frames whose boundaries are multiples of
32 frames - packed occupancy words are never shared between ranks.  The only exchange is at the end:
one all-gather of each rank's sorted distinct bond keys (NCCL over NVLink when the process group is
NCCL, gloo in the CPU tests) makes bond ids global, then the frame-sharded word columns are gathered so
every rank holds the complete matrix, exactly as a single-GPU run would.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames, rank, world):
    """Frame block [lo, hi) of `rank`; boundaries are multiples of 32."""
    n_words = (n_frames + 31) // 32
    lo = 32 * ((rank * n_words) // world)
    hi = 32 * (((rank + 1) * n_words) // world)
    return min(lo, n_frames), min(hi, n_frames)


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("shard=(rank, world) with world > 1 needs an initialised torch.distributed group")
    return dist


_GATHER_CAP = {}      # per process group: keys per rank the gather buffer holds (grows, identical on all ranks)


def merge_keys_device(local_keys, group=None, ctx=None, sync=True):
    """All-gather sorted distinct keys held in a device tensor (int64 view of the uint64 keys); returns the global
    sorted distinct key list as a device tensor.

    One all-gather of a fixed-capacity block per rank: slot 0 carries the rank's key count, the rest its keys
    (real keys are non-negative as int64: group ids < 2^31). Every rank sees every count, so if some rank did not
    fit, all ranks agree to repeat with a larger capacity - no separate size exchange.

    With `ctx` (a _native.Context on the same device) the merge runs in libhbgpu's k_merge_keys on the context's
    compute stream; with sync=False nothing waits on the host: the result is (padded keys, info) where info is a
    device tensor [n_keys, overflow] - the caller checks it later (used inside timed loops once the capacity has
    been learnt by a synchronous call)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    dev = local_keys.device
    n = int(local_keys.numel())
    cap = _GATHER_CAP.get(group, 2048)
    while True:
        buf = torch.empty((cap + 1,), dtype=torch.int64, device=dev)
        buf[0].fill_(n)
        m = min(n, cap)
        if m:
            buf[1:1 + m] = local_keys[:m]
        gathered = torch.empty((world, cap + 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gathered, buf, group=group)      # the one key exchange
        if ctx is not None and world * cap <= 16384:
            out = torch.empty((world * cap,), dtype=torch.int64, device=dev)
            info = torch.empty((2,), dtype=torch.int64, device=dev)
            ctx.merge_keys_device(gathered.data_ptr(), world, cap, out.data_ptr(), info.data_ptr())
            if not sync:
                return out, info
            n_out, over = info.tolist()
            if not over:
                return out[:n_out]
            counts = gathered[:, 0].tolist()
        else:
            counts = gathered[:, 0].tolist()
            if max(counts) <= cap:
                valid = torch.arange(cap, device=dev)[None, :] < gathered[:, :1]
                return torch.unique(gathered[:, 1:][valid])
        cap = 1 << (int(max(counts)) - 1).bit_length()
        _GATHER_CAP[group] = cap


def exchange_keys_queued(ctx, group=None):
    """The whole key exchange queued on the context's compute stream, nothing waits on the host:
    hb_pack_keys_device (this rank's [count, keys...] block straight from the queued finalisation) -> one
    ncclAllGather -> hb_merge_keys_device. Call between ctx.finalize_async() and ctx.finalize(); torch's current
    stream must be the context's compute stream. Returns (keys, info): padded global sorted distinct keys and the
    device tensor [n_keys, overflow] to be checked after the next synchronisation (overflow: repeat with the
    synchronous merge_keys_device, which learns a larger capacity)."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    dev = torch.device("cuda", ctx.device)
    cap = _GATHER_CAP.get(group, 2048)
    if world * cap > 16384:
        raise RuntimeError("exchange_keys_queued: more than 16384 gathered keys; use merge_keys_device")
    block = torch.empty((cap + 1,), dtype=torch.int64, device=dev)
    ctx.pack_keys_device(block.data_ptr(), cap)
    gathered = torch.empty((world, cap + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(gathered, block, group=group)        # the one key exchange
    out = torch.empty((world * cap,), dtype=torch.int64, device=dev)
    info = torch.empty((2,), dtype=torch.int64, device=dev)
    ctx.merge_keys_device(gathered.data_ptr(), world, cap, out.data_ptr(), info.data_ptr())
    return out, info


def merge_keys(local_keys, device=None, group=None):
    """All-gather sorted distinct uint64 keys; returns the global sorted distinct key list (numpy uint64).

    With an NCCL group the gather runs on `device` (keys travel GPU to GPU); with gloo on the CPU.
    """
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device if device is not None else torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    mine = torch.from_numpy(np.ascontiguousarray(local_keys).view(np.int64)).to(dev)
    if on_gpu:
        return merge_keys_device(mine, group).cpu().numpy().view(np.uint64)
    count = torch.tensor([mine.numel()], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    padded[:mine.numel()] = mine
    gathered = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(gathered, padded, group=group)            # the one key exchange
    allk = torch.cat([g[:c] for g, c in zip(gathered, counts)])
    allk = allk.cpu().numpy().view(np.uint64)
    return np.unique(allk)


def merge_shards(keys, words, lo, hi, n_frames, shard, ctx=None, group=None):
    """Combine per-rank results into the global (keys, words) every rank returns.

    keys: this rank's sorted distinct keys; words: uint32 [len(keys), ceil((hi-lo)/32)].
    """
    import torch
    dist = _dist()
    rank, world = shard
    device = getattr(ctx, "device", None)
    gkeys = merge_keys(keys, device=device, group=group)
    n_words_total = (n_frames + 31) // 32
    spans = [shard_range(n_frames, r, world) for r in range(world)]
    widths = [((b - a) + 31) // 32 for a, b in spans]
    wmax = max(max(widths), 1)
    # my word columns, rows aligned to the global key list, padded to the widest shard
    local = np.zeros((len(gkeys), wmax), dtype=np.uint32)
    if len(keys):
        local[np.searchsorted(gkeys, keys), :words.shape[1]] = words
    on_gpu = dist.get_backend(group) == "nccl"
    dev = torch.device("cuda", device if device is not None else torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    t = torch.from_numpy(local.view(np.int32)).to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    out = np.zeros((len(gkeys), n_words_total), dtype=np.uint32)
    for (a, b), w, part in zip(spans, widths, parts):
        if w:
            out[:, a // 32:a // 32 + w] = part[:, :w].cpu().numpy().view(np.uint32)
    return gkeys, out


def merge_wire_shards(keys, cols, first_direct, lo, hi, n_frames, group=None):
    """Water-wire results of frame block [lo, hi) of this rank -> the complete result on every rank.

    keys: sorted distinct uint64 residue-pair keys; cols: uint8 [n, hi - lo] per-frame bytes; first_direct: uint64 [n]
    with GLOBAL frame numbers (waterwire.split_wire_rows). One object all-gather (the result of a finished run, not a
    data-path collective); the global key list is the sorted union, columns are placed at their frame block and the
    first-direct word is the minimum over ranks."""
    dist = _dist()
    world = dist.get_world_size(group)
    parts = [None] * world
    dist.all_gather_object(parts, (int(lo), int(hi), np.asarray(keys), np.asarray(cols), np.asarray(first_direct)), group=group)
    all_keys = np.unique(np.concatenate([p[2] for p in parts])) if parts else np.zeros(0, np.uint64)
    out_cols = np.zeros((len(all_keys), n_frames), dtype=np.uint8)
    out_fd = np.full(len(all_keys), np.uint64(0xFFFFFFFFFFFFFFFF), dtype=np.uint64)
    for p_lo, p_hi, p_keys, p_cols, p_fd in parts:
        if len(p_keys) == 0:
            continue
        idx = np.searchsorted(all_keys, p_keys)
        out_cols[idx, p_lo:p_hi] = p_cols
        out_fd[idx] = np.minimum(out_fd[idx], p_fd)
    return all_keys, out_cols, out_fd
