"""Multi-GPU plumbing for the one exchange step of the path.

Units are (clip b, segment s) pairs: each needs one CLIP key-frame and one 8-frame InternVideo2 segment and yields a
self-contained [tokens_per_seg, D] block (llava_next_video.py:503-564 -- nothing crosses segments before the LLM).
Units are block-partitioned over ranks, encoded locally, and exchanged with ONE all-gather of the projected visual
tokens (NCCL over NVLink on GPUs; gloo in the CPU tests). After the exchange the LLM work is clip-sharded.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def partition_units(n_units, world_size):
    """Contiguous block partition; the first (n_units % world_size) ranks get one extra unit.
    Returns list of (start, count) per rank."""
    base, extra = divmod(n_units, world_size)
    out, start = [], 0
    for r in range(world_size):
        cnt = base + (1 if r < extra else 0)
        out.append((start, cnt))
        start += cnt
    return out


def clips_for_rank(n_clips, rank, world_size):
    """LLM work: clip b runs on rank b % world_size (for n_clips < world_size the low ranks run one clip each)."""
    return [b for b in range(n_clips) if b % world_size == rank]


def allgather_units(local_block, n_units, group=None):
    """local_block: [units_local, T, D] (this rank's units, in global unit order). Returns [n_units, T, D] on every rank.
    Uneven tails are handled by padding every rank's block to the maximum unit count (one collective)."""
    rank, ws = world()
    if ws == 1:
        return local_block
    parts = partition_units(n_units, ws)
    max_cnt = max(c for _, c in parts)
    T, D = local_block.shape[1], local_block.shape[2]
    send = local_block
    if send.shape[0] < max_cnt:
        pad = torch.zeros((max_cnt - send.shape[0], T, D), dtype=send.dtype, device=send.device)
        send = torch.cat([send, pad], dim=0)
    send = send.contiguous()
    recv = torch.empty((ws * max_cnt, T, D), dtype=send.dtype, device=send.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(ws, max_cnt, T, D)
    return torch.cat([recv[r, :c] for r, (_, c) in enumerate(parts)], dim=0)


def gather_strings(local, group=None):
    """Collect per-rank python objects (generated texts) on every rank."""
    rank, ws = world()
    if ws == 1:
        return [local]
    out = [None] * ws
    dist.all_gather_object(out, local, group=group)
    return out
