"""Multi-GPU plumbing for the one exchange step of the path.

Units are (clip b, segment s) pairs: each needs one CLIP key-frame and one 8-frame InternVideo2 segment and yields a
self-contained [tokens_per_seg, D] block (llava_next_video.py:503-564 -- nothing crosses segments before the LLM).
Units are block-partitioned over ranks, encoded locally, and exchanged with ONE collective of the projected visual
tokens (NCCL over NVLink on GPUs; gloo in the CPU tests). After the exchange the LLM work is clip-sharded.

Two forms of the exchange:
  allgather_units      every rank ends up with every unit -- what `encode_images` returns in the reference ([B, 3420, D]);
  exchange_units       every rank receives only the units of the clips IT decodes (clip b -> rank b % world): the same single
                       collective as an all-to-all with per-peer split sizes; at 1 clip per GPU with block partitioning nothing
                       leaves the GPU at all. Used by `generate`.
Collectives run on a side stream fenced with events (SURVEY 8b threading contract), results of `generate` come back as ONE
device-side int64 all-gather (`gather_tokens`), not as pickled python objects.
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


def _unit_owner(n_units, ws):
    owner = []
    for r, (_, c) in enumerate(partition_units(n_units, ws)):
        owner += [r] * c
    return owner


def exchange_plan(n_clips, segs, ws):
    """Static plan of the all-to-all: for every (src, dst) the list of global unit ids src sends to dst, in unit order.
    Unit u = b * segs + s is encoded on its block-partition owner and consumed on rank b % ws."""
    owner = _unit_owner(n_clips * segs, ws)
    plan = [[[] for _ in range(ws)] for _ in range(ws)]
    for u, src in enumerate(owner):
        plan[src][(u // segs) % ws].append(u)
    return plan


_side = {}


def _side_stream(device):
    key = str(device)
    if key not in _side:
        _side[key] = torch.cuda.Stream(device=device)
    return _side[key]


def _on_side_stream(fn, *tensors):
    """Run a collective on the dedicated side stream, fenced against the caller's stream on both sides."""
    t0 = tensors[0]
    if not t0.is_cuda:
        return fn()
    cur = torch.cuda.current_stream(t0.device)
    side = _side_stream(t0.device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        out = fn()
    cur.wait_stream(side)
    for t in tensors:
        t.record_stream(side)
    return out


def exchange_units(local_block, n_clips, segs, group=None):
    """local_block: [units_local, T, D] (this rank's block-partition units, global unit order).
    Returns (feats [n_mine, segs*T, D], mine): the visual tokens of the clips this rank decodes (clips_for_rank order)."""
    rank, ws = world()
    T, D = local_block.shape[1], local_block.shape[2]
    mine = clips_for_rank(n_clips, rank, ws)
    if ws == 1:
        return local_block.reshape(n_clips, segs * T, D), mine
    plan = exchange_plan(n_clips, segs, ws)
    start = partition_units(n_clips * segs, ws)[rank][0]
    send_ids = [u - start for dst in range(ws) for u in plan[rank][dst]]
    in_splits = [len(plan[rank][dst]) for dst in range(ws)]
    out_splits = [len(plan[src][rank]) for src in range(ws)]
    recv_ids = [u for src in range(ws) for u in plan[src][rank]]
    if all(not plan[src][dst] for src in range(ws) for dst in range(ws) if src != dst):
        got = local_block                                   # global decision (same plan on every rank): nothing crosses ranks,
    else:                                                   # e.g. 1 clip per GPU -- every unit is decoded where it was encoded
        send = local_block if send_ids == list(range(local_block.shape[0])) else local_block[torch.tensor(send_ids, device=local_block.device, dtype=torch.long)]
        send = send.contiguous()
        got = torch.empty((sum(out_splits), T, D), dtype=send.dtype, device=send.device)
        _on_side_stream(lambda: dist.all_to_all_single(got, send, out_splits, in_splits, group=group), send, got)
    # received in (src, unit) order; units of one clip are consecutive within a source and sources are ordered by unit id
    order = sorted(range(len(recv_ids)), key=lambda i: recv_ids[i])
    if order != list(range(len(order))):
        got = got[torch.tensor(order, device=got.device, dtype=torch.long)]
    return got.reshape(len(mine), segs * T, D), mine


def gather_tokens(local, n_clips, width, pad_id, device, group=None):
    """local: {clip index: int64 device tensor [<= width]} of the clips this rank decoded. ONE device-side all-gather of a
    [clips_per_rank, 1 + width] int64 block (column 0 = length). Returns a list of n_clips int64 tensors (on `device`)."""
    rank, ws = world()
    per = (n_clips + ws - 1) // ws
    blk = torch.full((per, 1 + width), int(pad_id), dtype=torch.int64, device=device)
    blk[:, 0] = -1
    for i, b in enumerate(clips_for_rank(n_clips, rank, ws)):
        t = local[b]
        blk[i, 0] = t.shape[0]
        blk[i, 1:1 + t.shape[0]] = t
    if ws == 1:
        allb = blk[None]
    else:
        allb = torch.empty((ws, per, 1 + width), dtype=torch.int64, device=device)
        _on_side_stream(lambda: dist.all_gather_into_tensor(allb.view(ws * per, 1 + width), blk, group=group), blk, allb)
    lens = allb[:, :, 0].cpu()                              # the one host read of the step (lengths; the tokens stay on the device)
    out = []
    for b in range(n_clips):
        r, i = b % ws, b // ws
        out.append(allb[r, i, 1:1 + int(lens[r, i])])
    return out
