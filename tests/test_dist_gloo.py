"""world_size-2 (and 3, uneven tail) CPU tests of the one exchange step on the path (gvl/dist.py) over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gvl import dist as gdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_units, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        T, D = 5, 8
        full = torch.arange(n_units * T * D, dtype=torch.float32).reshape(n_units, T, D).to(torch.bfloat16)
        start, cnt = gdist.partition_units(n_units, ws)[rank]
        got = gdist.allgather_units(full[start:start + cnt].clone(), n_units)
        ok = torch.equal(got, full)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,n_units", [(2, 12), (2, 5), (3, 7), (2, 1)])
def test_allgather_units_gloo(ws, n_units):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_units, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _worker_exchange(rank, ws, port, n_clips, segs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        T, D = 3, 4
        n_units = n_clips * segs
        full = torch.arange(n_units * T * D, dtype=torch.float32).reshape(n_units, T, D).to(torch.bfloat16)
        start, cnt = gdist.partition_units(n_units, ws)[rank]
        feats, mine = gdist.exchange_units(full[start:start + cnt].clone(), n_clips, segs)
        want = full.reshape(n_clips, segs * T, D)
        ok = mine == gdist.clips_for_rank(n_clips, rank, ws) and feats.shape[0] == len(mine)
        ok = ok and all(torch.equal(feats[i], want[b]) for i, b in enumerate(mine))
        # results: variable-length token rows decoded on their owners, ONE int64 all-gather
        local = {b: torch.arange(b + 1, dtype=torch.int64) + 100 * b for b in mine}
        toks = gdist.gather_tokens(local, n_clips, 8, 0, "cpu")
        ok = ok and all(torch.equal(toks[b], torch.arange(b + 1, dtype=torch.int64) + 100 * b) for b in range(n_clips))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,n_clips,segs", [(2, 2, 12), (2, 3, 4), (3, 4, 5), (2, 1, 12), (3, 2, 2), (2, 8, 3)])
def test_exchange_units_and_gather_tokens_gloo(ws, n_clips, segs):
    """The all-to-all form of the exchange (only the clips a rank decodes travel to it) and the device-side token gather:
    1 clip per rank (nothing crosses ranks, no collective issued), uneven tails, fewer clips than ranks."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_exchange, args=(r, ws, port, n_clips, segs, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_exchange_plan_covers_every_unit_once():
    for ws, n_clips, segs in ((8, 32, 12), (8, 1, 12), (4, 6, 12), (3, 7, 5)):
        plan = gdist.exchange_plan(n_clips, segs, ws)
        seen = sorted(u for src in range(ws) for dst in range(ws) for u in plan[src][dst])
        assert seen == list(range(n_clips * segs))
        for dst in range(ws):
            for src in range(ws):
                assert all((u // segs) % ws == dst for u in plan[src][dst])
