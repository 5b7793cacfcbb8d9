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
        texts = gdist.gather_strings({b: "clip%d" % b for b in gdist.clips_for_rank(4, rank, ws)})
        merged = {}
        for d in texts:
            merged.update(d)
        ok = ok and merged == {b: "clip%d" % b for b in range(4)}
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
