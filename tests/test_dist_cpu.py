"""world_size-2 gloo test of the N>1 host logic: size-balanced partition, padded all-gather of the
verdict bitmaps (the path's only collective), scatter back to global unit order.  The per-rank
"scan" is the oracle here (no GPU): what is under test is the sharding/collective plumbing."""
import os
import re
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mcp_context_forge_b200 import dist as cfd


def test_partition_is_balanced_and_complete():
    sizes = [2048] * 70 + [16384] * 25 + [262144] * 5
    parts = cfd.partition_units(sizes, 4)
    assert sorted(i for p in parts for i in p) == list(range(100))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) - min(loads) <= 262144
    assert all(p == sorted(p) for p in parts)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, units, q):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import hook_chain_ref as ref

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = cfd.partition_units([len(u.encode()) for u in units], world)
    mine = [units[i] for i in parts[rank]]
    pats = [(p, re.I) for v in ref.DEFAULT_LEXICONS.values() for p in v]
    local = torch.tensor(np.array(ref.scan_bitmaps(mine, pats, ["crap"], []), dtype=np.uint64).view(np.int64))
    gathered = cfd.gather_verdicts(local, [len(p) for p in parts], 1)
    full = cfd.scatter_back(parts, gathered, 1, len(units))
    if rank == 0:
        q.put(full.tolist())
    dist.destroy_process_group()


def test_two_rank_gloo_allgather_of_verdicts():
    from oracle import hook_chain_ref as ref

    units = [("kill him " if i % 7 == 0 else "") + ("crap " if i % 5 == 0 else "") + "x" * (10 + 37 * (i % 13)) for i in range(41)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, units, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    pats = [(p, re.I) for v in ref.DEFAULT_LEXICONS.values() for p in v]
    assert got == ref.scan_bitmaps(units, pats, ["crap"], [])
    assert any(got)
