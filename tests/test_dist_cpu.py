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


def _worker_chain(rank, world, port, units, q):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mcp_context_forge_b200 import engine
    from oracle import hook_chain_ref as ref, toon_ref

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class OracleShard(cfd.ShardedChain):       # the per-rank stage is the CPU oracle here: the plumbing around it is what is under test
        def _local(self, us, stage_mask, unit_stages, toon_flags):
            v = np.zeros(len(us), dtype=engine.VERDICT_DTYPE)
            outs = []
            bms = ref.scan_bitmaps(us, [(p, re.I) for pats in ref.DEFAULT_LEXICONS.values() for p in pats], ["crap"], [])
            for i, u in enumerate(us):
                v["match_bitmap"][i] = bms[i]
                t = toon_ref.process_text(u, 0, 1 << 30)
                if t is not None:
                    v["flags"][i] = 2
                    v["out_len"][i] = len(t.encode())
                    outs.append(t.encode())
                else:
                    outs.append(b"")
            oo = np.zeros(len(us) + 1, dtype=np.uint64)
            np.cumsum([len(o) for o in outs], out=oo[1:])
            return v, np.frombuffer(b"".join(outs), dtype=np.uint8), oo

    sc = OracleShard(None, device=None)
    parts = sc.partition([len(u.encode()) for u in units])
    full, mine, out, oo = sc.run(units, parts, 9)
    if rank == 0:
        q.put((full["match_bitmap"].tolist(), full["flags"].tolist(), full["out_len"].tolist(), [list(p) for p in parts]))
    dist.destroy_process_group()


def test_sharded_chain_two_ranks_mixed_sizes():
    """BASELINE configs[3] plumbing: a mixed 2 / 16 / 256 KiB batch, size-balanced across two ranks, verdict records of every
    unit on every rank after ONE all-gather."""
    from mcp_context_forge_b200 import synth
    from oracle import hook_chain_ref as ref, toon_ref

    units = [synth.payload("A", 2048, seed=i) for i in range(14)] + [synth.payload("A", 16384, seed=i) for i in range(5)] + [synth.payload("A", 262144, seed=1)]
    units[3] = "crap " + units[3]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_chain, args=(r, 2, port, units, q)) for r in range(2)]
    for p in procs:
        p.start()
    bms, flags, lens, parts = q.get(timeout=300)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    exp = [toon_ref.process_text(u, 0, 1 << 30) for u in units]
    assert flags == [2 if e is not None else 0 for e in exp]
    assert lens == [len(e.encode()) if e is not None else 0 for e in exp]
    assert bms[3] != 0 and sum(1 for b in bms if b) == 1
    loads = [sum(len(units[i]) for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 262144 and sorted(i for p in parts for i in p) == list(range(len(units)))
