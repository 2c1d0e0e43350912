"""Multi-GPU check of dist.ShardedChain (SURVEY §8(e) / `cf_run_batch_sharded`) on a mixed 2 / 16 / 256 KiB batch:
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_chain_check.py
Every rank partitions the same batch (dist.partition_units), runs the fused chain on its shard on its GPU, all-gathers the verdict
records over NCCL; rank 0 compares the gathered verdicts and every owner's texts with the oracle.  Prints one JSON line."""
import json
import os
import re
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mcp_context_forge_b200 import engine, synth  # noqa: E402
from mcp_context_forge_b200._native import CF_STAGE_SCAN, CF_STAGE_SUB, CF_STAGE_TOON, CF_V_REWRITTEN, CF_V_TOON  # noqa: E402
from mcp_context_forge_b200.dist import ShardedChain  # noqa: E402
from oracle import hook_chain_ref as ref, toon_ref  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
prog = engine.Program()
pats = [(p, re.I) for ps in ref.DEFAULT_LEXICONS.values() for p in ps]
for p, f in pats:
    prog.add_search(p, f)
SUBS = [("crap", 0, "crud"), ("crud", 0, "yikes")]
for s_, f_, r_ in SUBS:
    prog.add_sub(s_, f_, r_)
# BASELINE configs[3] sizes: 2 / 16 / 256 KiB tabular payloads (80 % / 19 % / 1 % by count), hits injected
N = 600
units = [synth.payload("A", 262144 if i % 100 == 0 else 16384 if i % 5 == 0 else 2048, seed=i, hit_rate=2e-3 if i % 7 == 0 else 0.0) for i in range(N)]
units = [u.replace("lorem", "crap", 2) if i % 11 == 0 else u.replace("ipsum", "kill him", 1) if i % 13 == 0 else u for i, u in enumerate(units)]
sc = ShardedChain(prog, device=local)
parts = sc.partition([len(u.encode()) for u in units])
t0 = time.perf_counter()
verdicts, mine, out, oo = sc.run(units, parts, CF_STAGE_SCAN | CF_STAGE_SUB | CF_STAGE_TOON)
dt = time.perf_counter() - t0
# each owner checks its texts, rank 0 checks every verdict
rules = ref.regex_compile_rules([{"search": s_, "replace": r_} for s_, _, r_ in SUBS])
bad = 0
for k, i in enumerate(mine):
    txt = out[int(oo[k]):int(oo[k + 1])].tobytes().decode()
    fl = int(verdicts[i]["flags"])
    if fl & CF_V_REWRITTEN:
        bad += txt != ref.regex_apply_str(rules, units[i])
    elif fl & CF_V_TOON:
        bad += txt != toon_ref.process_text(units[i])
    else:
        bad += bool(txt) or toon_ref.process_text(units[i]) is not None
if rank == 0:
    exp = ref.scan_bitmaps(units, pats, [], [(s_, f_) for s_, f_, _ in SUBS])
    bad += int(sum(int(v) != e for v, e in zip(verdicts["match_bitmap"], exp)))
t = torch.tensor([bad], dtype=torch.int64, device="cuda")
dist.all_reduce(t)
sizes = [sum(len(units[i].encode()) for i in p) for p in parts]
if rank == 0:
    print(json.dumps({"check": "ShardedChain on a mixed 2/16/256 KiB batch", "world": world, "units": N, "mismatches": int(t.item()),
                      "shard_units": [len(p) for p in parts], "shard_bytes": sizes, "imbalance": round(max(sizes) / (sum(sizes) / world), 3),
                      "rewritten": int(((verdicts["flags"] & CF_V_REWRITTEN) != 0).sum()), "toon": int(((verdicts["flags"] & CF_V_TOON) != 0).sum()),
                      "first_call_s": round(dt, 3)}))
dist.destroy_process_group()
sys.exit(1 if t.item() else 0)
