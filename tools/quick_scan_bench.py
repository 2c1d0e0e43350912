"""Quick device-resident timing of the scan kernel (development aid; bench.py is the contract)."""
import re
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from mcp_context_forge_b200 import engine, synth
from oracle import hook_chain_ref as ref

# usage: quick_scan_bench.py [units] [payload_bytes] [shape A|B|C] [hit_rate]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
size = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
shape = sys.argv[3] if len(sys.argv) > 3 else "A"
hit = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
ctx = engine.Context.get()
p = engine.Program()
for pats in ref.DEFAULT_LEXICONS.values():
    for pat in pats:
        p.add_search(pat, re.I)
for w in ["innovative", "groundbreaking", "revolutionary"]:
    p.add_literal(w)
# CF_QS_EXTRA=N adds N more deny words that do not occur in the payloads (large-rule-set probe)
import os
RARE = """xylophone quagmire zephyr jackhammer blizzard buzzword quizzical jukebox wizardry zigzagging
kumquat vortex pixelate fjord gazebo haphazard ivory jinx kiosk larynx mnemonic nymph oxygen pajama
quartz rhythm sphinx topaz unzip vixen waltz yacht zodiac abyss bayou crypt dwarves espionage fishhook
galaxy hyphen icebox jaundice keyhole luxury matrix nightclub ovary pneumonia queue rickshaw strength
transcript uptown vaporize whiskey xenon youthful zombie absurd bikini cobweb duplex embezzle fluffy
glowworm hymn injury jogging knapsack lucky microwave numbskull onyx peekaboo quorum razzmatazz subway
thumbscrew unknown voodoo wave wheezy yippee zilch askew bagpipes cycle disavow equip frizzled gossip""".split()
for w in RARE[:int(os.environ.get("CF_QS_EXTRA", "0"))]:
    p.add_literal(w)
p.add_sub("crap", 0, "crud")
p.add_sub("crud", 0, "yikes")
p.compile(ctx)
base = [synth.payload(shape, size, seed=s, hit_rate=hit) for s in range(32)]
units = [base[i % 32] for i in range(n)]
stream, offs = engine.pack_units(units)
print("stream bytes", len(stream), "units", n)
batch = engine.Batch(ctx, len(stream), n)
batch.upload(stream, offs)
bm = torch.zeros(n * p.words, dtype=torch.int64, device="cuda")
lib = ctx.lib
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.check(lib.cf_scan(ctx.h, p.h, batch.h, bm.data_ptr(), st), "scan")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
e0.record()
for _ in range(reps):
    ctx.check(lib.cf_scan(ctx.h, p.h, batch.h, bm.data_ptr(), st), "scan")
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"scan: {ms:.3f} ms/step  {len(stream)/ms/1e6:.1f} GB/s  {n/ms*1e3:.0f} payloads/s  counters={ctx.scan_counters()}")
flagged = int((bm != 0).sum())
print("flagged units", flagged)
