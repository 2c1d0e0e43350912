"""Device-resident timing of json_index_kernel (development aid).  usage: quick_index_bench.py [shape size n]..."""
import sys

import torch

sys.path.insert(0, ".")
from mcp_context_forge_b200 import engine, synth

ctx = engine.Context.get()
a = sys.argv[1:] or ["A", "16384", "32768"]
for i in range(0, len(a), 3):
    shape, size, n = a[i], int(a[i + 1]), int(a[i + 2])
    base = [synth.payload(shape, size, seed=s).encode() for s in range(64)]
    stream, offs = engine.pack_units([base[k % 64] for k in range(n)])
    batch = engine.Batch(ctx, len(stream), n)
    batch.upload(stream, offs)
    toks = torch.empty((len(stream) + 64, 2), dtype=torch.int32, device="cuda")
    cnt = torch.empty(n, dtype=torch.int32, device="cuda")
    for flags in (0, 1):
        for _ in range(2):
            ctx.check(ctx.lib.cf_json_index(ctx.h, batch.h, flags, toks.data_ptr(), cnt.data_ptr(), None), "index")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.check(ctx.lib.cf_json_index(ctx.h, batch.h, flags, toks.data_ptr(), cnt.data_ptr(), None), "index")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ntok = int((cnt & 0x7FFFFFFF).sum())
        print(shape, size, n, "classify" if flags else "stage1  ", "ms", round(ms, 3), "GB/s", round(len(stream) / ms / 1e6, 1),
              "tokens", ntok, "token bytes/input byte", round(ntok * 8 / len(stream), 2))
