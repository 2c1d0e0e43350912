"""Quick device-resident timing of toon_kernel (development aid)."""
import sys

import torch

sys.path.insert(0, ".")
from mcp_context_forge_b200 import engine, synth
from mcp_context_forge_b200.batching import GpuBatcher

b = GpuBatcher.get()
lib = b.ctx.lib
# usage: quick_toon_bench.py [flags] [shape size n]...
FLAGS = int(sys.argv[1]) if len(sys.argv) > 1 else 0
CASES = [("A", 16384, 32768), ("A", 2048, 131072), ("B", 16384, 32768)]
if len(sys.argv) > 2:
    a = sys.argv[2:]
    CASES = [(a[i], int(a[i + 1]), int(a[i + 2])) for i in range(0, len(a), 3)]
for shape, size, n in CASES:
    if shape == "M":   # BASELINE configs[3]: mixed 2 / 16 / 256 KiB tabular payloads (by count 80 % / 19 % / 1 %)
        pool = {k: [synth.payload("A", k, seed=s).encode() for s in range(8)] for k in (2048, 16384, 262144)}
        texts = [pool[262144 if i % 100 == 0 else 16384 if i % 5 == 0 else 2048][i % 8] for i in range(n)]
    elif shape == "P":   # prose inside a JSON document (one long string value), as in bench.py's chain mix
        import json
        base = [json.dumps({"title": f"document {s}", "lang": "en", "body": synth.payload("C", size - 64, seed=s)}, ensure_ascii=False, separators=(",", ":")).encode() for s in range(64)]
        texts = [base[i % 64] for i in range(n)]
    elif shape == "N":   # prose with newline escapes
        import json
        base = [json.dumps({"log": synth.payload("C", size - 64, seed=s).replace(". ", ".\n").replace(" a", "\na", 40)}, ensure_ascii=True).encode() for s in range(64)]
        texts = [base[i % 64] for i in range(n)]
    else:
        base = [synth.payload(shape, size, seed=s).encode() for s in range(64)]
        texts = [base[i % 64] for i in range(n)]
    stream, offs = engine.pack_units(texts)
    batch = engine.Batch(b.ctx, len(stream), n)
    batch.upload(stream, offs)
    d_out = torch.empty(len(stream) + 16, dtype=torch.uint8, device="cuda")
    d_len = torch.empty(n, dtype=torch.int32, device="cuda")
    d_st = torch.empty(n, dtype=torch.int32, device="cuda")
    for _ in range(2):
        lib.cf_toon(b.ctx.h, batch.h, FLAGS, d_out.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.cf_toon(b.ctx.h, batch.h, FLAGS, d_out.data_ptr(), d_len.data_ptr(), d_st.data_ptr(), None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(shape, size, n, "kernel ms", round(ms, 3), "GB/s", round(len(stream) / ms / 1e6, 1), "payloads/s", int(n / ms * 1e3), "converted", int((d_st == 0).sum()))
