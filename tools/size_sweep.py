"""BASELINE configs[4]: payload-size sweep of the fused scan (device-resident), 512 B - 1 MiB.
Batch = min(65536 units, 1 GiB).  Prints one JSON object; development/measurement aid, bench.py is the contract."""
import json
import re
import sys

import torch

sys.path.insert(0, ".")
import bench
from mcp_context_forge_b200 import engine, synth
from mcp_context_forge_b200.plugins.harmful_content_detector import DEFAULT_LEXICONS

ctx = engine.Context.get()
p = engine.Program()
for pats in DEFAULT_LEXICONS.values():
    for pat in pats:
        p.add_search(pat, re.I)
for w in bench.DENY:
    p.add_literal(w)
for s, f, r in bench.SUBS:
    p.add_sub(s, f, r)
p.compile(ctx)
peak = bench.read_peaks()[0]
rows = []
for size in [512, 2048, 16384, 65536, 262144, 1048576]:
    n = min(65536, (1 << 30) // size)
    shapes = ["A", "A", "B", "C"]
    base = [synth.payload(shapes[s % 4], size if shapes[s % 4] != "B" else max(256, int(size * 0.6)), seed=s, hit_rate=1e-4)
            for s in range(32 if size <= 65536 else 8)]
    units = [base[i % len(base)] for i in range(n)]
    stream, offs = engine.pack_units(units)
    batch = engine.Batch(ctx, len(stream), n)
    batch.upload(stream, offs)
    bm = torch.zeros(n * p.words, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        ctx.check(ctx.lib.cf_scan(ctx.h, p.h, batch.h, bm.data_ptr(), st), "scan")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ctx.check(ctx.lib.cf_scan(ctx.h, p.h, batch.h, bm.data_ptr(), st), "scan")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # the row is labelled by what the payloads actually measure: shape B (nested config) has a floor of a few KB, so the nominal
    # 512 B / 2 KiB rows are larger than their nominal size (round-1 ADVICE)
    row = {"nominal_payload_bytes": size, "mean_payload_bytes": round((len(stream) - n) / n, 1), "min_payload_bytes": min(len(u.encode()) for u in base),
           "max_payload_bytes": max(len(u.encode()) for u in base), "units": n, "stream_bytes": len(stream), "ms": round(ms, 4),
           "gb_per_s": round(len(stream) / ms / 1e6, 1), "payloads_per_s": round(n / ms * 1e3),
           "candidates": ctx.scan_counters()[0]}
    if peak:
        row["frac_of_measured_hbm_peak"] = round(row["gb_per_s"] / peak, 3)
    rows.append(row)
    print(row, file=sys.stderr)
    del batch, bm
print(json.dumps({"sweep": "fused scan, device-resident, cf_scan (memset + scan_kernel)", "rows": rows}))
