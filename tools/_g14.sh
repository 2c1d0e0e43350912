cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, json
sys.path.insert(0, ".")
import numpy as np
from mcp_context_forge_b200 import engine, synth
from mcp_context_forge_b200._native import CF_STAGE_SCAN, CF_STAGE_SUB, CF_STAGE_TOON
import re
ctx = engine.Context.get()
prog = engine.Program()
prog.add_search(r"\bsuicide\b", re.I); prog.add_sub("crap", 0, "crud"); prog.compile(ctx)
units = [synth.payload("A", 3000, seed=s) for s in range(24)] + [synth.payload("B", 2500, seed=s) for s in range(16)] + [json.dumps({"body": synth.payload("C", 3000, seed=s) + " crap é"}) for s in range(8)] + ["not json", "", "[1,2", '{"a":"x\\ny"}']
stream, offs = engine.pack_units([engine.encode_unit(u) for u in units])
b = engine.Batch(ctx, len(stream), len(units))
for _ in range(2):
    v, out, oo, _ = engine.run_batch(prog, b, stream, offs, CF_STAGE_SCAN | CF_STAGE_SUB | CF_STAGE_TOON)
print("ok", int((v["flags"] != 0).sum()))
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/r02_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py > gpurun_out/r02_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v "^ERROR\|^WARNING" | tail -3
