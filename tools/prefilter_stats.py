"""Candidate count of the compiled prefilter on the bench mix, computed on the CPU with the
test-only host build of the same tables (development aid)."""
import re
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import bench
import hostsim_util as h
from mcp_context_forge_b200.plugins.harmful_content_detector import DEFAULT_LEXICONS

p = h.HostProgram()
for pats in DEFAULT_LEXICONS.values():
    for pat in pats:
        p.add(pat, re.I)
for w in bench.DENY:
    p.add(re.escape(w))
for s, f, r in bench.SUBS:
    p.add(s, f)
units = bench.make_payloads(256) if hasattr(bench, "make_payloads") else None
_, st = p.scan(units)
nbytes = sum(len(u.encode()) for u in units)
print(f"bytes {nbytes}  candidates {st[0]}  ({st[0] / nbytes * 2**30:.0f} per GiB)  dfa steps {st[1]}")
