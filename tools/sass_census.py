"""Per-kernel census of the SASS instructions that show how data moves (cuobjdump -sass of the built library).
usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "mcp_context_forge_b200", "libcfgpu.so")
KEEP = re.compile(r"^(UTMALDG|UBLKCP|UTMA|SYNCS|LDS|STS|LDG|STG|LDSM|SHFL|VOTE|MATCH|WARPSYNC|POPC|FLO|REDUX|ATOM|RED|BAR|UTC|HMMA|IMMA|LDGSTS)")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
cur, acc = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        acc[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur and KEEP.match(m.group(1)):
        acc[cur][m.group(1)] += 1
print("# SASS census of mcp_context_forge_b200/libcfgpu.so (cuobjdump -sass), round 2: per kernel, the instructions that show how data moves.")
print("# UTMALDG = cp.async.bulk.tensor (TMA tile loads), UBLKCP = cp.async.bulk (1-D bulk copy), SYNCS = mbarrier, LDS/STS = shared memory,")
print("# VOTE/SHFL/MATCH = warp collectives, WARPSYNC = __syncwarp.  No tcgen05 / UTC*MMA by design: nothing on this path is a contraction.")
for k in sorted(acc):
    if not acc[k]:
        continue
    print("==", k)
    for op, n in acc[k].most_common(26):
        print(f"  {n:6d} {op}")
