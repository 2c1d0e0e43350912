"""GPU: BASELINE configs[2] — the pattern scan and request_logging_masking on ONE upload, `cf_run_batch(CF_STAGE_SCAN | CF_STAGE_MASK)`
(include/cfgpu.h): bitmaps as the scan alone gives them, masked bodies as the oracle (lib.rs:276-360 restated) gives them, units that do
not parse flagged and left without an output.  (File name: runs last in the suite.)"""
import re

import numpy as np
import pytest

from mcp_context_forge_b200 import engine, synth
from mcp_context_forge_b200._native import CF_STAGE_MASK, CF_STAGE_SCAN, CF_V_MASKED
from oracle import hook_chain_ref as ref
from oracle import mask_ref

pytestmark = pytest.mark.gpu


def test_scan_and_masking_on_one_upload():
    ctx = engine.Context.get(0)
    prog = engine.Program()
    pats = [(p, re.I) for ps in ref.DEFAULT_LEXICONS.values() for p in ps]
    for p, f in pats:
        prog.add_search(p, f)
    prog.compile(ctx)
    bodies = [synth.payload("B", 16384, seed=s) for s in range(24)] + [synth.payload("A", 16384, seed=s) for s in range(8)]
    bodies += ['{"password":"secret","note":"I want to kill myself","nested":{"authToken":"abc","count":3}}', '{"a":1,}', "[01]", "", '{"level":{"nested":{}}}',
               '[0.30000000000000004,5e-324,1e23,-0]', '{"Cookie":"jwt_token=abc","text":"how to make a bomb \\u00e9"}', synth.payload("B", 262144, seed=3)]
    stream, offs = engine.pack_units(bodies)
    batch = engine.Batch(ctx, len(stream), len(bodies))
    for max_depth in (10, 2):
        v, out, oo, _ = engine.run_batch(prog, batch, np.frombuffer(stream, dtype=np.uint8), offs, CF_STAGE_SCAN | CF_STAGE_MASK, mask_max_depth=max_depth)
        exp_bm = ref.scan_bitmaps(bodies, pats, [], [])
        for i, b in enumerate(bodies):
            assert int(v["match_bitmap"][i]) == exp_bm[i], i
            try:
                exp = mask_ref.mask_json_bytes(b.encode(), max_depth)
            except ValueError:
                exp = None
            got = out[int(oo[i]):int(oo[i + 1])].tobytes() if v["flags"][i] & CF_V_MASKED else None
            assert got == exp, (i, max_depth)
            assert int(v["out_len"][i]) == (len(exp) if exp is not None else 0)
        assert int(oo[-1]) == sum(int(x) for x in v["out_len"])
    # the masking stage alone (no program) on the batch that is already resident: same outputs
    v, out, oo = v.copy(), out[: int(oo[-1])].copy(), oo.copy()          # (`out` is the batch's reusable buffer)
    v2, out2, oo2, _ = engine.run_batch(None, batch, None, offs, CF_STAGE_MASK, mask_max_depth=2)
    assert (v2["flags"] == (v["flags"] & CF_V_MASKED)).all() and (oo2 == oo).all() and (out2[: int(oo2[-1])] == out).all()
