"""code_safety_linter: oracle vs golden vectors recorded from the reference's plugin file (tools/gen_golden.py
code_safety); the drop-in's patterns through the TEST-ONLY host build of the scan engine (CPU); the drop-in end to end on the GPU,
stand-alone and inside the chain-level manager."""
import asyncio
import json
import os
import re

import pytest

from oracle import code_safety_ref as ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "code_safety.json")


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)


def test_oracle_matches_reference_golden(gold):
    n = 0
    for block in gold:
        pats = None if block["config"] is None else block["config"]["blocked_patterns"]
        for c in block["cases"]:
            got = ref.hook(c["result"], pats)
            assert got["continue_processing"] == c["continue_processing"] and got["violation"] == c["violation"], c
            n += 1
    assert n >= 300


def test_patterns_on_the_host_build_of_the_engine(gold):
    """The drop-in's automaton (same front-end, same tables as the GPU program) against the golden verdicts, on the CPU simulator."""
    from hostsim_util import HostProgram

    for block in gold:
        pats = ref.DEFAULTS if block["config"] is None else block["config"]["blocked_patterns"]
        if not pats:
            continue
        hp = HostProgram()
        for p in pats:
            c = re.compile(p)
            hp.add(c.pattern, c.flags)
        texts, exps = [], []
        for c in block["cases"]:
            r = c["result"]
            t = r if isinstance(r, str) else r.get("text") if isinstance(r, dict) and isinstance(r.get("text"), str) else None
            if t:
                texts.append(t)
                exps.append(c["violation"]["details"]["patterns"] if c["violation"] else [])
        got, _ = hp.scan(texts)
        for t, g, e in zip(texts, got, exps):
            assert [p for i, p in enumerate(pats) if g >> i & 1] == e, t


@pytest.mark.gpu
def test_dropin_matches_reference_golden_gpu(gold):
    from mcp_context_forge_b200 import framework as fw
    from mcp_context_forge_b200.plugins.code_safety_linter import CodeSafetyLinterPlugin

    ctx = fw.PluginContext(global_context=fw.GlobalContext(request_id="t"))
    for block in gold:
        plug = CodeSafetyLinterPlugin(fw.PluginConfig(name="cs", kind="x", hooks=["tool_post_invoke"], config=block["config"]))

        async def go():
            return await asyncio.gather(*[plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), ctx) for c in block["cases"]])

        for c, r in zip(block["cases"], run(go())):
            assert r.continue_processing == c["continue_processing"], c
            assert (r.violation.model_dump(include={"reason", "description", "code", "details"}) if r.violation else None) == c["violation"], c


@pytest.mark.gpu
def test_dropin_inside_the_batched_manager_gpu(gold, tmp_path):
    """code_safety_linter speaks the chain protocol: its patterns join the chain's shared program and one fused launch serves the wave."""
    from mcp_context_forge_b200 import framework as fw
    from mcp_context_forge_b200.manager import BatchedPluginManager

    cfg = tmp_path / "plugins.yaml"
    cfg.write_text("""
plugins:
  - name: "CodeSafetyLinter"
    kind: "mcp_context_forge_b200.plugins.code_safety_linter.CodeSafetyLinterPlugin"
    hooks: ["tool_post_invoke"]
    mode: "sequential"
    priority: 10
  - name: "HarmfulContentDetector"
    kind: "mcp_context_forge_b200.plugins.harmful_content_detector.HarmfulContentDetectorPlugin"
    hooks: ["tool_post_invoke"]
    mode: "sequential"
    priority: 96
plugin_settings:
  plugin_timeout: 120
""")
    seq, bat = fw.PluginManager(str(cfg), timeout=120), BatchedPluginManager(str(cfg), timeout=120)
    loop = asyncio.new_event_loop()
    loop.run_until_complete(seq.initialize())
    loop.run_until_complete(bat.initialize())
    cases = gold[0]["cases"]
    gc = fw.GlobalContext(request_id="r")

    async def wave(m):
        return await asyncio.gather(*[m.invoke_hook("tool_post_invoke", fw.ToolPostInvokePayload(name="t", result=c["result"]), gc) for c in cases])

    a, b = loop.run_until_complete(wave(seq)), loop.run_until_complete(wave(bat))
    assert bat.launch_calls == 1
    for c, (ra, _), (rb, _) in zip(cases, a, b):
        assert ra.continue_processing == rb.continue_processing == c["continue_processing"], c
        va = ra.violation.model_dump() if ra.violation else None
        vb = rb.violation.model_dump() if rb.violation else None
        assert va == vb
