"""The chain-level batched executor (mcp_context_forge_b200.manager.BatchedPluginManager) against the sequential
executor it replaces: same YAML, same payloads, every result compared field by field — violations, short-circuits, rewritten
payloads, TOON conversions, `violations_as_exceptions` — and the whole wave must be ONE fused launch."""
import asyncio
import json
import os
import random
import tempfile

import pytest

from mcp_context_forge_b200 import framework as fw, synth
from mcp_context_forge_b200.cpex_compat.framework import HookPayloadPolicy
from mcp_context_forge_b200.manager import BatchedPluginManager

pytestmark = pytest.mark.gpu

YAML = """
plugins:
  - name: "HarmfulContentDetector"
    kind: "mcp_context_forge_b200.plugins.harmful_content_detector.HarmfulContentDetectorPlugin"
    hooks: ["prompt_pre_fetch", "tool_post_invoke"]
    mode: "%(harm_mode)s"
    priority: 96
  - name: "DenyListPlugin"
    kind: "mcp_context_forge_b200.plugins.deny_filter.DenyListPlugin"
    hooks: ["prompt_pre_fetch"]
    mode: "sequential"
    priority: 100
    config:
      words: [innovative, groundbreaking, revolutionary]
  - name: "ReplaceBadWordsPlugin"
    kind: "mcp_context_forge_b200.plugins.regex_filter.SearchReplacePlugin"
    hooks: ["prompt_pre_fetch", "tool_pre_invoke", "tool_post_invoke"]
    mode: "sequential"
    priority: %(regex_prio)d
    config:
      words:
        - {search: crap, replace: crud}
        - {search: crud, replace: yikes}
        - {search: '(?i)(kill) (him|her)', replace: '\\2 <\\1>'}
  - name: "SQLSanitizer"
    kind: "mcp_context_forge_b200.plugins.sql_sanitizer.SQLSanitizerPlugin"
    hooks: ["prompt_pre_fetch", "tool_pre_invoke"]
    mode: "sequential"
    priority: 45
    config:
      block_on_violation: false
  - name: "ToonEncoder"
    kind: "mcp_context_forge_b200.plugins.toon_encoder.ToonEncoderPlugin"
    hooks: ["tool_post_invoke"]
    mode: "sequential"
    priority: 900
    conditions:
      - tools: ["t", "u"]
plugin_settings:
  plugin_timeout: 120
"""
POL = {"tool_pre_invoke": HookPayloadPolicy(writable_fields=frozenset({"name", "args", "headers"})),
       "tool_post_invoke": HookPayloadPolicy(writable_fields=frozenset({"result"})),
       "prompt_pre_fetch": HookPayloadPolicy(writable_fields=frozenset({"args"}))}


def managers(td, **kw):
    cfg = os.path.join(td, "plugins.yaml")
    with open(cfg, "w") as f:
        f.write(YAML % kw)
    return fw.PluginManager(cfg, timeout=120, hook_policies=POL), BatchedPluginManager(cfg, timeout=120, hook_policies=POL)


def norm(res):
    if isinstance(res, BaseException):
        v = getattr(res, "violation", None)
        return ("EXC", type(res).__name__, v.model_dump() if v is not None else str(res)[:80])
    r = res[0]
    md = dict(r.metadata or {})
    md.pop("conversion_time_ms", None)
    return (r.continue_processing, r.modified_payload.model_dump() if r.modified_payload is not None else None, r.violation.model_dump() if r.violation else None, md)


def payloads(n, seed):
    rng = random.Random(seed)
    words = ["hello", "crap", "crud", "innovative", "kill him", "suicide", "normal text", "Kill her", "revolutionary idea", "I want to die", "racial slur", "fine",
             "DROP table t -- crap", "select 1 /* c */", "delete from t"]
    pre, tpre, post = [], [], []
    for i in range(n):
        args = {f"k{j}": " ".join(rng.choice(words) for _ in range(rng.randint(1, 6))) for j in range(rng.randint(0, 3))}
        if rng.random() < 0.2:
            args["n"] = rng.randint(0, 9)
        pre.append(fw.PromptPrehookPayload(prompt_id="p", args=args))
        tpre.append(fw.ToolPreInvokePayload(name="t", args=dict(args)))
        text = synth.payload("A", rng.choice([300, 1500, 6000]), seed=i) if rng.random() < 0.7 else json.dumps({"note": rng.choice(words), "x": [1, 2, {"y": rng.choice(words)}]})
        if rng.random() < 0.15:
            text = "not json " + rng.choice(words)
        result = {"content": [{"type": "text", "text": text}, {"type": "text", "text": rng.choice(words) * 3}], "summary": rng.choice(words) + " crap"}
        if rng.random() < 0.1:
            result = rng.choice(words)
        post.append(fw.ToolPostInvokePayload(name=rng.choice(["t", "u", "other"]), result=result))
    return pre, tpre, post


@pytest.mark.parametrize("harm_mode,regex_prio", [("sequential", 150), ("transform", 50)])
def test_batched_chain_equals_sequential_chain(harm_mode, regex_prio):
    with tempfile.TemporaryDirectory() as td:
        seq, bat = managers(td, harm_mode=harm_mode, regex_prio=regex_prio)
        loop = asyncio.new_event_loop()
        loop.run_until_complete(seq.initialize())
        loop.run_until_complete(bat.initialize())
        pre, tpre, post = payloads(160, 7)
        gcs = [fw.GlobalContext(request_id=f"r{i}") for i in range(len(pre))]
        for hook, pls in (("prompt_pre_fetch", pre), ("tool_pre_invoke", tpre), ("tool_post_invoke", post)):
            for vae in (False, True):
                async def wave(m):
                    return await asyncio.gather(*[m.invoke_hook(hook, p, g, None, vae) for p, g in zip(pls, gcs)], return_exceptions=True)
                a = loop.run_until_complete(wave(seq))
                before = bat.launch_calls
                b = loop.run_until_complete(wave(bat))
                assert bat.launch_calls - before == 1, "one fused launch per wave"
                bad = [(i, norm(x), norm(y)) for i, (x, y) in enumerate(zip(a, b)) if norm(x) != norm(y)]
                assert not bad, (hook, vae, bad[:2])
        # something of everything happened
        if regex_prio == 50:
            assert bat.slow_path_calls > 0                  # regex_filter ran first and rewrote values: the plugins behind it ran their own hook
        # (regex_prio == 150: only SQLSanitizer — a plugin without the chain protocol — and the plugins behind it on the requests whose
        #  args it changed run their own hook; everything else is served from the fused launch)
        assert bat.slow_path_calls < 6 * 160 * 3
        loop.run_until_complete(seq.shutdown())
        loop.run_until_complete(bat.shutdown())


def test_run_batch_stage_masks_and_resubmit():
    """cf_run_batch directly: per-unit stage masks, rewritten units carry CF_V_RESUBMIT when TOON was also requested."""
    import numpy as np
    from mcp_context_forge_b200 import engine
    from mcp_context_forge_b200._native import CF_STAGE_SCAN, CF_STAGE_SUB, CF_STAGE_TOON, CF_V_RESUBMIT, CF_V_REWRITTEN, CF_V_TOON
    ctx = engine.Context.get()
    prog = engine.Program()
    prog.add_sub("crap", 0, "crud")
    prog.add_sub("crud", 0, "yikes")
    b_h = prog.add_search(r"\bsuicide\b", 2)
    prog.compile(ctx)
    doc = synth.payload("A", 2000, seed=1)
    units = ["plain crap here", doc, "suicide note", json.dumps({"t": "crap " * 30, "rows": [{"a": 1}, {"a": 2}]}), "untouched crap"]
    stages = np.array([CF_STAGE_SCAN | CF_STAGE_SUB, CF_STAGE_SCAN | CF_STAGE_TOON, CF_STAGE_SCAN, CF_STAGE_SCAN | CF_STAGE_SUB | CF_STAGE_TOON, CF_STAGE_SCAN], dtype=np.uint8)
    enc = [engine.encode_unit(u) for u in units]
    stream, offs = engine.pack_units(enc)
    batch = engine.Batch(ctx, len(stream), len(units))
    v, out, oo, _ = engine.run_batch(prog, batch, stream, offs, CF_STAGE_SCAN | CF_STAGE_SUB | CF_STAGE_TOON, stages)
    text = lambda i: out[int(oo[i]):int(oo[i + 1])].tobytes().decode()
    assert v["flags"][0] == CF_V_REWRITTEN and text(0) == "plain yikes here"
    assert v["flags"][1] == CF_V_TOON and text(1).startswith("results[")
    assert v["flags"][2] == 0 and (int(v["match_bitmap"][2]) >> b_h) & 1
    assert v["flags"][3] == CF_V_REWRITTEN | CF_V_RESUBMIT and "yikes" in text(3)
    assert v["flags"][4] == 0 and int(v["match_bitmap"][4]) & 3          # matched, but SUB was not requested for this unit
