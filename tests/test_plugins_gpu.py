"""GPU drop-in plugins vs golden vectors recorded from the reference's own plugin files
(tests/golden/pattern_plugins.json, tools/gen_golden.py) and vs the oracle.  Exercises the whole
product path: Plugin hook -> batch coalescer -> packed stream -> C ABI -> CUDA kernels."""
import asyncio
import json
import os
import random
import re

import pytest

from mcp_context_forge_b200 import framework as fw
from mcp_context_forge_b200.plugins.deny_filter import DenyListPlugin
from mcp_context_forge_b200.plugins.harmful_content_detector import HarmfulContentDetectorPlugin
from mcp_context_forge_b200.plugins.regex_filter import SearchReplacePlugin
from oracle import hook_chain_ref as ref

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CTX = fw.PluginContext(global_context=fw.GlobalContext(request_id="t"))


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLD, "pattern_plugins.json"), encoding="utf-8") as f:
        return json.load(f)


def _templates_gold():
    with open(os.path.join(GOLD, "regex_filter_templates.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.mark.parametrize("which", ["plain", "templates"])
def test_regex_filter_matches_reference_golden(gold, which):
    """`templates`: rules that can match "" and replacement templates with group references (tests/golden/regex_filter_templates.json,
    recorded from the reference's plugin file)."""
    for block in (gold["regex_filter"] if which == "plain" else _templates_gold()):
        plug = SearchReplacePlugin(fw.PluginConfig(name="rf", kind="x", hooks=["tool_pre_invoke", "tool_post_invoke"], config={"words": block["words"]}))

        async def all_cases():
            pre = [plug.tool_pre_invoke(fw.ToolPreInvokePayload(name="t", args=c["args"]), CTX) for c in block["cases"] if c["hook"] == "tool_pre_invoke"]
            post = [plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), CTX) for c in block["cases"] if c["hook"] == "tool_post_invoke"]
            return await asyncio.gather(*pre), await asyncio.gather(*post)   # concurrent -> coalesced launches

        pre, post = run(all_cases())
        for c, r in zip([c for c in block["cases"] if c["hook"] == "tool_pre_invoke"], pre):
            assert r.modified_payload.args == c["out_args"], c["args"]
        for c, r in zip([c for c in block["cases"] if c["hook"] == "tool_post_invoke"], post):
            assert r.modified_payload.result == c["out_result"], c["result"]


def test_deny_filter_matches_reference_golden(gold):
    for block in gold["deny_filter"]:
        plug = DenyListPlugin(fw.PluginConfig(name="dl", kind="x", hooks=["prompt_pre_fetch"], config={"words": block["words"]}))
        for c in block["cases"]:
            r = run(plug.prompt_pre_fetch(fw.PromptPrehookPayload(prompt_id="p", args=c["args"]), CTX))
            assert (not r.continue_processing) == c["blocked"], (block["words"], c["args"])
            if c["blocked"]:
                got = r.violation.model_dump(exclude={"plugin_name", "http_status_code", "mcp_error_code", "http_headers"})
                assert got == c["violation"]
            assert r.modified_payload.args == c["args"]


def test_harmful_matches_reference_golden(gold):
    for block in gold["harmful"]:
        plug = HarmfulContentDetectorPlugin(fw.PluginConfig(name="hc", kind="x", hooks=["tool_post_invoke"], config=block["config"]))

        async def all_cases():
            return await asyncio.gather(*[plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), CTX) for c in block["cases"]])

        for c, r in zip(block["cases"], run(all_cases())):
            assert r.continue_processing == c["continue_processing"], c["result"]
            assert (r.metadata or {}) == (c["metadata"] or {})
            if c["violation"]:
                got = json.loads(json.dumps(r.violation.model_dump(exclude={"plugin_name", "http_status_code", "mcp_error_code", "http_headers"})))
                assert got == c["violation"]
            else:
                assert r.violation is None
            assert r.modified_payload is None


SUB_RULE_SETS = [
    [("crap", "crud"), ("crud", "yikes")], [(r"\bkill\b", "[k]"), (r"\d+", "#")], [(r"cr[au]p+", "X"), (r"a|ab|abc", "<>")],
    [(r"x+?", "y"), (r"\w+-\w+", "é")], [(r"k.l+", "日本"), (r"to (?:die|live)\b", "—")], [("a", "aaaa"), ("aa", "b")],
]


def test_sub_engine_vs_oracle_fuzz_and_sizes():
    from test_regex_engine_cpu import rand_text
    from mcp_context_forge_b200 import synth
    from mcp_context_forge_b200.batching import GpuBatcher

    rng = random.Random(9)
    for rules in SUB_RULE_SETS:
        plug = SearchReplacePlugin(fw.PluginConfig(name="rf", kind="x", config={"words": [{"search": s, "replace": r} for s, r in rules]}))
        comp = ref.regex_compile_rules([{"search": s, "replace": r} for s, r in rules])
        units = [rand_text(rng, rng.randint(0, 40)) for _ in range(1500)]
        units += [synth.payload("C", 16384, seed=3, hit_rate=5e-3), synth.payload("C", 300000, seed=4, hit_rate=2e-3), synth.payload("A", 2048, seed=5),
                  "crap" * 5000, "a" * 3000, "", "x" * 70000 + "crap" + "y" * 70000]
        got = run(plug._apply(units))
        exp = [ref.regex_apply_str(comp, u) for u in units]
        bad = [(u[:80], g[:80], e[:80]) for u, g, e in zip(units, got, exp) if g != e]
        assert not bad, (rules, bad[:2])
    assert GpuBatcher.get().launches > 0


def test_sub_engine_empty_matches_and_group_templates_fuzz():
    """The substitution kernel's nullable branch (every position a candidate, must_advance after an empty match) and the capture
    pass (Pike VM on lane 0) against CPython `re.sub`, on short fuzz texts and on long units (window boundaries, many matches)."""
    from test_regex_engine_cpu import GROUP_RULES, NULLABLE_RULES, rand_text
    from mcp_context_forge_b200 import synth

    rng = random.Random(31)
    rule_sets = [[(p, t)] for p, f, t in GROUP_RULES if f == 0] + [[(p, r.replace("\\", "\\\\"))] for p, f, r in NULLABLE_RULES if f == 0]
    rule_sets += [[(r"(\w+)@(\w+)", r"\2 at \1"), (r"x*", "-"), (r"(-)(-)?", r"\2\1")], [(r"\b", "|"), (r"(\|)(\w)", r"\2\1")]]
    for rules in rule_sets:
        words = [{"search": s, "replace": r} for s, r in rules]
        plug = SearchReplacePlugin(fw.PluginConfig(name="rf", kind="x", config={"words": words}))
        comp = ref.regex_compile_rules(words)
        units = [rand_text(rng, rng.randint(0, 30)) for _ in range(300)]
        units += ["", "x", "abxd", "a" * 700 + "b", "xy" * 400, synth.payload("C", 6000, seed=7), "user@host " * 200, "é日" * 300]
        got = run(plug._apply(units))
        exp = [ref.regex_apply_str(comp, u) for u in units]
        bad = [(u[:60], g[:80], e[:80]) for u, g, e in zip(units, got, exp) if g != e]
        assert not bad, (rules, bad[:2])


def test_chain_through_plugin_manager_yaml(tmp_path):
    cfg = tmp_path / "plugins.yaml"
    cfg.write_text("""
plugins:
  - name: "HarmfulContentDetector"
    kind: "mcp_context_forge_b200.plugins.harmful_content_detector.HarmfulContentDetectorPlugin"
    hooks: ["prompt_pre_fetch", "tool_post_invoke"]
    mode: "sequential"
    priority: 96
  - name: "ReplaceBadWordsPlugin"
    kind: "mcp_context_forge_b200.plugins.regex_filter.SearchReplacePlugin"
    hooks: ["prompt_pre_fetch", "tool_pre_invoke", "tool_post_invoke"]
    mode: "sequential"
    priority: 150
    config:
      words:
        - search: crap
          replace: crud
        - search: crud
          replace: yikes
  - name: "DenyListPlugin"
    kind: "mcp_context_forge_b200.plugins.deny_filter.DenyListPlugin"
    hooks: ["prompt_pre_fetch"]
    mode: "sequential"
    priority: 100
    config:
      words: [innovative, groundbreaking, revolutionary]
plugin_settings:
  plugin_timeout: 30
""")
    from mcp_context_forge_b200.cpex_compat.framework import HookPayloadPolicy

    pol = {"tool_pre_invoke": HookPayloadPolicy(writable_fields=frozenset({"name", "args", "headers"})), "tool_post_invoke": HookPayloadPolicy(writable_fields=frozenset({"result"})),
           "prompt_pre_fetch": HookPayloadPolicy(writable_fields=frozenset({"args"}))}
    m = fw.PluginManager(str(cfg), timeout=30, hook_policies=pol)
    run(m.initialize())
    gc = fw.GlobalContext(request_id="r1")
    # integration semantic pinned by the reference: "crap" -> "crud" -> "yikes"
    # (/root/reference/tests/integration/test_plugin_dynamic_behavior_bad_words.py:52-58)
    res, _ = run(m.invoke_hook("tool_pre_invoke", fw.ToolPreInvokePayload(name="echo", args={"text": "this is crap", "n": 3}), gc))
    assert res.continue_processing and res.modified_payload.args == {"text": "this is yikes", "n": 3}
    res, _ = run(m.invoke_hook("tool_post_invoke", fw.ToolPostInvokePayload(name="echo", result={"content": [{"type": "text", "text": "I want to die"}]}), gc))
    assert not res.continue_processing and res.violation.code == "HARMFUL_CONTENT" and res.violation.details["categories"] == ["self_harm"]
    res, _ = run(m.invoke_hook("prompt_pre_fetch", fw.PromptPrehookPayload(prompt_id="p", args={"q": "a revolutionary idea"}), gc))
    assert not res.continue_processing and res.violation.code == "deny" and res.violation.plugin_name == "DenyListPlugin"
    res, _ = run(m.invoke_hook("prompt_pre_fetch", fw.PromptPrehookPayload(prompt_id="p", args={"q": "plain crap"}), gc))
    assert res.continue_processing and res.modified_payload.args == {"q": "plain yikes"}
    with pytest.raises(fw.PluginViolationError):
        run(m.invoke_hook("prompt_pre_fetch", fw.PromptPrehookPayload(prompt_id="p", args={"q": "kill myself"}), gc, violations_as_exceptions=True))
