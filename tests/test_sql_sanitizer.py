"""sql_sanitizer: the oracle against golden vectors recorded from the reference's own plugin file
(tools/gen_golden.py sql_sanitizer), the drop-in's host logic on the CPU (verdict bits supplied by a stand-in batcher built on
stdlib `re`: TEST ONLY), and the drop-in end to end on the GPU."""
import asyncio
import json
import os
import re

import pytest

from oracle import sql_sanitizer_ref as ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sql_sanitizer.json")


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)


def strip_violation(v):
    return None if v is None else {k: v[k] for k in ("reason", "description", "code", "details")}


def test_oracle_matches_reference_golden(gold):
    n = 0
    for block in gold:
        c = ref.config(block["config"])
        for case in block["cases"]:
            got = ref.hook(case["args"], c, "tool args" if case["hook"] == "tool_pre_invoke" else "prompt args")
            assert got["continue_processing"] == case["continue_processing"]
            assert got["violation"] == strip_violation(case["violation"])
            assert got["out_args"] == case["out_args"]
            assert got["metadata"] == case["metadata"]
            n += 1
    assert n >= 700


class ReBatcher:
    """TEST-ONLY stand-in for GpuBatcher: the same interface, verdict bits from stdlib `re`."""

    def __init__(self, plugin):
        cfg = plugin._cfg
        self.search = [p for p in cfg.blocked_statements] + [ref.DELETE_FROM, ref.UPDATE, ref.WHERE]
        self.lits = ["+", "%.", "{", "}"]
        self.rules = [ref.LINE_COMMENT, ref.BLOCK_COMMENT] if cfg.strip_comments else []

    def bits(self, t):
        v = 0
        for i, p in enumerate(self.search):
            v |= bool(p.search(t)) << i
        for j, w in enumerate(self.lits):
            v |= (w in t) << (len(self.search) + j)
        for k, p in enumerate(self.rules):
            v |= bool(p.search(t)) << (len(self.search) + len(self.lits) + k)
        return v

    async def scan(self, prog, units):
        return [self.bits(u) for u in units]

    async def scan_sub(self, prog, units, rule_mask):
        out = []
        for u in units:
            b = self.bits(u)
            new = None
            if b & rule_mask:
                t = u
                for p in self.rules:
                    t = p.sub("", t)
                new = t.encode("utf-8", "surrogatepass")
            out.append((b, new))
        return out


class HostSimBatcher:
    """TEST-ONLY stand-in for GpuBatcher on the CPU simulator of the engine (same front-end, same automata and substitution
    routine as the GPU program, tests/hostsim): verdict bits and stripped texts come from the product's tables, not from `re`."""

    def __init__(self, plugin):
        from hostsim_util import HostProgram

        cfg = plugin._cfg
        self.hp = hp = HostProgram()
        for p in cfg.blocked_statements:
            hp.add(p.pattern, p.flags)
        for pat, fl in ((r"\bDELETE\b\s+\bFROM\b", re.I), (r"\bUPDATE\b\s+\w+", re.I), (r"\bWHERE\b", re.I)):
            hp.add(pat, fl)
        for w in ("+", "%.", "{", "}"):
            hp.add_ast(hp.fe.literal_ast(w))
        self.n_rules = 0
        if cfg.strip_comments:
            hp.add(r"--.*?$", re.M, ordered=True, repl="")
            hp.add(r"/\*.*?\*/", re.S, ordered=True, repl="")
            self.n_rules = 2
        hp.compile()

    async def scan(self, prog, units):
        return self.hp.scan(list(units))[0]

    async def scan_sub(self, prog, units, rule_mask):
        out = []
        for u, b in zip(units, self.hp.scan(list(units))[0]):
            new = None
            if b & rule_mask:
                t = u
                for r in range(self.n_rules):
                    t = self.hp.sub(r, t)[0]
                new = t.encode("utf-8", "surrogatepass")
            out.append((b, new))
        return out


def check_plugin(gold, make_batcher):
    from mcp_context_forge_b200 import framework as fw
    from mcp_context_forge_b200.plugins.sql_sanitizer import SQLSanitizerPlugin

    ctx = fw.PluginContext(global_context=fw.GlobalContext(request_id="t"))
    n = 0
    for block in gold:
        plug = SQLSanitizerPlugin(fw.PluginConfig(name="sql", kind="x", hooks=["prompt_pre_fetch", "tool_pre_invoke"], config=block["config"]))
        if make_batcher:
            plug._batcher = make_batcher(plug)
        for case in block["cases"]:
            if case["hook"] == "tool_pre_invoke":
                r = run(plug.tool_pre_invoke(fw.ToolPreInvokePayload(name="t", args=case["args"]), ctx))
            else:
                r = run(plug.prompt_pre_fetch(fw.PromptPrehookPayload(prompt_id="p", args=case["args"]), ctx))
            assert r.continue_processing == case["continue_processing"], case
            assert (r.violation.model_dump(include={"reason", "description", "code", "details"}) if r.violation else None) == strip_violation(case["violation"]), case
            assert (r.modified_payload.args if r.modified_payload is not None else None) == case["out_args"], case
            assert r.metadata == case["metadata"], case
            n += 1
    assert n >= 700


def test_dropin_host_logic_cpu(gold):
    check_plugin(gold, ReBatcher)


def test_dropin_on_the_host_build_of_the_engine(gold):
    check_plugin(gold, HostSimBatcher)


def test_dropin_rejects_what_the_engine_cannot_express():
    from mcp_context_forge_b200 import framework as fw
    from mcp_context_forge_b200.plugins.sql_sanitizer import SQLSanitizerPlugin
    from mcp_context_forge_b200.regex_frontend import UnsupportedPattern

    with pytest.raises(UnsupportedPattern):
        SQLSanitizerPlugin(fw.PluginConfig(name="sql", kind="x", hooks=["tool_pre_invoke"], config={"blocked_statements": [r"(a)\1"]}))


@pytest.mark.gpu
def test_dropin_matches_reference_golden_gpu(gold):
    check_plugin(gold, None)


@pytest.mark.gpu
def test_dropin_concurrent_hooks_coalesce_gpu(gold):
    from mcp_context_forge_b200 import framework as fw
    from mcp_context_forge_b200.batching import GpuBatcher
    from mcp_context_forge_b200.plugins.sql_sanitizer import SQLSanitizerPlugin

    ctx = fw.PluginContext(global_context=fw.GlobalContext(request_id="t"))
    block = gold[2]
    plug = SQLSanitizerPlugin(fw.PluginConfig(name="sql", kind="x", hooks=["tool_pre_invoke"], config=block["config"]))
    c = ref.config(block["config"])
    cases = [x for x in block["cases"]] * 4

    async def go():
        return await asyncio.gather(*[plug.tool_pre_invoke(fw.ToolPreInvokePayload(name="t", args=x["args"]), ctx) for x in cases])

    before = GpuBatcher.get().launches
    res = run(go())
    assert GpuBatcher.get().launches - before <= 4          # scan + sub of the wave, rescan of the stripped strings
    for x, r in zip(cases, res):
        exp = ref.hook(x["args"], c, "tool args")
        assert (r.modified_payload.args if r.modified_payload is not None else None) == exp["out_args"] and r.metadata == exp["metadata"]
