"""json_repair (SURVEY §8 row f-2's named consumer of the shared JSON parse; hook tool_post_invoke): the oracle vs golden vectors recorded from
the reference's plugin file (tools/gen_golden.py json_repair); the drop-in end to end on the engine's CPU simulator — stand-alone and inside
the chain-level manager next to toon_encoder / regex_filter — and (tests/test_zz_json_repair_gpu.py) on the GPU."""
import asyncio
import json
import os
import tempfile

import pytest

import hostsim_batcher
from mcp_context_forge_b200 import framework as fw
from oracle import json_repair_ref as ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "json_repair.json")

YAML = """
plugins:
  - name: "ReplaceBadWordsPlugin"
    kind: "mcp_context_forge_b200.plugins.regex_filter.SearchReplacePlugin"
    hooks: ["tool_post_invoke"]
    mode: "sequential"
    priority: %(regex_prio)d
    config:
      words:
        - {search: "crap", replace: "crud"}
  - name: "JSONRepairPlugin"
    kind: "mcp_context_forge_b200.plugins.json_repair.JSONRepairPlugin"
    hooks: ["tool_post_invoke"]
    mode: "sequential"
    priority: 145
    config: {}
  - name: "ToonEncoder"
    kind: "mcp_context_forge_b200.plugins.toon_encoder.ToonEncoderPlugin"
    hooks: ["tool_post_invoke"]
    mode: "sequential"
    priority: 900
    config: {min_size_bytes: 10}
plugin_dirs: []
plugin_settings: {plugin_timeout: 120, fail_on_plugin_error: true}
"""


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


@pytest.fixture(scope="module")
def cases():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)["cases"]


def test_oracle_matches_reference_golden(cases):
    assert len(cases) >= 300 and sum(c["modified"] for c in cases) >= 60
    for c in cases:
        g = ref.hook(c["result"])
        assert (g["modified"], g["out_result"], g["metadata"], g["continue_processing"]) == (c["modified"], c["out_result"], c["metadata"], c["continue_processing"]), c


def check_dropin(cases, plugin_cls):
    """Every golden case through the drop-in's hook, all calls concurrent (they share launches)."""
    plug = plugin_cls(fw.PluginConfig(name="jr", kind="x", hooks=["tool_post_invoke"]))
    ctx = fw.PluginContext(global_context=fw.GlobalContext(request_id="t"))

    async def go():
        return await asyncio.gather(*[plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), ctx) for c in cases])

    for c, r in zip(cases, run(go())):
        assert r.continue_processing == c["continue_processing"], c
        assert (r.modified_payload is not None) == c["modified"], c
        if c["modified"]:
            assert r.modified_payload.result == c["out_result"] and r.modified_payload.name == "t", c
        assert (r.metadata or {}) == c["metadata"], c
    return plug


def check_in_manager(cases, regex_prio):
    """BatchedPluginManager (json_repair reads the parse status of the chain's fused launch) == the sequential executor over the same chain."""
    from mcp_context_forge_b200.manager import BatchedPluginManager

    with tempfile.TemporaryDirectory() as td:
        cfg = os.path.join(td, "plugins.yaml")
        with open(cfg, "w") as f:
            f.write(YAML % {"regex_prio": regex_prio})
        seq, bat = fw.PluginManager(cfg, timeout=120), BatchedPluginManager(cfg, timeout=120)
        loop = asyncio.new_event_loop()
        loop.run_until_complete(seq.initialize())
        loop.run_until_complete(bat.initialize())
        results = [c["result"] for c in cases] + ["{'a': 'crap',}", '{"a": "crap"}', "[1, 2, 3,] crap", {"content": [{"type": "text", "text": '[{"a":1,"b":2},{"a":3,"b":4}]'}]}]
        pls = [fw.ToolPostInvokePayload(name="t", result=r) for r in results]
        gcs = [fw.GlobalContext(request_id=f"r{i}") for i in range(len(pls))]

        async def wave(m):
            return await asyncio.gather(*[m.invoke_hook("tool_post_invoke", p, g) for p, g in zip(pls, gcs)])

        a = loop.run_until_complete(wave(seq))
        b = loop.run_until_complete(wave(bat))

        def norm(x):
            res = x[0]
            return (res.continue_processing, res.modified_payload.result if res.modified_payload is not None else None,
                    {k: v for k, v in (res.metadata or {}).items() if k != "conversion_time_ms"})

        bad = [(i, norm(x), norm(y)) for i, (x, y) in enumerate(zip(a, b)) if norm(x) != norm(y)]
        assert not bad, bad[:3]
        assert bat.waves >= 1 and bat.slow_path_calls > 0          # (texts that do not parse run the hook itself)
        n_mod = sum(1 for x in b if x[0].modified_payload is not None)
        assert n_mod >= sum(c["modified"] for c in cases)
        return bat


def test_dropin_on_the_cpu_simulator(cases, monkeypatch):
    from mcp_context_forge_b200.plugins.json_repair import JSONRepairPlugin

    sim = hostsim_batcher.install(monkeypatch)
    check_dropin(cases, JSONRepairPlugin)
    assert sim.launches < len(cases)                               # concurrent hook calls were coalesced


@pytest.mark.parametrize("regex_prio", [50, 150])
def test_dropin_inside_the_batched_manager_on_the_cpu_simulator(cases, monkeypatch, regex_prio):
    hostsim_batcher.install(monkeypatch)
    check_in_manager(cases, regex_prio)
