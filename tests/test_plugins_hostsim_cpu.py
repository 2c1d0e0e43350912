"""The drop-in plugins and the chain-level manager END TO END on the CPU simulator of the engine (tests/hostsim_batcher.py): the
plugins' own host logic and the manager's speculate / replay against the golden vectors recorded from the reference's plugin files and
against the sequential executor.  The same checks run against the real kernels in tests/test_plugins_gpu.py / test_manager_gpu.py."""
import asyncio
import json
import os
import tempfile

import pytest

import hostsim_batcher
from mcp_context_forge_b200 import framework as fw

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CTX = fw.PluginContext(global_context=fw.GlobalContext(request_id="t"))


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


def load(name):
    with open(os.path.join(GOLD, name), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture()
def sim(monkeypatch):
    return hostsim_batcher.install(monkeypatch)


@pytest.mark.parametrize("which", ["pattern_plugins.json", "regex_filter_templates.json"])
def test_regex_filter_golden(sim, which):
    from mcp_context_forge_b200.plugins.regex_filter import SearchReplacePlugin

    blocks = load(which)
    blocks = blocks["regex_filter"] if isinstance(blocks, dict) else blocks
    n = 0
    for block in blocks:
        plug = SearchReplacePlugin(fw.PluginConfig(name="rf", kind="x", hooks=["tool_pre_invoke", "tool_post_invoke"], config={"words": block["words"]}))

        async def go():
            return await asyncio.gather(*[plug.tool_pre_invoke(fw.ToolPreInvokePayload(name="t", args=c["args"]), CTX) if c["hook"] == "tool_pre_invoke"
                                          else plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), CTX) for c in block["cases"]])

        for c, r in zip(block["cases"], run(go())):
            if c["hook"] == "tool_pre_invoke":
                assert r.modified_payload.args == c["out_args"], c
            else:
                assert r.modified_payload.result == c["out_result"], c
            n += 1
    assert n >= 200 and sim.launches > 0


def test_deny_and_harmful_golden(sim):
    from mcp_context_forge_b200.plugins.deny_filter import DenyListPlugin
    from mcp_context_forge_b200.plugins.harmful_content_detector import HarmfulContentDetectorPlugin

    gold = load("pattern_plugins.json")
    for block in gold["deny_filter"]:
        plug = DenyListPlugin(fw.PluginConfig(name="dl", kind="x", hooks=["prompt_pre_fetch"], config={"words": block["words"]}))
        for c in block["cases"]:
            r = run(plug.prompt_pre_fetch(fw.PromptPrehookPayload(prompt_id="p", args=c["args"]), CTX))
            assert (not r.continue_processing) == c["blocked"], c
            if c["violation"]:
                assert r.violation.model_dump(include=set(c["violation"])) == c["violation"]
    for block in gold["harmful"]:
        plug = HarmfulContentDetectorPlugin(fw.PluginConfig(name="hc", kind="x", hooks=["tool_post_invoke"], config=block["config"]))
        for c in block["cases"]:
            r = run(plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), CTX))
            assert r.continue_processing == c["continue_processing"] and r.metadata == c["metadata"], c
            got = json.loads(json.dumps(r.violation.model_dump(include=set(c["violation"])))) if r.violation else None     # tuples -> lists, like the recorded vectors
            assert got == c["violation"], c


def test_toon_encoder_plugin_golden(sim):
    from mcp_context_forge_b200.plugins.toon_encoder import ToonEncoderPlugin

    n = 0
    for block in load("toon.json")["plugin"]:
        plug = ToonEncoderPlugin(fw.PluginConfig(name="te", kind="x", hooks=["tool_post_invoke"], config=block["config"]))
        for c in block["cases"]:
            if "raises" in c:                                   # skip_on_error: false — the reference's exception type (the simulator honours CF_TOON_REPORT_ERRORS)
                with pytest.raises((ValueError, AttributeError)) as ei:
                    run(plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), CTX))
                assert (type(ei.value).__name__, str(ei.value)) == (c["raises"], c["message"]), c["result"]
                continue
            r = run(plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=c["result"]), CTX))
            if c["modified"] is None:
                assert r.modified_payload is None and r.continue_processing, c["result"]
            else:
                assert r.modified_payload.result == c["modified"], c["result"]
                md = dict(r.metadata)
                assert isinstance(md.pop("conversion_time_ms"), float)
                assert md == c["metadata"]
            n += 1
        assert plug.get_stats() == block["stats"]
    assert n >= 300


def test_batched_manager_equals_sequential_manager(sim):
    """BatchedPluginManager (one fused launch per wave, speculate / replay) == PluginManager (plugin after plugin) — on the simulator."""
    import test_manager_gpu as tm
    from mcp_context_forge_b200.manager import BatchedPluginManager

    for harm_mode, regex_prio in (("sequential", 150), ("transform", 50)):
        with tempfile.TemporaryDirectory() as td:
            cfg = os.path.join(td, "plugins.yaml")
            with open(cfg, "w") as f:
                f.write(tm.YAML % {"harm_mode": harm_mode, "regex_prio": regex_prio})
            seq, bat = fw.PluginManager(cfg, timeout=120, hook_policies=tm.POL), BatchedPluginManager(cfg, timeout=120, hook_policies=tm.POL)
            loop = asyncio.new_event_loop()
            loop.run_until_complete(seq.initialize())
            loop.run_until_complete(bat.initialize())
            pre, tpre, post = tm.payloads(60, 11)
            gcs = [fw.GlobalContext(request_id=f"r{i}") for i in range(len(pre))]
            for hook, pls in (("prompt_pre_fetch", pre), ("tool_pre_invoke", tpre), ("tool_post_invoke", post)):
                for vae in (False, True):
                    async def wave(m):
                        return await asyncio.gather(*[m.invoke_hook(hook, p, g, None, vae) for p, g in zip(pls, gcs)], return_exceptions=True)
                    a = loop.run_until_complete(wave(seq))
                    before = bat.launch_calls
                    b = loop.run_until_complete(wave(bat))
                    assert bat.launch_calls - before == 1
                    bad = [(i, tm.norm(x), tm.norm(y)) for i, (x, y) in enumerate(zip(a, b)) if tm.norm(x) != tm.norm(y)]
                    assert not bad, (hook, vae, bad[:2])
            if regex_prio == 50:
                assert bat.slow_path_calls > 0


def test_batched_manager_serves_two_event_loops(sim):
    """One BatchedPluginManager, two threads with an event loop each (a process may run one loop per thread): waves are per loop — every future is
    resolved on the loop that made it — and the launches of both are serialised; results equal the sequential manager's."""
    import threading

    import test_manager_gpu as tm
    from mcp_context_forge_b200.manager import BatchedPluginManager

    with tempfile.TemporaryDirectory() as td:
        cfg = os.path.join(td, "plugins.yaml")
        with open(cfg, "w") as f:
            f.write(tm.YAML % {"harm_mode": "sequential", "regex_prio": 150})
        seq, bat = fw.PluginManager(cfg, timeout=120, hook_policies=tm.POL), BatchedPluginManager(cfg, timeout=120, hook_policies=tm.POL)
        run(seq.initialize())
        run(bat.initialize())
        sets = [tm.payloads(40, 21), tm.payloads(40, 22)]
        out, errs = {}, []

        def worker(k):
            try:
                loop = asyncio.new_event_loop()
                pre, tpre, post = sets[k]
                gcs = [fw.GlobalContext(request_id=f"w{k}r{i}") for i in range(len(pre))]
                res = []
                for _ in range(3):
                    for hook, pls in (("prompt_pre_fetch", pre), ("tool_pre_invoke", tpre), ("tool_post_invoke", post)):
                        async def wave():
                            return await asyncio.gather(*[bat.invoke_hook(hook, p, g) for p, g in zip(pls, gcs)], return_exceptions=True)
                        res.append(loop.run_until_complete(wave()))
                out[k] = res
                loop.close()
            except BaseException as exc:  # noqa: BLE001
                errs.append(exc)

        ts = [threading.Thread(target=worker, args=(k,)) for k in (0, 1)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(300)
        assert not errs, errs
        for k in (0, 1):
            pre, tpre, post = sets[k]
            gcs = [fw.GlobalContext(request_id=f"w{k}r{i}") for i in range(len(pre))]
            exp = []
            for hook, pls in (("prompt_pre_fetch", pre), ("tool_pre_invoke", tpre), ("tool_post_invoke", post)):
                async def wave():
                    return await asyncio.gather(*[seq.invoke_hook(hook, p, g) for p, g in zip(pls, gcs)], return_exceptions=True)
                exp.append(run(wave()))
            for r in range(3):
                for h in range(3):
                    assert [tm.norm(x) for x in out[k][r * 3 + h]] == [tm.norm(x) for x in exp[h]], (k, r, h)
        assert not bat._pending and not bat._busy


def test_toon_encoder_raises_with_the_reference_wording(sim):
    """`skip_on_error: false`: the kernel's status says WHICH error (ValueError / AttributeError, in the encoder's order); the exception's wording —
    the code point of the first control character a quoted string holds, the type that has no `.keys()` — is the reference's (toon.py:279, :480),
    checked against the reference itself by tools/fuzz_vs_reference.py and fuzz_plugins_vs_reference.py."""
    from mcp_context_forge_b200.plugins.toon_encoder import ToonEncoderPlugin, _first_error

    plug = ToonEncoderPlugin(fw.PluginConfig(name="te", kind="x", hooks=["tool_post_invoke"], config={"min_size_bytes": 5, "skip_on_error": False}))

    def call(doc):
        res = {"content": [{"type": "text", "text": json.dumps(doc)}]}
        return run(plug.tool_post_invoke(fw.ToolPostInvokePayload(name="t", result=res), CTX))

    for doc, exc, msg in [({"k": "ctrl\x01"}, ValueError, "Cannot encode control character U+0001 in TOON"),
                          ([{"c": "ok"}, {"c": "tab\tbell\x07 and \x02"}], ValueError, "Cannot encode control character U+0007 in TOON"),
                          ({"a\x1fb": 1, "z": "\x00"}, ValueError, "Cannot encode control character U+001F in TOON"),
                          ([{"x": [1, 2]}, 5], AttributeError, "'int' object has no attribute 'keys'"),
                          ({"m": [{"x": ["s"], "y": "\x01"}, 5]}, AttributeError, "'str' object has no attribute 'keys'"),
                          ([2, {"A.b": "null"}, {"A\b": [0, "trail ", True]}], ValueError, "Cannot encode control character U+0008 in TOON")]:
        with pytest.raises(exc) as ei:
            call(doc)
        assert str(ei.value) == msg, doc
    # the keys of a columnar table are never quoted: no error, the table is written
    assert _first_error([{"a\x01": 1, "b": 2}, {"a\x01": 3, "b": 4}]) is None
    assert call([{"a\x01": 1, "bbbbbbbb": 2}, {"a\x01": 3, "bbbbbbbb": 4}]).modified_payload is not None
