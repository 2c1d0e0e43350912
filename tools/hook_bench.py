"""Throughput of the drop-in path as the gateway drives it: concurrent `PluginManager.invoke_hook`
calls on one asyncio loop (one gateway worker), the GPU plugins coalescing whatever is in flight
into one launch per plugin per wave.  Beside it: the same hooks computed by the oracle chain
(the reference plugins' loops on CPython `re` / the TOON restatement) on one core, which is what one
gateway worker's event loop gives the reference.  Prints one JSON object.  Measurement aid; the
contract metric is bench.py's."""
import asyncio
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, ".")
import bench
from mcp_context_forge_b200 import framework as fw, synth
from mcp_context_forge_b200.batching import GpuBatcher
from mcp_context_forge_b200.cpex_compat.framework import HookPayloadPolicy
from oracle import hook_chain_ref as ref, toon_ref

YAML = """
plugins:
  - name: "HarmfulContentDetector"
    kind: "mcp_context_forge_b200.plugins.harmful_content_detector.HarmfulContentDetectorPlugin"
    hooks: ["prompt_pre_fetch", "tool_post_invoke"]
    mode: "sequential"
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
    priority: 150
    config:
      words:
        - {search: crap, replace: crud}
        - {search: crud, replace: yikes}
  - name: "ToonEncoder"
    kind: "mcp_context_forge_b200.plugins.toon_encoder.ToonEncoderPlugin"
    hooks: ["tool_post_invoke"]
    mode: "sequential"
    priority: 900
plugin_settings:
  plugin_timeout: 120
"""


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    with tempfile.TemporaryDirectory() as td:
        cfg = os.path.join(td, "plugins.yaml")
        open(cfg, "w").write(YAML)
        pol = {"tool_pre_invoke": HookPayloadPolicy(writable_fields=frozenset({"name", "args", "headers"})),
               "tool_post_invoke": HookPayloadPolicy(writable_fields=frozenset({"result"})),
               "prompt_pre_fetch": HookPayloadPolicy(writable_fields=frozenset({"args"}))}
        m = fw.PluginManager(cfg, timeout=120, hook_policies=pol)
        loop = asyncio.new_event_loop()
        loop.run_until_complete(m.initialize())
        gc = fw.GlobalContext(request_id="bench")

        small = [synth.payload("C", 2048, seed=i, hit_rate=1e-4) for i in range(64)]
        big = [synth.payload("A", 16384, seed=i, hit_rate=1e-4) for i in range(64)]
        pre = [fw.PromptPrehookPayload(prompt_id="p", args={"q": small[i % 64]}) for i in range(n)]
        post = [fw.ToolPostInvokePayload(name="t", result={"content": [{"type": "text", "text": big[i % 64]}]}) for i in range(n)]

        async def wave(hook, payloads):
            return await asyncio.gather(*[m.invoke_hook(hook, p, gc) for p in payloads])

        out = {"requests_per_wave": n}
        for name, hook, payloads in (("prompt_pre_fetch 2 KiB (harmful+deny+regex_filter)", "prompt_pre_fetch", pre),
                                     ("tool_post_invoke 16 KiB JSON (harmful+regex_filter+toon)", "tool_post_invoke", post)):
            loop.run_until_complete(wave(hook, payloads))           # warm-up at full size (compiles, buffer growth)
            b = GpuBatcher.get()
            l0 = b.launches
            prof = None
            if os.environ.get("HOOK_PROF"):
                import cProfile
                prof = cProfile.Profile()
                prof.enable()
            t0 = time.perf_counter()
            res = loop.run_until_complete(wave(hook, payloads))
            dt = time.perf_counter() - t0
            if prof is not None:
                import pstats
                prof.disable()
                pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(22)
            out[name] = {"gpu_plugins_payloads_per_s": round(n / dt, 1), "ms_per_wave": round(dt * 1e3, 2), "coalesced_launches": b.launches - l0}
            # oracle chain, one core, on a bounded sample; also the parity spot check
            rules = ref.regex_compile_rules([{"search": "crap", "replace": "crud"}, {"search": "crud", "replace": "yikes"}])
            cats = ref.harmful_compile()
            k = min(n, 128)
            t0 = time.perf_counter()
            for i in range(k):
                if hook == "prompt_pre_fetch":
                    a = pre[i].args
                    [ref.harmful_scan_text(v, cats) for v in a.values()]
                    ref.deny_first_hit(bench.DENY, a)
                    exp = ref.regex_apply_dict(rules, a)
                else:
                    r = post[i].result
                    ref.harmful_tool_post(r, cats)
                    r2 = ref.regex_apply_dict(rules, r) if isinstance(r, dict) else r
                    exp = toon_ref.process_text(r["content"][0]["text"])
            dt_ref = time.perf_counter() - t0
            out[name]["oracle_chain_1core_payloads_per_s"] = round(k / dt_ref, 1)
            # parity spot check on the last sample element
            got = res[k - 1][0]
            if hook == "prompt_pre_fetch":
                ok = got.modified_payload is None or got.modified_payload.args == exp or not got.continue_processing
            else:
                mp = got.modified_payload
                txt = mp.result["content"][0]["text"] if mp is not None else post[k - 1].result["content"][0]["text"]
                ok = (exp is None and txt == post[k - 1].result["content"][0]["text"]) or txt == exp
            out[name]["spot_check_vs_oracle"] = bool(ok)
        loop.run_until_complete(m.shutdown())
        print(json.dumps(out))


if __name__ == "__main__":
    main()
