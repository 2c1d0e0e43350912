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


def latency(manager_cls, cfg, pol, hook, payloads, window_us, rate_per_s, seconds=2.0):
    """Open-loop arrivals at `rate_per_s` for `seconds`: added latency per request (submit -> result) p50 / p99 in microseconds."""
    from mcp_context_forge_b200.manager import BatchedPluginManager
    kw = {"window_us": window_us} if manager_cls is BatchedPluginManager else {}
    m = manager_cls(cfg, timeout=120, hook_policies=pol, **kw)
    loop = asyncio.new_event_loop()
    loop.run_until_complete(m.initialize())
    gc = fw.GlobalContext(request_id="lat")
    lats = []

    async def one(p):
        t0 = time.perf_counter()
        await m.invoke_hook(hook, p, gc)
        lats.append(time.perf_counter() - t0)

    async def drive():
        await asyncio.gather(*[one(p) for p in payloads[:64]])       # warm-up
        lats.clear()
        tasks = []
        t0 = time.perf_counter()
        i = 0
        while True:
            now = time.perf_counter() - t0
            if now >= seconds:
                break
            due = int(now * rate_per_s)
            while i < due:
                tasks.append(asyncio.ensure_future(one(payloads[i % len(payloads)])))
                i += 1
            await asyncio.sleep(0)
        await asyncio.gather(*tasks)
        return i / (time.perf_counter() - t0)

    achieved = loop.run_until_complete(drive())
    loop.run_until_complete(m.shutdown())
    lats.sort()
    return {"window_us": window_us, "offered_per_s": rate_per_s, "completed_per_s": round(achieved, 1), "p50_us": round(lats[len(lats) // 2] * 1e6, 1),
            "p99_us": round(lats[int(len(lats) * 0.99)] * 1e6, 1), "requests": len(lats)}


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
        # ---- the chain-level batched executor (mcp_context_forge_b200.manager): one fused launch per wave
        from mcp_context_forge_b200.manager import BatchedPluginManager
        bm = BatchedPluginManager(cfg, timeout=120, hook_policies=pol)
        loop.run_until_complete(bm.initialize())

        async def bwave(hook, payloads):
            return await asyncio.gather(*[bm.invoke_hook(hook, p, gc) for p in payloads])

        for name, hook, payloads in (("prompt_pre_fetch 2 KiB (harmful+deny+regex_filter)", "prompt_pre_fetch", pre),
                                     ("tool_post_invoke 16 KiB JSON (harmful+regex_filter+toon)", "tool_post_invoke", post)):
            loop.run_until_complete(bwave(hook, payloads))
            c0, a0, d0, r0 = bm.launch_calls, bm.assemble_s, bm.device_s, bm.replay_s
            t0 = time.perf_counter()
            for _ in range(3):
                res2 = loop.run_until_complete(bwave(hook, payloads))
            dt = (time.perf_counter() - t0) / 3
            out[name]["batched_manager"] = {"payloads_per_s": round(n / dt, 1), "ms_per_wave": round(dt * 1e3, 2), "fused_launches_per_wave": (bm.launch_calls - c0) / 3,
                                            "ms_assemble_python": round((bm.assemble_s - a0) / 3 * 1e3, 2), "ms_pack_upload_kernels_download": round((bm.device_s - d0) / 3 * 1e3, 2),
                                            "ms_replay_python": round((bm.replay_s - r0) / 3 * 1e3, 2), "slow_path_calls": bm.slow_path_calls}
        loop.run_until_complete(bm.shutdown())
        if os.environ.get("HOOK_LATENCY", "1") == "1":
            out["latency_tool_post_invoke_16KiB"] = [latency(BatchedPluginManager, cfg, pol, "tool_post_invoke", post, w, 4000) for w in (0, 50, 200, 1000)]
            out["latency_tool_post_invoke_16KiB_per_plugin_coalescer"] = latency(fw.PluginManager, cfg, pol, "tool_post_invoke", post, 0, 4000)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
