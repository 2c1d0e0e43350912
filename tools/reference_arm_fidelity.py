"""Container-only: is the CPU arm of bench.py (`--impl reference` / `cpu_baseline`: the oracle port of the chain) as fast as the reference's
own plugin classes?  Times, on one core and the bench's payload mix, (a) HarmfulContentDetectorPlugin + SearchReplacePlugin + ToonEncoderPlugin
imported unmodified from /root/reference (orjson stood in for by stdlib json) and (b) bench._cpu_chain (oracle/hook_chain_ref + oracle/toon_ref).
usage: python tools/reference_arm_fidelity.py [payloads]"""
import asyncio
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_golden  # noqa: E402


def main() -> int:
    if not os.path.isdir(gen_golden.REF):
        print("reference_arm_fidelity: /root/reference is not here (container-only tool)")
        return 0
    gen_golden.install_shims()
    logging.disable(logging.ERROR)
    import bench
    from cpex.framework import GlobalContext, PluginConfig, PluginContext, ToolPostInvokePayload
    from plugins.harmful_content_detector.harmful_content_detector import HarmfulContentDetectorPlugin
    from plugins.regex_filter.search_replace import SearchReplacePlugin
    from plugins.toon_encoder.toon_encoder import ToonEncoderPlugin

    payloads = bench.make_payloads(int(sys.argv[1]) if len(sys.argv) > 1 else 48)
    ctx = PluginContext(global_context=GlobalContext(request_id="x"))
    h = HarmfulContentDetectorPlugin(PluginConfig(name="h", kind="x", hooks=["tool_post_invoke"]))
    r = SearchReplacePlugin(PluginConfig(name="r", kind="x", hooks=["tool_post_invoke"], config={"words": [{"search": s, "replace": t} for s, _f, t in bench.SUBS]}))
    t = ToonEncoderPlugin(PluginConfig(name="t", kind="x", hooks=["tool_post_invoke"]))
    loop = asyncio.new_event_loop()

    def ref_chain(p):
        pl = ToolPostInvokePayload(name="t", result={"content": [{"type": "text", "text": p}]})
        loop.run_until_complete(h.tool_post_invoke(pl, ctx))
        b = loop.run_until_complete(r.tool_post_invoke(pl, ctx))
        loop.run_until_complete(t.tool_post_invoke(b.modified_payload or pl, ctx))

    bench._cpu_init()
    out = {}
    for name, fn in (("reference_plugin_classes_ms_per_payload", ref_chain), ("oracle_chain_ms_per_payload", lambda p: bench._cpu_chain([p]))):
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            for p in payloads:
                fn(p)
            best = min(best, (time.perf_counter() - t0) / len(payloads) * 1e3)
        out[name] = round(best, 3)
    print(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
