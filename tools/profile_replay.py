"""Development aid: where does the host time of a BatchedPluginManager wave go?  The device call is replaced by a stub (every
unit: no match, TOON text = a fixed string), so this runs without a GPU and times/profiles the Python around the launch only.
usage: python tools/profile_replay.py [requests] [--profile]"""
import asyncio
import cProfile
import logging
import os
import pstats
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from mcp_context_forge_b200 import engine, framework as fw, manager as mgr  # noqa: E402
from mcp_context_forge_b200.cpex_compat.framework import HookPayloadPolicy  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 1024
logging.disable(logging.WARNING)


class _Ctx:
    pass


engine.Context.get = classmethod(lambda cls, device=0: _Ctx())
engine.Program.compile = lambda self, ctx: self.compile_host() and self


def fake_launch(self, chain, units, stages):
    return [mgr.UnitResult(0, None, 0, b"rows[2]{a,b}:\n  1,2\n  3,4") for _ in units]


mgr.BatchedPluginManager._launch = fake_launch
payloads = bench.make_payloads(64)
with tempfile.TemporaryDirectory() as td:
    cfg = os.path.join(td, "plugins.yaml")
    open(cfg, "w").write(bench.CHAIN_YAML)
    m = mgr.BatchedPluginManager(cfg, timeout=300, hook_policies={"tool_post_invoke": HookPayloadPolicy(writable_fields=frozenset({"result"}))})
    loop = asyncio.new_event_loop()
    loop.run_until_complete(m.initialize())
    gc = fw.GlobalContext(request_id="bench")
    posts = [fw.ToolPostInvokePayload(name="t", result={"content": [{"type": "text", "text": payloads[i % len(payloads)]}]}) for i in range(N)]

    async def wave():
        return await asyncio.gather(*[m.invoke_hook("tool_post_invoke", p, gc) for p in posts])

    for _ in range(2):
        loop.run_until_complete(wave())
    a0, r0 = m.assemble_s, m.replay_s
    t0 = time.perf_counter()
    W = 5
    if "--profile" in sys.argv:
        pr = cProfile.Profile()
        pr.enable()
    for _ in range(W):
        loop.run_until_complete(wave())
    if "--profile" in sys.argv:
        pr.disable()
        pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
    dt = (time.perf_counter() - t0) / W
    print(f"{N} requests/wave: {dt * 1e3:.1f} ms/wave = {dt / N * 1e6:.1f} us/request; assemble {(m.assemble_s - a0) / W / N * 1e6:.1f} us, replay {(m.replay_s - r0) / W / N * 1e6:.1f} us")
