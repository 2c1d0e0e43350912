"""Host-side logic on CPU: the cpex-compatible surface (contract from
/root/reference/tests/acceptance/plugins/test_cpex_contract.py), the executor semantics, the C ABI
symbol table, and plugin construction (pattern compilation happens without a GPU)."""
import asyncio
import ctypes
import os
import re

import pytest

from mcp_context_forge_b200 import _native
from mcp_context_forge_b200 import framework as fw
from mcp_context_forge_b200.cpex_compat import install_as_cpex


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


def test_c_abi_exports_every_declared_symbol():
    lib = _native.load()
    header = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "include", "cfgpu.h")).read()
    declared = set(re.findall(r"\b(cf_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cfgpu.h but not exported by libcfgpu.so"
    assert set(_native.exported_symbols()) <= declared


def test_builder_usable_without_gpu():
    from mcp_context_forge_b200 import engine

    p = engine.Program()
    p.add_search(r"\bsuicide\b", re.I)
    p.add_literal("crap")
    p.add_sub("crud", 0, "yikes")
    st = p.compile_host()
    assert st.n_patterns == 3 and st.n_ordered == 1 and st.n_states > 3


def test_contract_surface_importable_as_cpex():
    install_as_cpex(force=True)
    from cpex.framework import (ConfigLoader, GlobalContext, HookRegistry, HttpHeaderPayload, Plugin, PluginCondition, PluginConfig, PluginContext,  # noqa: F401
                                PluginContextTable, PluginError, PluginErrorModel, PluginLoader, PluginManager, PluginMode, PluginPayload, PluginResult,
                                PluginViolation, PluginViolationError, PromptHookType, PromptPosthookPayload, PromptPrehookPayload, ToolHookType,
                                ToolPostInvokePayload, ToolPostInvokeResult, ToolPreInvokePayload, ToolPreInvokeResult, get_attr, get_hook_registry)
    from cpex.framework.constants import GATEWAY_METADATA, TOOL_METADATA  # noqa: F401
    from cpex.framework.hooks.policies import HookPayloadPolicy
    from cpex.framework.models import OnError
    from cpex.framework.settings import PluginsSettings

    assert PluginPayload.model_config.get("frozen") is True
    assert {"SEQUENTIAL", "TRANSFORM", "AUDIT", "CONCURRENT", "FIRE_AND_FORGET", "DISABLED"} <= {m.name for m in PluginMode}
    assert hasattr(OnError, "IGNORE") and PluginMode.SEQUENTIAL.value == "sequential"
    assert "name" in HookPayloadPolicy(writable_fields=frozenset({"name"})).writable_fields
    p = ToolPreInvokePayload(name="t", args={"k": "v"})
    assert ToolPreInvokePayload.model_validate(p.model_dump()).args == {"k": "v"}
    with pytest.raises(Exception):
        p.name = "other"
    os.environ["PLUGINS_ENABLED"] = "false"
    os.environ["PLUGINS_PLUGIN_TIMEOUT"] = "60"
    s = PluginsSettings()
    assert s.enabled is False and s.plugin_timeout == 60


class _Upper(fw.Plugin):
    async def tool_pre_invoke(self, payload, context):
        return fw.ToolPreInvokeResult(modified_payload=payload.model_copy(update={"args": {k: v.upper() for k, v in payload.args.items()}, "name": "hijack"}))


class _Block(fw.Plugin):
    async def tool_pre_invoke(self, payload, context):
        return fw.ToolPreInvokeResult(continue_processing=False, violation=fw.PluginViolation(reason="r", description="d", code="c", details={}))


class _Boom(fw.Plugin):
    async def tool_pre_invoke(self, payload, context):
        raise ValueError("boom")


class _Seen(fw.Plugin):
    seen = []

    async def tool_pre_invoke(self, payload, context):
        _Seen.seen.append(dict(payload.args))
        return fw.ToolPreInvokeResult(metadata={"seen": True})


def _mgr(specs, policies=None, fail=False):
    cfg = fw.Config(plugins=[fw.PluginConfig(name=n, kind=f"test_framework_cpu.{k}", hooks=["tool_pre_invoke"], priority=pr, mode=mode, on_error=oe)
                             for n, k, pr, mode, oe in specs], plugin_settings=fw.PluginSettings(fail_on_plugin_error=fail))
    m = fw.PluginManager(cfg, timeout=5, hook_policies=policies)
    run(m.initialize())
    return m


def test_executor_priority_chain_and_policy():
    _Seen.seen.clear()
    pol = {"tool_pre_invoke": fw.HookPayloadPolicy(writable_fields=frozenset({"args"}))}
    m = _mgr([("seen", "_Seen", 200, fw.PluginMode.SEQUENTIAL, fw.OnError.FAIL), ("upper", "_Upper", 10, fw.PluginMode.SEQUENTIAL, fw.OnError.FAIL)], pol)
    res, ctxs = run(m.invoke_hook(fw.ToolHookType.TOOL_PRE_INVOKE, fw.ToolPreInvokePayload(name="t", args={"a": "x"}), fw.GlobalContext(request_id="1")))
    assert res.continue_processing and res.modified_payload.args == {"a": "X"}
    assert res.modified_payload.name == "t"          # `name` is not writable under this policy
    assert _Seen.seen == [{"a": "X"}]                # lower priority number ran first, payload chained
    assert res.metadata == {"seen": True} and len(ctxs) == 2
    assert m.has_hooks_for("tool_pre_invoke") and not m.has_hooks_for("tool_post_invoke")


def test_executor_violation_and_error_modes():
    gc = fw.GlobalContext(request_id="2")
    pay = fw.ToolPreInvokePayload(name="t", args={"a": "x"})
    _Seen.seen.clear()
    m = _mgr([("block", "_Block", 1, fw.PluginMode.SEQUENTIAL, fw.OnError.FAIL), ("seen", "_Seen", 2, fw.PluginMode.SEQUENTIAL, fw.OnError.FAIL)])
    res, _ = run(m.invoke_hook("tool_pre_invoke", pay, gc))
    assert not res.continue_processing and res.violation.code == "c" and res.violation.plugin_name == "block" and _Seen.seen == []
    with pytest.raises(fw.PluginViolationError):
        run(m.invoke_hook("tool_pre_invoke", pay, gc, violations_as_exceptions=True))
    m = _mgr([("block", "_Block", 1, fw.PluginMode.TRANSFORM, fw.OnError.FAIL), ("seen", "_Seen", 2, fw.PluginMode.SEQUENTIAL, fw.OnError.FAIL)])
    res, _ = run(m.invoke_hook("tool_pre_invoke", pay, gc))
    assert res.continue_processing and _Seen.seen == [{"a": "x"}]           # transform: violation only logged
    with pytest.raises(fw.PluginError):
        run(_mgr([("boom", "_Boom", 1, fw.PluginMode.SEQUENTIAL, fw.OnError.FAIL)]).invoke_hook("tool_pre_invoke", pay, gc))
    res, _ = run(_mgr([("boom", "_Boom", 1, fw.PluginMode.SEQUENTIAL, fw.OnError.IGNORE)]).invoke_hook("tool_pre_invoke", pay, gc))
    assert res.continue_processing
    res, _ = run(_mgr([("boom", "_Boom", 1, fw.PluginMode.TRANSFORM, fw.OnError.FAIL)]).invoke_hook("tool_pre_invoke", pay, gc))
    assert res.continue_processing
    with pytest.raises(fw.PluginError):
        run(_mgr([("boom", "_Boom", 1, fw.PluginMode.TRANSFORM, fw.OnError.FAIL)], fail=True).invoke_hook("tool_pre_invoke", pay, gc))
    r = run(_mgr([("seen", "_Seen", 1, fw.PluginMode.SEQUENTIAL, fw.OnError.FAIL)]).invoke_hook_for_plugin("seen", "tool_pre_invoke", pay, context=gc))
    assert r.metadata == {"seen": True}


def test_gpu_plugins_construct_and_reject_unsupported():
    from mcp_context_forge_b200.plugins.deny_filter import DenyListPlugin
    from mcp_context_forge_b200.plugins.harmful_content_detector import HarmfulContentDetectorPlugin
    from mcp_context_forge_b200.plugins.regex_filter import SearchReplacePlugin
    from mcp_context_forge_b200.regex_frontend import UnsupportedPattern

    p = SearchReplacePlugin(fw.PluginConfig(name="a", kind="x", config={"words": [{"search": "crap", "replace": "crud"}, {"search": "(bad", "replace": "x"}, {"search": "crud", "replace": r"y\\n"}]}))
    assert bin(p._rule_mask).count("1") == 2          # the invalid pattern is skipped like the reference does
    g = SearchReplacePlugin(fw.PluginConfig(name="a", kind="x", config={"words": [{"search": "(a)b", "replace": r"\1"}, {"search": "x*", "replace": "-"}]}))
    assert bin(g._rule_mask).count("1") == 2          # group references and empty matches are part of the engine
    with pytest.raises(UnsupportedPattern):
        SearchReplacePlugin(fw.PluginConfig(name="a", kind="x", config={"words": [{"search": r"(a)\1", "replace": "x"}]}))      # back-reference in the pattern
    with pytest.raises(UnsupportedPattern):
        SearchReplacePlugin(fw.PluginConfig(name="a", kind="x", config={"words": [{"search": r"(?:a?){2}", "replace": "x"}]}))   # repeat of a nullable body
    with pytest.raises(re.error):
        SearchReplacePlugin(fw.PluginConfig(name="a", kind="x", config={"words": [{"search": "(a)b", "replace": r"\2"}]}))       # bad template: as pattern.sub raises
    with pytest.raises(UnsupportedPattern):
        HarmfulContentDetectorPlugin(fw.PluginConfig(name="a", kind="x", config={"categories": {"c": [r"(?<=x)y"]}}))
    with pytest.raises(re.error):
        HarmfulContentDetectorPlugin(fw.PluginConfig(name="a", kind="x", config={"categories": {"c": [r"(unclosed"]}}))
    h = HarmfulContentDetectorPlugin(fw.PluginConfig(name="a", kind="x"))
    assert len(h._bits) == 9
    DenyListPlugin(fw.PluginConfig(name="a", kind="x", config={"words": ["a", ""]}))


def test_fast_construct_and_copy_match_pydantic():
    """The executor's constructor-free paths: same objects as `model_construct` / `model_copy`, also when the installed framework
    declares more fields than the call names (a real cpex may)."""
    from typing import Any, Optional

    from pydantic import BaseModel, Field

    from mcp_context_forge_b200.cpex_compat.framework import fast_construct, fast_copy

    class Wider(BaseModel):
        a: int
        b: Optional[dict[str, Any]] = Field(default_factory=dict)
        c: str = "dflt"
        d: Optional[int] = None

    x = fast_construct(Wider, {"a": 1})
    assert x == Wider.model_construct(a=1) and x.model_dump() == {"a": 1, "b": {}, "c": "dflt", "d": None}
    y = fast_construct(Wider, {"a": 2})
    assert y.b is not x.b                                            # default factories run per object
    z = fast_copy(x, {"c": "new"})
    assert z == x.model_copy(update={"c": "new"}) and x.c == "dflt" and z.b is x.b
    r = fast_construct(fw.PluginResult, {"continue_processing": True, "modified_payload": None, "violation": None, "metadata": {}, "retry_delay_ms": 0})
    assert r == fw.PluginResult() and r.model_dump() == fw.PluginResult().model_dump()
