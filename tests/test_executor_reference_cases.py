"""Row a-1 (hook-chain executor): the scenarios the REFERENCE'S OWN tests hold for the real cpex executor, restated with in-file fixture
plugins so that they run anywhere (the reference's files themselves run unmodified against the same classes in this container:
tools/run_reference_tests.py, outcome committed as tests/golden/reference_tests_run_framework.json, test_reference_own_tests_cpu.py).
Each case cites the reference test whose expectation it carries."""
import asyncio
import logging
import sys

import pytest

from mcp_context_forge_b200 import framework as fw
from mcp_context_forge_b200.cpex_compat import install_as_cpex
from mcp_context_forge_b200.cpex_compat.framework import (AgentPostInvokePayload, AgentPostInvokeResult, AgentPreInvokePayload, AgentPreInvokeResult, Config,
                                                          CopyOnWriteDict, HookRef, PluginExecutor, PluginRef, PromptPosthookPayload)


def run(coro):
    return asyncio.new_event_loop().run_until_complete(coro)


class Msg:
    def __init__(self, text):
        self.text = text


class MessageFilter(fw.Plugin):
    """tests/unit/mcpgateway/plugins/fixtures/plugins/agent_plugins.py `MessageFilterAgentPlugin`, restated: drops messages with a blocked word;
    all dropped => violation BLOCKED_CONTENT; some dropped => modified payload."""

    def _filter(self, payload, result_cls, payload_cls, what):
        words = self.config.config.get("blocked_words", [])
        kept = [m for m in payload.messages if not any(w in m.text.lower() for w in words)]
        if not kept and payload.messages:
            return result_cls(continue_processing=False, violation=fw.PluginViolation(code="BLOCKED_CONTENT", reason=f"All {what} contained blocked content", description="d"))
        if len(kept) != len(payload.messages):
            return result_cls(modified_payload=payload_cls(agent_id=payload.agent_id, messages=kept))
        return result_cls(continue_processing=True)

    async def agent_pre_invoke(self, payload, context):
        return self._filter(payload, AgentPreInvokeResult, AgentPreInvokePayload, "messages")

    async def agent_post_invoke(self, payload, context):
        return self._filter(payload, AgentPostInvokeResult, AgentPostInvokePayload, "response messages")


class ContextTracking(fw.Plugin):
    """… `ContextTrackingAgentPlugin`: counts in the plugin's local context in the pre hook, verifies it in the post hook."""

    async def agent_pre_invoke(self, payload, context):
        context.metadata["invocation_count"] = context.metadata.get("invocation_count", 0) + 1
        context.metadata["agent_id"] = payload.agent_id
        return AgentPreInvokeResult(continue_processing=True)

    async def agent_post_invoke(self, payload, context):
        context.metadata["context_verified"] = context.metadata.get("invocation_count", 0) > 0 and context.metadata.get("agent_id", "") == payload.agent_id
        return AgentPostInvokeResult(continue_processing=True)


class SecondCallViolates(fw.Plugin):
    """Stands in for the external rate limiter of tests/integration/test_rate_limiter.py:644-745 (limit 1/s): the second call violates."""

    def __init__(self, config):
        super().__init__(config)
        self.calls = 0

    async def tool_pre_invoke(self, payload, context):
        self.calls += 1
        if self.calls > 1:
            return fw.ToolPreInvokeResult(continue_processing=False, violation=fw.PluginViolation(code="RATE_LIMIT", reason="Rate limit exceeded", description="d", http_status_code=429))
        return fw.ToolPreInvokeResult(continue_processing=True)


def _manager(cls, mode="sequential", **config):
    pc = fw.PluginConfig(name="P", kind=f"{__name__}.{cls.__name__}", hooks=["agent_pre_invoke", "agent_post_invoke"], mode=mode, priority=50, config=config)
    m = fw.PluginManager(Config(plugins=[pc]))
    run(m.initialize())
    return m


def _msgs(*texts):
    return [Msg(t) for t in texts]


def test_filter_plugin_violation_raises_with_its_code():
    """test_agent_plugins.py:64-92 and :96-124 — clean messages pass untouched; all blocked => PluginViolationError carrying code / reason."""
    m = _manager(MessageFilter, blocked_words=["spam", "malware", "phishing"])
    g = fw.GlobalContext(request_id="test-req-2")
    res, _ = run(m.invoke_hook(fw.AgentHookType.AGENT_PRE_INVOKE, AgentPreInvokePayload(agent_id="a", messages=_msgs("Hello agent!")), global_context=g))
    assert res.continue_processing is True and res.modified_payload is None and res.violation is None
    with pytest.raises(fw.PluginViolationError) as ei:
        run(m.invoke_hook(fw.AgentHookType.AGENT_PRE_INVOKE, AgentPreInvokePayload(agent_id="a", messages=_msgs("Click here for spam offers!")), global_context=g,
                          violations_as_exceptions=True))
    assert ei.value.violation.code == "BLOCKED_CONTENT" and "blocked content" in ei.value.violation.reason.lower()
    with pytest.raises(fw.PluginViolationError) as ei:
        run(m.invoke_hook(fw.AgentHookType.AGENT_POST_INVOKE, AgentPostInvokePayload(agent_id="a", messages=_msgs("This looks like malware to me.")), global_context=g,
                          violations_as_exceptions=True))
    assert ei.value.violation.code == "BLOCKED_CONTENT"
    # without violations_as_exceptions the same violation comes back in the result (tool_service.py:5866-5889 reads it there)
    res, _ = run(m.invoke_hook(fw.AgentHookType.AGENT_PRE_INVOKE, AgentPreInvokePayload(agent_id="a", messages=_msgs("spam")), global_context=g))
    assert res.continue_processing is False and res.violation.code == "BLOCKED_CONTENT" and res.violation.plugin_name == "P"


def test_filter_plugin_partial_filtering_modifies_payload():
    """test_agent_plugins.py:127-150 — only the blocked message is removed, order kept, the chain result carries the modified payload."""
    m = _manager(MessageFilter, blocked_words=["spam"])
    res, _ = run(m.invoke_hook(fw.AgentHookType.AGENT_PRE_INVOKE, AgentPreInvokePayload(agent_id="a", messages=_msgs("Hello agent!", "Check out this spam!", "What's the weather?")),
                               global_context=fw.GlobalContext(request_id="test-req-4")))
    assert res.modified_payload is not None
    assert [x.text for x in res.modified_payload.messages] == ["Hello agent!", "What's the weather?"]


def test_local_context_persists_from_pre_to_post_hook():
    """test_agent_plugins.py:154-183 — the contexts returned by the pre hook, passed as `local_contexts`, are the ones the post hook sees."""
    m = _manager(ContextTracking)
    g = fw.GlobalContext(request_id="test-req-5")
    pre, contexts = run(m.invoke_hook(fw.AgentHookType.AGENT_PRE_INVOKE, AgentPreInvokePayload(agent_id="test-agent-123", messages=_msgs("Hello!")), global_context=g))
    assert pre.continue_processing is True and contexts
    post, contexts2 = run(m.invoke_hook(fw.AgentHookType.AGENT_POST_INVOKE, AgentPostInvokePayload(agent_id="test-agent-123", messages=_msgs("Hi there!")), global_context=g,
                                        local_contexts=contexts))
    assert post.continue_processing is True
    (ctx,) = contexts2.values()
    assert ctx is next(iter(contexts.values()))
    assert ctx.metadata == {"invocation_count": 1, "agent_id": "test-agent-123", "context_verified": True}
    # without the table the post hook starts from an empty context
    _, fresh = run(m.invoke_hook(fw.AgentHookType.AGENT_POST_INVOKE, AgentPostInvokePayload(agent_id="test-agent-123", messages=[]), global_context=g))
    assert next(iter(fresh.values())).metadata == {"context_verified": False}


def _hook_ref(mode):
    plugin = SecondCallViolates(fw.PluginConfig(name="RateLimiter", kind="x", hooks=["tool_pre_invoke"], priority=100, mode=mode, config={}))
    return plugin, HookRef("tool_pre_invoke", PluginRef(plugin))


def test_executor_transform_mode_suppresses_the_violation_and_logs_it(caplog):
    """tests/integration/test_rate_limiter.py:665-697 — `execute_plugin` in TRANSFORM mode: no PluginViolationError even with
    violations_as_exceptions, `result.violation is None`, and a WARNING containing "raised violation" on `cpex.framework.manager`."""
    _plugin, ref = _hook_ref(fw.PluginMode.TRANSFORM)
    ex = PluginExecutor(timeout=5)
    ctx = fw.PluginContext(global_context=fw.GlobalContext(request_id="r1", user="alice"))
    payload = fw.ToolPreInvokePayload(name="tool", arguments={})
    run(ex.execute_plugin(ref, payload, ctx, violations_as_exceptions=True))
    with caplog.at_level("WARNING", logger="cpex.framework.manager"):
        result = run(ex.execute_plugin(ref, payload, ctx, violations_as_exceptions=True))
    assert result.violation is None and result.continue_processing is True
    assert any("raised violation" in r.getMessage() and r.name == "cpex.framework.manager" and r.levelno == logging.WARNING for r in caplog.records)


def test_executor_sequential_mode_raises():
    """tests/integration/test_rate_limiter.py:699-718 — the same plugin in SEQUENTIAL (enforce) mode raises on its second call."""
    _plugin, ref = _hook_ref(fw.PluginMode.SEQUENTIAL)
    ex = PluginExecutor(timeout=5)
    ctx = fw.PluginContext(global_context=fw.GlobalContext(request_id="r1", user="alice"))
    payload = fw.ToolPreInvokePayload(name="tool", arguments={})
    run(ex.execute_plugin(ref, payload, ctx, violations_as_exceptions=True))
    with pytest.raises(fw.PluginViolationError) as ei:
        run(ex.execute_plugin(ref, payload, ctx, violations_as_exceptions=True))
    assert ei.value.violation.http_status_code == 429


def test_executor_execute_skips_a_disabled_plugin():
    """tests/integration/test_rate_limiter.py:721-749 — `execute()` (the chain) skips a DISABLED plugin: never a violation, plugin never called."""
    plugin, ref = _hook_ref(fw.PluginMode.DISABLED)
    ex = PluginExecutor(timeout=5)
    g = fw.GlobalContext(request_id="r1", user="alice")
    for _ in range(10):
        result, _ctxs = run(ex.execute([ref], fw.ToolPreInvokePayload(name="tool", arguments={}), g, "tool_pre_invoke", violations_as_exceptions=True))
        assert result.violation is None
    assert plugin.calls == 0 and plugin.mode == fw.PluginMode.DISABLED


def test_observability_setter_and_shared_state_of_the_bare_manager(tmp_path):
    """tests/unit/mcpgateway/plugins/test_observability_adapter.py:258-293 — `manager.observability = x` is visible through the manager, its
    executor and a second bare `PluginManager()`; `None` clears it; `PluginManager.reset()` ends the sharing."""
    cfg = tmp_path / "valid_no_plugin.yaml"
    cfg.write_text("plugins: []\nplugin_dirs: []\nplugin_settings:\n  plugin_timeout: 30\n")
    fw.PluginManager.reset()
    m = fw.PluginManager(str(cfg))
    assert m.observability is None
    obs = object()
    m.observability = obs
    assert m.observability is obs and m._executor.observability is obs
    assert fw.PluginManager().observability is obs
    m.observability = None
    assert fw.PluginManager().observability is None
    fw.PluginManager.reset()
    other = fw.PluginManager(str(cfg), observability=obs)             # a configured manager owns its state
    assert other.observability is obs and m.observability is None
    fw.PluginManager.reset()


def test_prompt_result_given_as_a_mapping_is_a_model():
    """tests/unit/mcpgateway/plugins/plugins/test_prompt_output_sentinel.py:61-81 — a dict `result` is addressable as `.messages[i].content.text`."""
    p = PromptPosthookPayload(prompt_id="prompt-2", result={"description": "d", "messages": [{"role": "user", "content": {"type": "text", "text": "Rendered dict body"}}]})
    assert p.result.messages[0].content.text == "Rendered dict body" and p.result.description == "d"
    keep = object()
    assert PromptPosthookPayload(prompt_id="p", result=keep).result is keep          # the gateway's own PromptResult instances pass through
    assert PromptPosthookPayload(prompt_id="p", result={"x": 1}).result == {"x": 1}


def test_module_paths_the_reference_imports_from():
    """Import paths used under /root/reference (tests/integration/test_rate_limiter.py:44-46, tests/unit/plugins/test_sql_sanitizer.py:4,
    plugins/*/…: `cpex.framework.hooks.resources`, `.hooks.http`)."""
    if not install_as_cpex():
        pytest.skip("a real cpex is installed")
    from cpex.framework.base import HookRef as H, PluginRef as R
    from cpex.framework.errors import PluginError, PluginViolationError
    from cpex.framework.hooks.http import HttpAuthResolveUserPayload, HttpHeaderPayload
    from cpex.framework.hooks.resources import ResourceHookType, ResourcePostFetchPayload, ResourcePreFetchPayload, ResourcePreFetchResult
    from cpex.framework.manager import PluginExecutor as E, PluginManager as M
    from cpex.framework.memory import CopyOnWriteDict as C

    assert (H, R, E, M, C) == (HookRef, PluginRef, PluginExecutor, fw.PluginManager, CopyOnWriteDict)
    assert PluginError is fw.PluginError and PluginViolationError is fw.PluginViolationError
    assert all(x is not None for x in (HttpAuthResolveUserPayload, HttpHeaderPayload, ResourceHookType, ResourcePostFetchPayload, ResourcePreFetchPayload, ResourcePreFetchResult))
    assert "cpex.framework.manager" in sys.modules


def test_copy_on_write_dict_is_a_dict_that_never_writes_through():
    """tests/unit/plugins/test_sql_sanitizer.py:39-57 builds tool arguments with it; the plugins walk them with `isinstance(x, dict)`."""
    src = {"path": "sql.txt", "edits": [{"new": "DROP table tab1;"}]}
    d = CopyOnWriteDict(src)
    assert isinstance(d, dict) and d == src and list(d.items()) == list(src.items())
    d["path"] = "other"
    del d["edits"]
    assert src == {"path": "sql.txt", "edits": [{"new": "DROP table tab1;"}]} and d == {"path": "other"} and d.modified == {"path", "edits"} and d.original is src
    assert fw.ToolPreInvokePayload(name="echo", args=CopyOnWriteDict({"message": "x"})).args == {"message": "x"}


def test_manager_internals_the_gateway_itself_reads():
    """The gateway reads two protected attributes of the manager (mcpgateway/services/tool_service.py:4142-4160: `_registry.get_hook_refs_for_hook(
    hook_type=...)`, then `hook_ref.plugin_ref.mode / .conditions / .name / .plugin.config.config`; mcpgateway/services/plugin_service.py:84-130:
    `_registry.get_all_plugins()`, `_config.plugins`, `plugin_ref.name / .mode / .priority / .hooks / .tags / .plugin.config.{description,author,
    version,kind,namespace,config}`) — the same ones BatchedPluginManager's replay is built on."""
    m = _manager(MessageFilter, blocked_words=["spam"])
    refs = m._registry.get_hook_refs_for_hook(hook_type=fw.AgentHookType.AGENT_PRE_INVOKE)
    assert len(refs) == 1
    pr = refs[0].plugin_ref
    assert (pr.name, pr.mode, pr.priority, list(pr.conditions), pr.plugin.config.config) == ("P", fw.PluginMode.SEQUENTIAL, 50, [], {"blocked_words": ["spam"]})
    (only,) = m._registry.get_all_plugins()
    assert only is pr and pr.hooks == ["agent_pre_invoke", "agent_post_invoke"] and pr.tags == []
    c = pr.plugin.config
    assert (c.description, c.author, c.version, c.namespace, c.kind) == (None, None, None, None, f"{__name__}.MessageFilter")
    assert [p.name for p in m._config.plugins] == ["P"] and m.plugin_count == 1 and m.has_hooks_for(fw.AgentHookType.AGENT_POST_INVOKE)
    assert not m.has_hooks_for(fw.ToolHookType.TOOL_POST_INVOKE)
