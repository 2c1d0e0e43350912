"""Stand-in for the third-party `cpex` package (see framework/__init__.py).

`install_as_cpex()` registers these modules under the `cpex.*` names in sys.modules when the real
package is not importable, so reference-style `from cpex.framework import ...` statements work.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types


def real_cpex_available() -> bool:
    mod = sys.modules.get("cpex")
    if mod is not None:
        return not getattr(mod, "__cpex_compat__", False)
    try:
        return importlib.util.find_spec("cpex") is not None
    except (ImportError, ValueError):
        return False


def install_as_cpex(force: bool = False) -> bool:
    """Returns True when the stand-in is (now) serving `cpex`."""
    if not force and real_cpex_available():
        return False
    if getattr(sys.modules.get("cpex"), "__cpex_compat__", False):
        return True
    from . import framework as fw

    def mod(name: str, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        m.__cpex_compat__ = True
        sys.modules[name] = m
        return m

    root = mod("cpex")
    root.__path__ = []  # mark as package
    sys.modules["cpex.framework"] = fw
    fw.__cpex_compat__ = True
    root.framework = fw
    models = mod("cpex.framework.models", **{k: getattr(fw, k) for k in fw.__all__})
    hooks = mod("cpex.framework.hooks")
    hooks.__path__ = []
    policies = mod("cpex.framework.hooks.policies", HookPayloadPolicy=fw.HookPayloadPolicy)
    tools = mod("cpex.framework.hooks.tools", ToolHookType=fw.ToolHookType, ToolPreInvokePayload=fw.ToolPreInvokePayload, ToolPostInvokePayload=fw.ToolPostInvokePayload,
                ToolPreInvokeResult=fw.ToolPreInvokeResult, ToolPostInvokeResult=fw.ToolPostInvokeResult)
    prompts = mod("cpex.framework.hooks.prompts", PromptHookType=fw.PromptHookType, PromptPrehookPayload=fw.PromptPrehookPayload, PromptPosthookPayload=fw.PromptPosthookPayload,
                  PromptPrehookResult=fw.PromptPrehookResult, PromptPosthookResult=fw.PromptPosthookResult)
    def names(prefix: str) -> dict:
        return {k: getattr(fw, k) for k in fw.__all__ if k.startswith(prefix)}

    resources = mod("cpex.framework.hooks.resources", **names("Resource"))
    agents = mod("cpex.framework.hooks.agents", **names("Agent"))
    http = mod("cpex.framework.hooks.http", **names("Http"))
    hooks.policies, hooks.tools, hooks.prompts, hooks.resources, hooks.agents, hooks.http = policies, tools, prompts, resources, agents, http
    # module paths the reference's tests import from (tests/integration/test_rate_limiter.py:44-46, tests/unit/plugins/test_sql_sanitizer.py:4)
    manager = mod("cpex.framework.manager", PluginManager=fw.PluginManager, PluginExecutor=fw.PluginExecutor, TenantPluginManager=fw.TenantPluginManager)
    base = mod("cpex.framework.base", Plugin=fw.Plugin, PluginRef=fw.PluginRef, HookRef=fw.HookRef)
    errors = mod("cpex.framework.errors", PluginError=fw.PluginError, PluginViolationError=fw.PluginViolationError)
    memory = mod("cpex.framework.memory", CopyOnWriteDict=fw.CopyOnWriteDict)
    fw.manager, fw.base, fw.errors, fw.memory = manager, base, errors, memory
    constants = mod("cpex.framework.constants", GATEWAY_METADATA="gateway_metadata", TOOL_METADATA="tool_metadata")
    utils = mod("cpex.framework.utils", payload_matches=fw.payload_matches, get_attr=fw.get_attr)
    obs = mod("cpex.framework.observability", current_trace_id=__import__("contextvars").ContextVar("current_trace_id", default=None), ObservabilityProvider=fw.ObservabilityProvider)

    class PluginsSettings:
        """PLUGINS_* environment settings (tests/acceptance/plugins/test_cpex_contract.py:253-266)."""

        def __init__(self) -> None:
            env = os.environ
            self.enabled = env.get("PLUGINS_ENABLED", "false").strip().lower() in ("1", "true", "yes", "on")
            self.plugin_timeout = int(env.get("PLUGINS_PLUGIN_TIMEOUT", "30"))
            self.config_file = env.get("PLUGINS_CONFIG_FILE", "plugins/config.yaml")

    settings_mod = mod("cpex.framework.settings", PluginsSettings=PluginsSettings, settings=PluginsSettings())
    tools_pkg = mod("cpex.tools")
    tools_pkg.__path__ = []
    cli = mod("cpex.tools.cli", main=lambda *a, **k: 0)
    tools_pkg.cli = cli
    root.tools = tools_pkg
    for name, m in (("models", models), ("hooks", hooks), ("constants", constants), ("utils", utils), ("observability", obs), ("settings", settings_mod)):
        setattr(fw, name, m)
    return True
