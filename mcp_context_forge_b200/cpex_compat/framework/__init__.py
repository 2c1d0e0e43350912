# -*- coding: utf-8 -*-
"""cpex-compatible plugin framework surface (models + sequential executor).

The reference's plugin framework lives in the third-party package `cpex` (pinned cpex 0.1.0,
/root/reference/uv.lock:906-907; upstream contextforge-org/contextforge-plugins-framework), which
is not vendored under /root/reference and is not installable here.  This module restates the part
of its API the hot path touches, from the reference's own call sites, docs and contract test:

  * symbols / behaviours            tests/acceptance/plugins/test_cpex_contract.py:24-79,102-147,158-196
  * Plugin / payload / result shape docs/docs/using/plugins/index.md:516-538,842-941
  * executor ordering & errors      docs/docs/architecture/plugins.md:985-1002,1040-1055;
                                    docs/docs/using/plugins/index.md:380-388,1126-1136
  * call sites                      mcpgateway/services/tool_service.py:4137-4160,4870-4876,5866-5872
  * hook payload policies           mcpgateway/plugins/policy.py:24-45

When the real `cpex` is importable the GPU plugins use it; this module is the stand-in otherwise and
the harness that lets the *unmodified* reference plugin files run for golden-vector generation
(tools/gen_golden.py).  Executor semantics are "parity unpinned" in-tree: only the
live-cluster integration test exercises them.
"""
from __future__ import annotations

import asyncio
import copy
import fnmatch
import importlib
import logging
import os
import uuid
from enum import Enum
from typing import Any, Dict, Generic, List, Optional, Protocol, TypeVar, Union, runtime_checkable

import yaml
from pydantic import BaseModel, ConfigDict, Field, RootModel, field_validator

logger = logging.getLogger("cpex_compat")

T = TypeVar("T")


# --------------------------------------------------------------------------------------------- enums
class PluginMode(str, Enum):
    SEQUENTIAL = "sequential"
    TRANSFORM = "transform"
    AUDIT = "audit"
    CONCURRENT = "concurrent"
    FIRE_AND_FORGET = "fire_and_forget"
    DISABLED = "disabled"
    # legacy gateway names (mcpgateway/plugins/gateway_plugin_manager.py:43-48 maps them)
    ENFORCE = "enforce"
    ENFORCE_IGNORE_ERROR = "enforce_ignore_error"
    PERMISSIVE = "permissive"


class OnError(str, Enum):
    FAIL = "fail"
    IGNORE = "ignore"
    DISABLE = "disable"


class ToolHookType(str, Enum):
    TOOL_PRE_INVOKE = "tool_pre_invoke"
    TOOL_POST_INVOKE = "tool_post_invoke"


class PromptHookType(str, Enum):
    PROMPT_PRE_FETCH = "prompt_pre_fetch"
    PROMPT_POST_FETCH = "prompt_post_fetch"


class ResourceHookType(str, Enum):
    RESOURCE_PRE_FETCH = "resource_pre_fetch"
    RESOURCE_POST_FETCH = "resource_post_fetch"


class AgentHookType(str, Enum):
    AGENT_PRE_INVOKE = "agent_pre_invoke"
    AGENT_POST_INVOKE = "agent_post_invoke"


class HttpHookType(str, Enum):
    HTTP_PRE_REQUEST = "http_pre_request"
    HTTP_POST_REQUEST = "http_post_request"
    HTTP_AUTH_RESOLVE_USER = "http_auth_resolve_user"
    HTTP_AUTH_CHECK_PERMISSION = "http_auth_check_permission"


class HookType(str, Enum):
    """Union of the hook names (older docstrings use `HookType.TOOL_POST_INVOKE`)."""

    TOOL_PRE_INVOKE = "tool_pre_invoke"
    TOOL_POST_INVOKE = "tool_post_invoke"
    PROMPT_PRE_FETCH = "prompt_pre_fetch"
    PROMPT_POST_FETCH = "prompt_post_fetch"
    RESOURCE_PRE_FETCH = "resource_pre_fetch"
    RESOURCE_POST_FETCH = "resource_post_fetch"
    AGENT_PRE_INVOKE = "agent_pre_invoke"
    AGENT_POST_INVOKE = "agent_post_invoke"


def _hook_name(hook: Any) -> str:
    return hook.value if isinstance(hook, Enum) else str(hook)


# --------------------------------------------------------------------------------------------- models
class PluginCondition(BaseModel):
    server_ids: Optional[set[str]] = None
    tenant_ids: Optional[set[str]] = None
    tools: Optional[set[str]] = None
    prompts: Optional[set[str]] = None
    resources: Optional[set[str]] = None
    agents: Optional[set[str]] = None
    user_patterns: Optional[list[str]] = None
    content_types: Optional[list[str]] = None


class MCPServerConfig(BaseModel):
    proto: Optional[str] = None
    url: Optional[str] = None
    script: Optional[str] = None


class PluginConfig(BaseModel):
    name: str
    kind: str
    description: Optional[str] = None
    author: Optional[str] = None
    namespace: Optional[str] = None
    version: Optional[str] = None
    hooks: list[str] = Field(default_factory=list)
    tags: list[str] = Field(default_factory=list)
    mode: PluginMode = PluginMode.SEQUENTIAL
    on_error: OnError = OnError.FAIL
    priority: int = 100
    conditions: list[PluginCondition] = Field(default_factory=list)
    applied_to: Optional[Any] = None
    config: Optional[dict[str, Any]] = None
    mcp: Optional[MCPServerConfig] = None

    model_config = ConfigDict(use_enum_values=False)


class PluginSettings(BaseModel):
    parallel_execution_within_band: bool = False
    plugin_timeout: int = 30
    fail_on_plugin_error: bool = False
    enable_plugin_api: bool = False
    plugin_health_check_interval: int = 60


class Config(BaseModel):
    plugins: Optional[list[PluginConfig]] = Field(default_factory=list)
    plugin_dirs: list[str] = Field(default_factory=list)
    plugin_settings: PluginSettings = Field(default_factory=PluginSettings)
    server_settings: Optional[Any] = None


class PluginViolation(BaseModel):
    reason: str
    description: str
    code: str
    details: Optional[dict[str, Any]] = Field(default_factory=dict)
    plugin_name: Optional[str] = None
    http_status_code: Optional[int] = None
    mcp_error_code: Optional[int] = None
    http_headers: Optional[dict[str, str]] = None


class PluginErrorModel(BaseModel):
    message: str
    code: Optional[str] = ""
    details: Optional[dict[str, Any]] = Field(default_factory=dict)
    plugin_name: str = ""
    mcp_error_code: Optional[int] = None


class PluginError(Exception):
    def __init__(self, error: PluginErrorModel):
        self.error = error
        super().__init__(error.message)


class PluginViolationError(Exception):
    def __init__(self, message: str, violation: Optional[PluginViolation] = None):
        self.message = message
        self.violation = violation
        super().__init__(message)


class PluginPayload(BaseModel):
    """Base of hook payloads; frozen (tests/acceptance/plugins/test_cpex_contract.py:117-120)."""

    model_config = ConfigDict(frozen=True)


class PluginResult(BaseModel, Generic[T]):
    continue_processing: bool = True
    modified_payload: Optional[T] = None
    violation: Optional[PluginViolation] = None
    metadata: Optional[dict[str, Any]] = Field(default_factory=dict)
    retry_delay_ms: int = 0


class HttpHeaderPayload(RootModel[dict[str, str]]):
    def __iter__(self):  # type: ignore[override]
        return iter(self.root)

    def __getitem__(self, item: str) -> str:
        return self.root[item]

    def __setitem__(self, key: str, value: str) -> None:
        self.root[key] = value

    def __len__(self) -> int:
        return len(self.root)


class ToolPreInvokePayload(PluginPayload):
    name: str
    args: Optional[dict[str, Any]] = Field(default_factory=dict)
    headers: Optional[HttpHeaderPayload] = None


class ToolPostInvokePayload(PluginPayload):
    name: str
    result: Any = None


class PromptPrehookPayload(PluginPayload):
    prompt_id: str
    args: Optional[dict[str, Any]] = Field(default_factory=dict)


class _PromptText(BaseModel):
    model_config = ConfigDict(extra="allow")
    type: str = "text"
    text: str


class _PromptMessage(BaseModel):
    model_config = ConfigDict(extra="allow")
    role: Any
    content: Union[_PromptText, Any]


class PromptResult(BaseModel):
    """A rendered prompt given as a plain mapping is read into a model (the reference's plugins address `result.messages[i].content.text` on
    the payload they built from a dict: plugins/test_prompt_output_sentinel.py:79-111, its test :61-81); the gateway's own
    `mcpgateway.common.models.PromptResult` instances pass through untouched."""
    model_config = ConfigDict(extra="allow")
    messages: List[_PromptMessage]
    description: Optional[str] = None


class PromptPosthookPayload(PluginPayload):
    prompt_id: str
    result: Any = None

    @field_validator("result", mode="before")
    @classmethod
    def _mapping_to_model(cls, v: Any) -> Any:
        if isinstance(v, dict) and isinstance(v.get("messages"), list):
            try:
                return PromptResult.model_validate(v)
            except Exception:  # noqa: BLE001 - not a prompt result after all: keep the mapping
                return v
        return v


class ResourcePreFetchPayload(PluginPayload):
    uri: str
    metadata: Optional[dict[str, Any]] = Field(default_factory=dict)


class ResourcePostFetchPayload(PluginPayload):
    uri: str
    content: Any = None


class AgentPreInvokePayload(PluginPayload):
    agent_id: str
    messages: list[Any] = Field(default_factory=list)
    tools: Optional[list[str]] = None
    headers: Optional[HttpHeaderPayload] = None
    model: Optional[str] = None
    system_prompt: Optional[str] = None
    parameters: Optional[dict[str, Any]] = Field(default_factory=dict)


class AgentPostInvokePayload(PluginPayload):
    agent_id: str
    messages: list[Any] = Field(default_factory=list)
    tool_calls: Optional[list[dict[str, Any]]] = None


class HttpPreRequestPayload(PluginPayload):
    path: str = ""
    method: str = ""
    client_host: Optional[str] = None
    client_port: Optional[int] = None
    headers: HttpHeaderPayload = Field(default_factory=lambda: HttpHeaderPayload({}))


class HttpPostRequestPayload(HttpPreRequestPayload):
    response_headers: Optional[HttpHeaderPayload] = None
    status_code: Optional[int] = None


class HttpAuthResolveUserPayload(PluginPayload):
    credentials: Optional[dict[str, Any]] = None
    headers: HttpHeaderPayload = Field(default_factory=lambda: HttpHeaderPayload({}))
    client_host: Optional[str] = None
    client_port: Optional[int] = None


class HttpAuthCheckPermissionPayload(PluginPayload):
    user_email: str = ""
    permission: str = ""
    resource_type: Optional[str] = None
    team_id: Optional[str] = None
    is_admin: bool = False
    auth_method: Optional[str] = None
    client_host: Optional[str] = None
    user_agent: Optional[str] = None


class HttpAuthCheckPermissionResultPayload(PluginPayload):
    granted: bool = False
    reason: Optional[str] = None


ToolPreInvokeResult = PluginResult[ToolPreInvokePayload]
ToolPostInvokeResult = PluginResult[ToolPostInvokePayload]
PromptPrehookResult = PluginResult[PromptPrehookPayload]
PromptPosthookResult = PluginResult[PromptPosthookPayload]
ResourcePreFetchResult = PluginResult[ResourcePreFetchPayload]
ResourcePostFetchResult = PluginResult[ResourcePostFetchPayload]
AgentPreInvokeResult = PluginResult[AgentPreInvokePayload]
AgentPostInvokeResult = PluginResult[AgentPostInvokePayload]
HttpPreRequestResult = PluginResult[HttpHeaderPayload]
HttpPostRequestResult = PluginResult[HttpHeaderPayload]
HttpAuthResolveUserResult = PluginResult[dict]
HttpAuthCheckPermissionResult = PluginResult[HttpAuthCheckPermissionResultPayload]


class UserContext(BaseModel):
    user_id: str
    teams: list[str] = Field(default_factory=list)
    roles: list[str] = Field(default_factory=list)
    is_admin: bool = False
    attributes: dict[str, Any] = Field(default_factory=dict)


class GlobalContext(BaseModel):
    request_id: str
    user: Optional[Union[str, dict, UserContext]] = None
    tenant_id: Optional[str] = None
    server_id: Optional[str] = None
    content_type: Optional[str] = None
    state: dict[str, Any] = Field(default_factory=dict)
    metadata: dict[str, Any] = Field(default_factory=dict)


class PluginContext(BaseModel):
    state: dict[str, Any] = Field(default_factory=dict)
    global_context: GlobalContext
    metadata: dict[str, Any] = Field(default_factory=dict)

    def get_state(self, key: str, default: Any = None) -> Any:
        return self.state.get(key, default)

    def set_state(self, key: str, value: Any) -> None:
        self.state[key] = value

    async def cleanup(self) -> None:
        self.state.clear()
        self.metadata.clear()


PluginContextTable = Dict[str, PluginContext]


class HookPayloadPolicy(BaseModel):
    writable_fields: frozenset[str] = frozenset()

    model_config = ConfigDict(frozen=True)


@runtime_checkable
class ObservabilityProvider(Protocol):
    def start_span(self, trace_id: str, name: str, kind: str = "internal", resource_type: Optional[str] = None,
                   resource_name: Optional[str] = None, attributes: Optional[dict] = None) -> Optional[str]: ...

    def end_span(self, span_id: Optional[str], status: str = "ok", attributes: Optional[dict] = None) -> None: ...


# --------------------------------------------------------------------------------------------- plugin base
class Plugin:
    """Base class of every plugin (docs/docs/architecture/plugins.md:555-595)."""

    def __init__(self, config: PluginConfig, hook_payloads: Optional[dict] = None, hook_results: Optional[dict] = None) -> None:
        self._config = config

    @property
    def config(self) -> PluginConfig:
        return self._config

    @property
    def name(self) -> str:
        return self._config.name

    @property
    def priority(self) -> int:
        return self._config.priority

    @property
    def mode(self) -> PluginMode:
        return self._config.mode

    @property
    def hooks(self) -> list[str]:
        return self._config.hooks

    @property
    def tags(self) -> list[str]:
        return self._config.tags

    @property
    def conditions(self) -> list[PluginCondition]:
        return self._config.conditions

    async def initialize(self) -> None:
        return None

    async def shutdown(self) -> None:
        return None


class PluginRef:
    def __init__(self, plugin: Plugin):
        self._plugin = plugin
        self._uuid = uuid.uuid4()
        self.disabled = False

    @property
    def plugin(self) -> Plugin:
        return self._plugin

    @property
    def uuid(self) -> str:
        return self._uuid.hex

    @property
    def name(self) -> str:
        return self._plugin.name

    @property
    def priority(self) -> int:
        return self._plugin.priority

    @property
    def mode(self) -> PluginMode:
        return PluginMode.DISABLED if self.disabled else self._plugin.mode

    @property
    def hooks(self) -> list[str]:
        return self._plugin.hooks

    @property
    def tags(self) -> list[str]:
        return self._plugin.tags

    @property
    def conditions(self) -> list[PluginCondition]:
        return self._plugin.conditions

    @property
    def manifest(self) -> Any:
        return None


class HookRef:
    def __init__(self, hook: str, plugin_ref: PluginRef):
        self.name = hook
        self.plugin_ref = plugin_ref
        self.hook = getattr(plugin_ref.plugin, hook)


class PluginInstanceRegistry:
    def __init__(self) -> None:
        self._plugins: dict[str, PluginRef] = {}
        self._hooks: dict[str, list[HookRef]] = {}

    def register(self, plugin: Plugin) -> None:
        if plugin.name in self._plugins:
            raise ValueError(f"Plugin {plugin.name} already registered")
        ref = PluginRef(plugin)
        self._plugins[plugin.name] = ref
        for hook in plugin.hooks:
            h = _hook_name(hook)
            if not callable(getattr(plugin, h, None)):
                continue
            self._hooks.setdefault(h, []).append(HookRef(h, ref))
            self._hooks[h].sort(key=lambda r: r.plugin_ref.priority)   # stable: ties keep config order

    def get_plugin(self, name: str) -> Optional[PluginRef]:
        return self._plugins.get(name)

    def get_all_plugins(self) -> list[PluginRef]:
        return list(self._plugins.values())

    def get_hook_refs_for_hook(self, hook_type: Any) -> list[HookRef]:
        return list(self._hooks.get(_hook_name(hook_type), []))

    @property
    def plugin_count(self) -> int:
        return len(self._plugins)

    async def shutdown(self) -> None:
        for ref in self._plugins.values():
            try:
                await ref.plugin.shutdown()
            except Exception as exc:  # pragma: no cover
                logger.error("error shutting down %s: %s", ref.name, exc)
        self._plugins.clear()
        self._hooks.clear()


class HookRegistry:
    def __init__(self) -> None:
        self._hooks: dict[str, tuple] = {}

    def register_hook(self, hook_type: str, payload_class: type, result_class: type) -> None:
        self._hooks[_hook_name(hook_type)] = (payload_class, result_class)

    def get_payload_type(self, hook_type: str):
        return self._hooks.get(_hook_name(hook_type), (None, None))[0]

    def get_result_type(self, hook_type: str):
        return self._hooks.get(_hook_name(hook_type), (None, None))[1]

    def is_registered(self, hook_type: str) -> bool:
        return _hook_name(hook_type) in self._hooks

    def get_registered_hooks(self) -> list[str]:
        return list(self._hooks)


_HOOK_REGISTRY = HookRegistry()
for _h, _p, _r in (
    ("tool_pre_invoke", ToolPreInvokePayload, ToolPreInvokeResult), ("tool_post_invoke", ToolPostInvokePayload, ToolPostInvokeResult),
    ("prompt_pre_fetch", PromptPrehookPayload, PromptPrehookResult), ("prompt_post_fetch", PromptPosthookPayload, PromptPosthookResult),
    ("resource_pre_fetch", ResourcePreFetchPayload, ResourcePreFetchResult), ("resource_post_fetch", ResourcePostFetchPayload, ResourcePostFetchResult),
    ("agent_pre_invoke", AgentPreInvokePayload, AgentPreInvokeResult), ("agent_post_invoke", AgentPostInvokePayload, AgentPostInvokeResult),
):
    _HOOK_REGISTRY.register_hook(_h, _p, _r)


def get_hook_registry() -> HookRegistry:
    return _HOOK_REGISTRY


def get_attr(obj: Any, attr: str, default: Any = "") -> Any:
    if isinstance(obj, dict):
        return obj.get(attr, default) or default
    return getattr(obj, attr, default) or default


# --------------------------------------------------------------------------------------------- loading
class ConfigLoader:
    @staticmethod
    def load_config(config: str, use_jinja: bool = True) -> Config:
        with open(os.path.normpath(config), "r", encoding="utf-8") as f:
            text = f.read()
        if use_jinja:
            try:  # the reference renders env defaults through jinja2 when present
                import jinja2  # type: ignore

                text = jinja2.Environment(loader=jinja2.BaseLoader(), autoescape=False).from_string(text).render(env=os.environ)
            except ImportError:
                pass
        data = yaml.safe_load(text) or {}
        if data.get("plugins") is None:
            data["plugins"] = []
        return Config(**data)


class PluginLoader:
    def __init__(self) -> None:
        self._types: dict[str, type] = {}

    def _import(self, kind: str) -> type:
        if kind not in self._types:
            mod, _, cls = kind.rpartition(".")
            self._types[kind] = getattr(importlib.import_module(mod), cls)
        return self._types[kind]

    async def load_and_instantiate_plugin(self, config: PluginConfig) -> Optional[Plugin]:
        cls = self._import(config.kind)
        plugin = cls(config)
        await plugin.initialize()
        return plugin

    async def shutdown(self) -> None:
        self._types.clear()


class ExternalPluginServer:  # pragma: no cover - surface only
    def __init__(self, *args: Any, **kwargs: Any) -> None:
        raise NotImplementedError("external MCP plugin servers are outside the hot path")


def payload_matches(payload: Any, hook_type: Any, conditions: list[PluginCondition], context: GlobalContext) -> bool:
    """cpex.framework.utils.payload_matches (call site mcpgateway/services/tool_service.py:4148)."""
    if not conditions:
        return True
    hook = _hook_name(hook_type)
    for cond in conditions:
        if cond.server_ids and context.server_id not in cond.server_ids:
            continue
        if cond.tenant_ids and context.tenant_id not in cond.tenant_ids:
            continue
        if cond.user_patterns:
            user = context.user if isinstance(context.user, str) else getattr(context.user, "user_id", None) or ""
            if not any(p in user or fnmatch.fnmatch(user, p) for p in cond.user_patterns):
                continue
        if cond.content_types and context.content_type not in cond.content_types:
            continue
        if hook.startswith("tool_") and cond.tools and getattr(payload, "name", None) not in cond.tools:
            continue
        if hook.startswith("prompt_") and cond.prompts and getattr(payload, "prompt_id", None) not in cond.prompts:
            continue
        if hook.startswith("resource_") and cond.resources and getattr(payload, "uri", None) not in cond.resources:
            continue
        if hook.startswith("agent_") and cond.agents and getattr(payload, "agent_id", None) not in cond.agents:
            continue
        return True
    return False


class CopyOnWriteDict(dict):
    """cpex.framework.memory.CopyOnWriteDict as far as this tree shows it (tests/unit/plugins/test_sql_sanitizer.py:39-52 builds tool
    arguments with it and the plugins walk them as a `dict`): a dict whose writes never reach the mapping it was made from — the items are
    taken at construction, `original` keeps the wrapped mapping, `modified` names the keys written or deleted since."""

    def __init__(self, original: Any = None, **kwargs: Any) -> None:
        super().__init__(original or {}, **kwargs)
        self.original = original if original is not None else {}
        self.modified: set = set()

    def __setitem__(self, key: Any, value: Any) -> None:
        self.modified.add(key)
        super().__setitem__(key, value)

    def __delitem__(self, key: Any) -> None:
        self.modified.add(key)
        super().__delitem__(key)


# --------------------------------------------------------------------------------------------- manager
_osa = object.__setattr__


_defaults_cache: dict = {}
_nfields: dict = {}
_policy_fields: dict = {}


def _missing_defaults(cls: type, given: dict) -> list:
    """(name, is_factory, default) of the model fields `given` does not name; cached per (class, given key set)."""
    key = (cls, frozenset(given))
    hit = _defaults_cache.get(key)
    if hit is None:
        hit = []
        for name, f in cls.model_fields.items():
            if name in given:
                continue
            if f.default_factory is not None:
                hit.append((name, True, f.default_factory))
            else:
                hit.append((name, False, None if f.is_required() else f.default))
        _defaults_cache[key] = hit
    return hit


def fast_construct(cls: type, values: dict) -> Any:
    """`cls.model_construct(**values)` for a model without private attributes: no validation (the values come from our own code),
    fields that are not given take their declared defaults (so the call stays correct when the installed cpex declares more fields
    than the ones named here), four attribute stores instead of pydantic's generic path (≈ 4x faster; the executor builds five
    such objects per request)."""
    n = _nfields.get(cls)
    if n is None:                                   # (`cls.model_fields` is a Python-level descriptor: 0.4 us per access, so it is asked once per class)
        n = _nfields[cls] = len(cls.model_fields)
    if len(values) != n:                            # (the callers name every field of the stand-in models: the common case skips this)
        for name, is_factory, d in _missing_defaults(cls, values):
            values[name] = d() if is_factory else d
    m = cls.__new__(cls)
    _osa(m, "__dict__", values)
    _osa(m, "__pydantic_fields_set__", set(values))
    _osa(m, "__pydantic_extra__", None)
    _osa(m, "__pydantic_private__", None)
    return m


def fast_copy(model: Any, updates: dict) -> Any:
    """`model.model_copy(update=updates)` (shallow) for a model without extras / private attributes."""
    if model.__pydantic_extra__ is not None or model.__pydantic_private__ is not None:
        return model.model_copy(update=updates)
    m = model.__class__.__new__(model.__class__)
    _osa(m, "__dict__", {**model.__dict__, **updates})
    _osa(m, "__pydantic_fields_set__", model.__pydantic_fields_set__ | updates.keys())
    _osa(m, "__pydantic_extra__", None)
    _osa(m, "__pydantic_private__", None)
    return m


_LEGACY_MODES = {
    PluginMode.ENFORCE: (PluginMode.SEQUENTIAL, OnError.FAIL),
    PluginMode.ENFORCE_IGNORE_ERROR: (PluginMode.SEQUENTIAL, OnError.IGNORE),
    PluginMode.PERMISSIVE: (PluginMode.TRANSFORM, None),
}


def _effective_mode(ref: PluginRef) -> tuple:
    mode = ref.mode
    on_error = ref.plugin.config.on_error
    if mode in _LEGACY_MODES:
        mode, oe = _LEGACY_MODES[mode]
        on_error = oe or on_error
    return mode, on_error


def _result_with(result: Any, updates: dict) -> Any:
    """A copy of a plugin's result with `updates` applied (a result that is not one of the pydantic result models is rebuilt as a PluginResult)."""
    try:
        return fast_copy(result, updates)
    except AttributeError:
        base = {"continue_processing": getattr(result, "continue_processing", True), "modified_payload": getattr(result, "modified_payload", None),
                "violation": getattr(result, "violation", None), "metadata": getattr(result, "metadata", None) or {}}
        base.update(updates)
        return PluginResult(**base)


_mlog = logging.getLogger("cpex.framework.manager")        # the logger name the reference's tests listen on (tests/integration/test_rate_limiter.py:672)


class PluginExecutor:
    """cpex.framework.manager.PluginExecutor: the loop behind `PluginManager.invoke_hook`.  Pinned by reference-held tests of the real
    cpex executor: tests/integration/test_rate_limiter.py:644-745 (`execute_plugin`: a TRANSFORM-mode violation is suppressed in the result
    and logged as "... raised violation ..." at WARNING, a SEQUENTIAL one raises with `violations_as_exceptions`; `execute`: DISABLED plugins
    are skipped) and tests/unit/mcpgateway/plugins/agent/test_agent_plugins.py (chaining, violations, contexts across hooks) — both run
    unmodified against this class by tools/run_reference_tests.py; tests/test_executor_reference_cases.py restates them."""

    def __init__(self, config: Optional[Config] = None, timeout: int = 30, observability: Optional[ObservabilityProvider] = None,
                 hook_policies: Optional[dict[str, HookPayloadPolicy]] = None) -> None:
        self.config = config
        self.timeout = timeout
        self.observability = observability
        self.hook_policies = hook_policies

    def apply_policy(self, hook: str, current: Any, modified: Any) -> Any:
        """Only policy-writable fields of `modified` are accepted (mcpgateway/plugins/policy.py:24-45)."""
        if modified is None or modified is current:
            return current
        policy = (self.hook_policies or {}).get(hook)
        if policy is None or not isinstance(current, BaseModel) or type(modified) is not type(current):
            return modified
        key = (frozenset(policy.writable_fields), type(current))
        fields = _policy_fields.get(key)
        if fields is None:                          # the policy's writable fields this payload type declares (asked once per policy and type)
            fields = _policy_fields[key] = tuple(f for f in policy.writable_fields if f in type(current).model_fields)
        updates = {}
        for f in fields:
            v = getattr(modified, f)
            if v is not getattr(current, f):
                updates[f] = v
        return fast_copy(current, updates) if updates else current

    async def run_one(self, ref: PluginRef, hook: str, payload: Any, ctx: PluginContext) -> PluginResult:
        fn = getattr(ref.plugin, hook)
        return await asyncio.wait_for(fn(payload, ctx), timeout=self.timeout)

    async def execute_plugin(self, hook_ref: HookRef, payload: Any, local_context: PluginContext, violations_as_exceptions: bool = False,
                             global_context: Optional[GlobalContext] = None, combined_metadata: Optional[dict] = None) -> PluginResult:
        """One plugin of a chain with its mode applied.  Returns what the chain may act on: a violation (or `continue_processing=False`)
        is only left in the result of a blocking (SEQUENTIAL) plugin; the payload of an AUDIT / FIRE_AND_FORGET plugin is dropped; a
        plugin that failed and may be skipped yields the empty result."""
        ref = hook_ref.plugin_ref
        hook = hook_ref.name
        mode, on_error = _effective_mode(ref)
        fail_all = bool(self.config and self.config.plugin_settings.fail_on_plugin_error)
        try:
            result = await self.run_one(ref, hook, payload, local_context)
        except (PluginViolationError, PluginError):
            raise
        except Exception as exc:  # timeout or plugin bug
            msg = f"Plugin {ref.name} exceeded {self.timeout}s timeout" if isinstance(exc, asyncio.TimeoutError) else str(exc)
            _mlog.error("Plugin %s failed in %s: %s", ref.name, hook, msg)
            if fail_all or (mode == PluginMode.SEQUENTIAL and on_error == OnError.FAIL):
                raise PluginError(error=PluginErrorModel(message=msg, plugin_name=ref.name)) from exc
            if on_error == OnError.DISABLE:
                ref.disabled = True
            return PluginResult(continue_processing=True)
        if result is None:
            return PluginResult(continue_processing=True)
        if combined_metadata is not None and result.metadata:
            combined_metadata.update(result.metadata)
        drop_payload = result.modified_payload is not None and mode in (PluginMode.AUDIT, PluginMode.FIRE_AND_FORGET)
        if not result.continue_processing or result.violation is not None:
            if result.violation is not None:
                result.violation.plugin_name = ref.name
            if mode == PluginMode.SEQUENTIAL:
                if violations_as_exceptions:
                    v = result.violation
                    raise PluginViolationError(f"{hook} blocked by plugin {ref.name}: {v.code} - {v.reason} ({v.description})" if v else f"{hook} blocked by plugin {ref.name}", violation=v)
                return _result_with(result, {"modified_payload": None}) if drop_payload else result
            v = result.violation
            _mlog.warning("Plugin %s (%s) raised violation in %s: %s; continuing", ref.name, mode.value, hook, f"{v.code} - {v.reason}" if v else "continue_processing=False")
            return _result_with(result, {"continue_processing": True, "violation": None, "modified_payload": None if drop_payload else result.modified_payload})
        return _result_with(result, {"modified_payload": None}) if drop_payload else result

    async def execute(self, hook_refs: list, payload: Any, global_context: GlobalContext, hook_type: Any, local_contexts: Optional[PluginContextTable] = None,
                      violations_as_exceptions: bool = False) -> tuple:
        """The chain: ascending priority (the registry's order), payload chained through `modified_payload`, `(PluginResult, contexts)` out."""
        hook = _hook_name(hook_type)
        contexts: PluginContextTable = {}
        current = payload
        changed = False
        metadata: dict[str, Any] = {}
        retry_delay_ms = 0
        for href in hook_refs:
            ref = href.plugin_ref
            if ref.mode == PluginMode.DISABLED:
                continue
            if ref.conditions and not payload_matches(current, hook, ref.conditions, global_context):
                continue
            key = global_context.request_id + ref.uuid
            ctx = (local_contexts or {}).get(key) or PluginContext(global_context=global_context)
            contexts[key] = ctx
            result = await self.execute_plugin(href, current, ctx, violations_as_exceptions, global_context, metadata)
            retry_delay_ms = max(retry_delay_ms, getattr(result, "retry_delay_ms", 0) or 0)
            if result.modified_payload is not None:
                new = self.apply_policy(hook, current, result.modified_payload)
                if new is not current:
                    current, changed = new, True
            if not result.continue_processing or result.violation is not None:
                return (PluginResult(continue_processing=False, modified_payload=current if changed else None, violation=result.violation, metadata=metadata,
                                     retry_delay_ms=retry_delay_ms), contexts)
        return (PluginResult(continue_processing=True, modified_payload=current if changed else None, violation=None, metadata=metadata, retry_delay_ms=retry_delay_ms), contexts)


class PluginManager:
    """Hook-chain executor front: configuration, plugin loading and registry; `invoke_hook` hands the hook's refs to a `PluginExecutor`
    (ascending priority, payload chained through `modified_payload`, writable fields enforced by the hook policy, violations / errors
    handled per mode — see the module docstring)."""

    def __init__(self, config: Union[str, Config] = "", timeout: int = 30, observability: Optional[ObservabilityProvider] = None,
                 hook_policies: Optional[dict[str, HookPayloadPolicy]] = None) -> None:
        # cpex's manager is a Borg; what this tree pins of it (tests/unit/mcpgateway/plugins/test_observability_adapter.py:258-276): a bare
        # `PluginManager()` is another reference to the state of the configured one, until `PluginManager.reset()`.  Managers that are
        # given a configuration (and subclasses: the gateway builds one TenantPluginManager per context) own their state.
        shared = PluginManager._shared
        if type(self) is PluginManager and not config and observability is None and hook_policies is None and shared is not None:
            self.__dict__ = shared
            return
        self._config: Optional[Config] = ConfigLoader.load_config(config) if isinstance(config, str) and config else (config if isinstance(config, Config) else Config())
        self._timeout = timeout
        self._observability = observability
        self._hook_policies = hook_policies
        self._registry = PluginInstanceRegistry()
        self._loader = PluginLoader()
        self._executor = PluginExecutor(self._config, timeout, observability, hook_policies)
        self._initialized = False
        if type(self) is PluginManager:
            PluginManager._shared = self.__dict__

    # -- lifecycle
    _shared: Optional[dict] = None

    @classmethod
    def reset(cls) -> None:
        PluginManager._shared = None

    @property
    def config(self) -> Optional[Config]:
        return self._config

    @property
    def plugin_count(self) -> int:
        return self._registry.plugin_count

    @property
    def initialized(self) -> bool:
        return self._initialized

    @property
    def observability(self) -> Optional[ObservabilityProvider]:
        return self._observability

    @observability.setter
    def observability(self, provider: Optional[ObservabilityProvider]) -> None:     # (tests/unit/mcpgateway/plugins/test_observability_adapter.py assigns it)
        self._observability = provider
        self._executor.observability = provider

    def get_plugin(self, name: str) -> Optional[PluginRef]:
        return self._registry.get_plugin(name)

    def has_hooks_for(self, hook_type: Any) -> bool:
        return any(r.plugin_ref.mode != PluginMode.DISABLED for r in self._registry.get_hook_refs_for_hook(hook_type))

    async def initialize(self) -> None:
        if self._initialized:
            return
        for pc in (self._config.plugins or []) if self._config else []:
            if pc.mode == PluginMode.DISABLED:
                continue
            try:
                plugin = await self._loader.load_and_instantiate_plugin(pc)
            except Exception as exc:
                raise RuntimeError(f"Unable to register and initialize plugin: {pc.name}: {exc}") from exc
            if plugin is not None:
                self._registry.register(plugin)
        self._initialized = True

    async def shutdown(self) -> None:
        await self._registry.shutdown()
        await self._loader.shutdown()
        self._initialized = False

    # -- execution
    def _apply_policy(self, hook: str, current: Any, modified: Any) -> Any:
        return self._executor.apply_policy(hook, current, modified)

    async def _run_one(self, ref: PluginRef, hook: str, payload: Any, ctx: PluginContext) -> PluginResult:
        return await self._executor.run_one(ref, hook, payload, ctx)

    async def invoke_hook(self, hook_type: Any, payload: Any, global_context: GlobalContext, local_contexts: Optional[PluginContextTable] = None,
                          violations_as_exceptions: bool = False) -> tuple:
        hook = _hook_name(hook_type)
        return await self._executor.execute(self._registry.get_hook_refs_for_hook(hook), payload, global_context, hook, local_contexts, violations_as_exceptions)

    async def invoke_hook_for_plugin(self, name: str, hook_type: Any, payload: Any, context: Union[GlobalContext, PluginContext, None] = None,
                                     violations_as_exceptions: bool = False, payload_as_json: bool = False) -> PluginResult:
        """Single-plugin invocation used by the reference's profiling harness
        (tests/performance/test_plugins_performance.py:187-190)."""
        hook = _hook_name(hook_type)
        ref = self._registry.get_plugin(name)
        if ref is None or not callable(getattr(ref.plugin, hook, None)):
            raise PluginError(error=PluginErrorModel(message=f"Unable to find {hook} for plugin {name}", plugin_name=name))
        if isinstance(context, PluginContext):
            ctx = context
        else:
            ctx = PluginContext(global_context=context or GlobalContext(request_id=uuid.uuid4().hex))
        result = await self._run_one(ref, hook, payload, ctx)
        if violations_as_exceptions and result.violation is not None:
            result.violation.plugin_name = name
            raise PluginViolationError(f"{hook} blocked by plugin {name}", violation=result.violation)
        return result


class TenantPluginManager(PluginManager):
    """Per-context manager built by mcpgateway/plugins/gateway_plugin_manager.py:167-173."""


__all__ = [n for n in dir() if not n.startswith("_")]
