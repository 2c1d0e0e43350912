# -*- coding: utf-8 -*-
"""GPU drop-in for /root/reference/plugins/regex_filter/search_replace.py.

Same class name, config schema (`words: [{search, replace}]`) and hooks.  `pattern.sub(replacement,
value)` applied rule after rule (reference :127-130, :147-155) runs on the GPU: one fused scan marks
the values some rule matches, the substitution kernel rewrites only those (leftmost-first,
non-overlapping, rules in order; rules that can match "" follow `re.sub`'s empty-match rules, replacement
templates may reference groups — `\\1`, `\\g<name>` — resolved by a capture pass on the GPU).  Invalid patterns are
skipped exactly as the reference does (:73-75); valid patterns the GPU engine cannot express (back-references,
look-around, a repeat of a nullable body) raise at construction (no CPU fallback).
"""
from __future__ import annotations

import copy
import re
import re._parser as _sre_parser  # type: ignore[import]
from typing import Any, Dict, List, Optional

from pydantic import BaseModel

from .. import engine
from ..batching import GpuBatcher
from ..framework import (Plugin, PluginConfig, PluginContext, PromptPosthookPayload, PromptPosthookResult, PromptPrehookPayload, PromptPrehookResult,
                         ToolPostInvokePayload, ToolPostInvokeResult, ToolPreInvokePayload, ToolPreInvokeResult)
from ..cpex_compat.framework import fast_construct, fast_copy
from ..regex_frontend import UnsupportedPattern, template_parts  # noqa: F401


class SearchReplace(BaseModel):
    search: str
    replace: str


class SearchReplaceConfig(BaseModel):
    words: list[SearchReplace]


def replacement_parts(template: str, pattern: "re.Pattern[str]"):
    """The `re.sub` replacement template of a rule, expanded by sre's own template parser: literal strings and group indices
    (group references are resolved on the GPU by a capture pass over the match, csrc/scan_core.h pike_captures)."""
    return template_parts(template, pattern)


class SearchReplacePlugin(Plugin):
    def __init__(self, config: PluginConfig):
        super().__init__(config)
        self._srconfig = SearchReplaceConfig.model_validate(self._config.config)
        self._prog = engine.Program()
        self._rule_mask = 0
        for word in self._srconfig.words:
            try:
                compiled = re.compile(word.search)
            except re.error:
                continue                                    # reference :73-75
            bit = self._prog.add_sub(word.search, 0, replacement_parts(word.replace, compiled))
            self._rule_mask |= 1 << bit
        if self._rule_mask:
            self._prog.compile_host()
        self._batcher: Optional[GpuBatcher] = None

    # ---- chain protocol (mcp_context_forge_b200.manager.BatchedPluginManager): the whole chain of a wave of requests in ONE launch
    CHAIN_HOOKS = ("prompt_pre_fetch", "tool_pre_invoke", "tool_post_invoke")

    def chain_register(self, prog: engine.Program) -> bool:
        """Add this plugin's rules to the chain's shared program (rule order = config order).  One regex_filter per chain:
        cf_sub applies every ordered rule of the program."""
        if getattr(prog, "_has_ordered_owner", None) not in (None, self):
            return False
        prog._has_ordered_owner = self
        self._chain_mask = 0
        for word in self._srconfig.words:
            try:
                compiled = re.compile(word.search)
            except re.error:
                continue
            self._chain_mask |= 1 << prog.add_sub(word.search, 0, replacement_parts(word.replace, compiled))
        return True

    def chain_stage(self) -> int:
        return 3      # CF_STAGE_SCAN | CF_STAGE_SUB

    def chain_units(self, hook: str, payload: Any) -> Optional[List[str]]:
        if hook == "tool_post_invoke":
            r = payload.result
            if r and isinstance(r, dict):
                return [v for v in r.values() if isinstance(v, str)]
            if r and isinstance(r, str):
                return [r]
            return []
        if hook in ("prompt_pre_fetch", "tool_pre_invoke"):
            return [v for v in payload.args.values() if isinstance(v, str)] if payload.args else []
        return None

    def chain_finish(self, hook: str, payload: Any, units: List[str], results: List[Any]) -> Any:
        new = [u if r.rewritten is None else r.rewritten for u, r in zip(units, results)]
        if hook == "tool_post_invoke":
            r = payload.result
            if r and isinstance(r, dict):
                it = iter(new)
                payload = fast_copy(payload, {"result": {k: (next(it) if isinstance(v, str) else v) for k, v in r.items()}})
            elif r and isinstance(r, str):
                payload = fast_copy(payload, {"result": new[0]})
            return fast_construct(ToolPostInvokeResult, {"continue_processing": True, "modified_payload": payload, "violation": None, "metadata": {}, "retry_delay_ms": 0})
        if payload.args:
            it = iter(new)
            payload = fast_copy(payload, {"args": {k: (next(it) if isinstance(v, str) else v) for k, v in payload.args.items()}})
        return (PromptPrehookResult if hook == "prompt_pre_fetch" else ToolPreInvokeResult)(modified_payload=payload)

    async def _apply(self, values: List[str]) -> List[str]:
        if not self._rule_mask or not values:
            return values
        if self._batcher is None:
            self._batcher = GpuBatcher.get()
        out = await self._batcher.sub(self._prog, values, self._rule_mask)
        return [v if o is None else o.decode("utf-8", "surrogatepass") for v, o in zip(values, out)]

    async def _apply_dict(self, d: Dict[str, Any]) -> Dict[str, Any]:
        modified = dict(d)
        keys = [k for k, v in modified.items() if isinstance(v, str)]
        for k, v in zip(keys, await self._apply([modified[k] for k in keys])):
            modified[k] = v
        return modified

    async def prompt_pre_fetch(self, payload: PromptPrehookPayload, context: PluginContext) -> PromptPrehookResult:
        if payload.args:
            payload = payload.model_copy(update={"args": await self._apply_dict(payload.args)})
        return PromptPrehookResult(modified_payload=payload)

    async def prompt_post_fetch(self, payload: PromptPosthookPayload, context: PluginContext) -> PromptPosthookResult:
        if payload.result.messages:
            modified_result = copy.deepcopy(payload.result)
            texts = await self._apply([m.content.text for m in modified_result.messages])
            for m, t in zip(modified_result.messages, texts):
                m.content.text = t
            payload = payload.model_copy(update={"result": modified_result})
        return PromptPosthookResult(modified_payload=payload)

    async def tool_pre_invoke(self, payload: ToolPreInvokePayload, context: PluginContext) -> ToolPreInvokeResult:
        if payload.args:
            payload = payload.model_copy(update={"args": await self._apply_dict(payload.args)})
        return ToolPreInvokeResult(modified_payload=payload)

    async def tool_post_invoke(self, payload: ToolPostInvokePayload, context: PluginContext) -> ToolPostInvokeResult:
        if payload.result and isinstance(payload.result, dict):
            payload = payload.model_copy(update={"result": await self._apply_dict(payload.result)})
        elif payload.result and isinstance(payload.result, str):
            payload = payload.model_copy(update={"result": (await self._apply([payload.result]))[0]})
        return ToolPostInvokeResult(modified_payload=payload)
