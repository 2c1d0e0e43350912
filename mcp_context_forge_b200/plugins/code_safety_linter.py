# -*- coding: utf-8 -*-
"""GPU drop-in for /root/reference/plugins/code_safety_linter/code_safety_linter.py.

Same class name, config schema (`blocked_patterns`: strings or compiled patterns, defaults :34-40) and hook
(`tool_post_invoke`, :87-123).  The reference runs `pat.search(text)` for every blocked pattern over the tool result's text;
here the patterns are one automaton and the texts of all concurrent hook calls one fused scan (`GpuBatcher`), or — under
`BatchedPluginManager` — part of the chain's single launch (chain protocol below).  Findings keep the configured order.
No CPU fallback: a pattern the engine cannot express raises `UnsupportedPattern` at construction.
"""
from __future__ import annotations

import re
from typing import Any, List, Optional, Pattern

from pydantic import BaseModel, ConfigDict, Field, field_validator

from .. import engine
from ..batching import GpuBatcher
from ..framework import Plugin, PluginConfig, PluginContext, PluginViolation, ToolPostInvokePayload, ToolPostInvokeResult

_DEFAULTS = [r"\beval\s*\(", r"\bexec\s*\(", r"\bos\.system\s*\(", r"\bsubprocess\.(Popen|call|run)\s*\(", r"\brm\s+-rf\b"]      # reference :34-40


class CodeSafetyConfig(BaseModel):
    """Reference :26-66."""

    blocked_patterns: List[Pattern[str]] = Field(default_factory=lambda: [re.compile(p) for p in _DEFAULTS])

    @field_validator("blocked_patterns", mode="before")
    @classmethod
    def compile_patterns(cls, v: Any) -> Any:
        if not isinstance(v, list):
            return v
        return [re.compile(item) if isinstance(item, str) else item for item in v]

    model_config = ConfigDict(arbitrary_types_allowed=True)


def _text_of(payload: ToolPostInvokePayload) -> Optional[str]:
    """The text the reference examines (:99-104); None or "" = nothing to scan."""
    r = payload.result
    if isinstance(r, str):
        return r
    if isinstance(r, dict) and isinstance(r.get("text"), str):
        return r.get("text")
    return None


class CodeSafetyLinterPlugin(Plugin):
    def __init__(self, config: PluginConfig) -> None:
        super().__init__(config)
        self._cfg = CodeSafetyConfig(**(config.config or {}))
        self._prog = engine.Program()
        self._bits = [(self._prog.add_search(p.pattern, p.flags), p.pattern) for p in self._cfg.blocked_patterns]
        if self._bits:
            self._prog.compile_host()
        self._batcher: Optional[GpuBatcher] = None

    def _result(self, bitmap: int, bits) -> ToolPostInvokeResult:
        findings = [pattern for bit, pattern in bits if bitmap >> bit & 1]
        if findings:
            return ToolPostInvokeResult(continue_processing=False, violation=PluginViolation(reason="Unsafe code pattern", description="Detected unsafe code constructs",
                                                                                              code="CODE_SAFETY", details={"patterns": findings}))
        return ToolPostInvokeResult(continue_processing=True)

    async def tool_post_invoke(self, payload: ToolPostInvokePayload, context: PluginContext) -> ToolPostInvokeResult:
        text = _text_of(payload)
        if not text or not self._bits:
            return ToolPostInvokeResult(continue_processing=True)
        if self._batcher is None:
            self._batcher = GpuBatcher.get()
        return self._result((await self._batcher.scan(self._prog, [text]))[0], self._bits)

    # ---- chain protocol (mcp_context_forge_b200.manager.BatchedPluginManager)
    CHAIN_HOOKS = ("tool_post_invoke",)

    def chain_register(self, prog: engine.Program) -> bool:
        self._chain_bits = [(prog.add_search(p.pattern, p.flags), p.pattern) for p in self._cfg.blocked_patterns]
        return True

    def chain_stage(self) -> int:
        return 1      # CF_STAGE_SCAN

    def chain_units(self, hook: str, payload: ToolPostInvokePayload):
        text = _text_of(payload)
        return [text] if text else []

    def chain_finish(self, hook: str, payload: ToolPostInvokePayload, units, results) -> ToolPostInvokeResult:
        if not units:
            return ToolPostInvokeResult(continue_processing=True)
        return self._result(results[0].bitmap, self._chain_bits)
