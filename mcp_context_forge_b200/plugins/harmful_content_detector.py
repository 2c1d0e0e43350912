# -*- coding: utf-8 -*-
"""GPU drop-in for /root/reference/plugins/harmful_content_detector/harmful_content_detector.py.

Same class name, config schema, hooks, violation/metadata shapes.  The 9 x `pat.search(s)` loop of
`_scan_text` (reference :92-107) is replaced by one fused GPU scan over every string of the payload
(bit i of a unit's bitmap == pattern i matched that string); findings are rebuilt on the host in
the reference's order: string order, then category order, then pattern order (:190-195).
"""
from __future__ import annotations

import re
from typing import Any, Dict, Iterable, List, Tuple

from pydantic import BaseModel

from .. import engine
from ..batching import NOOP_RESULT, GpuBatcher
from ..cpex_compat.framework import fast_construct
from ..framework import (Plugin, PluginConfig, PluginContext, PluginViolation, PromptPrehookPayload, PromptPrehookResult,
                         ToolPostInvokePayload, ToolPostInvokeResult)

# reference :36-52
DEFAULT_LEXICONS: Dict[str, List[str]] = {
    "self_harm": [r"\bkill myself\b", r"\bsuicide\b", r"\bself-harm\b", r"\bwant to die\b"],
    "violence": [r"\bkill (?:him|her|them|someone)\b", r"\bshoot (?:him|her|them|someone)\b", r"\bstab (?:him|her|them|someone)\b"],
    "hate": [r"\b(?:kill|eradicate) (?:[a-z]+) people\b", r"\b(?:racial slur|hate speech)\b"],
}


class HarmfulContentConfig(BaseModel):
    """reference :55-89 (string patterns are kept as source text and compiled with IGNORECASE for the GPU engine; a pre-compiled
    `re.Pattern` keeps ITS OWN flags, like the reference's `p if not isinstance(p, str)` :72-75)."""

    categories: Dict[str, List[Any]] = {}
    block_on: List[str] = ["self_harm", "violence", "hate"]
    redact: bool = False
    redaction_text: str = "[REDACTED]"

    def __init__(self, **data):
        if "categories" not in data:
            data["categories"] = {c: list(p) for c, p in DEFAULT_LEXICONS.items()}
        else:
            data["categories"] = {c: list(pats) for c, pats in data["categories"].items()}
        super().__init__(**data)

    model_config = {"arbitrary_types_allowed": True}


def _iter_strings(value: Any) -> Iterable[Tuple[str, str]]:
    """reference :110-139."""
    def walk(obj: Any, path: str):
        if isinstance(obj, str):
            yield path, obj
        elif isinstance(obj, dict):
            for k, v in obj.items():
                yield from walk(v, f"{path}.{k}" if path else str(k))
        elif isinstance(obj, list):
            for i, v in enumerate(obj):
                yield from walk(v, f"{path}[{i}]")
    yield from walk(value, "")


class HarmfulContentDetectorPlugin(Plugin):
    def __init__(self, config: PluginConfig) -> None:
        super().__init__(config)
        self._cfg = HarmfulContentConfig(**(config.config or {}))
        self._bits: List[Tuple[str, str]] = []          # bit index -> (category, pattern source)
        self._prog = engine.Program()
        self._flags: List[int] = []
        for cat, pats in self._cfg.categories.items():
            for p in pats:
                if isinstance(p, str):
                    re.compile(p, re.IGNORECASE)        # same validation (and re.error) as the reference :76,85
                    src, fl = p, int(re.IGNORECASE)
                else:
                    src, fl = p.pattern, int(p.flags)   # a pre-compiled pattern is used as it is (reference :72-75)
                self._prog.add_search(src, fl)          # raises UnsupportedPattern loudly; no CPU fallback
                self._bits.append((cat, src))
                self._flags.append(fl)
        self._prog.compile_host()
        self._batcher: GpuBatcher | None = None

    # ---- chain protocol (mcp_context_forge_b200.manager.BatchedPluginManager)
    CHAIN_HOOKS = ("prompt_pre_fetch", "tool_post_invoke")

    def chain_register(self, prog: engine.Program) -> bool:
        self._chain_bits = [prog.add_search(p, fl) for (_, p), fl in zip(self._bits, self._flags)]
        self._chain_mask = 0
        for b in self._chain_bits:
            self._chain_mask |= 1 << b
        return True

    def chain_stage(self) -> int:
        return 1      # CF_STAGE_SCAN

    def chain_units(self, hook: str, payload: Any):
        if hook == "prompt_pre_fetch":
            return [s for _, s in _iter_strings(payload.args or {})]
        text = payload.result
        if isinstance(text, (dict, list)):
            return [s for _, s in _iter_strings(text)]
        return [text] if isinstance(text, str) else []

    def chain_finish(self, hook: str, payload: Any, units: List[str], results: List[Any]):
        findings: List[Tuple[str, str]] = []
        mask = self._chain_mask
        for r in results:
            bm = r.bitmap & mask
            if bm:
                for i, sb in enumerate(self._chain_bits):
                    if (bm >> sb) & 1:
                        findings.append(self._bits[i])
        if not findings:
            return NOOP_RESULT
        return self._result(PromptPrehookResult if hook == "prompt_pre_fetch" else ToolPostInvokeResult, findings)

    def _gpu(self) -> GpuBatcher:
        if self._batcher is None:
            self._batcher = GpuBatcher.get()
        return self._batcher

    def _findings_from_bitmaps(self, bitmaps: List[int]) -> List[Tuple[str, str]]:
        findings: List[Tuple[str, str]] = []
        for bm in bitmaps:
            if bm:
                for i, (cat, pat) in enumerate(self._bits):
                    if (bm >> i) & 1:
                        findings.append((cat, pat))
        return findings

    async def _scan(self, strings: List[str]) -> List[Tuple[str, str]]:
        if not strings or not self._bits:
            return []
        return self._findings_from_bitmaps(await self._gpu().scan(self._prog, strings))

    def _result(self, cls, findings: List[Tuple[str, str]]):
        cats = sorted(set(c for c, _ in findings))
        if any(c in self._cfg.block_on for c in cats):
            return cls(continue_processing=False,
                       violation=PluginViolation(reason="Harmful content", description=f"Detected categories: {', '.join(cats)}", code="HARMFUL_CONTENT",
                                                 details={"categories": cats, "findings": findings[:5]}))
        if cats:
            return cls(metadata={"harmful_categories": cats})
        if not findings:         # the overwhelmingly common result: nothing found (same field values as `cls()`, without pydantic's generic constructor)
            return fast_construct(cls, {"continue_processing": True, "modified_payload": None, "violation": None, "metadata": {}, "retry_delay_ms": 0})
        return cls()

    async def prompt_pre_fetch(self, payload: PromptPrehookPayload, context: PluginContext) -> PromptPrehookResult:
        """reference :157-181."""
        strings = [s for _, s in _iter_strings(payload.args or {})]
        return self._result(PromptPrehookResult, await self._scan(strings))

    async def tool_post_invoke(self, payload: ToolPostInvokePayload, context: PluginContext) -> ToolPostInvokeResult:
        """reference :183-213."""
        text = payload.result
        if isinstance(text, (dict, list)):
            strings = [s for _, s in _iter_strings(text)]
        elif isinstance(text, str):
            strings = [text]
        else:
            strings = []
        return self._result(ToolPostInvokeResult, await self._scan(strings))
