# -*- coding: utf-8 -*-
"""GPU drop-in for /root/reference/plugins/sql_sanitizer/sql_sanitizer.py.

Same class name, config schema and hooks (`prompt_pre_fetch`, `tool_pre_invoke`).  What the reference does per string
(`_find_issues` :131-163, `_strip_sql_comments` :102-114) is two `re.sub` passes and up to 5 + 3 `search` calls on the
stripped text.  Here all strings of a hook call (of all concurrent hook calls: `GpuBatcher`) are ONE fused scan:

  * pass 1 over the original strings: every search pattern (blocked statements, DELETE..FROM, UPDATE x, WHERE), the four
    interpolation literals of `_has_interpolation` (:117-128, evaluated on the ORIGINAL text there too) and the two comment
    rules (`--.*?$` MULTILINE, `/\\*.*?\\*/` DOTALL) as ordered substitution rules with an empty replacement;
  * only strings that contain a comment are rewritten (substitution kernel, rule after rule like :112-113) and scanned a
    second time — for every other string stripped == original, so pass 1's bits are the reference's answer.

Issue strings, their order, `scanned` (incl. the quirk that nested keys land at the top level of the new args, :186-187) and
the result shapes are rebuilt on the host exactly as the reference builds them.  No CPU fallback: patterns the engine cannot
express raise `UnsupportedPattern` at construction.
"""
from __future__ import annotations

import re
from typing import Any, List, Optional, Pattern, Tuple

from pydantic import BaseModel, ConfigDict, field_validator

from .. import engine
from ..batching import GpuBatcher
from ..framework import (Plugin, PluginConfig, PluginContext, PluginViolation, PromptPrehookPayload, PromptPrehookResult, ToolPreInvokePayload,
                         ToolPreInvokeResult)

_DEFAULT_BLOCKED = [r"\bDROP\b", r"\bTRUNCATE\b", r"\bALTER\b", r"\bGRANT\b", r"\bREVOKE\b"]          # reference :33-39
_LINE_COMMENT = (r"--.*?$", re.MULTILINE)                                                              # :40
_BLOCK_COMMENT = (r"/\*.*?\*/", re.DOTALL)                                                             # :41
_DELETE_FROM = (r"\bDELETE\b\s+\bFROM\b", re.IGNORECASE)                                               # :42
_UPDATE = (r"\bUPDATE\b\s+\w+", re.IGNORECASE)                                                         # :43
_WHERE = (r"\bWHERE\b", re.IGNORECASE)                                                                 # :44


class SQLSanitizerConfig(BaseModel):
    """Reference :47-99 (same fields, same validator behaviour)."""

    fields: Optional[list[str]] = None
    blocked_statements: list[Pattern[str]] = [re.compile(pat, re.IGNORECASE) for pat in _DEFAULT_BLOCKED]
    block_delete_without_where: bool = True
    block_update_without_where: bool = True
    strip_comments: bool = True
    require_parameterization: bool = False
    block_on_violation: bool = True

    @field_validator("blocked_statements", mode="before")
    @classmethod
    def compile_patterns(cls, v: Any) -> Any:
        if not isinstance(v, list):
            return v
        return [re.compile(item, re.IGNORECASE) if isinstance(item, str) else item for item in v]

    model_config = ConfigDict(arbitrary_types_allowed=True)


class _Slot:
    """One string the reference would hand to `_find_issues`: `label` is the issue prefix (`key` or `key[]`), `strip_key`
    the key `scanned` gets when comment stripping changes the text (None for list items, :194-198)."""

    __slots__ = ("label", "strip_key", "text")

    def __init__(self, label: str, strip_key: Optional[str], text: str):
        self.label, self.strip_key, self.text = label, strip_key, text


class SQLSanitizerPlugin(Plugin):
    def __init__(self, config: PluginConfig) -> None:
        super().__init__(config)
        self._cfg = SQLSanitizerConfig(**(config.config or {}))
        cfg = self._cfg
        p = self._prog = engine.Program()
        self._blocked: List[Tuple[int, str]] = [(p.add_search(pat.pattern, pat.flags & ~re.UNICODE), pat.pattern) for pat in cfg.blocked_statements]
        self._b_delete = p.add_search(*_DELETE_FROM)
        self._b_update = p.add_search(*_UPDATE)
        self._b_where = p.add_search(*_WHERE)
        self._b_plus, self._b_pct, self._b_lb, self._b_rb = (p.add_literal(w) for w in ("+", "%.", "{", "}"))
        self._rule_mask = 0
        if cfg.strip_comments:
            for pat, fl in (_LINE_COMMENT, _BLOCK_COMMENT):
                self._rule_mask |= 1 << p.add_sub(pat, fl, "")
        p.compile_host()
        self._batcher: Optional[GpuBatcher] = None

    # ---- host walk: the strings `_scan_value` (:166-198) examines, in its order
    def _slots(self, args: Any) -> List[_Slot]:
        out: List[_Slot] = []
        fields = self._cfg.fields

        def visit(key: str, value: Any) -> None:
            if isinstance(value, str):
                if fields is None or key in fields:
                    out.append(_Slot(key, key, value))
            elif isinstance(value, dict):
                for k, v in value.items():
                    visit(k, v)
            elif isinstance(value, list):
                for item in value:
                    if isinstance(item, dict):
                        for k, v in item.items():
                            visit(k, v)
                    elif isinstance(item, str):
                        if fields is None or key in fields:
                            out.append(_Slot(f"{key}[]", None, item))

        if args:
            for k, v in args.items():
                visit(k, v)
        return out

    def _issues_of(self, bits: int, orig_bits: int) -> List[str]:
        """`_find_issues` (:131-163) from the verdict bits of the stripped text (`bits`) and of the original (`orig_bits`)."""
        cfg = self._cfg
        issues = [f"Blocked statement matched: {pattern}" for bit, pattern in self._blocked if bits >> bit & 1]
        where = bits >> self._b_where & 1
        if cfg.block_delete_without_where and bits >> self._b_delete & 1 and not where:
            issues.append("DELETE without WHERE clause")
        if cfg.block_update_without_where and bits >> self._b_update & 1 and not where:
            issues.append("UPDATE without WHERE clause")
        if cfg.require_parameterization:
            o = orig_bits
            if o >> self._b_plus & 1 or o >> self._b_pct & 1 or (o >> self._b_lb & 1 and o >> self._b_rb & 1):
                issues.append("Possible non-parameterized interpolation detected")
        return issues

    async def _scan_args(self, args: Any) -> Tuple[List[str], dict]:
        """`_scan_args` (:201-222): (issues, scanned)."""
        slots = self._slots(args)
        if not slots:
            return [], {}
        if self._batcher is None:
            self._batcher = GpuBatcher.get()
        texts = [s.text for s in slots]
        first = await self._batcher.scan_sub(self._prog, texts, self._rule_mask)          # [(bits, rewritten bytes | None)]
        stripped: List[Optional[str]] = [None if new is None else new.decode("utf-8", "surrogatepass") for _, new in first]
        again = [i for i, t in enumerate(stripped) if t is not None]
        bits = [b for b, _ in first]
        if again:
            for i, b in zip(again, await self._batcher.scan(self._prog, [stripped[i] for i in again])):
                bits[i] = b
        issues: List[str] = []
        scanned: dict = {}
        for s, b, (orig, _), clean in zip(slots, bits, first, stripped):
            issues.extend(f"{s.label}: {m}" for m in self._issues_of(b, orig))
            if s.strip_key is not None and clean is not None and clean != s.text:
                scanned[s.strip_key] = clean
        return issues, scanned

    def _violation(self, issues: List[str], where: str) -> PluginViolation:
        return PluginViolation(reason="Risky SQL detected", description=f"Potentially dangerous SQL detected in {where}", code="SQL_SANITIZER", details={"issues": issues})

    async def prompt_pre_fetch(self, payload: PromptPrehookPayload, context: PluginContext) -> PromptPrehookResult:
        issues, scanned = await self._scan_args(payload.args or {})
        if issues and self._cfg.block_on_violation:
            return PromptPrehookResult(continue_processing=False, violation=self._violation(issues, "prompt args"))
        if scanned:
            new_args = {**(payload.args or {}), **scanned}
            return PromptPrehookResult(modified_payload=PromptPrehookPayload(prompt_id=payload.prompt_id, args=new_args), metadata={"sql_sanitized": True})
        return PromptPrehookResult(metadata={"sql_issues": issues} if issues else {})

    async def tool_pre_invoke(self, payload: ToolPreInvokePayload, context: PluginContext) -> ToolPreInvokeResult:
        issues, scanned = await self._scan_args(payload.args or {})
        if issues and self._cfg.block_on_violation:
            return ToolPreInvokeResult(continue_processing=False, violation=self._violation(issues, "tool args"))
        if scanned:
            new_args = {**(payload.args or {}), **scanned}
            return ToolPreInvokeResult(modified_payload=ToolPreInvokePayload(name=payload.name, args=new_args), metadata={"sql_sanitized": True})
        return ToolPreInvokeResult(metadata={"sql_issues": issues} if issues else {})
