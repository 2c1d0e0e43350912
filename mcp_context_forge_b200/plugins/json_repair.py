# -*- coding: utf-8 -*-
"""GPU drop-in for /root/reference/plugins/json_repair/json_repair.py (hook `tool_post_invoke`; SURVEY §8 row f-2 names it as a consumer of
the shared JSON parse).

The reference asks `orjson.loads(text)` whether a string tool result is JSON (`_try_parse`, :36-50) — the whole cost of the hook on the
common path — and, when it is not, tries up to three candidates in order (`_repair`, :53-78): single quotes → double quotes (bracketed text
without any double quote), trailing commas before a closing bracket removed, a bare `key: value` text wrapped in braces; the first candidate
that parses replaces the result.  Here "does it parse" is the parse status of the JSON kernel (`CF_TOON_NOT_JSON` or not — the same
orjson-exact validation the toon_encoder stage runs, csrc/json_tp.h) over the texts of all concurrent hook calls in one launch; under
`BatchedPluginManager` it is read off the chain's single fused launch (a result shared with toon_encoder is parsed once).  Candidates of the
few texts that do not parse are built on the host (strip / replace / brace wrapping: string methods, as in the reference), the trailing-comma
rule is one substitution rule of the engine (`sub_kernel`), and the candidates are validated together in one more launch.
No CPU fallback: without the CUDA library the hook fails like every other launch.  A text beyond the kernel's limits (nesting > 64,
numbers > 3200 bits: `CF_TOON_UNSUPPORTED`) is left as it is, with a warning — orjson accepts such texts up to depth 1024, and a broken one
is the only case this differs in.
"""
from __future__ import annotations

import logging
from typing import Any, List, Optional

from .. import engine
from ..batching import NOOP_RESULT, RUN_HOOK, GpuBatcher
from ..framework import Plugin, PluginConfig, PluginContext, ToolPostInvokePayload, ToolPostInvokeResult

logger = logging.getLogger(__name__)

_NOT_JSON, _UNSUPPORTED, _SKIPPED = 2, 6, 8            # CF_TOON_* (include/cfgpu.h)
_TRAILING_COMMA = r",(\s*[}\]])"                       # reference :33, replacement r"\1"


def _parsed(status: int) -> Optional[bool]:
    """True / False = orjson.loads succeeds / raises; None = the kernel cannot tell (beyond its limits)."""
    if status == _NOT_JSON:
        return False
    return None if status == _UNSUPPORTED else True


class JSONRepairPlugin(Plugin):
    def __init__(self, config: PluginConfig) -> None:
        super().__init__(config)
        self._prog = engine.Program()
        self._rule_mask = 1 << self._prog.add_sub(_TRAILING_COMMA, 0, [1])
        self._prog.compile_host()
        self._batcher: Optional[GpuBatcher] = None
        self.unsupported_kept = 0

    def _gpu(self) -> GpuBatcher:
        if self._batcher is None:
            self._batcher = GpuBatcher.get()
        return self._batcher

    async def _parse(self, texts: List[str]) -> List[Optional[bool]]:
        return [_parsed(st) for st, _ in await self._gpu().toon(texts, False)] if texts else []

    async def _repair(self, s: str) -> Optional[str]:
        """Reference :53-78 with the two `_try_parse` rounds as launches (candidate 1; then candidates 2 and 3 together)."""
        t = s.strip()
        base = t
        if len(t) >= 2 and t[0] in "[{" and t[-1] in "]}" and "'" in t and '"' not in t:          # _JSON_BRACKETS_RE.match(t) on a stripped text
            base = t.replace("'", '"')
            ok = (await self._parse([base]))[0]
            if ok:
                return base
            if ok is None:
                return self._keep(s)
        cands: List[str] = []
        out = (await self._gpu().sub(self._prog, [base], self._rule_mask))[0]
        if out is not None:
            cand = out.decode("utf-8", "surrogatepass")
            if cand != base:
                cands.append(cand)
        if not t.startswith("{") and ":" in t and "{" not in t and "}" not in t:
            cands.append("{" + t + "}")
        for cand, ok in zip(cands, await self._parse(cands)):
            if ok:
                return cand
            if ok is None:
                return self._keep(s)
        return None

    def _keep(self, s: str) -> None:
        self.unsupported_kept += 1
        logger.warning("json_repair: a text of %d characters is beyond the JSON kernel's limits (nesting > 64 or a number > 3200 bits); left as it is", len(s))
        return None

    async def tool_post_invoke(self, payload: ToolPostInvokePayload, context: PluginContext) -> ToolPostInvokeResult:
        text = payload.result
        if isinstance(text, str):
            ok = (await self._parse([text]))[0]
            if ok is None:
                self._keep(text)
            elif not ok:
                repaired = await self._repair(text)
                if repaired is not None:
                    return ToolPostInvokeResult(modified_payload=ToolPostInvokePayload(name=payload.name, result=repaired), metadata={"repaired": True})
        return ToolPostInvokeResult(continue_processing=True)

    # ---- chain protocol (mcp_context_forge_b200.manager.BatchedPluginManager): the parse status comes with the chain's fused launch
    CHAIN_HOOKS = ("tool_post_invoke",)

    def chain_register(self, prog: engine.Program) -> bool:
        return True                                     # no pattern of its own in the shared program: the repair rule lives in this plugin's program

    def chain_stage(self) -> int:
        return 8      # CF_STAGE_TOON: the JSON stage (its parse status is what this plugin reads)

    def chain_units(self, hook: str, payload: ToolPostInvokePayload) -> Optional[List[str]]:
        return [payload.result] if isinstance(payload.result, str) else []

    def chain_finish(self, hook: str, payload: ToolPostInvokePayload, units: List[str], results: List[Any]) -> Any:
        if not units:
            return NOOP_RESULT
        st = results[0].toon_status
        if st in (_NOT_JSON, _SKIPPED):                 # does not parse (or another plugin's rewrite skipped the JSON stage): the hook itself, on its own launches
            return RUN_HOOK
        if st == _UNSUPPORTED:
            self._keep(units[0])
        return NOOP_RESULT
