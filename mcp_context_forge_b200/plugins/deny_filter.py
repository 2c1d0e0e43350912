# -*- coding: utf-8 -*-
"""GPU drop-in for /root/reference/plugins/deny_filter/deny.py (prompt_pre_fetch only, like the
reference).  `any(word in payload.args[key] ...)` (reference :59-60) becomes one fused substring
scan over all argument values; the first offending key in dict order wins (:58-68)."""
from __future__ import annotations

import logging
from typing import List

from pydantic import BaseModel

from .. import engine
from ..batching import GpuBatcher
from ..framework import Plugin, PluginConfig, PluginContext, PluginViolation, PromptPrehookPayload, PromptPrehookResult

logger = logging.getLogger(__name__)


class DenyListConfig(BaseModel):
    words: list[str]


class DenyListPlugin(Plugin):
    def __init__(self, config: PluginConfig):
        super().__init__(config)
        self._dconfig = DenyListConfig.model_validate(self._config.config)
        self._deny_list: List[str] = list(self._dconfig.words)
        self._prog = engine.Program() if self._deny_list else None
        for w in self._deny_list:
            self._prog.add_literal(w)        # "" matches every value, exactly like `"" in value`
        if self._prog is not None:
            self._prog.compile_host()
        self._batcher: GpuBatcher | None = None

    # ---- chain protocol (mcp_context_forge_b200.manager.BatchedPluginManager)
    CHAIN_HOOKS = ("prompt_pre_fetch",)

    def chain_register(self, prog: engine.Program) -> bool:
        self._chain_mask = 0
        for w in self._deny_list:
            self._chain_mask |= 1 << prog.add_literal(w)
        return True

    def chain_stage(self) -> int:
        return 1      # CF_STAGE_SCAN

    def chain_units(self, hook: str, payload):
        return [v for v in payload.args.values() if isinstance(v, str)] if payload.args else []

    def chain_finish(self, hook: str, payload, units, results) -> PromptPrehookResult:
        if payload.args and self._deny_list:
            it = iter(results)
            for key, value in payload.args.items():
                hit = (next(it).bitmap & self._chain_mask) != 0 if isinstance(value, str) else any(word in value for word in self._deny_list)
                if hit:
                    violation = PluginViolation(reason="Prompt not allowed", description="A deny word was found in the prompt", code="deny", details={})
                    logger.warning(f"Deny word detected in prompt argument '{key}'")
                    return PromptPrehookResult(modified_payload=payload, violation=violation, continue_processing=False)
        return PromptPrehookResult(modified_payload=payload)

    async def prompt_pre_fetch(self, payload: PromptPrehookPayload, context: PluginContext) -> PromptPrehookResult:
        if payload.args and self._prog is not None:
            keys = list(payload.args)
            str_keys = [k for k in keys if isinstance(payload.args[k], str)]
            if self._batcher is None:
                self._batcher = GpuBatcher.get()
            bitmaps = dict(zip(str_keys, await self._batcher.scan(self._prog, [payload.args[k] for k in str_keys]))) if str_keys else {}
            for key in keys:
                value = payload.args[key]
                # non-str values keep Python's `in` semantics (container membership / TypeError), reference :59
                hit = bitmaps[key] != 0 if isinstance(value, str) else any(word in value for word in self._deny_list)
                if hit:
                    violation = PluginViolation(reason="Prompt not allowed", description="A deny word was found in the prompt", code="deny", details={})
                    logger.warning(f"Deny word detected in prompt argument '{key}'")
                    return PromptPrehookResult(modified_payload=payload, violation=violation, continue_processing=False)
        return PromptPrehookResult(modified_payload=payload)

    async def shutdown(self) -> None:
        logger.info("Deny list plugin shutting down")
