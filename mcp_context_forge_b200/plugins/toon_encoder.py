# -*- coding: utf-8 -*-
"""GPU drop-in for /root/reference/plugins/toon_encoder/toon_encoder.py.

Same class name, config keys, hook, result/metadata/stats shapes.  The per-item work of
`_process_content_item` — orjson.loads, toon.encode, "keep only if strictly smaller"
(reference :277-303) — runs on the GPU for all eligible items of all in-flight requests in one launch
(toon_kernel, csrc/json_toon.h).  Tool filters, size gates, annotation handling and statistics are
host logic identical to the reference (:221-235, :260-275, :305-326, :328-363).
"""
from __future__ import annotations

import json
import logging
import re
import time
from typing import Any, Dict, List, Optional

from .. import engine
from ..batching import NOOP_RESULT, GpuBatcher
from ..cpex_compat.framework import fast_construct
from ..framework import Plugin, PluginConfig, PluginContext, ToolPostInvokePayload, ToolPostInvokeResult

logger = logging.getLogger(__name__)


_KEY_RAW = re.compile(r"^[A-Za-z_][A-Za-z0-9_.]*$")             # toon.py `_VALID_KEY_RE`: such a key is written as it is, any other is quoted
_RESERVED = ("null", "true", "false")


def _first_error(v: Any) -> Optional[BaseException]:
    """The exception `toon.encode(v)` raises first, found by walking `v` in the encoder's own order (toon.py:82-565): the `ValueError` of the
    first quoted string that holds a control character other than \\n \\r \\t (:264-279; keys of a columnar table are never quoted, its
    values are visited row by row in the first row's key order), or the `AttributeError` of the unchecked `.keys()` when the FIRST field of
    an object inside a list is a non-empty list that is not made of objects (:400-404 -> :479-497).  Only called on the error path of
    `skip_on_error: false`, to word the exception like the reference does — the kernel's status already says WHICH kind it is."""
    def string(s: str) -> None:
        for ch in s:
            if ord(ch) < 32 and ch not in "\n\r\t":
                raise ValueError(f"Cannot encode control character U+{ord(ch):04X} in TOON")

    def key(k: str) -> None:
        if not (k and _KEY_RAW.match(k) and k not in _RESERVED):
            string(k)

    def simple(x: Any) -> bool:
        return x is None or isinstance(x, (bool, int, float, str))

    def columnar(arr: list) -> bool:                            # _try_columnar_encoding up to its decision; visits the rows when it says yes
        first = list(arr[0].keys())                             # AttributeError of a non-dict, worded by Python itself
        if not first:
            return False
        for obj in arr[1:]:
            if set(obj.keys()) != set(first):
                return False
        for obj in arr:
            if not all(simple(x) for x in obj.values()):
                return False
        for obj in arr:
            for k in first:
                value(obj[k])
        return True

    def array(arr: list) -> None:
        if not arr:
            return
        if all(isinstance(x, dict) for x in arr) and columnar(arr):
            return
        for x in arr:                                           # primitives in order; an object is a list item; a list is an array of its own
            if isinstance(x, dict):
                list_item(x)
            else:
                value(x)

    def list_item(obj: dict) -> None:
        for i, (k, x) in enumerate(obj.items()):
            key(k)
            if isinstance(x, list) and x:
                if i == 0 and columnar(x):
                    continue
                array(x)
            else:
                value(x)

    def value(x: Any) -> None:
        if isinstance(x, str):
            string(x)
        elif isinstance(x, list):
            array(x)
        elif isinstance(x, dict):
            for k, y in x.items():
                key(k)
                value(y)

    try:
        value(v)
    except (ValueError, AttributeError) as exc:
        return exc
    return None


def _encode_error(status: int, text: str) -> BaseException:
    """The exception of a unit the kernel reported as TOON_VALUE_ERROR / TOON_ATTR_ERROR, with the reference's wording."""
    kind = ValueError if status == engine.TOON_VALUE_ERROR else AttributeError
    try:
        exc = _first_error(json.loads(text))
    except (ValueError, RecursionError):
        exc = None
    if isinstance(exc, kind):
        return exc
    return kind("Cannot encode control character in TOON" if kind is ValueError else "object has no attribute 'keys'")


class ToonEncoderPlugin(Plugin):
    def __init__(self, config: PluginConfig) -> None:
        super().__init__(config)
        plugin_config = config.config or {}
        self._min_size_bytes: int = plugin_config.get("min_size_bytes", 100)
        self._max_size_bytes: int = plugin_config.get("max_size_bytes", 1024 * 1024)
        self._exclude_tools: List[str] = plugin_config.get("exclude_tools", [])
        self._include_tools: Optional[List[str]] = plugin_config.get("include_tools")
        self._add_format_marker: bool = plugin_config.get("add_format_marker", True)
        self._skip_on_error: bool = plugin_config.get("skip_on_error", True)
        self._tools_processed = 0
        self._tools_converted = 0
        self._items_attempted = 0
        self._items_converted = 0
        self._total_bytes_saved = 0
        self._items_unsupported = 0
        self._batcher: Optional[GpuBatcher] = None
        self._elig_memo: Dict[int, tuple] = {}      # id(text) -> (text, eligibility) for non-ASCII texts (see _eligible)
        self._elig_bytes = 0

    def _should_process_tool(self, tool_name: str) -> bool:
        if self._include_tools is not None:
            return tool_name in self._include_tools
        return tool_name not in self._exclude_tools

    def _eligible(self, item: Any) -> Optional[bytes]:
        """UTF-8 bytes of the item's text when the reference would attempt a conversion (:246-275)."""
        if not isinstance(item, dict):
            return None
        text = item.get("text", "")
        if item.get("type") != "text" or not isinstance(text, str):
            return None
        if text.isascii():                  # len(text.encode("utf-8")) without the copy (the common case)
            n = len(text)
            return text if self._min_size_bytes <= n <= self._max_size_bytes else None
        # non-ASCII: the size test needs the UTF-8 length.  The chain-level manager asks about the same `str` object up to three times
        # per request (speculate, identity check, finish): remember the last answers by object identity.
        memo = self._elig_memo
        hit = memo.get(id(text))
        if hit is not None and hit[0] is text:
            return hit[1]
        raw = text.encode("utf-8")          # raises UnicodeEncodeError on lone surrogates, like the reference's len(text.encode("utf-8"))
        res = None if len(raw) < self._min_size_bytes or len(raw) > self._max_size_bytes else raw
        self._elig_bytes += len(raw)
        if self._elig_bytes > (32 << 20):   # bounded: a wave's worth of texts, never more than 32 MB of them
            memo.clear()
            self._elig_bytes = len(raw)
        memo[id(text)] = (text, res)
        return res

    def _new_item(self, item: Dict[str, Any], toon_text: str) -> Dict[str, Any]:
        """reference :305-324."""
        new_item: Dict[str, Any] = {"type": "text", "text": toon_text}
        existing = item.get("annotations", {})
        if isinstance(existing, dict) and existing:
            new_item["annotations"] = {**existing}
        elif existing:
            new_item["annotations"] = existing
        if self._add_format_marker:
            if "annotations" not in new_item:
                new_item["annotations"] = {}
            if isinstance(new_item["annotations"], dict):
                new_item["annotations"]["format"] = "toon"
        return new_item

    # ---- chain protocol (mcp_context_forge_b200.manager.BatchedPluginManager)
    CHAIN_HOOKS = ("tool_post_invoke",)

    def chain_register(self, prog) -> bool:
        return True

    def chain_stage(self) -> int:
        return 8      # CF_STAGE_TOON

    def chain_toon_flags(self) -> int:
        return 0 if self._skip_on_error else 1      # CF_TOON_REPORT_ERRORS

    def _content(self, payload: ToolPostInvokePayload):
        if not self._should_process_tool(payload.name) or not isinstance(payload.result, dict):
            return None
        content = payload.result.get("content", [])
        return content if content and isinstance(content, list) else None

    def chain_units(self, hook: str, payload: ToolPostInvokePayload):
        content = self._content(payload)
        if content is None:
            return []
        return [item["text"] for item in content if self._eligible(item) is not None]

    def chain_finish(self, hook: str, payload: ToolPostInvokePayload, units, results) -> ToolPostInvokeResult:
        start_time = time.monotonic()
        content = self._content(payload)
        if content is None:
            return NOOP_RESULT
        self._tools_processed += 1
        raws = [self._eligible(item) for item in content]
        it = iter(results)
        outcomes = {i: (r.toon_status, r.toon_text) for i, raw in enumerate(raws) if raw is not None for r in (next(it),)}
        return self._finish(payload, content, raws, outcomes, start_time, NOOP_RESULT)

    async def tool_post_invoke(self, payload: ToolPostInvokePayload, _context: PluginContext) -> ToolPostInvokeResult:
        start_time = time.monotonic()
        tool_name = payload.name
        if not self._should_process_tool(tool_name):
            return ToolPostInvokeResult(continue_processing=True)
        result = payload.result
        if not isinstance(result, dict):
            return ToolPostInvokeResult(continue_processing=True)
        content = result.get("content", [])
        if not content or not isinstance(content, list):
            return ToolPostInvokeResult(continue_processing=True)

        self._tools_processed += 1
        raws = [self._eligible(item) for item in content]
        idx = [i for i, r in enumerate(raws) if r is not None]
        outcomes = {}
        if idx:
            if self._batcher is None:
                self._batcher = GpuBatcher.get()
            for i, oc in zip(idx, await self._batcher.toon([raws[i] for i in idx], report_errors=not self._skip_on_error)):
                outcomes[i] = oc
        return self._finish(payload, content, raws, outcomes, start_time)

    def _finish(self, payload: ToolPostInvokePayload, content, raws, outcomes, start_time, unchanged=None) -> ToolPostInvokeResult:
        """Everything after the per-item conversion (reference :285-326, :170-219): identical for the per-plugin and the chain path."""
        tool_name = payload.name
        result = payload.result

        new_content = []
        modified = False
        total_original = total_new = 0
        for i, item in enumerate(content):
            oc = outcomes.get(i)
            if oc is None:
                new_content.append(item)
                continue
            self._items_attempted += 1
            status, toon_bytes = oc
            if status == engine.TOON_NOT_JSON:     # counted as attempted, like the reference (:278-284)
                new_content.append(item)
            elif status == engine.TOON_CONVERTED:
                new_content.append(self._new_item(item, toon_bytes.decode("utf-8")))
                modified = True
                self._items_converted += 1
                total_original += len(raws[i])
                total_new += len(toon_bytes)
            elif status in (engine.TOON_VALUE_ERROR, engine.TOON_ATTR_ERROR):
                if not self._skip_on_error:
                    raise _encode_error(status, raws[i])
                logger.warning(f"ToonEncoder: Failed to encode '{tool_name}' to TOON")
                new_content.append(item)
            elif status == engine.TOON_NOT_SMALLER:
                new_content.append(item)
            else:  # TOON_UNSUPPORTED: outside the device limits (nesting > 64, number > 3200 bits) — never guessed: the item keeps its JSON,
                   # loudly (warning + counter), instead of turning a deeply nested tool result into a plugin error
                self._items_unsupported += 1
                logger.warning(f"ToonEncoder(GPU): an item of tool '{tool_name}' exceeds the device encoder's limits; left as JSON")
                new_content.append(item)

        if modified:
            self._tools_converted += 1
            bytes_saved = total_original - total_new
            self._total_bytes_saved += bytes_saved
            duration_ms = (time.monotonic() - start_time) * 1000
            savings_pct = (bytes_saved / total_original * 100) if total_original > 0 else 0
            new_result = {**result, "content": new_content}
            # (constructor-free: every value is ours and well-typed; same objects as ToolPostInvokePayload(name=, result=) / ...Result(...))
            return fast_construct(ToolPostInvokeResult, {
                "continue_processing": True,
                "modified_payload": fast_construct(ToolPostInvokePayload, {"name": tool_name, "result": new_result}),
                "violation": None,
                "metadata": {"toon_encoded": True, "bytes_saved": bytes_saved, "savings_percent": round(savings_pct, 2), "conversion_time_ms": round(duration_ms, 2)},
                "retry_delay_ms": 0})
        if unchanged is not None:           # (chain path: the manager's shared empty result)
            return unchanged
        return fast_construct(ToolPostInvokeResult, {"continue_processing": True, "modified_payload": None, "violation": None, "metadata": {}, "retry_delay_ms": 0})

    def get_stats(self) -> Dict[str, Any]:
        """reference :328-363."""
        return {
            "tools_processed": self._tools_processed,
            "tools_converted": self._tools_converted,
            "tool_conversion_rate": (self._tools_converted / self._tools_processed * 100 if self._tools_processed > 0 else 0.0),
            "items_attempted": self._items_attempted,
            "items_converted": self._items_converted,
            "item_conversion_rate": (self._items_converted / self._items_attempted * 100 if self._items_attempted > 0 else 0.0),
            "total_bytes_saved": self._total_bytes_saved,
        }
