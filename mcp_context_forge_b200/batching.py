"""Batch coalescer: concurrent in-flight hook calls -> one packed stream -> one GPU launch.

Every request runs on the gateway's asyncio loop and awaits its plugins one after another
(/root/reference/mcpgateway/services/tool_service.py:5866-5872), so many requests sit at `await`
points at the same time.  Plugins submit their units here; the coalescer collects whatever arrives
within `window_us` (0 = everything already queued in this loop iteration), packs it with
engine.pack_units and issues ONE cf_scan_host for the whole group, then hands every caller its
slice of the verdicts.  Per-request ordering is unaffected: a caller only resumes after its own
future resolves, exactly as if its plugin had computed inline.
"""
from __future__ import annotations

import asyncio
import threading
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import engine

Unit = Union[str, bytes]



class _NoopResult:
    """What a chain member returns from `chain_finish` when its hook would return the empty result (continue, no payload, no violation, no
    metadata): one shared immutable stand-in instead of a model per request; the replay skips it like the executor skips an empty result."""
    __slots__ = ()
    continue_processing = True
    modified_payload = None
    violation = None
    metadata: dict = {}
    retry_delay_ms = 0


NOOP_RESULT = _NoopResult()
RUN_HOOK = object()         # `chain_finish` -> "the fused launch does not settle this request: run my hook itself" (the manager awaits the plugin's own hook)


class GpuBatcher:
    _instances: Dict[int, "GpuBatcher"] = {}
    _ilock = threading.Lock()

    def __init__(self, ctx: Optional[engine.Context] = None, window_us: int = 0, max_units: int = 1 << 16):
        self.ctx = ctx or engine.Context.get()
        self.window_us = window_us
        self.max_units = max_units
        self._batch: Optional[engine.Batch] = None
        # per event loop (a process may run several, e.g. one per thread): (id(prog), op, arg) -> (prog, op, arg, [(units, future)]).
        # A wave is flushed on the loop that owns its futures; the launches themselves (which share the device batch) are serialised.
        self._pending: Dict[int, Dict[tuple, tuple]] = {}
        self._scheduled: Dict[int, bool] = {}
        self._launch_lock = threading.RLock()
        self.launches = 0
        self.units_seen = 0

    @classmethod
    def get(cls, device: int = 0) -> "GpuBatcher":
        with cls._ilock:
            b = cls._instances.get(device)
            if b is None:
                b = cls._instances[device] = GpuBatcher(engine.Context.get(device))
            return b

    # ---- device batch buffer, grown geometrically
    def _ensure_batch(self, nbytes: int, nunits: int) -> engine.Batch:
        b = self._batch
        if b is None or nbytes > b.max_bytes or nunits > b.max_units:
            cap_b = max(nbytes * 2, 1 << 20, b.max_bytes if b else 0)
            cap_u = max(nunits * 2, 1024, b.max_units if b else 0)
            self._batch = b = engine.Batch(self.ctx, cap_b, cap_u)
        return b

    # ---- synchronous core: one launch for a list of unit lists
    def scan_groups(self, prog: engine.Program, groups: Sequence[Sequence[Unit]]) -> List[List[int]]:
        flat: List[bytes] = [engine.encode_unit(u) for g in groups for u in g]
        if not flat:
            return [[] for _ in groups]
        if prog.h is None:
            prog.compile(self.ctx)
        stream, offs = engine.pack_units(flat)
        batch = self._ensure_batch(len(stream), len(flat))
        bm = engine.scan_host(prog, batch, stream, offs)
        self.launches += 1
        self.units_seen += len(flat)
        ints = engine.bitmaps_to_ints(bm, len(flat), prog.words)
        out, i = [], 0
        for g in groups:
            out.append(ints[i:i + len(g)])
            i += len(g)
        return out

    def scan_sync(self, prog: engine.Program, units: Sequence[Unit]) -> List[int]:
        return self.scan_groups(prog, [units])[0]

    def sub_groups(self, prog: engine.Program, groups: Sequence[Sequence[Unit]], rule_mask: int) -> List[List[Optional[bytes]]]:
        """Fused scan, then the substitution kernel on the units some rule matched.  Returns, per
        unit, the rewritten UTF-8 bytes or None when no rule touched it."""
        flat: List[bytes] = [engine.encode_unit(u) for g in groups for u in g]
        if not flat:
            return [[] for _ in groups]
        if prog.h is None:
            prog.compile(self.ctx)
        stream, offs = engine.pack_units(flat)
        batch = self._ensure_batch(len(stream), len(flat))
        bm = engine.bitmaps_to_ints(engine.scan_host(prog, batch, stream, offs), len(flat), prog.words)
        self.launches += 1
        self.units_seen += len(flat)
        dirty = [i for i, v in enumerate(bm) if v & rule_mask]
        res: List[Optional[bytes]] = [None] * len(flat)
        if dirty:
            for i, new in zip(dirty, engine.sub_host(prog, batch, dirty)):
                res[i] = new
            self.launches += 1
        out, i = [], 0
        for g in groups:
            out.append(res[i:i + len(g)])
            i += len(g)
        return out

    def scan_sub_groups(self, prog: engine.Program, groups: Sequence[Sequence[Unit]], rule_mask: int) -> List[List[tuple]]:
        """Like sub_groups, but every unit also gets its verdict bitmap: [(bits, rewritten bytes | None)] (sql_sanitizer)."""
        flat: List[bytes] = [engine.encode_unit(u) for g in groups for u in g]
        if not flat:
            return [[] for _ in groups]
        if prog.h is None:
            prog.compile(self.ctx)
        stream, offs = engine.pack_units(flat)
        batch = self._ensure_batch(len(stream), len(flat))
        bm = engine.bitmaps_to_ints(engine.scan_host(prog, batch, stream, offs), len(flat), prog.words)
        self.launches += 1
        self.units_seen += len(flat)
        dirty = [i for i, v in enumerate(bm) if v & rule_mask]
        res: List[tuple] = [(v, None) for v in bm]
        if dirty:
            for i, new in zip(dirty, engine.sub_host(prog, batch, dirty)):
                res[i] = (bm[i], new)
            self.launches += 1
        out, i = [], 0
        for g in groups:
            out.append(res[i:i + len(g)])
            i += len(g)
        return out

    def toon_groups(self, groups: Sequence[Sequence[Unit]], report_errors: bool = True) -> List[List[tuple]]:
        """JSON texts -> [(status, toon_bytes_or_None)] per text, one launch for all groups."""
        flat: List[bytes] = [engine.encode_unit(u) for g in groups for u in g]
        if not flat:
            return [[] for _ in groups]
        stream, offs = engine.pack_units(flat)
        batch = self._ensure_batch(len(stream), len(flat))
        status, texts = engine.toon_host(batch, stream, offs, report_errors)
        self.launches += 1
        self.units_seen += len(flat)
        res = [(int(s), t) for s, t in zip(status, texts)]
        out, i = [], 0
        for g in groups:
            out.append(res[i:i + len(g)])
            i += len(g)
        return out

    async def toon(self, texts: Sequence[Unit], report_errors: bool = True) -> List[tuple]:
        return await self._submit(None, "toon", 1 if report_errors else 0, texts)

    # ---- asyncio front: coalesce concurrent callers
    async def scan(self, prog: engine.Program, units: Sequence[Unit]) -> List[int]:
        return await self._submit(prog, "scan", 0, units)

    async def sub(self, prog: engine.Program, units: Sequence[Unit], rule_mask: int) -> List[Optional[bytes]]:
        return await self._submit(prog, "sub", rule_mask, units)

    async def scan_sub(self, prog: engine.Program, units: Sequence[Unit], rule_mask: int) -> List[tuple]:
        return await self._submit(prog, "scan_sub", rule_mask, units)

    async def _submit(self, prog: engine.Program, op: str, arg: int, units: Sequence[Unit]):
        if not units:
            return []
        loop = asyncio.get_running_loop()
        fut: asyncio.Future = loop.create_future()
        key = (id(prog) if prog is not None else 0, op, arg)
        lid = id(loop)
        mine = self._pending.setdefault(lid, {})
        if key not in mine:
            mine[key] = (prog, op, arg, [])
        mine[key][3].append((list(units), fut))
        if not self._scheduled.get(lid):
            self._scheduled[lid] = True
            if self.window_us > 0:
                loop.call_later(self.window_us / 1e6, self._flush, lid)
            else:
                loop.call_soon(self._flush, lid)
        return await fut

    def _dispatch(self, prog, op: str, arg: int, groups):
        with self._launch_lock:                      # the launches share the device batch: one at a time
            if op == "scan":
                return self.scan_groups(prog, groups)
            if op == "sub":
                return self.sub_groups(prog, groups, arg)
            if op == "scan_sub":
                return self.scan_sub_groups(prog, groups, arg)
            return self.toon_groups(groups, bool(arg))

    def _flush(self, lid: int) -> None:
        self._scheduled[lid] = False
        pending = self._pending.pop(lid, {})
        for prog, op, arg, waiters in pending.values():
            # a wave never outgrows max_units: split it between callers (one caller's units stay together)
            chunks, cur, cur_n = [], [], 0
            for w in waiters:
                if cur and cur_n + len(w[0]) > self.max_units:
                    chunks.append(cur)
                    cur, cur_n = [], 0
                cur.append(w)
                cur_n += len(w[0])
            if cur:
                chunks.append(cur)
            for chunk in chunks:
                try:
                    results = self._dispatch(prog, op, arg, [w[0] for w in chunk])
                except Exception as exc:  # surface the failure to every caller (no silent fallback)
                    for _, fut in chunk:
                        if not fut.done():
                            fut.set_exception(exc)
                    continue
                for (_, fut), res in zip(chunk, results):
                    if not fut.done():
                        fut.set_result(res)
