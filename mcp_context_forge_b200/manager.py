# -*- coding: utf-8 -*-
"""Chain-level batched executor: a drop-in `PluginManager` that coalesces WHOLE hook chains across the
requests in flight on one gateway worker.

The reference awaits one plugin after another per request (/root/reference/mcpgateway/services/tool_service.py:5866-5872;
fixed overhead per plugin per hook ≈ 0.03 ms, PLUGIN_PROFILING.md:99-127).  Here every `invoke_hook` call of a hook whose
plugins speak the chain protocol (the four GPU plugins) parks on a future; once per event-loop turn (or coalescing window) the
manager

  1. asks every plugin of the chain which strings it will examine for every parked request — on the ORIGINAL payloads —
     and de-duplicates them by object identity (harmful's string walk and toon's content texts are the same `str` objects),
  2. packs them into ONE stream and issues ONE fused `cf_run_batch` (scan for all plugins' patterns from one shared
     program + regex_filter rewriting + TOON) on a worker thread — the event loop keeps serving,
  3. replays the chain per request exactly as `PluginManager.invoke_hook` would — ascending priority, conditions, modes,
     violations, `violations_as_exceptions`, payload policies — feeding each plugin its slice of the verdicts.  A plugin whose
     inputs an earlier plugin of the same request changed (regex_filter rewrote a value) does not get a speculated verdict:
     it runs its own hook on the current payload, like the reference.  So the result is the sequential chain's, bit for bit.
"""
from __future__ import annotations

import asyncio
import concurrent.futures
import gc
import logging
import threading
import time
from typing import Any, Dict, List, Optional

import numpy as np

from . import engine
from .batching import NOOP_RESULT, RUN_HOOK
from ._native import CF_STAGE_SCAN, CF_STAGE_SUB, CF_STAGE_TOON, CF_V_REWRITTEN, CF_V_TOON
from .framework import (GlobalContext, OnError, PluginContext, PluginError, PluginErrorModel, PluginManager, PluginMode, PluginResult,
                        PluginViolationError)
from .cpex_compat.framework import _effective_mode, _hook_name, _mlog, fast_construct, payload_matches  # same helpers the sequential executor uses

logger = logging.getLogger(__name__)


class _gc_paused:
    """Cyclic GC off for the span of a wave's speculate / replay loops.  A wave allocates ~25 small containers per request (contexts,
    result models, dicts), none of them cyclic; with the collector on, every 700 allocations start a young-generation pass and the
    surviving wave is promoted and re-traversed by the older generations — measured 23 % of the host time of a 2 048-request wave
    (`tools/profile_replay.py`).  Re-entrant (waves of different hooks interleave at their awaits); the collector's previous state is
    restored when the last pause ends, and it then sees the wave's objects once."""
    _depth = 0
    _was_enabled = False
    _lock = threading.Lock()        # (managers on several event-loop threads share the process-wide collector switch)

    def __enter__(self):
        cls = _gc_paused
        with cls._lock:
            if cls._depth == 0:
                cls._was_enabled = gc.isenabled()
                gc.disable()
            cls._depth += 1

    def __exit__(self, *exc):
        cls = _gc_paused
        with cls._lock:
            cls._depth -= 1
            if cls._depth == 0 and cls._was_enabled:
                gc.enable()
        return False


class UnitResult:
    __slots__ = ("bitmap", "rewritten", "toon_status", "toon_text")

    def __init__(self, bitmap: int, rewritten: Optional[str], toon_status: int, toon_text: Optional[bytes]):
        self.bitmap, self.rewritten, self.toon_status, self.toon_text = bitmap, rewritten, toon_status, toon_text


class _Chain:
    """The plugins of one hook, one shared program."""

    def __init__(self, hook: str, refs: list):
        self.hook = hook
        self.refs = refs
        self.prog = engine.Program()
        self.uuids = [h.plugin_ref.uuid for h in refs]   # PluginRef.uuid formats a UUID on every access
        self.modes = [_effective_mode(h.plugin_ref) for h in refs]   # (mode, on_error) as configured; `ref.disabled` is checked per request
        self.member = {}                      # ref.uuid -> plugin speaks the protocol for this hook
        self.toon_flags = 0
        n_pat = 0
        for href in refs:
            plug = href.plugin_ref.plugin
            ok = hook in getattr(plug, "CHAIN_HOOKS", ()) and plug.chain_register(self.prog)
            self.member[href.plugin_ref.uuid] = bool(ok)
            if ok and hasattr(plug, "chain_toon_flags"):
                self.toon_flags |= plug.chain_toon_flags()
        self.has_patterns = self.prog.n_patterns > 0
        # what the replay loop reads per plugin and request, resolved once: (ref, plugin, uuid, mode, on_error, its payloads are applied, conditions)
        self.steps = [(h.plugin_ref, h.plugin_ref.plugin, u, m, oe, m not in (PluginMode.AUDIT, PluginMode.FIRE_AND_FORGET), h.plugin_ref.conditions or None)
                      for h, u, (m, oe) in zip(refs, self.uuids, self.modes)]
        self.usable = any(self.member.values())


class BatchedPluginManager(PluginManager):
    """Same constructor / API as `PluginManager` plus `window_us` (how long a wave may collect requests; 0 = whatever is parked
    in this event-loop turn) and `max_wave` (requests per fused launch)."""

    def __init__(self, *args: Any, window_us: int = 0, max_wave: int = 8192, device: int = 0, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        # built on the executor's internals as the in-tree restatement has them (registry of hook refs, config, timeout).  Under an installed
        # cpex whose manager is shaped differently this must fail at construction, loudly — not at the first request.
        reg = getattr(self, "_registry", None)
        if not callable(getattr(reg, "get_hook_refs_for_hook", None)) or not hasattr(self, "_config") or not hasattr(self, "_timeout"):
            raise RuntimeError("BatchedPluginManager: the installed cpex PluginManager does not expose the registry / config internals the chain replay "
                               "reads; run the drop-in plugins under cpex's own manager (per-plugin coalescing, batching.GpuBatcher) instead")
        self.window_us = window_us
        self.max_wave = max_wave
        self._device = device
        self._chains: Dict[str, Optional[_Chain]] = {}
        self._pending: Dict[str, list] = {}
        self._scheduled: Dict[str, bool] = {}
        self._busy: Dict[str, bool] = {}
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=1, thread_name_prefix="cfgpu-chain")   # ONE launch at a time: the waves of every loop share the device batch
        self._chain_lock = threading.Lock()
        self._ctx: Optional[engine.Context] = None
        self._batch: Optional[engine.Batch] = None
        self.waves = 0
        self.launch_calls = 0
        self.units_seen = 0
        self.slow_path_calls = 0
        self.assemble_s = 0.0
        self.device_s = 0.0
        self.replay_s = 0.0

    # ---- chain plans
    def _chain_for(self, hook: str) -> Optional[_Chain]:
        if hook in self._chains:
            return self._chains[hook]
        with self._chain_lock:                          # (first use from two event-loop threads at once: one builds and compiles the chain)
            if hook in self._chains:
                return self._chains[hook]
            refs = [h for h in self._registry.get_hook_refs_for_hook(hook) if h.plugin_ref.mode != PluginMode.DISABLED]
            ch = _Chain(hook, refs) if refs else None
            if ch is not None and not ch.usable:
                ch = None
            if ch is not None:
                self._ctx = self._ctx or engine.Context.get(self._device)
                if ch.has_patterns:
                    ch.prog.compile(self._ctx)
            self._chains[hook] = ch
            return ch

    async def shutdown(self) -> None:
        await super().shutdown()
        self._chains.clear()

    # ---- the public entry point: park, coalesce, replay
    async def invoke_hook(self, hook_type: Any, payload: Any, global_context: GlobalContext, local_contexts: Optional[dict] = None,
                          violations_as_exceptions: bool = False) -> tuple:
        hook = _hook_name(hook_type)
        if self._chain_for(hook) is None:
            return await super().invoke_hook(hook_type, payload, global_context, local_contexts, violations_as_exceptions)
        loop = asyncio.get_running_loop()
        fut: asyncio.Future = loop.create_future()
        key = (id(loop), hook)          # waves are per event loop (a process may run one per thread): a future is only ever resolved on the loop that made it
        self._pending.setdefault(key, []).append((payload, global_context, local_contexts, violations_as_exceptions, fut))
        if not self._scheduled.get(key):
            self._scheduled[key] = True
            if self.window_us > 0:
                loop.call_later(self.window_us / 1e6, lambda: loop.create_task(self._flush(key)))
            else:
                loop.call_soon(lambda: loop.create_task(self._flush(key)))
        return await fut

    async def invoke_hook_batch(self, hook_type: Any, payloads: List[Any], global_contexts: List[GlobalContext], violations_as_exceptions: bool = False) -> list:
        """A caller that already holds many requests (a batch endpoint, a replay job): one wave, results in order; a request
        that raised carries its exception in place of the (PluginResult, contexts) tuple."""
        return await asyncio.gather(*[self.invoke_hook(hook_type, p, g, None, violations_as_exceptions) for p, g in zip(payloads, global_contexts)],
                                    return_exceptions=True)

    async def _flush(self, key: tuple) -> None:
        """One wave at a time per (event loop, hook): requests that arrive while a wave is on the GPU park and form the next wave the moment
        it returns (the batch size adapts to the load by itself; `window_us` only adds a floor)."""
        hook = key[1]
        self._scheduled[key] = False
        if self._busy.get(key):
            return
        self._busy[key] = True
        try:
            while self._pending.get(key):
                waiting = self._pending[key]
                wave, self._pending[key] = waiting[: self.max_wave], waiting[self.max_wave:]
                try:
                    await self._run_wave(hook, wave)
                except Exception as exc:  # noqa: BLE001 - a failed launch fails every parked request loudly (no CPU fallback)
                    for *_, fut in wave:
                        if not fut.done():
                            fut.set_exception(exc)
                except BaseException as exc:  # the flush task itself is being cancelled: the parked requests learn it, the cancellation goes on
                    for *_, fut in wave:
                        if not fut.done():
                            fut.set_exception(RuntimeError(f"BatchedPluginManager: wave aborted ({type(exc).__name__})"))
                    raise
        finally:
            self._busy.pop(key, None)
            if not self._pending.get(key):              # (loops come and go: no entry outlives its last request)
                self._pending.pop(key, None)
                if not self._scheduled.get(key):
                    self._scheduled.pop(key, None)

    # ---- one wave
    def _speculate(self, chain: _Chain, wave: list):
        """units (de-duplicated by identity), their stage bits, and per request / plugin the unit indices it reads."""
        units: List[str] = []
        stages: List[int] = []
        index: Dict[int, int] = {}
        plans = []                                        # per request: {ref.uuid: (unit strings, unit indices)}
        hook = chain.hook
        for payload, gctx, _lc, _vae, _fut in wave:
            plan = {}
            for href, uid in zip(chain.refs, chain.uuids):
                if not chain.member[uid]:
                    continue
                ref = href.plugin_ref
                if ref.conditions and not payload_matches(payload, hook, ref.conditions, gctx):
                    continue
                plug = ref.plugin
                us = plug.chain_units(hook, payload)
                if us is None:
                    continue
                st = plug.chain_stage()
                idx = []
                for u in us:
                    k = index.get(id(u))
                    if k is None:
                        k = index[id(u)] = len(units)
                        units.append(u)
                        stages.append(st)
                    else:
                        stages[k] |= st
                    idx.append(k)
                plan[uid] = (us, idx)
            plans.append(plan)
        return units, stages, plans

    def _launch(self, chain: _Chain, units: List[str], stages: List[int]):
        """Worker thread: pack, ONE cf_run_batch, unpack.  Returns per-unit UnitResult."""
        enc = [u.encode("utf-8", "surrogatepass") for u in units]
        stream, offs = engine.pack_units(enc)
        ctx = self._ctx
        b = self._batch
        if b is None or len(stream) > b.max_bytes or len(enc) > b.max_units:
            self._batch = b = engine.Batch(ctx, max(len(stream) * 2, 1 << 20, b.max_bytes if b else 0), max(len(enc) * 2, 1024, b.max_units if b else 0))
        stage_all = 0
        for s in stages:
            stage_all |= s
        if not chain.has_patterns:
            stage_all &= ~(CF_STAGE_SCAN | CF_STAGE_SUB)
        W = chain.prog.words if chain.has_patterns else 1
        verdicts, out, out_offs, full = engine.run_batch(chain.prog if chain.has_patterns else None, b, stream, offs, stage_all,
                                                         np.asarray(stages, dtype=np.uint8), chain.toon_flags, want_full_bitmaps=W > 1)
        flags = verdicts["flags"]
        res: List[UnitResult] = []
        bm0 = verdicts["match_bitmap"].tolist()
        aux = verdicts["aux"].tolist()
        hot = np.nonzero(flags)[0].tolist()                      # units with an output
        texts: Dict[int, bytes] = {}
        if hot:
            raw = out[: int(out_offs[-1])].tobytes()          # `out` is the batch's reusable buffer: copy what this wave produced
            for i in hot:
                texts[i] = raw[int(out_offs[i]):int(out_offs[i + 1])]
        fl = flags.tolist()
        for i in range(len(units)):
            bm = bm0[i]
            if W > 1:
                for w in range(1, W):
                    bm |= int(full[i * W + w]) << (64 * w)
            f = fl[i]
            rew = texts[i].decode("utf-8", "surrogatepass") if f & CF_V_REWRITTEN else None
            res.append(UnitResult(bm, rew, aux[i], texts.get(i) if f & CF_V_TOON else None))
        return res

    async def _run_wave(self, hook: str, wave: list) -> None:
        chain = self._chains[hook]
        t0 = time.perf_counter()
        with _gc_paused():
            units, stages, plans = self._speculate(chain, wave)
        t1 = time.perf_counter()
        results: List[UnitResult] = []
        if units:
            results = await asyncio.get_running_loop().run_in_executor(self._pool, self._launch, chain, units, stages)
            self.launch_calls += 1
        t2 = time.perf_counter()
        self.waves += 1
        self.units_seen += len(units)
        with _gc_paused():
            for (payload, gctx, lctx, vae, fut), plan in zip(wave, plans):
                if fut.done():
                    continue
                try:
                    fut.set_result(await self._replay(chain, payload, gctx, lctx, vae, plan, results))
                except BaseException as exc:  # noqa: BLE001 - PluginViolationError / PluginError of this request only
                    fut.set_exception(exc)
        t3 = time.perf_counter()
        self.assemble_s += t1 - t0
        self.device_s += t2 - t1
        self.replay_s += t3 - t2

    async def _replay(self, chain: _Chain, payload: Any, global_context: GlobalContext, local_contexts: Optional[dict], violations_as_exceptions: bool,
                      plan: dict, results: List[UnitResult]) -> tuple:
        """`PluginManager.invoke_hook` for one request, every chain member fed from the wave's verdicts."""
        hook = chain.hook
        contexts: dict = {}
        current = payload
        changed = False
        metadata: dict[str, Any] = {}
        retry_delay_ms = 0
        fail_all = bool(self._config and self._config.plugin_settings.fail_on_plugin_error)
        rid = global_context.request_id
        for ref, plugin, uid, mode, on_error, applies, conditions in chain.steps:
            if ref.disabled or mode == PluginMode.DISABLED:
                continue
            if conditions is not None and not payload_matches(current, hook, conditions, global_context):
                continue
            key = rid + uid
            ctx = local_contexts.get(key) if local_contexts else None
            if ctx is None:
                ctx = fast_construct(PluginContext, {"state": {}, "global_context": global_context, "metadata": {}})
            contexts[key] = ctx
            try:
                spec = plan.get(uid)
                result = None
                if spec is not None:
                    us, idx = spec
                    if current is not payload:                      # an earlier plugin replaced the payload: are my inputs untouched?
                        # (list equality: identity first, then value — a unit that is a different object with the SAME text has the same verdict)
                        if plugin.chain_units(hook, current) != us:
                            spec = None
                    if spec is not None:
                        result = plugin.chain_finish(hook, current, us, [results[k] for k in idx])
                        if result is RUN_HOOK:                      # (json_repair: the text does not parse — its repair rounds are the hook's own launches)
                            spec = None
                if spec is None:
                    self.slow_path_calls += 1
                    result = await self._run_one(ref, hook, current, ctx)
            except (PluginViolationError, PluginError):
                raise
            except Exception as exc:  # timeout or plugin bug
                msg = f"Plugin {ref.name} exceeded {self._timeout}s timeout" if isinstance(exc, asyncio.TimeoutError) else str(exc)
                logger.error("Plugin %s failed in %s: %s", ref.name, hook, msg)
                if fail_all or (mode == PluginMode.SEQUENTIAL and on_error == OnError.FAIL):
                    raise PluginError(error=PluginErrorModel(message=msg, plugin_name=ref.name)) from exc
                if on_error == OnError.DISABLE:
                    ref.disabled = True
                continue
            if result is None or result is NOOP_RESULT:
                continue
            if result.metadata:
                metadata.update(result.metadata)
            rd = getattr(result, "retry_delay_ms", 0)
            if rd and rd > retry_delay_ms:
                retry_delay_ms = rd
            mp = result.modified_payload
            if mp is not None and applies:
                new = self._apply_policy(hook, current, mp)
                if new is not current:
                    current, changed = new, True
            if not result.continue_processing or result.violation is not None:
                if result.violation is not None:
                    result.violation.plugin_name = ref.name
                if mode == PluginMode.SEQUENTIAL:
                    if violations_as_exceptions:
                        v = result.violation
                        raise PluginViolationError(f"{hook} blocked by plugin {ref.name}: {v.code} - {v.reason} ({v.description})" if v else f"{hook} blocked by plugin {ref.name}", violation=v)
                    return (PluginResult(continue_processing=False, modified_payload=current if changed else None, violation=result.violation, metadata=metadata,
                                         retry_delay_ms=retry_delay_ms), contexts)
                v = result.violation                  # (same record as the sequential executor's: logger `cpex.framework.manager`, "... raised violation ...")
                _mlog.warning("Plugin %s (%s) raised violation in %s: %s; continuing", ref.name, mode.value, hook, f"{v.code} - {v.reason}" if v else "continue_processing=False")
        return (fast_construct(PluginResult, {"continue_processing": True, "modified_payload": current if changed else None, "violation": None,
                                              "metadata": metadata, "retry_delay_ms": retry_delay_ms}), contexts)
