"""TEST-ONLY: the plugins' device interface (`GpuBatcher`, `engine.Program`, the manager's fused launch) on the CPU simulator of
the engine (tests/hostsim: the same front-end, automata, substitution / capture routines and JSON->TOON encoder the kernels are
built from).  With `install(monkeypatch)` the drop-in plugins and `BatchedPluginManager` run end to end in the `-m "not gpu"`
suite: their host logic — unit extraction, verdict -> result assembly, speculate / replay — is the product's own code; only the
launches are simulated.  Nothing here is reachable from the product package."""
from __future__ import annotations

from typing import List, Optional, Sequence

import hostsim_util as hs
from mcp_context_forge_b200 import batching, engine, manager as mgr
from mcp_context_forge_b200 import regex_frontend as fe


class RecordingProgram(engine.Program):
    """engine.Program that also remembers what was added, so that the same program can be rebuilt on the simulator's builder."""

    def __init__(self):
        super().__init__()
        self.rec: List[tuple] = []

    def add_search(self, pattern: str, flags: int = 0) -> int:
        self.rec.append(("search", pattern, flags))
        return super().add_search(pattern, flags)

    def add_literal(self, word: str) -> int:
        self.rec.append(("literal", word))
        return super().add_literal(word)

    def add_sub(self, pattern: str, flags: int, replacement) -> int:
        self.rec.append(("sub", pattern, flags, replacement))
        return super().add_sub(pattern, flags, replacement)

    def compile(self, ctx):                       # no device here: the host half of compilation is what exists on a CPU box
        self.compile_host()
        self.h, self.ctx = object(), ctx
        return self


class SimProgram:
    def __init__(self, rec: Sequence[tuple]):
        self.hp = hp = hs.HostProgram()
        self.rule_bits: List[int] = []
        for r in rec:
            if r[0] == "search":
                hp.add(r[1], r[2])
            elif r[0] == "literal":
                hp.add_ast(fe.literal_ast(r[1]))
            else:
                repl = r[3]
                if isinstance(repl, str):
                    repl = [repl] if repl else []
                refs = any(isinstance(p, int) for p in repl)
                self.rule_bits.append(hp.add(r[1], r[2], ordered=True, repl=repl if refs else "".join(repl)))
        self.empty = not rec
        if rec:
            hp.compile()

    def scan(self, units: Sequence[str]) -> List[int]:
        return [0] * len(units) if self.empty else self.hp.scan(list(units))[0]

    def sub(self, text: str) -> str:
        for r in range(len(self.rule_bits)):
            text = self.hp.sub(r, text)[0]
        return text

    @property
    def rule_mask(self) -> int:
        m = 0
        for b in self.rule_bits:
            m |= 1 << b
        return m




def sim_of(prog) -> SimProgram:
    """The simulator's build of `prog`, kept ON the program object (an id()-keyed cache would hand a dead program's tables to a new one)."""
    held = getattr(prog, "_sim", None)
    if held is None or held[0] != len(prog.rec):
        held = prog._sim = (len(prog.rec), SimProgram(prog.rec))
    return held[1]


def _text(u) -> str:
    return u if isinstance(u, str) else bytes(u).decode("utf-8", "surrogatepass")


def _toon(t: str, report_errors: bool):
    """cf_toon's contract per unit.  With CF_TOON_REPORT_ERRORS the encoder runs to the end even when the output has outgrown the input, so a
    ValueError / AttributeError anywhere in the document is reported; the text is kept only when strictly smaller (include/cfgpu.h)."""
    if not report_errors:
        return hs.toon_host(t)
    st, txt = hs.toon_host(t, unlimited=True)
    if st == 0 and len(txt.encode("utf-8", "surrogatepass")) >= len(t.encode("utf-8", "surrogatepass")):
        return 1, None                                          # CF_TOON_NOT_SMALLER
    return st, txt


class SimBatcher(batching.GpuBatcher):
    """GpuBatcher whose synchronous cores run on the simulator (the asyncio coalescing above them is the product's)."""

    def __init__(self):
        self.window_us, self.max_units = 0, 1 << 16
        self._pending, self._scheduled = {}, {}
        import threading

        self._launch_lock = threading.RLock()
        self.launches = self.units_seen = 0
        self.ctx = None

    def _account(self, groups):
        self.launches += 1
        self.units_seen += sum(len(g) for g in groups)

    def scan_groups(self, prog, groups):
        self._account(groups)
        sp = sim_of(prog)
        return [sp.scan([_text(u) for u in g]) for g in groups]

    def sub_groups(self, prog, groups, rule_mask):
        self._account(groups)
        sp = sim_of(prog)
        out = []
        for g in groups:
            ts = [_text(u) for u in g]
            out.append([sp.sub(t).encode("utf-8", "surrogatepass") if b & rule_mask else None for t, b in zip(ts, sp.scan(ts))])
        return out

    def scan_sub_groups(self, prog, groups, rule_mask):
        self._account(groups)
        sp = sim_of(prog)
        out = []
        for g in groups:
            ts = [_text(u) for u in g]
            out.append([(b, sp.sub(t).encode("utf-8", "surrogatepass") if b & rule_mask else None) for t, b in zip(ts, sp.scan(ts))])
        return out

    def toon_groups(self, groups, report_errors: bool = True):
        self._account(groups)
        res = []
        for g in groups:
            row = []
            for u in g:
                st, txt = _toon(_text(u), report_errors)
                row.append((st, txt.encode("utf-8") if st == 0 else None))
            res.append(row)
        return res


def sim_launch(self, chain, units, stages):
    """BatchedPluginManager._launch on the simulator: the contract of cf_run_batch (include/cfgpu.h) per unit."""
    sp = sim_of(chain.prog)
    bms = sp.scan(units) if chain.has_patterns else [0] * len(units)
    out = []
    for u, st, bm in zip(units, stages, bms):
        rew: Optional[str] = None
        if (st & 2) and (bm & sp.rule_mask):
            rew = sp.sub(u)
        tstat, ttxt = 8, None                                   # CF_TOON_SKIPPED
        if st & 8 and rew is None:
            tstat, t = _toon(u, bool(chain.toon_flags & 1))
            ttxt = t.encode("utf-8") if tstat == 0 else None
        out.append(mgr.UnitResult(bm, rew, tstat, ttxt))
    return out


def sim_mask_host(batch, stream, offsets, max_depth: int = 10):
    """engine.mask_host on the host build of csrc/json_mask.h: (status int32[n], masked bytes or None per unit)."""
    import numpy as np

    raw = bytes(stream)
    st, outs = [], []
    for i in range(len(offsets) - 1):
        s, o = hs.mask_host(raw[int(offsets[i]):int(offsets[i + 1]) - 1], max_depth)
        st.append(s)
        outs.append(o if s == 0 else None)
    return np.asarray(st, dtype=np.int32), outs


def sim_classify_keys_host(batch, keys):
    return [hs.key_sensitive_host(k if isinstance(k, str) else bytes(k).decode("utf-8", "surrogatepass")) for k in keys]


def install(monkeypatch) -> SimBatcher:
    sim = SimBatcher()
    from mcp_context_forge_b200 import masking

    monkeypatch.setattr(engine, "mask_host", sim_mask_host)                       # the masking module's three launches, on the simulator
    monkeypatch.setattr(engine, "classify_keys_host", sim_classify_keys_host)
    monkeypatch.setattr(masking, "_batch", lambda nbytes, nunits: None)
    monkeypatch.setattr(masking, "_fallback_prog", None)
    monkeypatch.setattr(engine, "Program", RecordingProgram)
    monkeypatch.setattr(batching.GpuBatcher, "get", classmethod(lambda cls, device=0: sim))
    monkeypatch.setattr(engine.Context, "get", classmethod(lambda cls, device=0: object()))
    monkeypatch.setattr(mgr.BatchedPluginManager, "_launch", sim_launch)
    return sim
