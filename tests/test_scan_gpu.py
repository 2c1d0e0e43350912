"""GPU parity of the fused multi-pattern scan kernel (through the C ABI) against the oracle
(oracle/hook_chain_ref.py — CPython `re`, the reference's own matcher).  Bit-exact."""
import random
import re

import numpy as np
import pytest

from mcp_context_forge_b200 import engine, synth
from oracle import hook_chain_ref as ref

pytestmark = pytest.mark.gpu

HARMFUL = [(p, re.I) for pats in ref.DEFAULT_LEXICONS.values() for p in pats]
DENY = ["innovative", "groundbreaking", "revolutionary"]
SUBS = [("crap", 0, "crud"), ("crud", 0, "yikes")]


def default_program():
    p = engine.Program()
    for pat, f in HARMFUL:
        p.add_search(pat, f)
    for w in DENY:
        p.add_literal(w)
    for pat, f, r in SUBS:
        p.add_sub(pat, f, r)
    return p.compile(engine.Context.get())


def oracle_bits(units):
    return ref.scan_bitmaps(units, HARMFUL, DENY, [(p, f) for p, f, _ in SUBS])


def test_curated_cases():
    prog = default_program()
    cases = ["I want to die", "kill myself", "ſuicide", "Kill him now", "ésuicide", "suicide_", "suicide١", "suicide", "x suicide.",
             "kill  him", "Kill all people", "kill ſome people", "eradicate é people", "hate speech!", "", "self-harm", "innovative crap",
             "crud", "Innovative", "shoot someonex", "stab Them", "WANT TO DIE", "a" * 70000 + " suicide", "suicide" + "b" * 33000]
    assert engine.scan_units(prog, cases) == oracle_bits(cases)


def test_fuzz_small_units():
    from test_regex_engine_cpu import rand_text

    rng = random.Random(5)
    prog = default_program()
    units = [rand_text(rng, rng.randint(0, 60)) for _ in range(20000)]
    got = engine.scan_units(prog, units)
    exp = oracle_bits(units)
    bad = [(u, g, e) for u, g, e in zip(units, got, exp) if g != e]
    assert not bad, bad[:3]
    assert sum(1 for g in got if g) > 100


@pytest.mark.parametrize("shape,size", [("A", 2048), ("A", 16384), ("C", 16384), ("C", 262144), ("B", 16384), ("C", 1 << 20)])
def test_payload_shapes(shape, size):
    prog = default_program()
    n = max(4, min(256, (8 << 20) // size))
    units = [synth.payload(shape, size, seed=s, hit_rate=2e-4) for s in range(n)]
    got = engine.scan_units(prog, units)
    assert got == oracle_bits(units)
    if shape == "C":
        assert any(got)


def test_hits_at_tile_and_lane_boundaries():
    """Place a hit at every offset around the 64-byte lane chunks and the 16 KiB tile edge."""
    prog = default_program()
    units = []
    for off in list(range(16384 - 40, 16384 + 40)) + list(range(0, 140)):
        units.append("x" * off + " suicide " + "y" * 50)
    # many units in one stream -> hits land at arbitrary stream offsets as well
    got = engine.scan_units(prog, units)
    assert got == oracle_bits(units)
    assert all(g == 1 << 1 for g in got)


def test_feature_patterns_gpu():
    from test_regex_engine_cpu import FEATURE_PATTERNS, rand_text

    rng = random.Random(11)
    pats = [pf for pf in FEATURE_PATTERNS]
    p = engine.Program()
    for pat, f in pats:
        p.add_search(pat, f)
    p.compile(engine.Context.get())
    units = [rand_text(rng, rng.randint(0, 30)) for _ in range(4000)] + ["", "\n", "a"]
    got = engine.scan_units(p, units)
    exp = ref.scan_bitmaps(units, pats, [], [])
    bad = [(u, bin(g ^ e)) for u, g, e in zip(units, got, exp) if g != e]
    assert not bad, bad[:3]


def test_many_patterns_multiword():
    rng = random.Random(3)
    vocab = ["w%03d" % i for i in range(200)]
    p = engine.Program()
    for w in vocab:
        p.add_literal(w)
    p.compile(engine.Context.get())
    units = [" ".join(rng.choice(vocab + ["zzz", "w0"]) for _ in range(rng.randint(0, 40))) for _ in range(2000)]
    got = engine.scan_units(p, units)
    exp = ref.scan_bitmaps(units, [], vocab, [])
    assert got == exp


def test_full_size_checksum_property():
    """BASELINE size (16 KiB x 4096 payloads): idempotence + count of flagged units equals the oracle's
    on a bounded sample, and a second scan gives identical bitmaps."""
    prog = default_program()
    base = [synth.payload("C", 16384, seed=s, hit_rate=1e-4) for s in range(64)]
    units = [base[i % 64] for i in range(4096)]
    got1 = engine.scan_units(prog, units)
    got2 = engine.scan_units(prog, units)
    assert got1 == got2
    exp = oracle_bits(base)
    assert got1[:64] == exp
    assert all(got1[i] == exp[i % 64] for i in range(4096))


def test_candidate_dense_input_overflows_queue():
    """Every few bytes is a prefilter candidate (several million in one launch, more than the
    candidate queue holds): the in-place verification path must give the same bitmaps."""
    prog = default_program()
    rng = random.Random(7)
    words = ["kill", "bomb", "crap", "crud", "killer", "skill", "innovative", "suicid", "bombs", "kil", "k1ll",
             "assault", "Kill yourself", "self-harm", "I hate", "innovativ"]
    base = []
    for s in range(16):
        parts = []
        n = 0
        while n < 16000:
            w = rng.choice(words)
            parts.append(w)
            n += len(w) + 1
        base.append(" ".join(parts))
    units = [base[i % 16] for i in range(2048)]
    got = engine.scan_units(prog, units)
    exp = oracle_bits(base)
    assert all(got[i] == exp[i % 16] for i in range(2048))
    cand, _ = engine.Context.get().scan_counters()
    assert cand > (1 << 20)          # really did exceed the queue capacity


@pytest.mark.parametrize("n_extra", [40, 400])
def test_large_rule_sets_pair_prefilter_kernel(n_extra, monkeypatch):
    """The pair-prefilter variant of the scan kernel (forced, so that both table kinds are covered at both rule
    set sizes) against the oracle: default rules + many more deny words, hits at tile / lane / pair boundaries."""
    monkeypatch.setenv("CF_PAIR_FILTER", "1")
    rng = random.Random(n_extra)
    syll = ["ba", "co", "di", "fu", "ge", "ha", "ki", "lo", "mu", "ne", "pi", "qua", "ro", "su", "ty", "vo", "wi", "xe", "yo", "zu", "sch", "tion"]
    words = set()
    while len(words) < n_extra:
        words.add("".join(rng.choice(syll) for _ in range(rng.randint(2, 4))))
    words = sorted(words)
    p = engine.Program()
    for pat, f in HARMFUL:
        p.add_search(pat, f)
    for w in DENY + words:
        p.add_literal(w)
    for pat, f, r in SUBS:
        p.add_sub(pat, f, r)
    st = p.compile_host()
    assert st.prefilter == 1
    p.compile(engine.Context.get())
    base = [synth.payload(s, 16384, seed=k, hit_rate=1e-3) for k, s in enumerate("ABCCAB")]
    units = []
    for i in range(600):
        t = base[i % 6]
        cut = rng.randrange(0, 4096)
        w = rng.choice(words + DENY + ["kill him", "Kill yourself", "crap"])
        units.append(t[:cut] + " " + w + " " + t[cut:cut + rng.randrange(100, 12000)])
    # exact 64-byte / 2048-byte alignment sweeps of one hit
    for off in range(0, 130):
        units.append("x" * off + words[off % len(words)] + " tail")
    got = engine.scan_units(p, units)
    exp = ref.scan_bitmaps(units, HARMFUL, DENY + words, [(q, f) for q, f, _ in SUBS])
    assert got == exp
    assert sum(1 for v in exp if v) > 600


def test_word_boundary_after_a_multibyte_character_at_the_very_start_of_the_stream():
    """The character before a candidate is decoded by walking back over continuation bytes; for a candidate within the first four
    bytes of unit 0 that walk must stop at the stream start (regression: the look-back bound wrapped around and the boundary
    assertion saw the continuation byte as a non-word character)."""
    p = engine.Program()
    p.add_search(r"\bDROP\b", re.I)
    p.add_search(r"\Bkill", re.I)
    p.compile(engine.Context.get())
    comp = [re.compile(r"\bDROP\b", re.I), re.compile(r"\Bkill", re.I)]
    for first in ["İDROP TABLE users", "éDROP x", "日DROP", "\U0001F600DROP", "aDROP", "DROP", "İkill", "日kill", "é kill"]:
        units = [first, "x DROP y", "İDROP"]
        got = engine.scan_units(p, units)
        exp = [sum(1 << i for i, c in enumerate(comp) if c.search(u)) for u in units]
        assert got == exp, (first, got, exp)
