"""Randomly generated patterns (literals, classes, assertions, alternation, greedy/lazy repeats, random
flags) x random texts: the compiled engine (front-end + automata + BOTH prefilters, through the TEST-ONLY
host build) must agree with CPython `re.search` on every (pattern, text).  Bounded for the CPU suite;
tools/fuzz_regex.py runs the same generator for as long as one likes."""
import random
import re

import pytest

from hostsim_util import HostProgram
from mcp_context_forge_b200.regex_frontend import UnsupportedPattern

ALPH = list("abcKkſsé19_- \nİıx")


def atom(rng, depth):
    r = rng.random()
    if r < 0.45:
        return re.escape(rng.choice(ALPH))
    if r < 0.55:
        return rng.choice([".", r"\d", r"\w", r"\s", r"\W", r"\D", r"\S"])
    if r < 0.65:
        items = "".join(re.escape(rng.choice(ALPH)) for _ in range(rng.randint(1, 3)))
        if rng.random() < 0.3:
            items += "a-c"
        return "[" + ("^" if rng.random() < 0.3 else "") + items + "]"
    if r < 0.72:
        return rng.choice([r"\b", r"\B", "^", "$"])
    if depth <= 0:
        return re.escape(rng.choice(ALPH))
    return "(?:" + "|".join(seq(rng, depth - 1) for _ in range(rng.randint(1, 3))) + ")"


def seq(rng, depth):
    out = []
    for _ in range(rng.randint(1, 4)):
        a = atom(rng, depth)
        if a not in (r"\b", r"\B", "^", "$") and rng.random() < 0.35:
            a += rng.choice(["*", "+", "?", "{2}", "{1,3}", "*?", "+?", "??"])
        out.append(a)
    return "".join(out)


def pattern(rng):
    p = seq(rng, 2)
    if rng.random() < 0.15:
        p += "$"
    if rng.random() < 0.1:
        p = r"\A" + p
    if rng.random() < 0.1:
        p += r"\Z"
    fl = 0
    for f in (re.I, re.M, re.S, re.A):
        if rng.random() < 0.25:
            fl |= f
    return p, fl


def run_round(seed):
    """Returns (number of patterns checked, list of mismatches)."""
    rng = random.Random(seed)
    hp = HostProgram()
    pats = []
    for _ in range(rng.randint(1, 12)):
        p, fl = pattern(rng)
        try:
            c = re.compile(p, fl)
        except re.error:
            continue
        try:
            hp.add(p, fl)
        except UnsupportedPattern:
            continue
        pats.append((p, fl, c))
    if not pats:
        return 0, []
    units = ["".join(rng.choice(ALPH) for _ in range(rng.randint(0, 30))) for _ in range(300)] + ["", "\n", "a\n"]
    try:
        got, _ = hp.scan(units)
    except RuntimeError as exc:            # a loud, documented limit (automaton size) is not a mismatch
        if "too large" in str(exc):
            return 0, []
        raise
    bad = []
    for u, g in zip(units, got):
        exp = 0
        for i, (_, _, c) in enumerate(pats):
            if c.search(u):
                exp |= 1 << i
        if g != exp:
            i = ((g ^ exp) & -(g ^ exp)).bit_length() - 1
            bad.append((pats[i][0], pats[i][1], u, (g >> i) & 1))
    return len(pats), bad


@pytest.mark.parametrize("mode", ["0", "1"])
def test_random_patterns_agree_with_cpython(mode, monkeypatch):
    monkeypatch.setenv("CF_PAIR_FILTER", mode)     # byte prefilter / pair prefilter
    n = 0
    for rd in range(10):
        k, bad = run_round(7000 + 10 * int(mode) + rd)
        assert not bad, bad[:3]
        n += k
    assert n >= 30


# ---- substitution: random rules (capturing groups, patterns that can match "", templates with group references) vs re.subn
def sub_pattern(rng):
    p = seq(rng, 2).replace("(?:", "(" if rng.random() < 0.7 else "(?:")
    fl = 0
    for f in (re.I, re.M, re.S):
        if rng.random() < 0.2:
            fl |= f
    return p, fl


def sub_round(seed):
    from mcp_context_forge_b200.regex_frontend import template_parts

    rng = random.Random(seed)
    checked, bad = 0, []
    for _ in range(12):
        p, fl = sub_pattern(rng)
        try:
            c = re.compile(p, fl)
        except re.error:
            continue
        tmpl = "".join(rng.choice(["-", "<", ">", "é", "\\\\", ""] + [f"\\{g}" for g in range(1, c.groups + 1)] + ["\\g<0>"]) for _ in range(rng.randint(0, 4)))
        hp = HostProgram()
        try:
            hp.add(p, fl, ordered=True, repl=template_parts(tmpl, c))
            hp.compile()
        except UnsupportedPattern:
            continue
        except RuntimeError as exc:
            if "too large" in str(exc):
                continue
            raise
        checked += 1
        for _ in range(60):
            u = "".join(rng.choice(ALPH) for _ in range(rng.randint(0, 16)))
            got, n = hp.sub(0, u)
            exp, en = c.subn(tmpl, u)
            if (got, n) != (exp, en):
                bad.append((p, fl, tmpl, u, got, exp))
                break
    return checked, bad


def test_random_substitution_rules_agree_with_cpython():
    n = 0
    for rd in range(25):
        k, bad = sub_round(9100 + rd)
        assert not bad, bad[:3]
        n += k
    assert n >= 100
