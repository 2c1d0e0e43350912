"""CPU validation of the regex front-end + C++ back-end tables (prefilter, search DFA, ordered
DFA) against CPython `re`, through the TEST-ONLY host simulator (tests/hostsim).  The CUDA kernels
run the very same tables and the same scan_core.h routines; their parity is tested in
test_scan_gpu.py."""
import random
import re

import pytest

from hostsim_util import HostProgram

DEFAULT_LEXICONS = {  # /root/reference/plugins/harmful_content_detector/harmful_content_detector.py:36-52
    "self_harm": [r"\bkill myself\b", r"\bsuicide\b", r"\bself-harm\b", r"\bwant to die\b"],
    "violence": [r"\bkill (?:him|her|them|someone)\b", r"\bshoot (?:him|her|them|someone)\b", r"\bstab (?:him|her|them|someone)\b"],
    "hate": [r"\b(?:kill|eradicate) (?:[a-z]+) people\b", r"\b(?:racial slur|hate speech)\b"],
}
HARMFUL = [p for v in DEFAULT_LEXICONS.values() for p in v]

WORDS = ["kill", "myself", "suicide", "self-harm", "want", "to", "die", "him", "her", "them", "someone", "shoot", "stab",
         "eradicate", "people", "racial", "slur", "hate", "speech", "crap", "crud", "innovative", "the", "a", "of", "x",
         "Kill", "KILL", "ſuicide", "Kill", "é", "ß", "naïve", "日本語", "😀", "İ", "ı", "ͅ", "µ", "μ", "١٢", "12", "0",
         "_", "-", ".", ",", ":", "\n", "\t", " ", " ", "\x1c", "ǅ", "ǆ", "ᾳ", "Σ", "ς", "σ", "\ud800", "\U00010400", "\U00010428"]
SEPS = [" ", " ", " ", "", "  ", "\n", "-", "_", ".", "é", "1"]


def rand_text(rng, n):
    out = []
    for _ in range(n):
        out.append(rng.choice(WORDS))
        out.append(rng.choice(SEPS))
    return "".join(out)


def expected_bits(compiled, s):
    v = 0
    for i, c in enumerate(compiled):
        if c.search(s):
            v |= 1 << i
    return v


def test_default_lexicon_curated():
    hp = HostProgram()
    for p in HARMFUL:
        hp.add(p, re.I)
    comp = [re.compile(p, re.I) for p in HARMFUL]
    cases = ["I want to die", "kill myself", "ſuicide", "Kill him now", "ésuicide", "suicide_", "suicide١", "suicide", "x suicide.",
             "kill  him", "kill him", "Kill all people", "kill ſome people", "eradicate é people", "hate speech!", "", "self-harm",
             "selfharm", "shoot someonex", "stab Them", "WANT TO DIE", "killmyself", "kill myselfé", "suicide\n", "\nsuicide"]
    got, _ = hp.scan(cases)
    for s, g in zip(cases, got):
        assert g == expected_bits(comp, s), repr(s)


def test_default_lexicon_fuzz():
    rng = random.Random(1)
    hp = HostProgram()
    for p in HARMFUL:
        hp.add(p, re.I)
    comp = [re.compile(p, re.I) for p in HARMFUL]
    units = [rand_text(rng, rng.randint(0, 40)) for _ in range(3000)]
    got, stats = hp.scan(units)
    bad = [(u, g, expected_bits(comp, u)) for u, g in zip(units, got) if g != expected_bits(comp, u)]
    assert not bad, bad[:3]
    assert any(got)


FEATURE_PATTERNS = [
    (r"crap", 0), (r"cr[au]p+", 0), (r"\bkill\b", re.I), (r"\Bill", re.I), (r"^kill", 0), (r"^kill", re.M), (r"die$", 0), (r"die$", re.M),
    (r"\Akill", 0), (r"die\Z", 0), (r"k.ll", 0), (r"k.ll", re.S), (r"\d+", 0), (r"\w+-\w+", 0), (r"\s{2,}", 0), (r"[^a-z\s]+", re.I),
    (r"(?i)STRASSE|straße", 0), (r"[a-zA-Z]{5,7}\b", 0), (r"a|ab|abc", 0), (r"(?:ab|a)(?:bc|c)?x", 0), (r"x*?y", 0), (r"h[ae]te? sp", 0),
    (r"\bto\b.*\bdie\b", 0), (r"s(?i:UI)cide", 0), (r"[à-ÿ]+", re.I), (r"[kK]ill", 0), (r"ı|İ", re.I), (r"[^\W\d_]+", 0),
    (r"(?a)\w+\d", 0), (r"µ", re.I), (r"ͅ", re.I), (r"ǆ", re.I), (r"σ+", re.I), (r"[\U00010400-\U00010410]", re.I), (r"\.", 0),
    (r"", 0), (r"\b", 0), (r"\B", 0), (r"x?", 0), (r"(want|need) to (die|live)", 0), (r"[-_.]{2}", 0), (r"\x1c", 0), (r"\ud800", 0),
    (r"die$\n", 0), (r"kill$.", re.S), (r"(?:die|kill$)\s", 0), (r"$\n\Z", 0),          # `$` inside a pattern: holds before the FINAL newline only
]


@pytest.mark.parametrize("chunk", range(4))
def test_feature_patterns_search_fuzz(chunk):
    rng = random.Random(100 + chunk)
    pats = FEATURE_PATTERNS[chunk::4]
    hp = HostProgram()
    for p, f in pats:
        hp.add(p, f)
    comp = [re.compile(p, f) for p, f in pats]
    units = [rand_text(rng, rng.randint(0, 25)) for _ in range(1500)] + ["", "\n", "a", "é"]
    got, _ = hp.scan(units)
    for u, g in zip(units, got):
        exp = expected_bits(comp, u)
        assert g == exp, (u, [pats[i] for i in range(len(pats)) if (g ^ exp) >> i & 1])


def test_many_patterns_multiword_bitmap():
    rng = random.Random(7)
    vocab = ["w%03d" % i for i in range(200)]
    hp = HostProgram()
    for w in vocab:
        hp.add_ast(hp.fe.literal_ast(w))
    stats = hp.compile()
    assert stats[1] == 4  # 200 patterns -> 4 u64 words
    units = [" ".join(rng.choice(vocab + ["zzz", "w", "w0"]) for _ in range(rng.randint(0, 12))) for _ in range(300)]
    got, _ = hp.scan(units)
    for u, g in zip(units, got):
        exp = sum(1 << i for i, w in enumerate(vocab) if w in u)
        assert g == exp


SUB_RULES = [
    (r"crap", 0, "crud"), (r"crud", 0, "yikes"), (r"cr[au]p+", 0, "X"), (r"\bkill\b", re.I, "[k]"), (r"a|ab|abc", 0, "<>"),
    (r"(?:ab|a)(?:bc|c)?x", 0, ""), (r"x+?", 0, "y"), (r"\d+", 0, "#"), (r"\w+-\w+", 0, "é"), (r"[^a-z\s]+", re.I, "·"),
    (r"k.l+", re.S, "日本"), (r"s(?i:UI)cide", 0, "ſ"), (r"to (?:die|live)\b", 0, "—"), (r"\s{2,}", 0, " "), (r"e\Z", 0, "E"),
]


@pytest.mark.parametrize("idx", range(len(SUB_RULES)))
def test_sub_leftmost_first_fuzz(idx):
    pat, flags, repl = SUB_RULES[idx]
    rng = random.Random(200 + idx)
    hp = HostProgram()
    hp.add(pat, flags, ordered=True, repl=repl)
    c = re.compile(pat, flags)
    units = [rand_text(rng, rng.randint(0, 25)) for _ in range(800)] + ["", "crapcrap", "abcx abx acx ax", "xxxx"]
    for u in units:
        got, n = hp.sub(0, u)
        exp, en = c.subn(repl.replace("\\", "\\\\"), u)
        assert got == exp and n == en, (u, got, exp)


def test_unsupported_and_invalid():
    from mcp_context_forge_b200.regex_frontend import UnsupportedPattern, compile_ast

    for bad in [r"(a)\1", r"(?=a)b", r"(?<!a)b", r"(?>a)b", r"a*+", r"(a)?(?(1)b|c)"]:
        with pytest.raises(UnsupportedPattern):
            compile_ast(bad)
    with pytest.raises(re.error):
        compile_ast(r"(unclosed")
    compile_ast(r"a$b")                        # a `$` in the middle is legal (and, without MULTILINE, can still hold before a final newline)
    with pytest.raises(UnsupportedPattern):
        compile_ast(r"(?:a*)*b", 0, "sub")     # unbounded repeat of a nullable body: sre's empty-iteration rule is not expressible


NULLABLE_RULES = [
    (r"x*", 0, "-"), (r"x*?", 0, "-"), (r"", 0, "·"), (r"\b", 0, "|"), (r"^", re.M, "> "), (r"$", re.M, ";"), (r"[ \t]*$", re.M, "!"), (r"\s*\Z", 0, ""), (r"^\s*", 0, ""),
    (r"a?", 0, "<>"), (r"a??", 0, "[]"), (r"(?:ab)?", 0, "é"), (r"\B", 0, "_"), (r"x*|y", 0, "#"), (r"|x", 0, "#"), (r"x|", 0, "#"), (r"\d*", 0, "N"),
    (r"(?:x|xy)?", 0, "Q"), (r"$", 0, "!"), (r"\s*$", 0, ""), (r"x*$", 0, "E"), (r"$\n?", 0, "<>"), (r"(?:a|$)b?", 0, "~"), (r"\Z", 0, "END"), (r"\A\s*", 0, "^"), (r"[ \t]*(?:\n|\Z)", 0, "/"), (r"k*\b", re.I, "."), (r"\w{0,2}", 0, "w"),
]


@pytest.mark.parametrize("idx", range(len(NULLABLE_RULES)))
def test_sub_rules_that_can_match_the_empty_string(idx):
    """`re.sub` with empty matches (Python >= 3.7: an empty match adjacent to a previous match is replaced; after an empty match
    the same position is retried with must_advance, Modules/_sre/sre.c pattern_subx)."""
    pat, flags, repl = NULLABLE_RULES[idx]
    rng = random.Random(900 + idx)
    hp = HostProgram()
    hp.add(pat, flags, ordered=True, repl=repl)
    c = re.compile(pat, flags)
    frag = ["x", "xx", "y", "a", "ab", "b", " ", "  ", "\n", "\t", "1", "23", "k", "K", "é", "日", "-", "_", "xy", "\n\n", " \n"]
    units = ["", "x", "abxd", "xxx", "\n", "a\n", "\na", " ", "ab ab"] + ["".join(rng.choice(frag) for _ in range(rng.randint(0, 14))) for _ in range(700)]
    for u in units:
        got, n = hp.sub(0, u)
        exp, en = c.subn(repl.replace("\\", "\\\\"), u)
        assert got == exp and n == en, (pat, u, got, exp)


GROUP_RULES = [
    (r"(\w+)@(\w+)", 0, r"\2 at \1"), (r"(a)(b)?", 0, r"[\1|\2]"), (r"(?P<k>k\w*)", re.I, r"<\g<k>>"), (r"(x+)(y*)", 0, r"\g<0>\g<0>"),
    (r"(?:(a)|b)+", 0, r"{\1}"), (r"(a|ab)(c|bcd)(d*)", 0, r"\3-\2-\1"), (r"(\d+)-(\d+)", 0, r"\2-\1"), (r"(\s*)(\S+)", 0, r"\2\1"),
    (r"(x*)", 0, r"(\1)"), (r"(a?)(b?)", 0, r"<\1\2>"), (r"\b(\w)(\w*)", 0, r"\2\1ay"), (r"((a)|(b))+?c", 0, r"\1\2\3"),
    (r"(é|日)(.)", re.S, r"\2\1"), (r"^(\s*)#(.*)$", re.M, r"\1//\2"), (r"(a+)(a*)", 0, r"\1,\2"), (r"(a+?)(a*)", 0, r"\1,\2"),
    (r"(?:(x)|(y)|(z)){2,3}", 0, r"\1\2\3"), (r"(a(b(c)?)?)", 0, r"\3\2\1"), (r"(?i:(k))(?:-(\d))?", 0, r"\2\1"),
]


@pytest.mark.parametrize("idx", range(len(GROUP_RULES)))
def test_sub_templates_with_group_references(idx):
    """`pattern.sub(template, text)` with `\\1` / `\\g<name>` / `\\g<0>` references: the ordered DFA finds the match, a Pike-VM pass over
    the rule's NFA (cf::pike_captures, the same function the kernel runs) finds the group spans of that match."""
    from mcp_context_forge_b200.regex_frontend import template_parts

    pat, flags, tmpl = GROUP_RULES[idx]
    rng = random.Random(1300 + idx)
    c = re.compile(pat, flags)
    hp = HostProgram()
    hp.add(pat, flags, ordered=True, repl=template_parts(tmpl, c))
    frag = ["a", "b", "c", "d", "ab", "abc", "bcd", "x", "xx", "y", "z", "k", "K", "-", "1", "23", "@", " ", "  ", "\n", "#", "é", "日", "w0", "_"]
    units = ["", "a", "ab", "abcd", "abcdd", "xxyy", "k-1", "user@host", " # c\n#d"] + ["".join(rng.choice(frag) for _ in range(rng.randint(0, 14))) for _ in range(700)]
    for u in units:
        got, n = hp.sub(0, u)
        exp, en = c.subn(tmpl, u)
        assert got == exp and n == en, (pat, tmpl, u, got, exp)


def test_template_errors_are_loud():
    from mcp_context_forge_b200.regex_frontend import template_parts

    with pytest.raises(re.error):
        template_parts(r"\2", re.compile(r"(a)"))            # invalid group reference, as pattern.sub would raise
    hp = HostProgram()
    hp.add_ast(hp.fe.compile_ast(r"(a)", 0, "sub", groups=False), True, ["x", 1])    # the AST carries no group 1
    with pytest.raises(RuntimeError):
        hp.compile()


def _big_rule_set(n_words, seed=0):
    rng = random.Random(seed)
    syll = ["ba", "co", "di", "fu", "ge", "ha", "ki", "lo", "mu", "ne", "pi", "qua", "ro", "su", "ty", "vo", "wi", "xe", "yo", "zu", "sch", "tion", "ing"]
    words = set()
    while len(words) < n_words:
        words.add("".join(rng.choice(syll) for _ in range(rng.randint(2, 4))))
    return sorted(words)


@pytest.mark.parametrize("n_words", [40, 300])
def test_large_rule_sets_switch_to_the_pair_prefilter_and_stay_exact(n_words):
    """Large deny lists: the compiler must pick the pair prefilter (the byte filter would admit far too many
    windows) and the verdicts must still equal CPython's, including case-insensitive and \\b rules mixed in."""
    import re as _re
    words = _big_rule_set(n_words)
    p = HostProgram()
    pats = []
    for i, w in enumerate(words):
        if i % 7 == 0:
            pat, fl = r"\b" + w + r"\b", _re.I
        elif i % 7 == 1:
            pat, fl = w[:-1] + "[" + w[-1] + w[-1].upper() + "]s?", 0
        else:
            pat, fl = _re.escape(w), 0
        pats.append((pat, fl))
        p.add(pat, fl)
    stats = p.compile()
    if n_words >= 300:
        assert stats[7] == 1, "pair prefilter expected for a large rule set"
    rng = random.Random(1)
    filler = "the quick brown fox jumps over the lazy dog while json payloads flow through gateways".split()
    units = []
    for u in range(400):
        toks = []
        for _ in range(rng.randint(0, 60)):
            r = rng.random()
            if r < 0.08:
                w = rng.choice(words)
                w = w.upper() if rng.random() < 0.3 else w
                toks.append(w + ("s" if rng.random() < 0.2 else ""))
            elif r < 0.12:
                toks.append(rng.choice(words)[:-1])                 # near miss
            elif r < 0.14:
                toks.append("ſK" + rng.choice(words))       # long s / Kelvin sign next to a word
            else:
                toks.append(rng.choice(filler))
        units.append((" " if rng.random() < 0.5 else "_").join(toks))
    got, st = p.scan(units)
    exp = []
    comp = [_re.compile(pt, fl) for pt, fl in pats]
    for u in units:
        v = 0
        for i, c in enumerate(comp):
            if c.search(u):
                v |= 1 << i
        exp.append(v)
    assert got == exp
    assert sum(1 for v in exp if v) > 100
