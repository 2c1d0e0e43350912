"""The token-parallel TOON kernel body (csrc/json_tp.h) on the CPU: the very source toon_tp_kernel compiles, executed
lane for lane by the TEST-ONLY 32-fibre warp emulator (tests/hostsim/warp_emu.cpp), against
  * the vectors recorded from the reference's own toon.py / toon_encoder.py (tests/golden/toon.json),
  * the sequential encoder (csrc/json_toon.h, pinned to the same vectors) on a seeded differential fuzz,
  * the oracle on the synthetic bench payloads.
Status 7 means "handed to the sequential encoder" (a GPU path too): allowed, never wrong."""
import json
import os
import random
import sys

import pytest

import hostsim_util as hs
from mcp_context_forge_b200 import synth
from oracle import toon_ref

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_toon_tp  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden", "toon.json")
ST = {"converted": 0, "not_smaller": 1, "not_json": 2, "value_error": 3, "attr_error": 4, "fallback": 7}


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)


@pytest.mark.parametrize("order", [0, 1])
def test_reference_golden_encode_cases(gold, order):
    n_ok = n_err = n_fb = 0
    for c in gold["encode"]:
        st, got = hs.toon_tp(c["json"], unlimited=True, order=order)
        if st == ST["fallback"]:
            n_fb += 1
            continue
        if "toon" in c:
            assert st == 0 and got == c["toon"], (c["json"][:300], st, got, c["toon"])
            n_ok += 1
        else:
            assert st == (ST["attr_error"] if c["error"] == "AttributeError" else ST["value_error"]), (c["json"][:200], st, c["error"])
            n_err += 1
    assert n_ok > 350 and n_err > 5, (n_ok, n_err, n_fb)


def test_plugin_decisions_golden(gold):
    n = conv = 0
    for block in gold["plugin"]:
        cfg = block["config"] or {}
        if cfg.get("exclude_tools") or cfg.get("include_tools") or cfg.get("skip_on_error") is False:
            continue
        lo, hi = cfg.get("min_size_bytes", 100), cfg.get("max_size_bytes", 1 << 20)
        for c in block["cases"]:
            item = c["result"]["content"][0]
            text = item["text"]
            size = len(text.encode("utf-8"))
            if size < lo or size > hi:
                continue
            new = (c["modified"] or c["result"])["content"][0]
            st, got = hs.toon_tp(text, report_errors=False, order=n & 1)
            if st == ST["fallback"]:
                continue
            if new == item:
                assert st != 0, text[:200]
            else:
                assert st == 0 and got == new["text"], text[:200]
                conv += 1
            n += 1
    assert n > 80 and conv > 15, (n, conv)


def test_differential_fuzz_vs_sequential_encoder():
    rng = random.Random(20260921)
    case = fuzz_toon_tp.make_gen(rng)
    bad = []
    nfb = 0
    for it in range(6000):
        t = case()
        r = fuzz_toon_tp.check(t, rng.random() < 0.5, rng.random() < 0.5, (it & 1) | (rng.randrange(16) << 4))
        if r == "fallback":
            nfb += 1
        elif r:
            bad.append((t[:300], r))
    assert not bad, bad[:3]
    assert nfb < 2000


@pytest.mark.parametrize("shape,size", [("A", 600), ("A", 16384), ("B", 16384), ("A", 262144), ("C", 4096)])
def test_synthetic_payloads_vs_oracle(shape, size):
    for seed in range(4):
        text = synth.payload(shape, size, seed=seed)
        if shape == "C":
            text = json.dumps({"doc": text, "n": seed})
        exp = toon_ref.process_text(text, 0, 1 << 30)
        st, got = hs.toon_tp(text, report_errors=False, order=(seed & 1) | ((seed * 5 % 16) << 4))
        if st == ST["fallback"]:
            continue
        assert (got if st == 0 else None) == exp
        if shape == "A":
            assert st == 0          # the tabular shape must stay on the fast path


def test_strict_json_rejections():
    bad = ['{"a":1,}', "[1,]", "{'a':1}", "[01]", "[1.]", "[.5]", "[+1]", "NaN", "[Infinity]", '"\\x"', '"\\ud800"', '"\\udc00\\ud800"', '"a\tb"', "[1] x", "", "  ",
           '{"a" 1}', "[1 2]", '"\\u12g4"', "tru", "nul", '{"a":}', "[", "{", '"abc', "{:1}", '{"a"::1}', '{"a":1 "b":2}', '{"a":"b":1}', "[,1]", "[1,,2]", '{"a":1}}',
           "[]]", "1 2", '{"a":1,:2}', '["a":1]', "{1:2}", '[1}', '{"a":1]', "-", "1e", "--1", "[1" + "," * 300 + "2]", '"\x7f\xff"']
    for t in bad:
        for order in (0, 1):
            st, _ = hs.toon_tp(t.encode("latin1") if "\xff" in t else t, unlimited=True, order=order)
            assert st in (ST["not_json"],), (t, st)
    deep = "[" * 70 + "]" * 70
    assert hs.toon_tp(deep, unlimited=True)[0] in (6, 7)        # beyond the limits, or handed to the sequential encoder which reports 6
    ok64 = "[" * 64 + "]" * 64
    assert hs.toon_tp(ok64, unlimited=True) == hs.toon_host(ok64, unlimited=True)


def test_long_strings_and_chunk_boundaries():
    """Strings, scalars and separators that straddle the 32-byte front-end chunks and the 32-token batches."""
    rng = random.Random(5)
    for pad in range(0, 70):
        doc = json.dumps({"p" * (pad % 7 + 1): "x" * pad, "rows": [{"id": i, "v": "y" * (pad + i), "f": i * 1.5, "t": i % 2 == 0} for i in range(40)],
                          "long": " ".join("w%d" % rng.randint(0, 999) for _ in range(60 + pad)), "esc": "a\\nb" * (pad % 5), "n": -pad}, separators=(",", ":"))
        for order in (0, 1):
            assert hs.toon_tp(doc, unlimited=True, order=order | ((pad % 16) << 4)) == hs.toon_host(doc, unlimited=True), pad


def test_long_non_ascii_strings_whole_warp_utf8():
    """Strings of >= 96 bytes with non-ASCII content take the whole-warp UTF-8 validator: valid text in several scripts, every
    class of malformed sequence at varying offsets, and the Unicode-whitespace / special-character quoting rules at both ends."""
    rng = random.Random(11)
    pieces = ["é", "ß", "日本語", "\U0001F600", "Ünï", "a", " ", "word ", "x-y", "\u00a0", "\u3000", "\u2028", "K", "ſ"]
    docs = []
    for _ in range(300):
        body = "".join(rng.choice(pieces) for _ in range(rng.randint(40, 160)))
        lead = rng.choice(["", "\u00a0", "\u3000", " ", "é", "x"])
        tail = rng.choice(["", "\u00a0", "\u2028", " ", "é", "x"])
        docs.append(json.dumps({"k": lead + body + tail, "n": 1, lead + "key" + body[:60] + "é" * 40: 2}, ensure_ascii=False).encode("utf-8"))
    bad_seqs = [b"\xc0\x80", b"\xc1\xbf", b"\xe0\x80\x80", b"\xe0\x9f\xbf", b"\xed\xa0\x80", b"\xed\xbf\xbf", b"\xf0\x80\x80\x80", b"\xf0\x8f\xbf\xbf", b"\xf4\x90\x80\x80",
                b"\xf5\x80\x80\x80", b"\x80", b"\xbf", b"\xc3", b"\xe2\x82", b"\xf0\x9f\x98", b"\xff", b"\xc3\xc3\xa9", b"\xe2\x28\xa1", b"\xf0\x28\x8c\xbc"]
    good_seqs = [b"\xc2\x80", b"\xdf\xbf", b"\xe0\xa0\x80", b"\xed\x9f\xbf", b"\xee\x80\x80", b"\xef\xbf\xbf", b"\xf0\x90\x80\x80", b"\xf4\x8f\xbf\xbf"]
    for seq in bad_seqs + good_seqs:
        for off in (0, 1, 31, 32, 33, 95, 150):
            for at_end in (False, True):
                filler = ("é" * 100).encode("utf-8")
                body = filler[:off * 2 // 2 * 1] if False else (b"a" * off + filler)
                body = (body + seq) if at_end else (body[:off] + seq + body[off:])
                docs.append(b'{"k":"' + body + b'","z":[1,2]}')
    n_checked = 0
    for d in docs:
        a = hs.toon_host(d, unlimited=True)
        for order in (0, 1 | (7 << 4)):
            b = hs.toon_tp(d, unlimited=True, order=order)
            assert b[0] == 7 or a == b, (d[:120], a, b)
            n_checked += b[0] != 7
    assert n_checked > 800


def test_long_escaped_strings_whole_warp_escape_check():
    """Strings of >= 96 bytes with backslash escapes take the whole-warp escape validator (two-character escapes) or, with \\uXXXX,
    the sequential one: every escape kind, invalid escapes, backslash runs across the 32-byte chunks, quoting at both ends."""
    rng = random.Random(12)
    esc = ['\\n', '\\t', '\\r', '\\"', '\\\\', '\\/', '\\b', '\\f', '\\u00e9', '\\u0041', '\\ud83d\\ude00', '\\u0009', '\\u002c']
    plain = ["word", " ", "é", "x", "lorem ipsum", "a/b", "日本"]
    docs = []
    for _ in range(400):
        kinds = rng.sample(esc, rng.randint(1, 3)) if rng.random() < 0.8 else ['\\/']
        body = "".join(rng.choice(plain + kinds) for _ in range(rng.randint(30, 120)))
        lead = rng.choice(["", "\\/", "\\n", " ", "x"])
        tail = rng.choice(["", "\\/", "\\t", " ", "x", "\\\\"])
        docs.append(('{"k":"' + lead + body + tail + '","n":[1,2]}').encode("utf-8"))
    for bad in ['\\x', '\\u12g4', '\\ud800', '\\udc00', '\\u12', '\\ ', '\\a']:
        for off in (0, 30, 31, 32, 33, 64, 100):
            docs.append(('{"k":"' + "a" * off + bad + "b" * 120 + '"}').encode("utf-8"))
    for run in range(1, 9):                                     # backslash runs straddling chunk boundaries
        for off in (28, 29, 30, 31, 32, 60, 61, 62, 63):
            docs.append(('{"k":"' + "a" * off + "\\" * run + ('"' if run % 2 else 'n') + "b" * 100 + '"}').encode("utf-8"))
    n_checked = 0
    for d in docs:
        a = hs.toon_host(d, unlimited=True)
        for order in (0, 1 | (9 << 4)):
            b = hs.toon_tp(d, unlimited=True, order=order)
            assert b[0] == 7 or a == b, (d[:160], a, b)
            n_checked += b[0] != 7
    assert n_checked > 900


def test_duplicate_keys_are_handed_over_wherever_they_sit():
    """The duplicate-key screen (FNV hashes of an open object's keys on a stack in shared memory): a repeated
    key in the same run, behind a nested container, 40 members later, inside a nested object or a table row hands the unit to the
    sequential encoder (status 7); no false alarm for equal keys in DIFFERENT objects."""
    dup = ['{"a":1,"b":2,"a":3}', '{"a":1,"n":{"x":1},"a":3}', '{"a":1,' + ",".join(f'"k{i}":{i}' for i in range(40)) + ',"a":2}',
           '{"o":{"p":1,"q":2,"p":3},"z":1}', '[{"a":1,"b":2},{"a":1,"a":2}]']
    ok = ['{"a":1,"n":{"a":1},"b":{"a":2}}', "{" + ",".join(f'"k{i}":{{"a":{i}}}' for i in range(50)) + "}"]
    for order in (0, 1):
        for t in dup:
            assert hs.toon_tp(t, order=order)[0] == 7, t
        for t in ok:
            assert hs.toon_tp(t, order=order) == hs.toon_host(t), t
