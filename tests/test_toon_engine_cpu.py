"""The JSON -> DOM -> TOON code the CUDA kernel runs (csrc/json_toon.h), executed on the CPU through
the TEST-ONLY host build, against vectors recorded from the reference's own toon.py / toon_encoder.py
and against the oracle on synthetic payloads."""
import json
import os
import random

import pytest

import hostsim_util as hs
from mcp_context_forge_b200 import synth
from oracle import toon_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden", "toon.json")
ST = {"converted": 0, "not_smaller": 1, "not_json": 2, "value_error": 3, "attr_error": 4}


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)


def test_encoder_output_matches_reference(gold):
    n_ok = n_err = 0
    for c in gold["encode"]:
        st, got = hs.toon_host(c["json"], unlimited=True)
        if "toon" in c:
            assert st == 0 and got == c["toon"], (c["json"][:300], got, c["toon"])
            n_ok += 1
        else:
            assert st == (ST["attr_error"] if c["error"] == "AttributeError" else ST["value_error"]), (c["json"][:200], st, c["error"])
            n_err += 1
    assert n_ok > 700 and n_err > 10


def test_item_decision_matches_reference_plugin(gold):
    """text -> converted? (strictly smaller) with the product's capacity rule."""
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
            new = (c["modified"] or c["result"])["content"][0]
            if size < lo or size > hi:
                assert new == item
                continue
            st, got = hs.toon_host(text)
            if new == item:
                assert st != 0, text[:200]
            else:
                assert st == 0 and got == new["text"], text[:200]
                conv += 1
            n += 1
    assert n > 250 and conv > 50


@pytest.mark.parametrize("shape,size", [("A", 600), ("A", 16384), ("B", 16384), ("A", 262144)])
def test_synthetic_payloads_vs_oracle(shape, size):
    for seed in range(6):
        text = synth.payload(shape, size, seed=seed)
        exp = toon_ref.process_text(text, 0, 1 << 30)
        st, got = hs.toon_host(text)
        assert (got if st == 0 else None) == exp


def test_number_formatting_vs_python():
    rng = random.Random(3)
    vals = []
    for _ in range(3000):
        k = rng.random()
        if k < 0.3:
            vals.append(rng.uniform(-1e6, 1e6))
        elif k < 0.5:
            vals.append(rng.random() * 10 ** rng.randint(-30, 30))
        elif k < 0.6:
            vals.append(float(rng.randint(-10 ** 18, 10 ** 18)))
        elif k < 0.7:
            vals.append(rng.choice([5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, 4.9406564584124654e-324, 1e22, 1e23, 9007199254740993.0, 0.1 + 0.2, 1 / 3]))
        elif k < 0.8:
            vals.append(round(rng.uniform(0, 1000), rng.randint(0, 6)))
        else:
            vals.append(rng.uniform(-1, 1) * 1e-5)
    texts = [repr(v) for v in vals]
    texts += ["1e5", "1E5", "1e+5", "1.5e-7", "-1e-20", "1e-20", "0.000001", "0.00001", "123456789012345678", "1234567890123456789012", "-9223372036854775808",
              "-9223372036854775809", "18446744073709551615", "18446744073709551616", "0.30000000000000004", "100.0", "-0.0", "-0", "0e0", "0.0e5", "1.000000000000000000000001",
              "0.1000000000000000055511151231257827021181583404541015625", "2.5e-5", "99999999999999.98", "123456.7890", "0.000123456789012345678", "4.35", "1e308", "1e-400"]
    doc = "[" + ",".join(texts) + "]"
    st, got = hs.toon_host(doc, unlimited=True)
    exp = toon_ref.encode(toon_ref.loads_strict(doc))
    assert st == 0
    g, e = got.split(": ", 1)[1].split(","), exp.split(": ", 1)[1].split(",")
    bad = [(t, a, b) for t, a, b in zip(texts, g, e) if a != b]
    assert not bad, bad[:5]
    assert hs.toon_host("[1e400]", unlimited=True)[0] == ST["not_json"]      # orjson/yyjson reject inf


def test_strict_json_rejections_and_edge_cases():
    bad = ['{"a":1,}', "[1,]", "{'a':1}", "[01]", "[1.]", "[.5]", "[+1]", "NaN", "[Infinity]", '"\\x"', '"\\ud800"', '"\\udc00\\ud800"', '"a\tb"', "[1] x", "", "  ",
           '{"a" 1}', "[1 2]", '"\\u12g4"', "tru", "nul", '{"a":}', "[", "{", '"abc']
    for t in bad:
        assert hs.toon_host(t, unlimited=True)[0] == ST["not_json"], t
    assert hs.toon_host(b'"\xff"'.decode("latin1").encode("latin1").decode("utf-8", "surrogateescape").encode("utf-8", "surrogatepass").decode("utf-8", "surrogatepass"), unlimited=True)[0] == ST["not_json"]
    good = {' {"a" : [ 1 , 2 ] } ': "a[2]: 1,2", '{"a":1,"b":2,"a":3}': "a: 3\nb: 2", '"\\u00e9\\ud83d\\ude00\\/"': "é😀/", '{"k\\u0061":1,"ka":2}': "ka: 2",
            '{"key\\n":1}': "key\n: 1", '["-"]': '[1]: "-"', '{"a-b":1}': '"a-b": 1', '[[]]': "[1]:\n  - [0]:"}
    for t, exp in good.items():
        st, got = hs.toon_host(t, unlimited=True)
        assert st == 0 and got == exp, (t, got)
    deep = "[" * 70 + "]" * 70
    assert hs.toon_host(deep, unlimited=True)[0] == 6          # deeper than the device handles: reported, never guessed


def test_duplicate_keys_last_wins_small_and_huge_objects():
    """Duplicate keys keep the first position and the last value (dict semantics of orjson.loads), for
    small objects (hash screen in thread-local memory) and for objects with more keys than the screen
    holds (falls back to the node walk), nested inside each other."""
    rng = random.Random(3)

    def obj_text(nkeys, ndup, depth):
        keys = [f"k{i}" for i in range(nkeys)]
        items = [(k, rng.randint(0, 999)) for k in keys]
        for _ in range(ndup):
            items.insert(rng.randint(1, len(items)), (rng.choice(keys), rng.randint(1000, 1999)))
        parts = []
        for j, (k, v) in enumerate(items):
            if depth and j % 97 == 5:
                parts.append(f'"{k}":' + obj_text(rng.choice([3, 8, 40]), rng.randint(0, 3), depth - 1))
            else:
                parts.append(f'"{k}":{v}')
        return "{" + ",".join(parts) + "}"

    for nkeys, ndup in ((2, 1), (6, 2), (50, 5), (511, 3), (512, 2), (513, 2), (700, 9), (1500, 4)):
        text = obj_text(nkeys, ndup, 2)
        exp = toon_ref.encode(toon_ref.loads_strict(text))
        st, got = hs.toon_host(text, unlimited=True)
        assert st == 0 and got == exp, (nkeys, ndup)
    # escaped vs raw spelling of the same key are duplicates too
    text = '{"a\\u0062c": 1, "x": 2, "abc": 3}'
    st, got = hs.toon_host(text, unlimited=True)
    assert st == 0 and got == toon_ref.encode(toon_ref.loads_strict(text)) == "abc: 3\nx: 2"


def test_error_precedence_key_before_columnar_crash():
    """A list item whose FIRST key cannot be encoded (control character -> ValueError, toon.py:396) and whose value is a list that would crash the
    unchecked `.keys()` (AttributeError, toon.py:400-404): the reference encodes the key first, so ValueError (status 3) wins — found by
    tools/fuzz_vs_reference.py against the reference's own toon.py (the sequential encoder used to report 4)."""
    for t in ['[2,{"A.b":"null"},{"A\\b":[0,"trail ",true,"Z1",null]}]', '[{"k\\u0001":[1,2]}]', '{"x":[{"\\u0007":[0]},3]}']:
        assert hs.toon_host(t, unlimited=True) == (3, None), t
        st, _ = hs.toon_tp(t, unlimited=True)
        assert st in (3, 7), t
        try:
            toon_ref.encode(toon_ref.loads_strict(t))
            raise AssertionError("the oracle should raise ValueError")
        except ValueError:
            pass
    assert hs.toon_host('[{"k":[1,2]}]', unlimited=True)[0] == 4           # the crash alone is still an AttributeError
