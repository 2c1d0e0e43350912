"""Structural index + token-driven DOM builder (csrc/json_index.h — the parse path of the CUDA JSON
kernels) against the sequential parser (csrc/json_toon.h json_parse) and against an independent Python
restatement of the token definition.  CPU only, through the TEST-ONLY host build."""
import json
import os
import random

import pytest

import hostsim_util as hs
from mcp_context_forge_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def py_tokens(data: bytes):
    """Token positions by a plain scanner: a byte is escaped when an odd-length run of backslashes
    precedes it (inside or outside strings — outside it only matters for invalid documents); real quotes
    are the unescaped ones.  Tokens: structural characters outside strings, opening and closing quotes,
    first byte of every run of other non-whitespace bytes outside strings."""
    n = len(data)
    run = 0
    real_quote = [False] * n
    for i, c in enumerate(data):
        if c == 0x22 and run % 2 == 0:
            real_quote[i] = True
        run = run + 1 if c == 0x5C else 0
    toks = []
    in_str = False
    prev_other = False
    for i, c in enumerate(data):
        if in_str:
            if real_quote[i]:
                toks.append((i, True))
                in_str = False
            prev_other = False
            continue
        if real_quote[i]:
            toks.append((i, False))
            in_str = True
            prev_other = False
        elif c in b"{}[]:,":
            toks.append((i, False))
            prev_other = False
        elif c in b" \t\n\r":
            prev_other = False
        else:
            if not prev_other:
                toks.append((i, False))
            prev_other = True
    return toks, in_str


def corpus():
    out = []
    with open(os.path.join(GOLD, "toon.json"), encoding="utf-8") as f:
        g = json.load(f)
    out += [c["json"] for c in g["encode"]]
    for block in g["plugin"]:
        for c in block["cases"]:
            t = c.get("text")
            if isinstance(t, str):
                out.append(t)
    with open(os.path.join(GOLD, "masking_twin.json"), encoding="utf-8") as f:
        m = json.load(f)

    def strings(o):
        if isinstance(o, str):
            yield o
        elif isinstance(o, dict):
            for v in o.values():
                yield from strings(v)
        elif isinstance(o, list):
            for v in o:
                yield from strings(v)
    out += [s for s in strings(m) if s[:1] in "{["]
    for shape in "ABC":
        out += [synth.payload(shape, size, seed=s) for size in (300, 2048, 16384) for s in range(3)]
    return out


EDGE = [
    b"", b" ", b"{}", b"[]", b"{", b"}", b"[", b"]", b'"', b'""', b'"a', b'a"', b"1", b"-", b"-0", b"01", b"1.", b"1e", b"1e+", b"1.5e-3",
    b"true", b"truex", b"tru", b"nul", b"null ", b" false", b"123abc", b"[1 2]", b"[1,]", b"[,1]", b'{"a"}', b'{"a":}', b'{"a":1,}', b'{,}',
    b'{"a":1 "b":2}', b'{"a" 1}', b'["a""b"]', b'"a"1', b'1"a"', b"[1]]", b"[[1]", b'{"a":[}', b'{"a":]}', b"[}", b"{]",
    b'"\\', b'"\\"', b'"\\\\"', b'"\\\\\\"', b'"\\\\\\\\"', b'["\\\\","x"]', b'["\\"",1]', b'["a\\\\\\"b"]', b'"\\u00e9"', b'"\\ud83d\\ude00"',
    b'"\\ud83d"', b'"\\udc00"', b'"\\x"', b'"\x01"', b'"\xff"', b'"\xc3\xa9"', b'"\xc3"', b"\xc3\xa9", b'{"k":"\\n"}', b"[1,\x002]", b"\\", b'\\"x"',
    b'{"a":1,"a":2}', b'{"a":{"a":1,"b":2,"a":3},"a":4}', b" \t\n\r[ \t1 ,\n2 ]\r ", b'["' + b"\\" * 31 + b'"]', b'["' + b"\\" * 32 + b'"]',
    b'["' + b"\\" * 33 + b'"]', b'["' + b"\\" * 64 + b'"]', b'["' + b"\\" * 65 + b'"]', b"[" * 64 + b"]" * 64, b"[" * 65 + b"]" * 65,
    b'"' + b"x" * 31 + b'"', b'"' + b"x" * 30 + b'"', b'["' + b"y" * 29 + b'","z"]', b" " * 31 + b'"q"', b" " * 32 + b"7", b"9" * 40, b"[" + b"9" * 31 + b"]",
]


def test_token_definition_vs_python_scanner():
    for data in EDGE + [t.encode("utf-8", "surrogatepass") for t in corpus()[:400]]:
        got, unt = hs.json_index(data)
        exp, exp_unt = py_tokens(data)
        assert [(p, c) for p, c, _ in got] == exp and unt == exp_unt, data[:120]


def test_builder_equals_sequential_parser_on_corpus_and_edges():
    n_ok = n_bad = 0
    for data in EDGE + [t.encode("utf-8", "surrogatepass") for t in corpus()]:
        rc, a, b = hs.index_equiv(data)
        assert rc == 0, (rc, a, b, data[:200])
        n_ok += a == 0
        n_bad += a != 0
    assert n_ok > 500 and n_bad > 40


def test_builder_equals_sequential_parser_on_mutations():
    """Byte-level mutations of valid documents: both parsers must agree on accept/reject and on every node."""
    rng = random.Random(11)
    base = [t.encode("utf-8", "surrogatepass") for t in corpus() if 20 < len(t) < 3000][:150]
    alphabet = b'{}[]:,"\\ \n\t0123456789-+.eEtrufalsn\x00\x1f\x7f\xc3\xa9\xe2\x82\xac\xf0\x9f\x98\x80x'
    checked = 0
    for data in base:
        for _ in range(40):
            d = bytearray(data)
            for _ in range(rng.randint(1, 3)):
                op = rng.random()
                i = rng.randrange(len(d) + 1)
                if op < 0.4 and d:
                    d[min(i, len(d) - 1)] = rng.choice(alphabet)
                elif op < 0.7:
                    d[i:i] = bytes([rng.choice(alphabet)])
                elif d:
                    del d[min(i, len(d) - 1)]
            rc, a, b = hs.index_equiv(bytes(d))
            assert rc == 0, (rc, a, b, bytes(d)[:200])
            checked += 1
    assert checked >= 6000


def test_outputs_through_the_index_path_match_golden():
    with open(os.path.join(GOLD, "toon.json"), encoding="utf-8") as f:
        g = json.load(f)
    for c in g["encode"]:
        st, got = hs.toon_host(c["json"], unlimited=True, indexed=True)
        st2, got2 = hs.toon_host(c["json"], unlimited=True)
        assert (st, got) == (st2, got2)
        if "toon" in c:
            assert st == 0 and got == c["toon"]
    rng = random.Random(5)
    for _ in range(200):
        body = synth.payload("B", rng.choice([200, 1000, 5000]), seed=rng.randrange(10 ** 6)).encode()
        assert hs.mask_host(body, 10, indexed=True) == hs.mask_host(body, 10)
