"""Pins oracle/mask_ref.py: the Rust crate's own unit tests (lib.rs:380-425) and vectors recorded
from the Python twin the reference ships for the same algorithm (tests/golden/masking_twin.json)."""
import json
import os

import pytest

from oracle import mask_ref as m

GOLD = os.path.join(os.path.dirname(__file__), "golden", "masking_twin.json")


@pytest.fixture(scope="module")
def gold():
    with open(GOLD, encoding="utf-8") as f:
        return json.load(f)


def test_crate_unit_tests():
    assert m.normalize_key("__ClientSecret__") == "client_secret"                      # lib.rs:389-394
    assert m.normalize_key("auth-token---") == "auth_token"
    assert m.mask_cookie_header("jwt_token=abc; theme=dark; session_id=xyz") == "jwt_token=******; theme=dark; session_id=******"   # lib.rs:398-404
    assert m.mask_cookie_header("theme=dark") == "theme=dark"
    assert m.mask_value({"password": "secret", "nested": {"authToken": "abc", "count": 3}}, 10) == {"password": "******", "nested": {"authToken": "******", "count": 3}}  # :408-419


def test_classifier_matches_python_twin(gold):
    for key, norm, sens in gold["classifier"]:
        assert m.normalize_key(key) == norm, key
        assert m.is_sensitive_key(key) == sens, key
    assert sum(1 for _, _, s in gold["classifier"] if s) > 25


def test_mask_data_matches_python_twin(gold):
    for c in gold["mask_sensitive_data"]:
        assert m.mask_value(c["data"], c["max_depth"]) == c["masked"]
    assert len(gold["mask_sensitive_data"]) > 500


def test_cookies_and_headers_match_twin_on_ascii(gold):
    for cookie, exp in gold["cookies"]:
        if cookie.isascii() and "\x1c" not in cookie:      # non-ASCII / \x1c: documented Rust-vs-Python deltas (SURVEY A-8)
            assert m.mask_cookie_header(cookie) == exp if cookie else True
    for c in gold["headers"]:
        if all(isinstance(v, str) and v.isascii() and "\x1c" not in v for v in c["headers"].values()) and all(k.isascii() for k in c["headers"]):
            assert m.mask_headers(c["headers"]) == c["masked"]


def test_json_bytes_format():
    f = m.mask_json_bytes
    assert f(b'{"b":1,"a":{"password":"x","n":[1,2.50,1e16,1e15,-0,0.00001,1e-6,"\\u00e9\\/"]},"b":2}') == \
        '{"a":{"n":[1,2.5,1e16,1000000000000000.0,-0.0,0.00001,1e-6,"é/"],"password":"******"},"b":2}'.encode()
    assert f(b'[18446744073709551615,18446744073709551616,-9223372036854775808,-9223372036854775809,1.0,100.0,0.1,123456789.125]') == \
        b'[18446744073709551615,1.8446744073709552e19,-9223372036854775808,-9.223372036854776e18,1.0,100.0,0.1,123456789.125]'
    assert f(b'{"level":{"nested":{}}}', 1) == b'{"level":"<nested too deep>"}'
    assert f(b'"x"', 0) == b'"<nested too deep>"'
    assert f(b' {"k":"\\u0001\\n\\"\\\\\\u007f"} ') == '{"k":"\\u0001\\n\\"\\\\\x7f"}'.encode()
    for bad in (b'{"a":1,}', b"[01]", b"NaN", b'"\\ud800"', b"\xff", b"[1e400]", b"", b"[" * 200 + b"]" * 200):
        with pytest.raises(ValueError):
            f(bad)


def test_serde_json_format_vectors():
    """Pins the bytes-API formatting (numbers through ryu's shortest round trip + pretty layout, string escapes, sorted keys, duplicate
    keys) of BOTH the oracle restatement and the code the CUDA kernel runs (host build of csrc/json_mask.h) to published vectors:
    tests/golden/serde_format.json (provenance inside)."""
    import hostsim_util as hs

    with open(os.path.join(os.path.dirname(__file__), "golden", "serde_format.json"), encoding="utf-8") as f:
        g = json.load(f)
    for src, exp in g["numbers"]:
        doc = f"[{src}]".encode()
        assert m.mask_json_bytes(doc, 10) == f"[{exp}]".encode(), (src, m.mask_json_bytes(doc, 10))
        st, out = hs.mask_host(doc, 10)
        assert st == 0 and out == f"[{exp}]".encode(), (src, st, out)
    for src in g["errors"]:
        doc = f"[{src}]".encode()
        with pytest.raises(ValueError):
            m.mask_json_bytes(doc, 10)
        assert hs.mask_host(doc, 10)[0] == 2, src
    for src, exp in g["strings"] + g["documents"]:
        assert m.mask_json_bytes(src.encode("utf-8"), 10) == exp.encode("utf-8"), src
        st, out = hs.mask_host(src.encode("utf-8"), 10)
        assert st == 0 and out == exp.encode("utf-8"), (src, st, out)
